// Microbenchmark: strided (2-D) PCIe copies between pinned host memory and HBM -- cudaMemcpy2DAsync on the copy
// engine vs an SM kernel that reads / writes the mapped host memory directly -- for the column-panel transfers of
// bof_csr_mm (row segments of `width` bytes out of rows `pitch` bytes apart).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/pcie2d.cu -o /tmp/pcie2d && /tmp/pcie2d
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// one warp per row segment of `w4` float4s; to_host selects the direction
__global__ void strided_copy(const float4* __restrict__ src, size_t spitch4, float4* __restrict__ dst, size_t dpitch4,
                             int w4, size_t rows) {
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (size_t r = warp; r < rows; r += nwarps)
    for (int j = lane; j < w4; j += 32) dst[r * dpitch4 + j] = src[r * spitch4 + j];
}

int main() {
  const size_t total = 2ull << 30;  // bytes moved per measurement
  float *h, *d;
  CK(cudaHostAlloc(&h, 2 * total, cudaHostAllocMapped));
  CK(cudaMalloc(&d, total));
  CK(cudaMemset(d, 0, total));
  cudaStream_t s; CK(cudaStreamCreate(&s));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto timeit = [&](auto fn) { fn(); CK(cudaStreamSynchronize(s)); CK(cudaEventRecord(e0, s)); fn(); fn(); CK(cudaEventRecord(e1, s));
                               CK(cudaStreamSynchronize(s)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return total * 2 / (ms * 1e-3) / 1e9; };
  printf("contiguous H2D %.1f GB/s, D2H %.1f GB/s\n",
         timeit([&] { CK(cudaMemcpyAsync(d, h, total, cudaMemcpyHostToDevice, s)); }),
         timeit([&] { CK(cudaMemcpyAsync(h, d, total, cudaMemcpyDeviceToHost, s)); }));
  for (size_t width : {128, 256, 512, 1024, 2048, 4096}) {
    const size_t pitch = 2 * width, rows = total / width;
    const double ce_h2d = timeit([&] { CK(cudaMemcpy2DAsync(d, width, h, pitch, width, rows, cudaMemcpyHostToDevice, s)); });
    const double ce_d2h = timeit([&] { CK(cudaMemcpy2DAsync(h, pitch, d, width, width, rows, cudaMemcpyDeviceToHost, s)); });
    double k_h2d[3], k_d2h[3];
    int gi = 0;
    for (int blocks : {148, 148 * 4, 148 * 8}) {
      k_h2d[gi] = timeit([&] { strided_copy<<<blocks, 256, 0, s>>>((const float4*)h, pitch / 16, (float4*)d, width / 16, (int)(width / 16), rows); });
      k_d2h[gi] = timeit([&] { strided_copy<<<blocks, 256, 0, s>>>((const float4*)d, width / 16, (float4*)h, pitch / 16, (int)(width / 16), rows); });
      ++gi;
    }
    printf("width %5zu B pitch %5zu B: copy engine H2D %.1f D2H %.1f | SM kernel (148/592/1184 blocks) H2D %.1f %.1f %.1f  D2H %.1f %.1f %.1f GB/s\n",
           width, pitch, ce_h2d, ce_d2h, k_h2d[0], k_h2d[1], k_h2d[2], k_d2h[0], k_d2h[1], k_d2h[2]);
  }
  // both directions at once, 512-byte segments: copy engine vs kernels
  {
    const size_t width = 512, pitch = 1024, rows = total / 2 / width;
    cudaStream_t s2; CK(cudaStreamCreate(&s2));
    float* d2 = d + total / 8;  // second half of the device buffer (floats)
    float* h2 = h + total / 4;
    auto both = [&](bool kernel) {
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0, s));
      for (int it = 0; it < 2; ++it) {
        if (kernel) {
          strided_copy<<<592, 256, 0, s>>>((const float4*)h, pitch / 16, (float4*)d, width / 16, (int)(width / 16), rows);
          strided_copy<<<592, 256, 0, s2>>>((const float4*)d2, width / 16, (float4*)h2, pitch / 16, (int)(width / 16), rows);
        } else {
          CK(cudaMemcpy2DAsync(d, width, h, pitch, width, rows, cudaMemcpyHostToDevice, s));
          CK(cudaMemcpy2DAsync(h2, pitch, d2, width, width, rows, cudaMemcpyDeviceToHost, s2));
        }
      }
      CK(cudaStreamSynchronize(s2));
      CK(cudaEventRecord(e1, s));
      CK(cudaStreamSynchronize(s));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      return (double)total * 2 / (ms * 1e-3) / 1e9;  // bytes in both directions together
    };
    printf("duplex 512 B segments: copy engine %.1f GB/s total, SM kernels %.1f GB/s total\n", both(false), both(true));
  }
  return 0;
}
