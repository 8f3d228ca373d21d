// TEST INFRASTRUCTURE — entry point that calls the reference's own flash::kmeans (src/blas/kmeans.cpp:27-198 ->
// KMeansTask::execute, include/tasks/kmeans_task.h:53-82; linked unmodified from oracle/_ref/libfblas.a) with the
// argument form of its only call site, drivers/kmeans.cpp:37-39, and writes the distance matrix it produces.
// The reference's drivers/kmeans.cpp itself cannot run as shipped: it never calls flash::flash_setup and issues
// flash::read_sync from OpenMP worker threads that have no AIO context (drivers/kmeans.cpp:80-86); the complete
// Lloyd iteration is pinned by drivers/in_mem_kmeans.cpp instead (oracle/_ref/in_mem_kmeans_driver).
//   ref_kmeans_dist <points> <centers> <dist_out> <npoints> <ndims> <ncenters>
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

#include "flash_blas.h"
#include "lib_funcs.h"

int main(int argc, char** argv) {
  if (argc != 7) {
    std::fprintf(stderr, "usage: %s points centers dist_out npoints ndims ncenters\n", argv[0]);
    return 2;
  }
  std::string fp = argv[1], fc = argv[2], fd = argv[3];
  FBLAS_UINT npoints = std::stoull(argv[4]), ndims = std::stoull(argv[5]), ncenters = std::stoull(argv[6]);
  std::vector<FPTYPE> pts(npoints * ndims), ctr(ncenters * ndims);
  std::ifstream in(fp, std::ios::binary);
  in.read((char*) pts.data(), pts.size() * sizeof(FPTYPE));
  in.close();
  in.open(fc, std::ios::binary);
  in.read((char*) ctr.data(), ctr.size() * sizeof(FPTYPE));
  in.close();
  std::vector<FPTYPE> p2(npoints), c2(ncenters), ones(std::max(npoints, ncenters), (FPTYPE) 1.0);
  for (FBLAS_UINT p = 0; p < npoints; p++) p2[p] = mkl_dot(ndims, &pts[p * ndims], 1, &pts[p * ndims], 1);
  for (FBLAS_UINT c = 0; c < ncenters; c++) c2[c] = mkl_dot(ndims, &ctr[c * ndims], 1, &ctr[c * ndims], 1);

  flash::flash_setup("./");
  auto points = flash::map_file<FPTYPE>(fp, flash::Mode::READWRITE);
  auto centers = flash::map_file<FPTYPE>(fc, flash::Mode::READWRITE);
  auto dist = flash::map_file<FPTYPE>(fd, flash::Mode::READWRITE);
  FBLAS_INT ret = flash::kmeans('C', 'T', 'N', ncenters, npoints, ndims, (FPTYPE) -2.0, (FPTYPE) 0.0, centers, points,
                                dist, ndims, ndims, ncenters, c2.data(), p2.data(), ones.data());
  // flash::kmeans returns with the result tiles still in the scheduler's cache (unlike flash::gemm, which ends with
  // sched.flush_cache(), src/blas/gemm.cpp:200); write them back before the file is read
  flash::sched.flush_cache();
  flash::unmap_file(points);
  flash::unmap_file(centers);
  flash::unmap_file(dist);
  flash::flash_destroy();
  return (int) ret;
}
