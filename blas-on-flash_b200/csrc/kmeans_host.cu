// k-means handle: this rank's shard of the points resident in HBM across Lloyd iterations
// (drivers/kmeans.cpp:103-217, drivers/in_mem_kmeans.cpp:89-152).
#include "host_internal.cuh"

using namespace bof;

namespace {
__global__ void __launch_bounds__(256) accumulate_kernel(float* __restrict__ y, const float* __restrict__ x, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] += x[i];
}
}  // namespace

extern "C" {

// ---- k-means: points shard resident across iterations -------------------------------------------

struct bof_kmeans {
  bof_ctx* ctx;
  int64_t npoints, ncenters, dim;
  float* points;        // P x dim
  void* point_planes;   // TF32 hi/lo planes of the points (filled once)
  float* p_l2sq;        // P
  float* centers;       // K x dim
  float* c_l2sq;        // K
  float* partial;       // [K*dim sums | K counts & 4095 | K counts >> 12]: the allreduce payload, every entry exact in fp32
  int32_t* assign;      // P
  void* ws_assign; size_t ws_assign_bytes;
  void* ws_reduce; size_t ws_reduce_bytes;
  int64_t* assign64;    // P, for bof_kmeans_get
  // Out-of-core mode (the points do not fit in HBM; the reference streams them from flash every iteration,
  // drivers/kmeans.cpp:114-117,143-145): the points stay in host memory and every iteration uploads them chunk by
  // chunk through two chunk buffers; `points`, `point_planes`, `p_l2sq` then hold ONE chunk pair, not the shard.
  const float* points_host = nullptr;
  int64_t chunk = 0;               // points per chunk (0: resident mode)
  float* cpoints[2] = {};          // chunk buffers
  void* cplanes[2] = {};
  float* cp2[2] = {};
  float* chunk_partial = nullptr;  // [K*dim | K | K] of one chunk, added to `partial` in chunk order
};

static void kmeans_free(bof_kmeans* km) {
  if (!km) return;
  void* ptrs[] = {km->points, km->point_planes, km->p_l2sq, km->centers, km->c_l2sq, km->partial,
                  km->assign, km->ws_assign, km->ws_reduce, km->assign64, km->cpoints[0], km->cpoints[1],
                  km->cplanes[0], km->cplanes[1], km->cp2[0], km->cp2[1], km->chunk_partial};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete km;
}

int bof_kmeans_open(bof_ctx* ctx, int64_t npoints, int64_t ncenters, int64_t dim, const float* points_host,
                    const float* centers_host, bof_kmeans** out) {
  if (!ctx || !out) return BOF_EINVAL;
  *out = nullptr;
  BOF_REQUIRE(ctx, npoints >= 0 && ncenters > 0 && dim > 0 && npoints < (1ll << 31), "kmeans: bad dimension");
  BOF_CUDA(ctx, cudaSetDevice(ctx->device));
  bof_kmeans* km = new bof_kmeans();
  km->ctx = ctx; km->npoints = npoints; km->ncenters = ncenters; km->dim = dim;
  const size_t P = (size_t)std::max<int64_t>(npoints, 1);
  // resident footprint: points + operand planes + norms + assignments + workspaces; beyond ~70 % of HBM (or when
  // BOF_KMEANS_CHUNK=<points> forces it: test knob) the shard is streamed in chunks instead
  static const int64_t forced_chunk = getenv("BOF_KMEANS_CHUNK") ? atoll(getenv("BOF_KMEANS_CHUNK")) : 0;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); total_b = (size_t)160 << 30; }
  const double resident = (double)P * dim * 4 + (double)bof_kmeans_point_planes_bytes(npoints, dim) + (double)P * 20 +
                          (double)bof_kmeans_workspace_bytes(npoints, ncenters, dim, 0) +
                          (double)kmeans_reduce_workspace_bytes(npoints, ncenters, dim);
  if (forced_chunk > 0 && forced_chunk < npoints) km->chunk = forced_chunk;
  else if (resident > 0.7 * (double)total_b) {
    const double per_point = resident / (double)P;
    km->chunk = std::max<int64_t>(1 << 16, ((int64_t)(0.25 * (double)total_b / per_point) >> 16) << 16);
  }
  const int64_t unit = km->chunk > 0 ? km->chunk : npoints;   // points the device-side buffers are sized for
  const size_t U = (size_t)std::max<int64_t>(unit, 1);
  km->ws_assign_bytes = bof_kmeans_workspace_bytes(unit, ncenters, dim, 0);
  km->ws_reduce_bytes = kmeans_reduce_workspace_bytes(unit, ncenters, dim);
  const size_t part_bytes = ((size_t)ncenters * dim + 2 * ncenters) * 4;
  std::vector<std::pair<void**, size_t>> reqs = {
      {(void**)&km->centers, (size_t)ncenters * dim * 4}, {(void**)&km->c_l2sq, (size_t)ncenters * 4},
      {(void**)&km->partial, part_bytes}, {(void**)&km->assign, P * 4}, {&km->ws_assign, km->ws_assign_bytes},
      {&km->ws_reduce, km->ws_reduce_bytes}, {(void**)&km->assign64, P * 8}};
  if (km->chunk == 0) {
    reqs.push_back({(void**)&km->points, P * dim * 4});
    reqs.push_back({&km->point_planes, bof_kmeans_point_planes_bytes(npoints, dim)});
    reqs.push_back({(void**)&km->p_l2sq, P * 4});
  } else {
    for (int g = 0; g < 2; ++g) {
      reqs.push_back({(void**)&km->cpoints[g], U * dim * 4});
      reqs.push_back({&km->cplanes[g], bof_kmeans_point_planes_bytes(unit, dim)});
      reqs.push_back({(void**)&km->cp2[g], U * 4});
    }
    reqs.push_back({(void**)&km->chunk_partial, part_bytes});
    km->points_host = points_host;
  }
  for (auto& r : reqs) {
    if (cudaMalloc(r.first, r.second) != cudaSuccess) {
      cudaGetLastError();
      kmeans_free(km);
      return fail(ctx, BOF_ENOMEM, "kmeans: cudaMalloc of %zu bytes failed", r.second);
    }
  }
  cudaStream_t s = ctx->compute;
  auto guard = [&](int rc) { if (rc != BOF_OK) { quiesce(ctx); kmeans_free(km); } return rc; };
  if (int rc = guard(copy1d(ctx, km->centers, centers_host, (size_t)ncenters * dim * 4, cudaMemcpyHostToDevice, s))) return rc;
  if (int rc = guard(launch_row_sqnorm(ctx, s, ncenters, dim, km->centers, dim, km->c_l2sq))) return rc;
  if (km->chunk == 0) {
    if (int rc = guard(copy1d(ctx, km->points, points_host, (size_t)npoints * dim * 4, cudaMemcpyHostToDevice, s))) return rc;
    if (int rc = guard(launch_row_sqnorm(ctx, s, npoints, dim, km->points, dim, km->p_l2sq))) return rc;
    if (int rc = guard(bof_kmeans_prepare_points(ctx, s, npoints, dim, km->points, km->point_planes))) return rc;
  }
  if (cudaStreamSynchronize(s) != cudaSuccess) { kmeans_free(km); return fail(ctx, BOF_ECUDA, "kmeans: upload failed"); }
  *out = km;
  return BOF_OK;
}

int bof_kmeans_local_step(bof_kmeans* km, void** dev_partial, size_t* partial_floats) {
  if (!km) return BOF_EINVAL;
  bof_ctx* ctx = km->ctx;
  cudaStream_t s = ctx->compute;
  float* cnt_lo = km->partial + km->ncenters * km->dim;
  const size_t part_floats = (size_t)km->ncenters * km->dim + 2 * (size_t)km->ncenters;
  if (km->chunk == 0) {
    BOF_TRY(bof_kmeans_assign(ctx, s, km->npoints, km->ncenters, km->dim, km->points, km->centers, km->c_l2sq,
                              km->p_l2sq, km->assign, km->point_planes, km->ws_assign, km->ws_assign_bytes));
    BOF_TRY(launch_kmeans_reduce_ws(ctx, s, km->npoints, km->ncenters, km->dim, km->points, km->assign, km->partial,
                                    cnt_lo, km->ws_reduce, km->ws_reduce_bytes, cnt_lo + km->ncenters));
  } else {
    // out-of-core: upload chunk c+1 while chunk c is assigned and reduced; chunk partials are added in chunk order
    BOF_CUDA(ctx, cudaMemsetAsync(km->partial, 0, part_floats * 4, s));
    const int nchunk = (int)ceil_div<int64_t>(km->npoints, km->chunk);
    float* ccnt = km->chunk_partial + km->ncenters * km->dim;
    for (int c = 0; c < nchunk; ++c) {
      const int g = c & 1;
      const int64_t p0 = (int64_t)c * km->chunk, cnt = std::min(km->chunk, km->npoints - p0);
      cudaEvent_t ev_up = get_event(ctx, 110 + g), ev_free = get_event(ctx, 112 + g);
      if (c >= 2) BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->h2d, ev_free, 0));   // chunk c-2 has been consumed
      BOF_TRY(copy1d(ctx, km->cpoints[g], km->points_host + p0 * km->dim, (size_t)cnt * km->dim * 4, cudaMemcpyHostToDevice, ctx->h2d));
      BOF_CUDA(ctx, cudaEventRecord(ev_up, ctx->h2d));
      BOF_CUDA(ctx, cudaStreamWaitEvent(s, ev_up, 0));
      BOF_TRY(launch_row_sqnorm(ctx, s, cnt, km->dim, km->cpoints[g], km->dim, km->cp2[g]));
      BOF_TRY(bof_kmeans_prepare_points(ctx, s, cnt, km->dim, km->cpoints[g], km->cplanes[g]));
      BOF_TRY(bof_kmeans_assign(ctx, s, cnt, km->ncenters, km->dim, km->cpoints[g], km->centers, km->c_l2sq, km->cp2[g],
                                km->assign + p0, km->cplanes[g], km->ws_assign, km->ws_assign_bytes));
      BOF_TRY(launch_kmeans_reduce_ws(ctx, s, cnt, km->ncenters, km->dim, km->cpoints[g], km->assign + p0, km->chunk_partial,
                                      ccnt, km->ws_reduce, km->ws_reduce_bytes, ccnt + km->ncenters));
      accumulate_kernel<<<(unsigned)std::min<int64_t>(ceil_div<int64_t>((int64_t)part_floats, 256), 1024), 256, 0, s>>>(
          km->partial, km->chunk_partial, (int64_t)part_floats);
      BOF_LAUNCH_CHECK(ctx, "accumulate_kernel");
      BOF_CUDA(ctx, cudaEventRecord(ev_free, s));
    }
  }
  if (dev_partial) *dev_partial = km->partial;
  if (partial_floats) *partial_floats = part_floats;
  return BOF_OK;
}

int bof_kmeans_update(bof_kmeans* km) {
  if (!km) return BOF_EINVAL;
  float* cnt_lo = km->partial + km->ncenters * km->dim;
  return launch_kmeans_finalize(km->ctx, km->ctx->compute, km->ncenters, km->dim, km->partial, cnt_lo, km->centers,
                                km->c_l2sq, cnt_lo + km->ncenters);
}

// The one exchange step of the path: in-place NCCL sum of [sums | counts] over the ranks of this
// context's communicator, on the stream the k-means kernels run on.  No-op without a communicator / at world 1.
int bof_kmeans_allreduce(bof_kmeans* km) {
  if (!km) return BOF_EINVAL;
  if (comm_world(km->ctx) <= 1) return BOF_OK;
  return comm_allreduce_sum_f32(km->ctx, km->partial, (size_t)km->ncenters * km->dim + 2 * (size_t)km->ncenters,
                                km->ctx->compute);
}

// `iters` Lloyd iterations on the resident shard: assign + local sums, allreduce over the ranks, centroid update
// (drivers/in_mem_kmeans.cpp:89-152 per iteration).  Asynchronous: bof_kmeans_get synchronises.
int bof_kmeans_lloyd(bof_kmeans* km, int64_t iters) {
  if (!km) return BOF_EINVAL;
  for (int64_t it = 0; it < iters; ++it) {
    BOF_TRY(bof_kmeans_local_step(km, nullptr, nullptr));
    BOF_TRY(bof_kmeans_allreduce(km));
    BOF_TRY(bof_kmeans_update(km));
  }
  return BOF_OK;
}

int bof_kmeans_get(bof_kmeans* km, float* centers_host, int64_t* assign_host) {
  if (!km) return BOF_EINVAL;
  bof_ctx* ctx = km->ctx;
  cudaStream_t s = ctx->compute;
  CallGuard call_guard(ctx);
  if (centers_host) BOF_TRY(copy1d(ctx, centers_host, km->centers, (size_t)km->ncenters * km->dim * 4, cudaMemcpyDeviceToHost, s));
  if (assign_host && km->npoints > 0) {
    BOF_TRY(launch_idx_widen(ctx, s, km->assign, km->assign64, km->npoints));
    BOF_TRY(copy1d(ctx, assign_host, km->assign64, (size_t)km->npoints * 8, cudaMemcpyDeviceToHost, s));
  }
  BOF_TRY(sync_all(ctx));
  return call_guard.done();
}

void* bof_kmeans_stream(bof_kmeans* km) { return km ? (void*)km->ctx->compute : nullptr; }

int bof_kmeans_close(bof_kmeans* km) {
  if (!km) return BOF_OK;
  cudaStreamSynchronize(km->ctx->compute);
  kmeans_free(km);
  return BOF_OK;
}

}  // extern "C"
