// k-means handle: this rank's shard of the points resident in HBM across Lloyd iterations
// (drivers/kmeans.cpp:103-217, drivers/in_mem_kmeans.cpp:89-152).
#include "host_internal.cuh"

using namespace bof;

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
};

static void kmeans_free(bof_kmeans* km) {
  if (!km) return;
  void* ptrs[] = {km->points, km->point_planes, km->p_l2sq, km->centers, km->c_l2sq, km->partial,
                  km->assign, km->ws_assign, km->ws_reduce, km->assign64};
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
  km->ws_assign_bytes = bof_kmeans_workspace_bytes(npoints, ncenters, dim, 0);
  km->ws_reduce_bytes = kmeans_reduce_workspace_bytes(npoints, ncenters, dim);
  struct Req { void** p; size_t bytes; } reqs[] = {
      {(void**)&km->points, P * dim * 4}, {&km->point_planes, bof_kmeans_point_planes_bytes(npoints, dim)},
      {(void**)&km->p_l2sq, P * 4}, {(void**)&km->centers, (size_t)ncenters * dim * 4},
      {(void**)&km->c_l2sq, (size_t)ncenters * 4}, {(void**)&km->partial, ((size_t)ncenters * dim + 2 * ncenters) * 4},
      {(void**)&km->assign, P * 4}, {&km->ws_assign, km->ws_assign_bytes}, {&km->ws_reduce, km->ws_reduce_bytes},
      {(void**)&km->assign64, P * 8}};
  for (auto& r : reqs) {
    if (cudaMalloc(r.p, r.bytes) != cudaSuccess) {
      cudaGetLastError();
      kmeans_free(km);
      return fail(ctx, BOF_ENOMEM, "kmeans: cudaMalloc of %zu bytes failed", r.bytes);
    }
  }
  cudaStream_t s = ctx->compute;
  auto guard = [&](int rc) { if (rc != BOF_OK) { quiesce(ctx); kmeans_free(km); } return rc; };
  if (int rc = guard(copy1d(ctx, km->points, points_host, (size_t)npoints * dim * 4, cudaMemcpyHostToDevice, s))) return rc;
  if (int rc = guard(copy1d(ctx, km->centers, centers_host, (size_t)ncenters * dim * 4, cudaMemcpyHostToDevice, s))) return rc;
  if (int rc = guard(launch_row_sqnorm(ctx, s, npoints, dim, km->points, dim, km->p_l2sq))) return rc;
  if (int rc = guard(launch_row_sqnorm(ctx, s, ncenters, dim, km->centers, dim, km->c_l2sq))) return rc;
  if (int rc = guard(bof_kmeans_prepare_points(ctx, s, npoints, dim, km->points, km->point_planes))) return rc;
  if (cudaStreamSynchronize(s) != cudaSuccess) { kmeans_free(km); return fail(ctx, BOF_ECUDA, "kmeans: upload failed"); }
  *out = km;
  return BOF_OK;
}

int bof_kmeans_local_step(bof_kmeans* km, void** dev_partial, size_t* partial_floats) {
  if (!km) return BOF_EINVAL;
  bof_ctx* ctx = km->ctx;
  cudaStream_t s = ctx->compute;
  BOF_TRY(bof_kmeans_assign(ctx, s, km->npoints, km->ncenters, km->dim, km->points, km->centers, km->c_l2sq,
                            km->p_l2sq, km->assign, km->point_planes, km->ws_assign, km->ws_assign_bytes));
  float* cnt_lo = km->partial + km->ncenters * km->dim;
  BOF_TRY(launch_kmeans_reduce_ws(ctx, s, km->npoints, km->ncenters, km->dim, km->points, km->assign, km->partial,
                                  cnt_lo, km->ws_reduce, km->ws_reduce_bytes, cnt_lo + km->ncenters));
  if (dev_partial) *dev_partial = km->partial;
  if (partial_floats) *partial_floats = (size_t)km->ncenters * km->dim + 2 * (size_t)km->ncenters;
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
