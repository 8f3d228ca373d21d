// One process, several GPUs: the flash:: entry points of a C++ application (and drivers/*.cpp) spread over the
// GPUs of the node without Python or a launcher.  A bof_mgpu owns one context + communicator rank per device;
// every call runs one host thread per device, each driving its rank's share through the same bof_dist_* /
// bof_host_* pipelines a torchrun rank would use (north_star (5): output row blocks sharded with no collective;
// the replicated dense operand crosses PCIe once per node and is broadcast over NVLink; k-means allreduces
// centroid sums and counts).  Host operands are shared by all threads: one copy of B in host memory instead of N.
#include "host_internal.cuh"

#include <thread>

using namespace bof;

extern "C" {
int bof_comm_unique_id(void* id_out);
int bof_comm_init(bof_ctx* ctx, int world, int rank, const void* id);
int bof_dist_gemm(bof_ctx* ctx, char ta, char tb, int64_t m_local, int64_t n, int64_t k, float alpha, float beta,
                  const float* a_local, const float* b, float* c_local, int64_t lda, int64_t ldb, int64_t ldc);
int bof_dist_csrmm(bof_ctx* ctx, int64_t m_local, int64_t n, int64_t k, float alpha, float beta, const float* a,
                   const int64_t* ia, const int64_t* ja, const float* b, float* c_local);
}

struct bof_mgpu {
  std::vector<bof_ctx*> ctx;
  std::string err;
};

namespace {

// run fn(rank) on one thread per device; first failure wins
template <class F>
int for_each_rank(bof_mgpu* mg, F fn) {
  const int world = (int)mg->ctx.size();
  std::vector<int> rc((size_t)world, BOF_OK);
  std::vector<std::thread> th;
  for (int r = 1; r < world; ++r) th.emplace_back([&, r] { rc[r] = fn(r); });
  rc[0] = fn(0);
  for (auto& t : th) t.join();
  for (int r = 0; r < world; ++r)
    if (rc[r] != BOF_OK) {
      mg->err = "rank " + std::to_string(r) + ": " + bof_last_error(mg->ctx[r]);
      return rc[r];
    }
  return BOF_OK;
}

// rows [r0, r1) of rank r: equal shares rounded up to `align` rows
void split_rows(int64_t m, int world, int r, int64_t align, int64_t* r0, int64_t* r1) {
  int64_t per = ceil_div<int64_t>(m, world);
  per = round_up<int64_t>(std::max<int64_t>(per, 1), align);
  *r0 = std::min(m, (int64_t)r * per);
  *r1 = std::min(m, *r0 + per);
}

// rows [r0, r1) of rank r holding ~nnz / world nonzeros (the nnz-budget idea of get_next_blk_size,
// include/blas_utils.h:72-82, at the granularity of a GPU)
void split_nnz(const int64_t* ia, int64_t m, int world, int r, int64_t* r0, int64_t* r1) {
  const int64_t nnz = ia[m] - ia[0];
  auto cut = [&](int g) -> int64_t {
    if (g <= 0) return 0;
    if (g >= world) return m;
    const int64_t target = ia[0] + (int64_t)((__int128)nnz * g / world);
    return std::lower_bound(ia, ia + m + 1, target) - ia;
  };
  *r0 = std::min(cut(r), m);
  *r1 = std::max(*r0, std::min(cut(r + 1), m));
}

}  // namespace

extern "C" {

int bof_mgpu_create(const bof_config* cfg, int ndev, const int* devices, bof_mgpu** out) {
  if (!out) return BOF_EINVAL;
  *out = nullptr;
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) { cudaGetLastError(); return BOF_ENODEV; }
  if (ndev <= 0) ndev = have;
  bof_mgpu* mg = new bof_mgpu();
  for (int i = 0; i < ndev; ++i) {
    bof_config c{};
    if (cfg) c = *cfg;
    c.device = devices ? devices[i] : i;
    bof_ctx* x = nullptr;
    const int rc = bof_ctx_create(&c, &x);
    if (rc != BOF_OK) {
      for (bof_ctx* y : mg->ctx) bof_ctx_destroy(y);
      delete mg;
      return rc;
    }
    mg->ctx.push_back(x);
  }
  if (ndev > 1) {
    char id[128];
    int rc = bof_comm_unique_id(id);
    if (rc == BOF_OK) rc = for_each_rank(mg, [&](int r) { return bof_comm_init(mg->ctx[r], ndev, r, id); });
    if (rc != BOF_OK) {
      for (bof_ctx* y : mg->ctx) bof_ctx_destroy(y);
      delete mg;
      return rc;
    }
  }
  *out = mg;
  return BOF_OK;
}

int bof_mgpu_destroy(bof_mgpu* mg) {
  if (!mg) return BOF_OK;
  for (bof_ctx* x : mg->ctx) bof_ctx_destroy(x);
  delete mg;
  return BOF_OK;
}

int bof_mgpu_count(const bof_mgpu* mg) { return mg ? (int)mg->ctx.size() : 0; }
bof_ctx* bof_mgpu_ctx(bof_mgpu* mg, int rank) { return mg && rank >= 0 && rank < (int)mg->ctx.size() ? mg->ctx[rank] : nullptr; }
const char* bof_mgpu_last_error(const bof_mgpu* mg) { return mg ? mg->err.c_str() : ""; }

// flash::gemm over the GPUs: rows of C (row-major) / columns of C (column-major) sharded, the other operand
// replicated by panel broadcast.
int bof_mgpu_gemm(bof_mgpu* mg, char ord, char ta, char tb, int64_t m, int64_t n, int64_t k, float alpha, float beta,
                  const float* a, const float* b, float* c, int64_t lda, int64_t ldb, int64_t ldc) {
  if (!mg || mg->ctx.empty()) return BOF_EINVAL;
  const int world = (int)mg->ctx.size();
  if (world == 1) return bof_host_gemm(mg->ctx[0], ord, ta, tb, m, n, k, alpha, beta, a, b, c, lda, ldb, ldc);
  if (!is_rc(ord) || !is_nt(ta) || !is_nt(tb)) { mg->err = "gemm: mat_ord must be 'R'/'C', trans 'N'/'T'"; return BOF_EINVAL; }
  // column-major gemm(ta, tb, m, n, k, A, B, C) is row-major gemm(tb, ta, n, m, k, B, A, C)
  if (ord == 'C') { std::swap(ta, tb); std::swap(m, n); std::swap(a, b); std::swap(lda, ldb); }
  // row-major from here on: op(A) is m x k, its tight leading dimension k ('N') or m ('T')
  if (lda == 0) lda = ta == 'N' ? k : m;
  if (ldb == 0) ldb = tb == 'N' ? n : k;
  if (ldc == 0) ldc = n;
  return for_each_rank(mg, [&](int r) {
    int64_t r0, r1;
    split_rows(m, world, r, 256, &r0, &r1);
    const float* a_loc = ta == 'N' ? a + r0 * lda : a + r0;   // 'T': a column range of the stored k x m matrix
    return bof_dist_gemm(mg->ctx[r], ta, tb, r1 - r0, n, k, alpha, beta, a_loc, b, c + r0 * ldc, lda, ldb, ldc);
  });
}

// flash::csrmm over the GPUs ('N', row-major: nnz-balanced row blocks per GPU; anything else runs on GPU 0)
int bof_mgpu_csrmm(bof_mgpu* mg, char trans_a, int64_t m, int64_t n, int64_t k, float alpha, float beta, const float* a,
                   const int64_t* ia, const int64_t* ja, char ord_b, const float* b, float* c) {
  if (!mg || mg->ctx.empty()) return BOF_EINVAL;
  const int world = (int)mg->ctx.size();
  if (world == 1 || trans_a != 'N' || ord_b != 'R' || m < world)
    return bof_host_csrmm(mg->ctx[0], trans_a, m, n, k, alpha, beta, a, ia, ja, ord_b, b, c);
  return for_each_rank(mg, [&](int r) {
    int64_t r0, r1;
    split_nnz(ia, m, world, r, &r0, &r1);
    const int64_t z0 = ia[r0] - ia[0];
    return bof_dist_csrmm(mg->ctx[r], r1 - r0, n, k, alpha, beta, a + z0, ia + r0, ja + z0, b, c + r0 * k);
  });
}

// flash::csrgemv over the GPUs.  'N': row blocks, disjoint y.  'T': every GPU produces a full-length partial y from
// its row block; the partials are added on the host in GPU order (deterministic; the reference adds task results
// in completion order under a mutex, include/tasks/csrgemv_task.h:170-176).
int bof_mgpu_csrgemv(bof_mgpu* mg, char trans_a, int64_t m, int64_t n, const float* a, const int64_t* ia,
                     const int64_t* ja, const float* x, float* y) {
  if (!mg || mg->ctx.empty()) return BOF_EINVAL;
  const int world = (int)mg->ctx.size();
  if (world == 1 || !is_nt(trans_a) || m < world) return bof_host_csrgemv(mg->ctx[0], trans_a, m, n, a, ia, ja, x, y);
  std::vector<std::vector<float>> part;
  if (trans_a == 'T') part.assign((size_t)world, std::vector<float>());
  int rc = for_each_rank(mg, [&](int r) {
    int64_t r0, r1;
    split_nnz(ia, m, world, r, &r0, &r1);
    const int64_t z0 = ia[r0] - ia[0];
    if (trans_a == 'N') return bof_host_csrgemv(mg->ctx[r], 'N', r1 - r0, n, a + z0, ia + r0, ja + z0, x, y + r0);
    float* dst = y;
    if (r > 0) { part[r].resize((size_t)n); dst = part[r].data(); }
    return bof_host_csrgemv(mg->ctx[r], 'T', r1 - r0, n, a + z0, ia + r0, ja + z0, x + r0, dst);
  });
  if (rc != BOF_OK || trans_a == 'N') return rc;
  for (int r = 1; r < world; ++r)
    for (int64_t j = 0; j < n; ++j) y[j] += part[r][j];
  return BOF_OK;
}

// `iters` Lloyd iterations (drivers/in_mem_kmeans.cpp:89-152), points sharded over the GPUs and resident, one NCCL
// allreduce of [sums | counts] per iteration.  centers_host: initial centres in, final centres out; assign_out
// (npoints entries, may be NULL): assignment of the last iteration.
int bof_mgpu_kmeans_lloyd(bof_mgpu* mg, int64_t npoints, int64_t ncenters, int64_t dim, const float* points_host,
                          float* centers_host, int64_t iters, int64_t* assign_out) {
  if (!mg || mg->ctx.empty()) return BOF_EINVAL;
  const int world = (int)mg->ctx.size();
  std::vector<float> c0((size_t)ncenters * dim);
  memcpy(c0.data(), centers_host, c0.size() * 4);   // every rank starts from the same centres; rank 0 writes the result
  return for_each_rank(mg, [&](int r) {
    int64_t p0, p1;
    split_rows(npoints, world, r, 1, &p0, &p1);
    bof_kmeans* km = nullptr;
    int rc = bof_kmeans_open(mg->ctx[r], p1 - p0, ncenters, dim, points_host + p0 * dim, c0.data(), &km);
    if (rc == BOF_OK) rc = bof_kmeans_lloyd(km, iters);
    if (rc == BOF_OK) rc = bof_kmeans_get(km, r == 0 ? centers_host : nullptr, assign_out ? assign_out + p0 : nullptr);
    if (km) bof_kmeans_close(km);
    return rc;
  });
}

}  // extern "C"
