// flash::gemm / flash::kmeans on the host side: canonical form of the 8 layout cases, the device-resident GEMM
// with a caller workspace, and the host pipeline (panelled upload of the resident operand, ring of row blocks,
// slab drain).  Reference: src/blas/gemm.cpp:27-202, include/tasks/gemm_task.h:67-93.
#include "host_internal.cuh"

namespace bof {

// Leading-dimension defaults and the row/col role table of src/blas/gemm.cpp:52-67.
int canon_gemm(bof_ctx* ctx, char ord, char ta, char tb, int64_t m, int64_t n, int64_t k,
               const float* A, int64_t lda, const float* B, int64_t ldb, int64_t ldc, Canon* out) {
  BOF_REQUIRE(ctx, is_rc(ord), "gemm: mat_ord must be 'R' or 'C' (got '%c')", ord);
  BOF_REQUIRE(ctx, is_nt(ta), "gemm: trans_a must be 'N' or 'T' (got '%c')", ta);
  BOF_REQUIRE(ctx, is_nt(tb), "gemm: trans_b must be 'N' or 'T' (got '%c')", tb);
  BOF_REQUIRE(ctx, m >= 0 && n >= 0 && k >= 0, "gemm: negative dimension");
  const bool col = ord == 'C', tA = ta == 'T', tB = tb == 'T';
  const int64_t a_cols = (tA != col) ? m : k;  // contiguous extent of A as stored
  const int64_t b_cols = (tB != col) ? k : n;
  const int64_t c_cols = col ? m : n;
  if (lda == 0) lda = a_cols;
  if (ldb == 0) ldb = b_cols;
  if (ldc == 0) ldc = c_cols;
  BOF_REQUIRE(ctx, lda >= a_cols && ldb >= b_cols && ldc >= c_cols, "gemm: leading dimension too small");
  // op(A)(i, kk) = A[i*a_si + kk*a_sk]: contiguous in kk iff A is stored with k as its inner extent
  const int64_t a_si = (tA == col) ? lda : 1, a_sk = (tA == col) ? 1 : lda;
  // op(B)(kk, j) = B[j*b_sj + kk*b_sk]: contiguous in kk iff B is stored with k as its inner extent
  const int64_t b_sj = (tB != col) ? ldb : 1, b_sk = (tB != col) ? 1 : ldb;
  Canon c{};
  c.K = k;
  c.ldc = ldc;
  if (!col) {  // C[i*ldc + j]
    c.Mo = m; c.No = n;
    c.psrc = A; c.p_sr = a_si; c.p_sk = a_sk;
    c.qsrc = B; c.q_sr = b_sj; c.q_sk = b_sk;
  } else {     // column-major C is the row-major transpose: C^T = op(B)^T op(A)^T
    c.Mo = n; c.No = m;
    c.psrc = B; c.p_sr = b_sj; c.p_sk = b_sk;
    c.qsrc = A; c.q_sr = a_si; c.q_sk = a_sk;
  }
  *out = c;
  return BOF_OK;
}

int64_t padded_k(int64_t k) { return std::max<int64_t>(32, round_up<int64_t>(k, 32)); }
size_t plane_bytes(int64_t rows, int64_t kp) { return round_up<size_t>((size_t)rows * kp * 4, 256); }

int pick_gemm_path(const bof_ctx* ctx, int64_t Mo, int64_t No, int64_t K) {
  if (ctx->cfg.gemm_force_path) return ctx->cfg.gemm_force_path;
  if ((double)Mo * No * K < 2e6) return 3;  // launch-latency territory: CUDA cores, no planes
  return 2;
}

int64_t k_chunk_of(const bof_ctx* ctx) {
  if (ctx->cfg.gemm_k_chunk < 0) return 0;
  return ctx->cfg.gemm_k_chunk == 0 ? 256 : ctx->cfg.gemm_k_chunk;
}

// GEMM on device-resident canonical operands with a caller-provided plane workspace.
int gemm_canon_device(bof_ctx* ctx, cudaStream_t s, const Canon& c, float alpha, float beta, float* C,
                      void* ws, size_t ws_bytes) {
  if (c.Mo == 0 || c.No == 0) return BOF_OK;
  const int path = pick_gemm_path(ctx, c.Mo, c.No, c.K);
  if (c.K == 0 || path == 3) {
    // K == 0 degenerates to C = beta*C, which the CUDA-core kernel handles as well
    return launch_gemm_ffma(ctx, s, c.Mo, c.No, c.K, alpha, c.psrc, c.p_sr, c.p_sk, c.qsrc, c.q_sk, c.q_sr,
                            beta, C, c.ldc);
  }
  const int64_t kp = padded_k(c.K);
  const size_t pb = plane_bytes(c.Mo, kp), qb = plane_bytes(c.No, kp);
  BOF_REQUIRE(ctx, ws != nullptr && ws_bytes >= 2 * pb + 2 * qb + 256, "gemm: workspace too small");
  uint8_t* base = reinterpret_cast<uint8_t*>(round_up<uintptr_t>(reinterpret_cast<uintptr_t>(ws), 256));
  float* p_hi = reinterpret_cast<float*>(base);
  float* p_lo = reinterpret_cast<float*>(base + pb);
  float* q_hi = reinterpret_cast<float*>(base + 2 * pb);
  float* q_lo = reinterpret_cast<float*>(base + 2 * pb + qb);
  BOF_TRY(launch_split_planes(ctx, s, c.Mo, c.K, c.psrc, c.p_sr, c.p_sk, p_hi, p_lo, kp));
  BOF_TRY(launch_split_planes(ctx, s, c.No, c.K, c.qsrc, c.q_sr, c.q_sk, q_hi, q_lo, kp));
  GemmEpilogue ep;
  ep.alpha = alpha; ep.beta = beta; ep.C = C; ep.ldc = c.ldc;
  return launch_gemm_tc(ctx, s, path == 1 ? 1 : 2, c.Mo, c.No, c.K, kp, p_hi, p_lo, q_hi, q_lo, ep, k_chunk_of(ctx));
}


}  // namespace bof

using namespace bof;

extern "C" {

// flash::gemm.  The canonical Q operand (op(B)^T for row-major problems) is uploaded and split
// into TF32 planes once; the canonical P operand and the output stream in row blocks,
// double-buffered (upload / split + MMA / download overlap).  The reference's k-dimension
// accumulate chain (src/blas/gemm.cpp:114-126) is an I/O artefact: the whole k extent is reduced
// on the device, so each C block crosses PCIe once.
static int host_gemm_impl(bof_ctx* ctx, char ord, char ta, char tb, int64_t m, int64_t n, int64_t k, float alpha,
                          float beta, const float* a, const float* b, float* c, int64_t lda, int64_t ldb, int64_t ldc,
                          bool q_on_device, const float* term_m = nullptr, const float* term_n = nullptr,
                          bool dist_q = false) {
  if (!ctx) return BOF_EINVAL;
  Canon cn;
  BOF_TRY(canon_gemm(ctx, ord, ta, tb, m, n, k, a, lda, b, ldb, ldc, &cn));
  BOF_REQUIRE(ctx, !q_on_device || ord == 'R', "gemm: a device-resident B is supported for mat_ord='R' only");
  BOF_CUDA(ctx, cudaSetDevice(ctx->device));
  stats_begin(ctx);
  CallGuard call_guard(ctx);
  // A collective call (bof_dist_gemm) must take the same decisions on every rank: N and K are common, the local M
  // is not -- a rank whose shard is empty or tiny still owns panels of Q and takes part in the exchange.
  const bool dist_call = dist_q && comm_world(ctx) > 1;
  if (cn.No == 0 || (cn.Mo == 0 && !dist_call)) { stats_end(ctx); return call_guard.done(); }
  const cudaMemcpyKind H2D = cudaMemcpyHostToDevice;
  const int path = dist_call ? (ctx->cfg.gemm_force_path ? ctx->cfg.gemm_force_path : 2) : pick_gemm_path(ctx, cn.Mo, cn.No, cn.K);
  const int64_t K = cn.K, kp = padded_k(K);
  const bool tensor = (K > 0 && path != 3);
  const int cg = path == 1 ? 1 : 2;

  // upload rows [r0, r1) of a canonical operand as a tight device matrix; returns its strides
  auto upload_rows = [&](const float* src, int64_t s_r, int64_t s_k, int64_t r0, int64_t r1, float* dst,
                         int64_t* d_sr, int64_t* d_sk, cudaStream_t s) -> int {
    const int64_t rows = r1 - r0;
    if (K == 0) { *d_sr = 1; *d_sk = 1; return BOF_OK; }
    if (s_k == 1) {  // rows contiguous in k
      *d_sr = K; *d_sk = 1;
      return copy2d(ctx, dst, (size_t)K * 4, src + r0 * s_r, (size_t)s_r * 4, (size_t)K * 4, (size_t)rows, H2D, s);
    }
    *d_sr = 1; *d_sk = rows;  // stored k-major: a column range of a [K x ld] matrix
    return copy2d(ctx, dst, (size_t)rows * 4, src + r0, (size_t)s_k * 4, (size_t)rows * 4, (size_t)K, H2D, s);
  };

  // ---- buffers ----
  float* qraw = nullptr;
  float* q_hi = nullptr;
  float* q_lo = nullptr;
  const bool dist_mode = dist_q && comm_world(ctx) > 1 && tensor;   // Q lives in the exchange buffer the peers push into
  if (q_on_device) qraw = const_cast<float*>(cn.qsrc);  // B already in HBM in its source layout; read-only here
  else if (dist_mode) BOF_TRY(comm_exchange_begin(ctx, (size_t)cn.No * std::max<int64_t>(K, 1) * sizeof(float), &qraw));
  else BOF_TRY(slot_reserve(ctx, S_DENSE, (size_t)cn.No * std::max<int64_t>(K, 1), &qraw));
  static const int64_t n_q_panels_env = getenv("BOF_GEMM_QPANELS") ? std::min(16, std::max(1, atoi(getenv("BOF_GEMM_QPANELS")))) : 0;
  const int64_t n_q_panels = n_q_panels_env ? n_q_panels_env : (dist_mode ? (comm_world(ctx) >= 8 ? 16 : 8) : 8);
  if (dist_mode && cn.Mo == 0) {
    // No rows of C on this rank: upload and push the Q panels it owns, and wait for the peers' panels so that no push
    // into this rank's exchange buffer is still in flight when the call returns.
    const int world = comm_world(ctx), rank = comm_rank(ctx);
    const int64_t rows_per = std::max<int64_t>(256, round_up<int64_t>(ceil_div<int64_t>(cn.No, n_q_panels), 256));
    const int npan = (int)ceil_div<int64_t>(cn.No, rows_per);
    constexpr int EV_OWN = 40;
    for (int j = 0; j < npan; ++j) {
      const int64_t n0 = (int64_t)j * rows_per, n1 = std::min(cn.No, n0 + rows_per);
      if (j % world == rank) {
        int64_t sr, sk;
        BOF_TRY(upload_rows(cn.qsrc, cn.q_sr, cn.q_sk, n0, n1, qraw + n0 * K, &sr, &sk, ctx->h2d));
        BOF_CUDA(ctx, cudaEventRecord(get_event(ctx, EV_OWN + j), ctx->h2d));
        BOF_TRY(comm_push(ctx, (size_t)(n0 * K), (size_t)(n1 - n0) * K, j, get_event(ctx, EV_OWN + j)));
      } else {
        BOF_TRY(comm_wait_item(ctx, ctx->compute, j));
      }
    }
    BOF_TRY(sync_all(ctx));
    stats_end(ctx);
    return call_guard.done();
  }
  if (cn.Mo == 0) { stats_end(ctx); return call_guard.done(); }   // collective call without an exchange (K == 0 or CUDA-core path)
  const size_t qb = plane_bytes(cn.No, kp);
  if (tensor) {
    void* p;
    BOF_TRY(slot_reserve(ctx, S_DENSE_T, 2 * qb, &p));
    q_hi = static_cast<float*>(p);
    q_lo = reinterpret_cast<float*>(static_cast<uint8_t*>(p) + qb);
  }
  int64_t rb = (int64_t)ctx->cfg.gemm_row_block;
  rb = std::max<int64_t>(256, round_up<int64_t>(rb, 256));
  rb = std::min(rb, round_up<int64_t>(cn.Mo, 256));
  // a rank's shard of a multi-GPU job can be a single block: cut it into four so that the first MMAs start after a
  // quarter of it has landed and the C slabs leave earlier
  if (dist_q && comm_world(ctx) > 1 && cn.Mo <= 4 * rb) rb = std::max<int64_t>(1024, round_up<int64_t>(ceil_div<int64_t>(cn.Mo, 4), 256));
  const int nblk = (int)ceil_div<int64_t>(cn.Mo, rb);
  const size_t pb = plane_bytes(rb, kp);
  constexpr int NB = kGemmRing;
  const int ngen = std::min(NB, nblk);
  // Ring of block generations.  Planes and C blocks of the ring are each ONE allocation (generation g at
  // row g*rb), so that consecutive generations can be multiplied in a single launch during the prologue.
  float* praw[NB];
  float* hi_all = nullptr; float* lo_all = nullptr; float* c_all = nullptr;
  for (int g = 0; g < ngen; ++g)
    BOF_TRY(slot_reserve(ctx, S_PRAW + g, (size_t)rb * std::max<int64_t>(K, 1), &praw[g]));
  if (tensor) {
    void* p;
    BOF_TRY(slot_reserve(ctx, S_PPLANES, 2 * (size_t)ngen * pb, &p));
    hi_all = static_cast<float*>(p);
    lo_all = reinterpret_cast<float*>(static_cast<uint8_t*>(p) + (size_t)ngen * pb);
  }
  BOF_TRY(slot_reserve(ctx, S_GCBLK, (size_t)ngen * rb * cn.No, &c_all));
  auto p_hi_of = [&](int g) { return hi_all + (size_t)g * rb * kp; };
  auto p_lo_of = [&](int g) { return lo_all + (size_t)g * rb * kp; };
  auto cblk_of = [&](int g) { return c_all + (size_t)g * rb * cn.No; };

  // flash::kmeans: C(i, j) += term_m[i], then += term_n[j] (i over m, j over n), applied to each block on the
  // device before it is downloaded.  In canonical (row-major output) form the rows are m for 'R', n for 'C'.
  float* term_rows_d = nullptr;
  float* term_cols_d = nullptr;
  const bool with_terms = term_m != nullptr && term_n != nullptr;
  const bool canon_rows_are_m = ord == 'R';
  if (with_terms) {
    float* t;
    BOF_TRY(slot_reserve(ctx, S_MISC, (size_t)(cn.Mo + cn.No), &t));
    term_rows_d = t;
    term_cols_d = t + cn.Mo;
    BOF_TRY(copy1d(ctx, term_rows_d, canon_rows_are_m ? term_m : term_n, (size_t)cn.Mo * 4, H2D, ctx->h2d));
    BOF_TRY(copy1d(ctx, term_cols_d, canon_rows_are_m ? term_n : term_m, (size_t)cn.No * 4, H2D, ctx->h2d));
    BOF_CUDA(ctx, cudaEventRecord(get_event(ctx, 3), ctx->h2d));
    BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, get_event(ctx, 3), 0));
  }

  // events: 8+g P block uploaded, 16+g P block split, 24+g block computed, 32+g block downloaded, 40+j Q panel
  // (reused for the column slabs of the last block)
  constexpr int EV_UP = 8, EV_SPLIT = 16, EV_DONE = 24, EV_DOWN = 32, EV_QPAN = 40;
  static_assert(kGemmRing <= 8, "event ids are spaced for at most 8 generations");
  bool used[NB] = {};
  uint64_t down_ticket[NB] = {};  // pageable C: the drainer enqueues the download and records EV_DOWN
  int64_t q_sr = 1, q_sk = 1;  // strides of the raw Q copy on the device

  auto upload_block = [&](int i) -> int {  // P rows (+ old C rows when beta != 0) of block i
    const int g = i % NB;
    const int64_t r0 = (int64_t)i * rb, r1 = std::min(cn.Mo, r0 + rb), rows = r1 - r0;
    if (used[g]) {
      d2h_fence(ctx, down_ticket[g]);  // EV_DOWN of block i-NB has been recorded; its EV_DONE may be re-recorded
      BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->h2d, get_event(ctx, (tensor ? EV_SPLIT : EV_DONE) + g), 0));  // raw P of block i-NB consumed
      if (beta != 0.f) BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->h2d, get_event(ctx, EV_DOWN + g), 0));       // C buffer free
    }
    int64_t p_sr, p_sk;
    trace_host(ctx, "caller: upload P block begin", i);
    BOF_TRY(upload_rows(cn.psrc, cn.p_sr, cn.p_sk, r0, r1, praw[g], &p_sr, &p_sk, ctx->h2d));
    trace_host(ctx, "caller: upload P block end", i);
    if (beta != 0.f)
      BOF_TRY(copy2d(ctx, cblk_of(g), (size_t)cn.No * 4, c + r0 * cn.ldc, (size_t)cn.ldc * 4, (size_t)cn.No * 4, (size_t)rows, H2D, ctx->h2d));
    BOF_CUDA(ctx, cudaEventRecord(get_event(ctx, EV_UP + g), ctx->h2d));
    trace_mark(ctx, ctx->h2d, "h2d: P block landed", i);
    return BOF_OK;
  };
  // compute of block i against Q rows [n0, n1) (the whole Q when not panelled); `first`/`last` bracket the block
  auto rows_of = [&](int i) { return std::min(cn.Mo, (int64_t)(i + 1) * rb) - (int64_t)i * rb; };
  // wait for block i's upload (and for its buffers), split its rows into planes
  auto prepare_block = [&](int i) -> int {
    const int g = i % NB;
    const int64_t rows = rows_of(i);
    const int64_t p_sr = cn.p_sk == 1 ? K : 1, p_sk = cn.p_sk == 1 ? 1 : rows;
    BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, get_event(ctx, EV_UP + g), 0));
    if (used[g]) {
      d2h_fence(ctx, down_ticket[g]);
      BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, get_event(ctx, EV_DOWN + g), 0));
    }
    if (tensor) {
      BOF_TRY(launch_split_planes(ctx, ctx->compute, rows, K, praw[g], p_sr, p_sk, p_hi_of(g), p_lo_of(g), kp));
      BOF_CUDA(ctx, cudaEventRecord(get_event(ctx, EV_SPLIT + g), ctx->compute));
    }
    return BOF_OK;
  };
  // blocks [i0, i0 + cnt) (consecutive generations, no ring wrap) against Q rows [n0, n1), one launch
  auto gemm_blocks = [&](int i0, int cnt, int64_t n0, int64_t n1) -> int {
    const int g = i0 % NB;
    int64_t rows = 0;
    for (int i = i0; i < i0 + cnt; ++i) rows += rows_of(i);
    if (tensor) {
      GemmEpilogue ep;
      ep.alpha = alpha; ep.beta = beta; ep.C = cblk_of(g) + n0; ep.ldc = cn.No;
      trace_mark(ctx, ctx->compute, "compute: gemm start, first block", i0);
      const int rc = launch_gemm_tc(ctx, ctx->compute, cg, rows, n1 - n0, K, kp, p_hi_of(g), p_lo_of(g), q_hi + n0 * kp,
                                    q_lo + n0 * kp, ep, k_chunk_of(ctx));
      trace_mark(ctx, ctx->compute, "compute: gemm end, rows", (int)rows);
      return rc;
    }
    const int64_t p_sr = cn.p_sk == 1 ? K : 1, p_sk = cn.p_sk == 1 ? 1 : rows;  // cnt == 1 on this path
    return launch_gemm_ffma(ctx, ctx->compute, rows, n1 - n0, K, alpha, praw[g], p_sr, p_sk, qraw + n0 * q_sr, q_sk, q_sr,
                            beta, cblk_of(g) + n0, cn.No);
  };
  auto finish_block = [&](int i) -> int {
    const int g = i % NB;
    if (with_terms)  // term_m is added first (kmeans_task.h:74-80): it is the row term iff the canonical rows are m
      BOF_TRY(launch_add_outer_terms(ctx, ctx->compute, cblk_of(g), rows_of(i), cn.No, cn.No, term_rows_d + (int64_t)i * rb,
                                     term_cols_d, canon_rows_are_m ? 1 : 0));
    BOF_CUDA(ctx, cudaEventRecord(get_event(ctx, EV_DONE + g), ctx->compute));
    used[g] = true;
    return BOF_OK;
  };
  auto fetch_block = [&](int i) -> int {
    const int g = i % NB;
    const int64_t r0 = (int64_t)i * rb, rows = std::min(cn.Mo, r0 + rb) - r0;
    BOF_TRY(d2h_transfer(ctx, c + r0 * cn.ldc, (size_t)cn.ldc * 4, cblk_of(g), (size_t)cn.No * 4, (size_t)cn.No * 4, (size_t)rows,
                         ctx->d2h, get_event(ctx, EV_DONE + g), get_event(ctx, EV_DOWN + g), &down_ticket[g]));
    if (down_ticket[g] == 0) trace_mark(ctx, ctx->d2h, "d2h: C block downloaded", i);
    return BOF_OK;
  };

  // ---- Q (resident) and block 0 ----
  // When Q is K-major in the source its rows upload as contiguous panels: panel 0, then P block 0, then the
  // remaining panels; block 0 is computed panel by panel as they land, so only one panel and one P block of
  // PCIe time are exposed before the tensor cores start.
  // dist_q (bof_dist_gemm): Q is replicated over the ranks of the communicator.  Panel j is uploaded by rank
  // j % world only and broadcast over NVLink on the collective stream, so Q crosses PCIe once per node; every
  // rank consumes the panels in the same order as they arrive, exactly like its own uploads.
  const int world = dist_q ? comm_world(ctx) : 1, rank = comm_rank(ctx);
  const bool dist = dist_mode;
  const bool q_panels = tensor && !q_on_device && (dist || (size_t)cn.No * K * 4 >= (256u << 20));
  const int64_t qpan_rows = q_panels ? std::max<int64_t>(256, round_up<int64_t>(ceil_div<int64_t>(cn.No, n_q_panels), 256)) : cn.No;
  const int n_qpan = (int)ceil_div<int64_t>(cn.No, qpan_rows);
  constexpr int EV_SLAB = 88;
  // Every panel is kept as its own tight block at qraw + n0*K: [rows x K] when Q is K-major in the source,
  // [K x rows] (a column range of the stored matrix, pitched copy) when it is not.
  std::vector<int64_t> pan_sr((size_t)n_qpan, 1), pan_sk((size_t)n_qpan, 1);
  auto upload_q_panel = [&](int j) -> int {
    const int64_t n0 = (int64_t)j * qpan_rows, n1 = std::min(cn.No, n0 + qpan_rows);
    if (dist) {
      const int owner = j % world;
      if (cn.q_sk == 1) { pan_sr[j] = K; pan_sk[j] = 1; } else { pan_sr[j] = 1; pan_sk[j] = n1 - n0; }  // what upload_rows produces
      if (owner == rank) {
        // upload my panel, then push it into every peer's exchange buffer with copy-engine peer copies (no SMs)
        BOF_TRY(upload_rows(cn.qsrc, cn.q_sr, cn.q_sk, n0, n1, qraw + n0 * K, &pan_sr[j], &pan_sk[j], ctx->h2d));
        BOF_CUDA(ctx, cudaEventRecord(get_event(ctx, EV_QPAN + j), ctx->h2d));
        trace_mark(ctx, ctx->h2d, "h2d: own Q panel landed", j);
        BOF_TRY(comm_push(ctx, (size_t)(n0 * K), (size_t)(n1 - n0) * K, j, get_event(ctx, EV_QPAN + j)));
      }
      return BOF_OK;
    }
    if (q_on_device) { pan_sr[j] = cn.q_sr; pan_sk[j] = cn.q_sk; }  // single panel, source strides
    else BOF_TRY(upload_rows(cn.qsrc, cn.q_sr, cn.q_sk, n0, n1, qraw + n0 * K, &pan_sr[j], &pan_sk[j], ctx->h2d));
    if (n_qpan == 1) { q_sr = pan_sr[0]; q_sk = pan_sk[0]; }
    BOF_CUDA(ctx, cudaEventRecord(get_event(ctx, EV_QPAN + j), ctx->h2d));
    trace_mark(ctx, ctx->h2d, "h2d: Q panel landed", j);
    return BOF_OK;
  };
  auto split_q_panel = [&](int j) -> int {
    const int64_t n0 = (int64_t)j * qpan_rows, n1 = std::min(cn.No, n0 + qpan_rows);
    if (dist && j % world != rank) {
      BOF_TRY(comm_wait_item(ctx, ctx->compute, j));   // pushed by its owner: gate on the arrival flag
      trace_mark(ctx, ctx->compute, "compute: peer Q panel arrived", j);
    } else {
      BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, get_event(ctx, EV_QPAN + j), 0));
    }
    if (!tensor) return BOF_OK;
    return launch_split_planes(ctx, ctx->compute, n1 - n0, K, qraw + n0 * K, pan_sr[j], pan_sk[j], q_hi + n0 * kp,
                               q_lo + n0 * kp, kp);
  };
  // Prologue: the first blocks ride between the Q panels (h2d order Q0 P0 Q1 P1 Q2 P2 Q3 P3 Q4 .. Qn) and are
  // computed against each panel as it lands (compute order = arrival order), so the tensor cores start after
  // one panel + one block of PCIe time and stay fed while the rest of Q uploads: every new panel unlocks one
  // tile per prologue block.
  trace_mark(ctx, ctx->h2d, "start", 0);
  // One generation stays out of the prologue: the prologue blocks all finish together (with the last panel), so
  // the first steady-state block would otherwise wait for a whole C block to be downloaded (11 ms at 32768^3,
  // seen with BOF_TRACE=1) before it could reuse generation 0.
  const int npro = n_qpan > 1 ? std::min({nblk, NB - 1, n_qpan}) : 1;  // blocks handled by the prologue
  auto pan = [&](int j, int64_t* n0, int64_t* n1) { *n0 = (int64_t)j * qpan_rows; *n1 = std::min(cn.No, *n0 + qpan_rows); };
  const bool merge = tensor;  // the CUDA-core path multiplies one block per launch
  // When every block rides in the prologue (a rank's shard of a multi-GPU job, or a small M), each (block, panel)
  // product completes a column slab of C for good: download slab by slab behind the compute instead of whole blocks
  // at the end, so only the last slab's download is exposed.
  const bool slab_download = tensor && !with_terms && n_qpan > 1 && nblk <= npro;
  uint64_t slab_ticket = 0;
  auto fetch_slab = [&](int i, int64_t n0, int64_t n1, cudaEvent_t after) -> int {
    const int g = i % NB;
    const int64_t r0 = (int64_t)i * rb;
    return d2h_transfer(ctx, c + r0 * cn.ldc + n0, (size_t)cn.ldc * 4, cblk_of(g) + n0, (size_t)cn.No * 4, (size_t)(n1 - n0) * 4,
                        (size_t)rows_of(i), ctx->d2h, after, nullptr, &slab_ticket);
  };
  // Uploads and launches are issued panel by panel: with pinned memory the order of issue is immaterial (everything
  // is asynchronous), but a pageable upload blocks this thread while it is staged, and launching only after the
  // whole prologue had been uploaded left the GPU idle for the first 170 ms at 32768^3 (BOF_TRACE).
  for (int t = 0; t < n_qpan; ++t) {
    int64_t n0, n1;
    pan(t, &n0, &n1);
    BOF_TRY(upload_q_panel(t));
    if (t < npro) BOF_TRY(upload_block(t));
    BOF_TRY(split_q_panel(t));
    // panel t against the blocks that landed before it: one launch over those consecutive generations
    const int older = std::min(t, npro);
    if (older > 0) {
      if (merge) BOF_TRY(gemm_blocks(0, older, n0, n1));
      else for (int i = 0; i < older; ++i) BOF_TRY(gemm_blocks(i, 1, n0, n1));
    }
    // block t landed right after panel t: all panels so far at once
    if (t < npro) {
      BOF_TRY(prepare_block(t));
      BOF_TRY(gemm_blocks(t, 1, 0, n1));
    }
    if (slab_download) {
      cudaEvent_t ev = get_event(ctx, EV_SLAB + t);
      BOF_CUDA(ctx, cudaEventRecord(ev, ctx->compute));
      for (int i = 0; i < older; ++i) BOF_TRY(fetch_slab(i, n0, n1, ev));      // panel t of the earlier blocks
      if (t < npro) BOF_TRY(fetch_slab(t, 0, n1, ev));                          // panels 0..t of block t
      if (slab_ticket == 0) trace_mark(ctx, ctx->d2h, "d2h: slabs of panel downloaded", t);
    } else if (t == n_qpan - 1) {
      for (int i = 0; i < npro; ++i) BOF_TRY(finish_block(i));
    }
  }
  if (slab_download) {
    trace_host(ctx, "caller: everything issued", 0);
    BOF_TRY(sync_all(ctx));
    trace_host(ctx, "caller: sync_all returned", 0);
    trace_dump(ctx, "bof_host_gemm");
    stats_end(ctx);
    return call_guard.done();
  }
  // ---- steady state: Q complete; keep NB blocks in flight, fetch the oldest before reusing its buffers ----
  int next_fetch = 0;
  int fetch_end = nblk;  // blocks [next_fetch, fetch_end) still have to be downloaded whole
  for (int i = npro; i < nblk; ++i) {
    if (next_fetch < i) BOF_TRY(fetch_block(next_fetch++));            // keeps the downloads flowing
    while (next_fetch <= i - NB) BOF_TRY(fetch_block(next_fetch++));   // block i-NB owned these buffers
    BOF_TRY(upload_block(i));
    BOF_TRY(prepare_block(i));
    if (i == nblk - 1 && tensor && cn.No >= 2048) {
      // Drain: the last block is multiplied and downloaded in four column slabs, so only the last slab's
      // download (a quarter of a block) is exposed after the tensor cores stop.
      const int g = i % NB;
      const int64_t r0 = (int64_t)i * rb, rows = rows_of(i);
      const int64_t slab = round_up<int64_t>(ceil_div<int64_t>(cn.No, 4), 256);
      while (next_fetch < i) BOF_TRY(fetch_block(next_fetch++));  // d2h is FIFO: earlier blocks first
      int j = 0;
      for (int64_t n0 = 0; n0 < cn.No; n0 += slab, ++j) {
        const int64_t n1 = std::min(cn.No, n0 + slab);
        BOF_TRY(gemm_blocks(i, 1, n0, n1));
        if (with_terms)
          BOF_TRY(launch_add_outer_terms(ctx, ctx->compute, cblk_of(g) + n0, rows, n1 - n0, cn.No, term_rows_d + r0,
                                         term_cols_d + n0, canon_rows_are_m ? 1 : 0));
        BOF_CUDA(ctx, cudaEventRecord(get_event(ctx, EV_QPAN + j), ctx->compute));
        BOF_TRY(d2h_transfer(ctx, c + r0 * cn.ldc + n0, (size_t)cn.ldc * 4, cblk_of(g) + n0, (size_t)cn.No * 4, (size_t)(n1 - n0) * 4,
                             (size_t)rows, ctx->d2h, get_event(ctx, EV_QPAN + j), n1 == cn.No ? get_event(ctx, EV_DOWN + g) : nullptr,
                             &down_ticket[g]));
        if (down_ticket[g] == 0) trace_mark(ctx, ctx->d2h, "d2h: last block, slab downloaded", j);
      }
      BOF_CUDA(ctx, cudaEventRecord(get_event(ctx, EV_DONE + g), ctx->compute));
      used[g] = true;
      fetch_end = i;
      break;
    }
    BOF_TRY(gemm_blocks(i, 1, 0, cn.No));
    BOF_TRY(finish_block(i));
  }
  while (next_fetch < fetch_end) BOF_TRY(fetch_block(next_fetch++));
  BOF_TRY(sync_all(ctx));
  trace_dump(ctx, "bof_host_gemm");
  stats_end(ctx);
  return call_guard.done();
}

int bof_host_gemm(bof_ctx* ctx, char ord, char ta, char tb, int64_t m, int64_t n, int64_t k, float alpha,
                  float beta, const float* a, const float* b, float* c, int64_t lda, int64_t ldb, int64_t ldc) {
  return host_gemm_impl(ctx, ord, ta, tb, m, n, k, alpha, beta, a, b, c, lda, ldb, ldc, false);
}

int bof_host_kmeans_dist(bof_ctx* ctx, char ord, char ta, char tb, int64_t m, int64_t n, int64_t k, float alpha,
                         float beta, const float* a, const float* b, float* c, int64_t lda, int64_t ldb, int64_t ldc,
                         const float* c_l2sq, const float* p_l2sq) {
  if (ctx && (c_l2sq == nullptr || p_l2sq == nullptr)) return fail(ctx, BOF_EINVAL, "kmeans: c_l2sq / p_l2sq is null");
  return host_gemm_impl(ctx, ord, ta, tb, m, n, k, alpha, beta, a, b, c, lda, ldb, ldc, false, c_l2sq, p_l2sq);
}

int bof_host_gemm_devb(bof_ctx* ctx, char ta, char tb, int64_t m, int64_t n, int64_t k, float alpha, float beta,
                       const float* a, const float* b_dev, float* c, int64_t lda, int64_t ldb, int64_t ldc) {
  return host_gemm_impl(ctx, 'R', ta, tb, m, n, k, alpha, beta, a, b_dev, c, lda, ldb, ldc, true);
}

int bof_dist_gemm(bof_ctx* ctx, char ta, char tb, int64_t m_local, int64_t n, int64_t k, float alpha, float beta,
                  const float* a_local, const float* b, float* c_local, int64_t lda, int64_t ldb, int64_t ldc) {
  return host_gemm_impl(ctx, 'R', ta, tb, m_local, n, k, alpha, beta, a_local, b, c_local, lda, ldb, ldc, false, nullptr,
                        nullptr, true);
}

}  // extern "C"
