/*
 * oracle.c -- CPU restatement of BLAS-on-Flash's tiled hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's library.  The product (libbof_b200.so, include/flash_blas.h) never does.
 *
 * PARITY PINNED TO THE REFERENCE.  The reference holds no golden vectors for this path (SURVEY.md
 * section 8c), so they are produced by running the reference itself: oracle/Makefile.ref compiles its
 * unmodified sources (in_mem_* drivers and the flash:: library with its scheduler / cache / libaio file
 * handles) behind two shim headers -- oracle/ref_shim/mkl.h, which forwards to the oneMKL 2024.2 inside
 * libtorch_cpu.so (SGEMM_64, mkl_sparse_s_mm/_mv) and restates mkl_scsrcsc / BLAS-1, and
 * oracle/ref_shim/libaio.h over the native-AIO syscalls -- into oracle/_ref/.  tests/golden/golden_ref.npz
 * holds outputs of those binaries (tests/golden/make_golden_ref.py), and tests/test_oracle_ref.py checks
 * every function of this file against them and against live runs of the binaries on fresh inputs.
 * What this file restates is the semantics of the MKL entry points at the reference's call sites, in
 * plain fp32 (and with fp64 accumulation to bound both sides); the integer/compare-only results
 * (csrcsc, isamin assignment) are fully determined without MKL.
 *
 * Every function cites the reference file:line it follows (paths relative to the reference tree).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;

/* ---------------------------------------------------------------------------------------------
 * csrmm: C = alpha * op(A) * B + beta * C
 * Follows drivers/in_mem_csrmm.cpp:96-120 (one monolithic mkl_scsrmm call) and the per-tile call
 * include/tasks/csrmm_task.h:219-228.  ord 'R' = matdescra[3]='C' (0-based, row-major B/C),
 * ord 'C' = matdescra[3]='F' (column-major B/C; indices here stay 0-based -- the +1 in
 * in_mem_csrmm.cpp:101-111 is an MKL calling convention, not data).
 *   A: m x n CSR (ia m+1 offsets, possibly un-rebased; ja/a indexed by ia[i]-ia[0])
 *   trans 'N': B n x k, C m x k.   trans 'T': B m x k, C n x k.
 *   ldb/ldc as in the driver: row-major -> k, column-major -> rows of the matrix.
 * beta == 0 => C is not read (csrmm_task.h:194-196: the C tile is only fetched when beta != 0).
 * acc64 != 0 accumulates each dot product in double (the "both sides" bound).
 * --------------------------------------------------------------------------------------------- */
int orc_csrmm(char trans, i64 m, i64 n, i64 k, float alpha, float beta, const float* a, const i64* ia,
              const i64* ja, char ord, const float* b, float* c, int acc64) {
  if ((trans != 'N' && trans != 'T') || (ord != 'R' && ord != 'C')) return -1; /* src/blas/csrmm.cpp:433-449 */
  const i64 crows = trans == 'N' ? m : n, brows = trans == 'N' ? n : m;
  const i64 base = ia[0];
#define BIDX(r, j) (ord == 'R' ? (r) * k + (j) : (j) * brows + (r))
#define CIDX(r, j) (ord == 'R' ? (r) * k + (j) : (j) * crows + (r))
  if (trans == 'N') {
#pragma omp parallel for schedule(dynamic, 256)
    for (i64 i = 0; i < m; ++i) {
      for (i64 j = 0; j < k; ++j) {
        float s32 = 0.f;
        double s64 = 0.0;
        for (i64 p = ia[i] - base; p < ia[i + 1] - base; ++p) {
          if (acc64) s64 += (double)a[p] * (double)b[BIDX(ja[p], j)];
          else s32 = fmaf(a[p], b[BIDX(ja[p], j)], s32);
        }
        const float s = acc64 ? (float)s64 : s32;
        c[CIDX(i, j)] = (beta == 0.f) ? alpha * s : alpha * s + beta * c[CIDX(i, j)];
      }
    }
  } else {
    /* C(n x k) = alpha * A^T B + beta*C: scatter form, row order of A (what a CSR 'T' kernel does) */
    double* acc = (double*)calloc((size_t)(n * k), sizeof(double));
    if (!acc) return -2;
    for (i64 i = 0; i < m; ++i)
      for (i64 p = ia[i] - base; p < ia[i + 1] - base; ++p)
        for (i64 j = 0; j < k; ++j) {
          if (acc64) acc[ja[p] * k + j] += (double)a[p] * (double)b[BIDX(i, j)];
          else acc[ja[p] * k + j] = (double)fmaf(a[p], b[BIDX(i, j)], (float)acc[ja[p] * k + j]);
        }
    for (i64 r = 0; r < n; ++r)
      for (i64 j = 0; j < k; ++j) {
        const float s = (float)acc[r * k + j];
        c[CIDX(r, j)] = (beta == 0.f) ? alpha * s : alpha * s + beta * c[CIDX(r, j)];
      }
    free(acc);
  }
#undef BIDX
#undef CIDX
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * gemm: C = alpha * op(A) * op(B) + beta * C, cblas_sgemm semantics.
 * Follows drivers/in_mem_gemm.cpp:58-67 (monolithic call) and include/tasks/gemm_task.h:87-90.
 * ord 'R'/'C', ta/tb 'N'/'T'; ld == 0 picks the tight default of src/blas/gemm.cpp:63-67.
 * beta == 0 => C is not read (gemm_task.h:49-53).
 * --------------------------------------------------------------------------------------------- */
static void gemm_strides(char ord, char t, i64 rows, i64 cols, i64* ld, i64* s_row, i64* s_col) {
  /* element (r, c) of op(X), X stored per (ord, t) */
  const int col = ord == 'C', tr = t == 'T';
  const i64 inner = (tr != col) ? rows : cols; /* contiguous extent as stored */
  if (*ld == 0) *ld = inner;
  if (tr != col) { *s_row = 1; *s_col = *ld; } else { *s_row = *ld; *s_col = 1; }
}

int orc_gemm(char ord, char ta, char tb, i64 m, i64 n, i64 k, float alpha, float beta, const float* a, i64 lda,
             const float* b, i64 ldb, float* c, i64 ldc, int acc64) {
  if ((ord != 'R' && ord != 'C') || (ta != 'N' && ta != 'T') || (tb != 'N' && tb != 'T')) return -1;
  i64 a_r, a_c, b_r, b_c, c_r, c_c;
  gemm_strides(ord, ta, m, k, &lda, &a_r, &a_c);
  gemm_strides(ord, tb, k, n, &ldb, &b_r, &b_c);
  gemm_strides(ord, 'N', m, n, &ldc, &c_r, &c_c);
#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < m; ++i) {
    double* row64 = acc64 ? (double*)calloc((size_t)n, sizeof(double)) : NULL;
    float* row32 = acc64 ? NULL : (float*)calloc((size_t)n, sizeof(float));
    for (i64 p = 0; p < k; ++p) {
      const float av = a[i * a_r + p * a_c];
      const float* brow = b + p * b_r;
      if (acc64) for (i64 j = 0; j < n; ++j) row64[j] += (double)av * (double)brow[j * b_c];
      else for (i64 j = 0; j < n; ++j) row32[j] = fmaf(av, brow[j * b_c], row32[j]);
    }
    for (i64 j = 0; j < n; ++j) {
      const float s = acc64 ? (float)row64[j] : row32[j];
      float* dst = c + i * c_r + j * c_c;
      *dst = (beta == 0.f) ? alpha * s : alpha * s + beta * *dst;
    }
    free(row64);
    free(row32);
  }
  return 0;
}

/* The reference's tiler (src/blas/gemm.cpp:46-129): blocks of `blk` with the tail-merge rule
 * (a remainder of < 128 elements joins the last block, :69-75), one compact-tile sgemm per
 * (l, i, j) and the k-dimension accumulate chain: l > 0 runs with beta = 1 on the same C tile
 * (:114-126).  Row-major/column-major and transposes are forwarded to orc_gemm on sub-views. */
int orc_gemm_tiled(char ord, char ta, char tb, i64 m, i64 n, i64 k, float alpha, float beta, const float* a,
                   i64 lda, const float* b, i64 ldb, float* c, i64 ldc, i64 blk) {
  if ((ord != 'R' && ord != 'C') || (ta != 'N' && ta != 'T') || (tb != 'N' && tb != 'T')) return -1;
  i64 a_r, a_c, b_r, b_c, c_r, c_c;
  gemm_strides(ord, ta, m, k, &lda, &a_r, &a_c);
  gemm_strides(ord, tb, k, n, &ldb, &b_r, &b_c);
  gemm_strides(ord, 'N', m, n, &ldc, &c_r, &c_c);
  const i64 dims[3] = {m, k, n};
  i64 bsz[3], nb[3];
  for (int d = 0; d < 3; ++d) {
    bsz[d] = dims[d] < blk ? dims[d] : blk;
    if (bsz[d] == 0) return 0;
    const i64 div = dims[d] / bsz[d];
    nb[d] = (dims[d] - div * bsz[d] < 128) ? div : div + 1; /* SECTOR_LEN / sizeof(float) = 128 */
    if (nb[d] == 0) nb[d] = 1;
  }
  for (i64 l = 0; l < nb[1]; ++l)
    for (i64 i = 0; i < nb[0]; ++i)
      for (i64 j = 0; j < nb[2]; ++j) {
        const i64 idx[3] = {i, l, j};
        i64 ext[3];
        for (int d = 0; d < 3; ++d) ext[d] = (idx[d] == nb[d] - 1) ? dims[d] - idx[d] * bsz[d] : bsz[d];
        const float* at = a + i * bsz[0] * a_r + l * bsz[1] * a_c;
        const float* bt = b + l * bsz[1] * b_r + j * bsz[2] * b_c;
        float* ct = c + i * bsz[0] * c_r + j * bsz[2] * c_c;
        const int rc = orc_gemm(ord, ta, tb, ext[0], ext[2], ext[1], alpha, l > 0 ? 1.f : beta, at, lda, bt, ldb,
                                ct, ldc, 0);
        if (rc) return rc;
      }
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * csrgemv: y = op(A) x, y overwritten (mkl_cspblas_scsrgemv, zero-based, no alpha/beta).
 * Follows drivers/in_mem_csrgemv.cpp:36-83; the square padding there (dim = max(m, n)) only adds
 * empty rows / zero entries, so the rectangular statement below is equivalent.
 * 'T' follows include/tasks/csrgemv_task.h:152-179 with the blocks added in row order.
 * --------------------------------------------------------------------------------------------- */
int orc_csrgemv(char trans, i64 m, i64 n, const float* a, const i64* ia, const i64* ja, const float* x,
                float* y, int acc64) {
  if (trans != 'N' && trans != 'T') return -1;
  const i64 base = ia[0];
  if (trans == 'N') {
#pragma omp parallel for schedule(dynamic, 1024)
    for (i64 i = 0; i < m; ++i) {
      float s32 = 0.f;
      double s64 = 0.0;
      for (i64 p = ia[i] - base; p < ia[i + 1] - base; ++p) {
        if (acc64) s64 += (double)a[p] * (double)x[ja[p]];
        else s32 = fmaf(a[p], x[ja[p]], s32);
      }
      y[i] = acc64 ? (float)s64 : s32;
    }
  } else {
    double* acc = (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    if (!acc) return -2;
    for (i64 i = 0; i < m; ++i)
      for (i64 p = ia[i] - base; p < ia[i + 1] - base; ++p) {
        if (acc64) acc[ja[p]] += (double)a[p] * (double)x[i];
        else acc[ja[p]] = (double)fmaf(a[p], x[i], (float)acc[ja[p]]);
      }
    for (i64 j = 0; j < n; ++j) y[j] = (float)acc[j];
    free(acc);
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * csrcsc: CSR(m x n) -> CSC == CSR of A^T, zero-based, all three output arrays
 * (mkl_scsrcsc with job = {0,0,0,-1,-1,1}, drivers/in_mem_csrcsc.cpp:72-78).
 * The result is the stable counting sort by column: inside an output row the entries appear in
 * ascending source row, duplicates in storage order.
 * --------------------------------------------------------------------------------------------- */
int orc_csrcsc(i64 m, i64 n, const i64* ia, const i64* ja, const float* a, i64* ia_tr, i64* ja_tr, float* a_tr) {
  const i64 base = ia[0];
  const i64 nnz = ia[m] - base;
  memset(ia_tr, 0, (size_t)(n + 1) * sizeof(i64));
  for (i64 p = 0; p < nnz; ++p) ia_tr[ja[p] + 1]++;
  for (i64 j = 0; j < n; ++j) ia_tr[j + 1] += ia_tr[j];
  i64* fill = (i64*)malloc((size_t)(n > 0 ? n : 1) * sizeof(i64));
  if (!fill) return -2;
  memcpy(fill, ia_tr, (size_t)n * sizeof(i64));
  for (i64 i = 0; i < m; ++i)
    for (i64 p = ia[i] - base; p < ia[i + 1] - base; ++p) {
      const i64 q = fill[ja[p]]++;
      ja_tr[q] = i;
      a_tr[q] = a[p];
    }
  free(fill);
  return 0;
}

/* Row-block size rule of include/blas_utils.h:72-82 (get_next_blk_size): grow from min_size while
 * the block's nnz stays <= max_nnzs, then cap at max_size. */
i64 orc_next_blk_size(const i64* offs, i64 nrows, i64 min_size, i64 max_size, i64 max_nnzs) {
  i64 blk = min_size;
  while (blk < nrows && (offs[blk] - offs[0]) <= max_nnzs) blk++;
  return blk < max_size ? blk : max_size;
}

/* The reference's out-of-core algorithm, restated to show it equals orc_csrcsc:
 * (1) per row block: transpose the block as a zero-row-padded square matrix and rebase the row
 *     indices by the block start (include/tasks/csrcsc_task.h:42-92);
 * (2) ia_tr = scan of the summed per-block column counts (src/blas/csrcsc.cpp:88-99);
 * (3) per output row concatenate the blocks' segments in block order (csrcsc_task.h:143-162).
 * Row blocks come from fill_blocks(min=10, max=rblk) (src/blas/csrcsc.cpp:44-45). */
int orc_csrcsc_blocked(i64 m, i64 n, const i64* ia, const i64* ja, const float* a, i64* ia_tr, i64* ja_tr,
                       float* a_tr, i64 rblk, i64 max_nnzs) {
  const i64 base = ia[0], nnz = ia[m] - base;
  i64 nblk = 0, cap = 16;
  i64* starts = (i64*)malloc((size_t)cap * sizeof(i64));
  for (i64 cur = 0; cur < m;) {
    i64 sz = orc_next_blk_size(ia + cur, m - cur, 10, rblk, max_nnzs);
    if (sz > m - cur) sz = m - cur; /* the reference relies on reads past the end being harmless */
    if (nblk + 2 > cap) { cap *= 2; starts = (i64*)realloc(starts, (size_t)cap * sizeof(i64)); }
    starts[nblk++] = cur;
    cur += sz;
  }
  starts[nblk] = m;
  i64** boffs = (i64**)malloc((size_t)(nblk > 0 ? nblk : 1) * sizeof(i64*));
  i64* bja = (i64*)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(i64));
  float* ba = (float*)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(float));
  for (i64 bi = 0; bi < nblk; ++bi) {
    const i64 r0 = starts[bi], r1 = starts[bi + 1], z0 = ia[r0] - base;
    boffs[bi] = (i64*)malloc((size_t)(n + 1) * sizeof(i64));
    orc_csrcsc(r1 - r0, n, ia + r0, ja + z0, a + z0, boffs[bi], bja + z0, ba + z0);
    for (i64 p = z0; p < ia[r1] - base; ++p) bja[p] += r0; /* idx += blk.start */
  }
  memset(ia_tr, 0, (size_t)(n + 1) * sizeof(i64));
  for (i64 bi = 0; bi < nblk; ++bi)
    for (i64 j = 1; j <= n; ++j) ia_tr[j] += boffs[bi][j] - boffs[bi][j - 1];
  for (i64 j = 1; j <= n; ++j) ia_tr[j] += ia_tr[j - 1];
  for (i64 j = 0; j < n; ++j) {
    i64 fill = ia_tr[j];
    for (i64 bi = 0; bi < nblk; ++bi) {
      const i64 z0 = ia[starts[bi]] - base;
      const i64 cnt = boffs[bi][j + 1] - boffs[bi][j];
      memcpy(ja_tr + fill, bja + z0 + boffs[bi][j], (size_t)cnt * sizeof(i64));
      memcpy(a_tr + fill, ba + z0 + boffs[bi][j], (size_t)cnt * sizeof(float));
      fill += cnt;
    }
  }
  for (i64 bi = 0; bi < nblk; ++bi) free(boffs[bi]);
  free(boffs); free(bja); free(ba); free(starts);
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * kmeans, following drivers/in_mem_kmeans.cpp in full.
 * --------------------------------------------------------------------------------------------- */
static float dot32(const float* x, const float* y, i64 d) {
  float s = 0.f;
  for (i64 j = 0; j < d; ++j) s = fmaf(x[j], y[j], s);
  return s;
}
static double dot64(const float* x, const float* y, i64 d) {
  double s = 0.0;
  for (i64 j = 0; j < d; ++j) s += (double)x[j] * (double)y[j];
  return s;
}

/* row squared norms: mkl_dot(x, x), in_mem_kmeans.cpp:75-78 and :179-182 */
void orc_row_sqnorm(i64 rows, i64 dim, const float* x, float* out) {
#pragma omp parallel for schedule(static)
  for (i64 r = 0; r < rows; ++r) out[r] = dot32(x + r * dim, x + r * dim, dim);
}

/* closest_centers (in_mem_kmeans.cpp:69-87): D[p, c] = fl(fl(-2*<x_p, mu_c> + c2[c]) + p2[p])
 * in the operation order of distsq_points_to_centers (:34-42: sgemm alpha=-2 beta=0, then the
 * two rank-1 updates), then center_index[p] = cblas_isamin(K, D[p, :]) = first index of the
 * minimum ABSOLUTE value (:84-85).  margin[p] (optional) = second-smallest |D| minus smallest,
 * so that tests can restrict bit-exact comparison to ties-free points.  acc64 computes the dot
 * product in double before rounding to fp32 (the inner product's rounding is MKL-internal). */
void orc_kmeans_assign(i64 npoints, i64 ncenters, i64 dim, const float* points, const float* centers,
                       const float* c_l2sq, const float* p_l2sq, i64* assign, float* margin, int acc64) {
#pragma omp parallel for schedule(static)
  for (i64 p = 0; p < npoints; ++p) {
    float best = INFINITY, second = INFINITY;
    i64 bi = 0;
    for (i64 c = 0; c < ncenters; ++c) {
      const float dot = acc64 ? (float)dot64(points + p * dim, centers + c * dim, dim)
                              : dot32(points + p * dim, centers + c * dim, dim);
      float d = -2.0f * dot;
      d = d + c_l2sq[c];
      d = d + p_l2sq[p];
      d = fabsf(d);
      if (d < best) { second = best; best = d; bi = c; }
      else if (d < second) second = d;
    }
    assign[p] = bi;
    if (margin) margin[p] = second - best;
  }
}

/* centroid update of lloyds_iter (in_mem_kmeans.cpp:105-125): bucket the points by center in
 * ascending p, zero the centers, then per cluster saxpy(1/n_c) in ascending p; an empty cluster
 * stays the zero vector (:112).  mode 0 = the reference's scale-then-add order in fp32;
 * mode 1 = sum in fp32 then divide (what a sharded reduction does); mode 2 = fp64 sum then divide. */
void orc_kmeans_update(i64 npoints, i64 ncenters, i64 dim, const float* points, const i64* assign,
                       float* centers, i64* counts_out, int mode) {
  i64* counts = (i64*)calloc((size_t)ncenters, sizeof(i64));
  for (i64 p = 0; p < npoints; ++p) counts[assign[p]]++;
  double* acc = (double*)calloc((size_t)(ncenters * dim), sizeof(double));
  memset(centers, 0, (size_t)(ncenters * dim) * sizeof(float));
  for (i64 p = 0; p < npoints; ++p) {
    const i64 c = assign[p];
    const float* x = points + p * dim;
    if (mode == 0) {
      const float w = 1.0f / (float)counts[c];
      for (i64 j = 0; j < dim; ++j) centers[c * dim + j] = fmaf(w, x[j], centers[c * dim + j]);
    } else if (mode == 1) {
      for (i64 j = 0; j < dim; ++j) centers[c * dim + j] += x[j];
    } else {
      for (i64 j = 0; j < dim; ++j) acc[c * dim + j] += (double)x[j];
    }
  }
  if (mode != 0)
    for (i64 c = 0; c < ncenters; ++c)
      for (i64 j = 0; j < dim; ++j) {
        if (counts[c] == 0) centers[c * dim + j] = 0.f;
        else if (mode == 1) centers[c * dim + j] = centers[c * dim + j] * (1.0f / (float)counts[c]);
        else centers[c * dim + j] = (float)(acc[c * dim + j] / (double)counts[c]);
      }
  if (counts_out) memcpy(counts_out, counts, (size_t)ncenters * sizeof(i64));
  free(acc);
  free(counts);
}

/* residual of lloyds_iter (in_mem_kmeans.cpp:127-151): sum_p distsq(x_p, mu_assign[p]) with
 * distsq = <x,x> + <mu,mu> - 2<x,mu> (:13-18), chunked by 8196 points. */
double orc_kmeans_residual(i64 npoints, i64 dim, const float* points, const float* centers, const i64* assign) {
  double total = 0.0;
  const i64 CH = 8196;
  for (i64 c0 = 0; c0 < npoints; c0 += CH) {
    float r = 0.f;
    for (i64 p = c0; p < npoints && p < c0 + CH; ++p) {
      const float* x = points + p * dim;
      const float* mu = centers + assign[p] * dim;
      r += dot32(x, x, dim) + dot32(mu, mu, dim) - 2 * dot32(x, mu, dim);
    }
    total += (double)r;
  }
  return total;
}

/* One Lloyd iteration exactly as in_mem_kmeans.cpp:89-152 strings the pieces together. */
double orc_lloyd_iter(i64 npoints, i64 ncenters, i64 dim, const float* points, float* centers,
                      const float* p_l2sq, i64* assign_out, int update_mode) {
  float* c2 = (float*)malloc((size_t)ncenters * sizeof(float));
  i64* assign = assign_out ? assign_out : (i64*)malloc((size_t)(npoints > 0 ? npoints : 1) * sizeof(i64));
  orc_row_sqnorm(ncenters, dim, centers, c2);
  orc_kmeans_assign(npoints, ncenters, dim, points, centers, c2, p_l2sq, assign, NULL, 0);
  orc_kmeans_update(npoints, ncenters, dim, points, assign, centers, NULL, update_mode);
  const double res = orc_kmeans_residual(npoints, dim, points, centers, assign);
  if (!assign_out) free(assign);
  free(c2);
  return res;
}

/* ---------------------------------------------------------------------------------------------
 * Data generators with the value patterns of the reference's tools, for "compat" tests:
 * misc/sparse_create.cpp:52-55 (values (i % 9) + 1) and misc/dense_create.cpp:28-32 (i % 10).
 * Column patterns use our own counter-based hash instead of rand_r.
 * --------------------------------------------------------------------------------------------- */
static uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
static int cmp_i64(const void* a, const void* b) {
  const i64 x = *(const i64*)a, y = *(const i64*)b;
  return (x > y) - (x < y);
}
/* exactly nnz_per_row sorted unique columns per row; vals: mode 0 = (i % 9) + 1, mode 1 = U[0,1) */
int orc_gen_csr(i64 m, i64 n, i64 nnz_per_row, uint64_t seed, int val_mode, i64* ia, i64* ja, float* a) {
  if (nnz_per_row > n) return -1;
#pragma omp parallel for schedule(dynamic, 1024)
  for (i64 r = 0; r < m; ++r) {
    i64* cols = ja + r * nnz_per_row;
    i64 have = 0;
    uint64_t ctr = 0;
    while (have < nnz_per_row) {
      for (i64 j = have; j < nnz_per_row; ++j)
        cols[j] = (i64)(mix64(seed ^ mix64((uint64_t)r * 0x100000001B3ull + ctr++)) % (uint64_t)n);
      qsort(cols, (size_t)nnz_per_row, sizeof(i64), cmp_i64);
      i64 w = 0;
      for (i64 j = 0; j < nnz_per_row; ++j)
        if (j == 0 || cols[j] != cols[j - 1]) cols[w++] = cols[j];
      have = w;
    }
    ia[r] = r * nnz_per_row;
    for (i64 j = 0; j < nnz_per_row; ++j) {
      const i64 i = r * nnz_per_row + j;
      a[i] = val_mode == 0 ? (float)((i % 9) + 1)
                           : (float)(mix64(seed * 31 + (uint64_t)i) >> 40) * (1.0f / 16777216.0f);
    }
  }
  ia[m] = m * nnz_per_row;
  return 0;
}
/* mode 0 = i % 10 (dense_create 's'), mode 1 = U[0,1), mode 2 = zeros ('z') */
void orc_gen_dense(i64 count, uint64_t seed, int mode, float* out) {
#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < count; ++i)
    out[i] = mode == 0 ? (float)(i % 10)
                       : mode == 1 ? (float)(mix64(seed + (uint64_t)i) >> 40) * (1.0f / 16777216.0f) : 0.f;
}
