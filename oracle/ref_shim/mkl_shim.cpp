// TEST INFRASTRUCTURE — implementation of oracle/ref_shim/mkl.h for building the unmodified reference.
//
// Where the oneMKL 2024.2 inside libtorch_cpu.so exports the operation, the shim forwards to it, so the
// arithmetic (blocking, threading, summation order) is MKL's own:
//   cblas_sgemm            -> SGEMM_64 (Fortran ILP64 interface; row-major handled by operand swap)
//   mkl_scsrmm             -> mkl_sparse_s_create_csr + mkl_sparse_s_mm   (LP64: indices copied to int32)
//   mkl_cspblas_scsrgemv   -> mkl_sparse_s_create_csr + mkl_sparse_s_mv
//   mkl_set_num_threads_local -> same symbol
// Where it exports nothing equivalent the shim RESTATES the documented MKL semantics:
//   mkl_scsrcsc  (stable counting sort by column; job = {0,0,0,-1,-1,1} only)
//   cblas_sdot, cblas_saxpy, cblas_sgemv, cblas_isamin (BLAS-1/2 reference loops, fp32)
// so csrcsc outputs of oracle/_ref pin the reference's *blocked algorithm and merge order*
// (src/blas/csrcsc.cpp, include/tasks/csrcsc_task.h), with the per-block transpose restated.
#include "mkl.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" {
// oneMKL symbols exported by libtorch_cpu.so (LP64 sparse, ILP64 `_64` dense)
void SGEMM_64(const char* ta, const char* tb, const long long* m, const long long* n, const long long* k,
              const float* alpha, const float* a, const long long* lda, const float* b, const long long* ldb,
              const float* beta, float* c, const long long* ldc);
void DGEMM_64(const char* ta, const char* tb, const long long* m, const long long* n, const long long* k,
              const double* alpha, const double* a, const long long* lda, const double* b, const long long* ldb,
              const double* beta, double* c, const long long* ldc);
struct shim_matrix_descr {
  int type, mode, diag;
};
int mkl_sparse_s_create_csr(void** A, int indexing, int rows, int cols, int* rows_start, int* rows_end,
                            int* col_indx, float* values);
int mkl_sparse_s_mm(int op, float alpha, void* A, shim_matrix_descr descr, int layout, const float* x, int columns,
                    int ldx, float beta, float* y, int ldy);
int mkl_sparse_s_mv(int op, float alpha, void* A, shim_matrix_descr descr, const float* x, float beta, float* y);
int mkl_sparse_destroy(void* A);
}

namespace {
  const int OP_N = 10, OP_T = 11, BASE0 = 0, GENERAL = 20, ROW_MAJOR = 101, COL_MAJOR = 102;

  void die(const char* what, long long v) {
    std::fprintf(stderr, "mkl_shim: %s (%lld)\n", what, v);
    std::abort();
  }

  // 4-array CSR with 64-bit, possibly un-rebased, possibly one-based offsets -> rebased zero-based int32 copy
  struct Csr32 {
    std::vector<int> offs, idx;
    const float* vals;
    Csr32(MKL_INT rows, const float* val, const MKL_INT* indx, const MKL_INT* pb, const MKL_INT* pe, MKL_INT base) {
      MKL_INT first = pb[0] - base;  // rows are contiguous at every call site (pntre = pntrb + 1)
      MKL_INT nnz = (rows ? pe[rows - 1] : pb[0]) - pb[0];
      if (nnz > 0x7fffffffLL) die("nnz exceeds the LP64 sparse interface", nnz);
      offs.resize(rows + 1);
      for (MKL_INT r = 0; r < rows; r++) offs[r] = (int) (pb[r] - pb[0]);
      offs[rows] = (int) nnz;
      idx.resize(nnz);
      for (MKL_INT i = 0; i < nnz; i++) idx[i] = (int) (indx[first + i] - base);
      vals = val + first;
    }
  };
}  // namespace

extern "C" {

void* mkl_malloc(size_t bytes, int alignment) {
  void* p = nullptr;
  if (alignment < (int) sizeof(void*)) alignment = sizeof(void*);
  if (posix_memalign(&p, alignment, bytes ? bytes : 1) != 0) return nullptr;
  return p;
}
void mkl_free(void* p) { std::free(p); }

void cblas_sgemm(CBLAS_LAYOUT layout, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, MKL_INT m, MKL_INT n, MKL_INT k,
                 float alpha, const float* a, MKL_INT lda, const float* b, MKL_INT ldb, float beta, float* c,
                 MKL_INT ldc) {
  char cta = ta == CblasNoTrans ? 'N' : 'T', ctb = tb == CblasNoTrans ? 'N' : 'T';
  if (layout == CblasColMajor)
    SGEMM_64(&cta, &ctb, &m, &n, &k, &alpha, a, &lda, b, &ldb, &beta, c, &ldc);
  else  // row-major C = op(A) op(B)  <=>  column-major C^T = op(B)^T op(A)^T
    SGEMM_64(&ctb, &cta, &n, &m, &k, &alpha, b, &ldb, a, &lda, &beta, c, &ldc);
}
void cblas_dgemm(CBLAS_LAYOUT layout, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, MKL_INT m, MKL_INT n, MKL_INT k,
                 double alpha, const double* a, MKL_INT lda, const double* b, MKL_INT ldb, double beta, double* c,
                 MKL_INT ldc) {
  char cta = ta == CblasNoTrans ? 'N' : 'T', ctb = tb == CblasNoTrans ? 'N' : 'T';
  if (layout == CblasColMajor)
    DGEMM_64(&cta, &ctb, &m, &n, &k, &alpha, a, &lda, b, &ldb, &beta, c, &ldc);
  else
    DGEMM_64(&ctb, &cta, &n, &m, &k, &alpha, b, &ldb, a, &lda, &beta, c, &ldc);
}

// restatement (BLAS-2 reference loop)
void cblas_sgemv(CBLAS_LAYOUT layout, CBLAS_TRANSPOSE ta, MKL_INT m, MKL_INT n, float alpha, const float* a,
                 MKL_INT lda, const float* x, MKL_INT incx, float beta, float* y, MKL_INT incy) {
  bool rows_are_outputs = (layout == CblasRowMajor) == (ta == CblasNoTrans);
  // view a as R x S with element (r, s) at a[r*lda + s]; y has R entries if rows_are_outputs, else S
  MKL_INT R = layout == CblasRowMajor ? m : n, S = layout == CblasRowMajor ? n : m;
  MKL_INT ny = rows_are_outputs ? R : S, nx = rows_are_outputs ? S : R;
  for (MKL_INT i = 0; i < ny; i++) {
    float acc = 0.f;
    for (MKL_INT j = 0; j < nx; j++) acc += (rows_are_outputs ? a[i * lda + j] : a[j * lda + i]) * x[j * incx];
    y[i * incy] = beta == 0.f ? alpha * acc : alpha * acc + beta * y[i * incy];
  }
}
// restatements (BLAS-1 reference loops)
void cblas_saxpy(MKL_INT n, float alpha, const float* x, MKL_INT incx, float* y, MKL_INT incy) {
  for (MKL_INT i = 0; i < n; i++) y[i * incy] += alpha * x[i * incx];
}
float cblas_sdot(MKL_INT n, const float* x, MKL_INT incx, const float* y, MKL_INT incy) {
  float acc = 0.f;
  for (MKL_INT i = 0; i < n; i++) acc += x[i * incx] * y[i * incy];
  return acc;
}
size_t cblas_isamin(MKL_INT n, const float* x, MKL_INT incx) {
  size_t best = 0;
  float bv = n > 0 ? std::fabs(x[0]) : 0.f;
  for (MKL_INT i = 1; i < n; i++) {
    float v = std::fabs(x[i * incx]);
    if (v < bv) bv = v, best = (size_t) i;
  }
  return best;
}

void mkl_scsrmm(const char* transa, const MKL_INT* m, const MKL_INT* n, const MKL_INT* k, const float* alpha,
                const char* matdescra, const float* val, const MKL_INT* indx, const MKL_INT* pntrb,
                const MKL_INT* pntre, const float* b, const MKL_INT* ldb, const float* beta, float* c,
                const MKL_INT* ldc) {
  if (matdescra[0] != 'G') die("only general matrices are used by the reference", matdescra[0]);
  // matdescra[3]: 'C' = zero-based indices AND row-major dense operands, 'F' = one-based AND column-major
  bool zero = matdescra[3] == 'C';
  if (*m > 0x7fffffffLL || *k > 0x7fffffffLL || *n > 0x7fffffffLL) die("dimension exceeds LP64", *m);
  Csr32 A(*m, val, indx, pntrb, pntre, zero ? 0 : 1);
  void* h = nullptr;
  int st = mkl_sparse_s_create_csr(&h, BASE0, (int) *m, (int) *k, A.offs.data(), A.offs.data() + 1, A.idx.data(),
                                   const_cast<float*>(A.vals));
  if (st) die("mkl_sparse_s_create_csr", st);
  shim_matrix_descr d{GENERAL, 0, 0};
  int op = (*transa == 'N' || *transa == 'n') ? OP_N : OP_T;
  st = mkl_sparse_s_mm(op, *alpha, h, d, zero ? ROW_MAJOR : COL_MAJOR, b, (int) *n, (int) *ldb, *beta, c, (int) *ldc);
  if (st) die("mkl_sparse_s_mm", st);
  mkl_sparse_destroy(h);
}

void mkl_cspblas_scsrgemv(const char* transa, const MKL_INT* m, const float* a, const MKL_INT* ia,
                          const MKL_INT* ja, const float* x, float* y) {
  if (*m > 0x7fffffffLL) die("dimension exceeds LP64", *m);
  Csr32 A(*m, a, ja, ia, ia + 1, 0);
  void* h = nullptr;
  int st = mkl_sparse_s_create_csr(&h, BASE0, (int) *m, (int) *m, A.offs.data(), A.offs.data() + 1, A.idx.data(),
                                   const_cast<float*>(A.vals));
  if (st) die("mkl_sparse_s_create_csr", st);
  shim_matrix_descr d{GENERAL, 0, 0};
  int op = (*transa == 'N' || *transa == 'n') ? OP_N : OP_T;
  st = mkl_sparse_s_mv(op, 1.0f, h, d, x, 0.0f, y);
  if (st) die("mkl_sparse_s_mv", st);
  mkl_sparse_destroy(h);
}

// RESTATEMENT of mkl_scsrcsc for job = {0, 0, 0, *, *, 1}: CSR -> CSC of a square order-n matrix, zero-based in
// and out, all three output arrays filled; within a column entries keep ascending row order (stable).
void mkl_scsrcsc(const MKL_INT* job, const MKL_INT* n, float* acsr, MKL_INT* ja, MKL_INT* ia, float* acsc,
                 MKL_INT* ja1, MKL_INT* ia1, MKL_INT* info) {
  if (job[0] != 0 || job[1] != 0 || job[2] != 0 || job[5] != 1) die("unsupported mkl_scsrcsc job", job[0]);
  MKL_INT N = *n, base = ia[0], nnz = ia[N] - base;
  std::vector<MKL_INT> cnt(N + 1, 0);
  for (MKL_INT i = 0; i < nnz; i++) cnt[ja[base + i] + 1]++;
  for (MKL_INT c = 0; c < N; c++) cnt[c + 1] += cnt[c];
  for (MKL_INT c = 0; c <= N; c++) ia1[c] = cnt[c];
  for (MKL_INT r = 0; r < N; r++)
    for (MKL_INT p = ia[r]; p < ia[r + 1]; p++) {
      MKL_INT dst = cnt[ja[p]]++;
      ja1[dst] = r;
      acsc[dst] = acsr[p];
    }
  if (info) *info = 0;
}
}  // extern "C"
