/* TEST INFRASTRUCTURE — shim `mkl.h` so that the UNMODIFIED reference sources under /root/reference
 * compile in an image that has no Intel MKL headers or ILP64 link libraries.
 *
 * The reference includes "mkl.h" from include/bof_types.h:8 and is built with -DMKL_ILP64
 * (CMakeLists.txt:104), i.e. MKL_INT is a 64-bit integer.  This header declares exactly the MKL
 * entry points the reference calls (bof_types.h:18-29 and the call sites listed per function
 * below); oracle/ref_shim/mkl_shim.cpp implements them on top of what the oneMKL 2024.2 build
 * inside libtorch_cpu.so exports (SGEMM_64, mkl_sparse_s_mm / _mv, mkl_set_num_threads_local),
 * and as a plain restatement where that MKL build exports nothing equivalent (mkl_scsrcsc,
 * cblas_sdot, cblas_isamin, cblas_saxpy, cblas_sgemv) — said so at each definition.
 * Nothing under blas-on-flash_b200/ includes or links this. */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifdef MKL_ILP64
typedef long long int MKL_INT; /* mkl_types.h: MKL_INT64 */
#else
typedef int MKL_INT;
#endif

typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_LAYOUT;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;

void* mkl_malloc(size_t bytes, int alignment);
void mkl_free(void* p);
/* as in the real mkl_service.h: the lower-case name is a macro for the C entry point (the lower-case SYMBOL in
 * the library is the Fortran by-reference variant) */
int MKL_Set_Num_Threads_Local(int nt);
#define mkl_set_num_threads_local MKL_Set_Num_Threads_Local

/* include/tasks/gemm_task.h:87-91, drivers/in_mem_gemm.cpp:64-67, drivers/in_mem_kmeans.cpp:34-42 */
void cblas_sgemm(CBLAS_LAYOUT layout, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, MKL_INT m, MKL_INT n, MKL_INT k,
                 float alpha, const float* a, MKL_INT lda, const float* b, MKL_INT ldb, float beta, float* c,
                 MKL_INT ldc);
void cblas_dgemm(CBLAS_LAYOUT layout, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, MKL_INT m, MKL_INT n, MKL_INT k,
                 double alpha, const double* a, MKL_INT lda, const double* b, MKL_INT ldb, double beta, double* c,
                 MKL_INT ldc);
void cblas_sgemv(CBLAS_LAYOUT layout, CBLAS_TRANSPOSE ta, MKL_INT m, MKL_INT n, float alpha, const float* a,
                 MKL_INT lda, const float* x, MKL_INT incx, float beta, float* y, MKL_INT incy);
void cblas_saxpy(MKL_INT n, float alpha, const float* x, MKL_INT incx, float* y, MKL_INT incy);
float cblas_sdot(MKL_INT n, const float* x, MKL_INT incx, const float* y, MKL_INT incy);
size_t cblas_isamin(MKL_INT n, const float* x, MKL_INT incx);

/* include/tasks/csrmm_task.h:224-228,306-312; drivers/in_mem_csrmm.cpp:113-117 */
void mkl_scsrmm(const char* transa, const MKL_INT* m, const MKL_INT* n, const MKL_INT* k, const float* alpha,
                const char* matdescra, const float* val, const MKL_INT* indx, const MKL_INT* pntrb,
                const MKL_INT* pntre, const float* b, const MKL_INT* ldb, const float* beta, float* c,
                const MKL_INT* ldc);
/* include/tasks/csrgemv_task.h:74,165; drivers/in_mem_csrgemv.cpp */
void mkl_cspblas_scsrgemv(const char* transa, const MKL_INT* m, const float* a, const MKL_INT* ia,
                          const MKL_INT* ja, const float* x, float* y);
/* include/tasks/csrcsc_task.h:68-74; drivers/in_mem_csrcsc.cpp:72-78 */
void mkl_scsrcsc(const MKL_INT* job, const MKL_INT* n, float* acsr, MKL_INT* ja, MKL_INT* ia, float* acsc,
                 MKL_INT* ja1, MKL_INT* ia1, MKL_INT* info);

#ifdef __cplusplus
}
#endif
