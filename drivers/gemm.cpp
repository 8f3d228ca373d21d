// gemm driver -- CLI of the reference's drivers/gemm.cpp:17-24,41-51:
//   <mat_A_file> <mat_B_file> <mat_C_file> <A_nrows> <A_ncols> <B_ncols> <alpha> <beta>
//   <a transpose?> <b transpose?> <matr order> <lda_a> <lda_b> <lda_c>
#include "driver_common.h"

int main(int argc, char** argv) {
  if (argc != 15)
    drv::usage_exit("gemm <A> <B> <C> <A_nrows> <A_ncols> <B_ncols> <alpha> <beta> <transA N|T> <transB N|T> "
                    "<order R|C> <lda_a> <lda_b> <lda_c>");
  flash::flash_setup("/tmp/gemm_driver_temps");
  auto A = flash::map_file<FPTYPE>(argv[1], flash::Mode::READWRITE);
  auto B = flash::map_file<FPTYPE>(argv[2], flash::Mode::READWRITE);
  auto C = flash::map_file<FPTYPE>(argv[3], flash::Mode::READWRITE);
  const FBLAS_UINT m = drv::to_u(argv[4]), k = drv::to_u(argv[5]), n = drv::to_u(argv[6]);
  drv::StopWatch sw;
  const FBLAS_INT rc = flash::gemm(argv[11][0], argv[9][0], argv[10][0], m, n, k, drv::to_f(argv[7]), drv::to_f(argv[8]),
                                   A, B, C, drv::to_u(argv[12]), drv::to_u(argv[13]), drv::to_u(argv[14]));
  drv::report("gemm()", sw.seconds(), rc);
  flash::unmap_file(A);
  flash::unmap_file(B);
  flash::unmap_file(C);
  flash::flash_destroy();
  return rc == 0 ? 0 : 1;
}
