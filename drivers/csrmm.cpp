// csrmm driver -- CLI of the reference's drivers/csrmm.cpp:12-16:
//   <vals_A> <indices_A> <offsets_A> <vals_B> <vals_C> <A_nrows> <A_ncols> <B_ncols> <alpha> <beta>
//   <trans_a> <ord_b>
// Note the argument swap the reference performs (drivers/csrmm.cpp:63-64): files come as values,
// indices, offsets but flash::csrmm takes (a, ia = offsets, ja = indices).
#include "driver_common.h"

int main(int argc, char** argv) {
  if (argc != 13)
    drv::usage_exit("csrmm <vals_A> <indices_A> <offsets_A> <vals_B> <vals_C> <A_nrows> <A_ncols> <B_ncols> "
                    "<alpha> <beta> <trans_a N|T> <ord_b R|C>");
  flash::flash_setup("/tmp/");
  auto a = flash::map_file<FPTYPE>(argv[1], flash::Mode::READWRITE);
  auto ja = flash::map_file<MKL_INT>(argv[2], flash::Mode::READWRITE);
  auto ia = flash::map_file<MKL_INT>(argv[3], flash::Mode::READWRITE);
  auto b = flash::map_file<FPTYPE>(argv[4], flash::Mode::READWRITE);
  auto c = flash::map_file<FPTYPE>(argv[5], flash::Mode::READWRITE);
  drv::StopWatch sw;
  const FBLAS_INT rc = flash::csrmm(argv[11][0], drv::to_u(argv[6]), drv::to_u(argv[7]), drv::to_u(argv[8]),
                                    drv::to_f(argv[9]), drv::to_f(argv[10]), a, ia, ja, argv[12][0], b, c);
  drv::report("csrmm()", sw.seconds(), rc);
  for (auto p : {a, b, c}) flash::unmap_file(p);
  flash::unmap_file(ia);
  flash::unmap_file(ja);
  flash::flash_destroy();
  return rc == 0 ? 0 : 1;
}
