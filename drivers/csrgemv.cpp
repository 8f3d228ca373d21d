// csrgemv driver -- CLI of the reference's drivers/csrgemv.cpp:12-16:
//   <vals_A> <indices_A> <offsets_A> <vals_B> <vals_C> <A_nrows> <A_ncols> <trans_a>
// b and c are host vectors read from / written to plain files (drivers/csrgemv.cpp:37-74).
#include "driver_common.h"

int main(int argc, char** argv) {
  if (argc != 9)
    drv::usage_exit("csrgemv <vals_A> <indices_A> <offsets_A> <vals_B> <vals_C> <A_nrows> <A_ncols> <trans_a N|T>");
  flash::flash_setup("/tmp/");
  const FBLAS_UINT m = drv::to_u(argv[6]), n = drv::to_u(argv[7]);
  const char trans = argv[8][0];
  auto a = flash::map_file<FPTYPE>(argv[1], flash::Mode::READWRITE);
  auto ja = flash::map_file<MKL_INT>(argv[2], flash::Mode::READWRITE);
  auto ia = flash::map_file<MKL_INT>(argv[3], flash::Mode::READWRITE);
  std::vector<FPTYPE> x = drv::read_file<FPTYPE>(argv[4], trans == 'N' ? n : m);
  std::vector<FPTYPE> y(trans == 'N' ? m : n, 0.f);
  drv::StopWatch sw;
  const FBLAS_INT rc = flash::csrgemv(trans, m, n, a, ia, ja, x.data(), y.data());
  drv::report("csrgemv()", sw.seconds(), rc);
  drv::write_file(argv[5], y);
  flash::unmap_file(a);
  flash::unmap_file(ia);
  flash::unmap_file(ja);
  flash::flash_destroy();
  return rc == 0 ? 0 : 1;
}
