// csrmm_pmem driver -- CLI of the reference's drivers/csrmm_pmem.cpp:14-17 (B and C live in host memory, A on
// flash; the in-memory overload include/flash_blas.h:43-46):
//   <vals_A> <indices_A> <offsets_A> <vals_B> <vals_C> <A_nrows> <A_ncols> <B_ncols> <alpha> <beta>
//   <trans_a> <ord_b> [<n_calls>]
// The optional last argument is ours: with n_calls > 1 the matrix is pinned in HBM (flash::csr_pin) and the
// product is repeated from the same B and the original C -- the access pattern of the eigensolver that
// re-multiplies one A -- and the time of the first and of the later calls is reported.  C on disk is the result
// of one product, as the reference's driver writes it (drivers/csrmm_pmem.cpp:75-78).
#include "driver_common.h"

int main(int argc, char** argv) {
  if (argc != 13 && argc != 14)
    drv::usage_exit("csrmm_pmem <vals_A> <indices_A> <offsets_A> <vals_B> <vals_C> <A_nrows> <A_ncols> <B_ncols> "
                    "<alpha> <beta> <trans_a N|T> <ord_b R|C> [<n_calls>]");
  const FBLAS_UINT m = drv::to_u(argv[6]), n = drv::to_u(argv[7]), k = drv::to_u(argv[8]);
  const FPTYPE alpha = drv::to_f(argv[9]), beta = drv::to_f(argv[10]);
  const CHAR trans_a = argv[11][0], ord_b = argv[12][0];
  const int n_calls = argc == 14 ? std::atoi(argv[13]) : 1;
  flash::flash_setup("/tmp/");
  auto a = flash::map_file<FPTYPE>(argv[1], flash::Mode::READWRITE);
  auto ja = flash::map_file<MKL_INT>(argv[2], flash::Mode::READWRITE);
  auto ia = flash::map_file<MKL_INT>(argv[3], flash::Mode::READWRITE);
  const FBLAS_UINT b_rows = trans_a == 'N' ? n : m, c_rows = trans_a == 'N' ? m : n;
  const std::vector<FPTYPE> b = drv::read_file<FPTYPE>(argv[4], b_rows * k);
  const std::vector<FPTYPE> c0 = drv::read_file<FPTYPE>(argv[5], c_rows * k);
  std::vector<FPTYPE> c = c0;
  FBLAS_INT rc = 0;
  if (n_calls > 1) {
    drv::StopWatch sw;
    rc = flash::csr_pin(m, n, a, ia, ja, trans_a == 'T');
    drv::report("csr_pin()", sw.seconds(), rc);
  }
  for (int it = 0; it < n_calls && rc == 0; ++it) {
    c = c0;
    drv::StopWatch sw;
    rc = flash::csrmm(trans_a, m, n, k, alpha, beta, a, ia, ja, ord_b, const_cast<FPTYPE*>(b.data()), c.data());
    drv::report(it == 0 ? "csrmm() first call" : "csrmm() repeat call", sw.seconds(), rc);
  }
  if (rc == 0) {
    // in-place overwrite of the leading c_rows * k values, like the reference's fstream write
    std::fstream out(argv[5], std::ios::binary | std::ios::in | std::ios::out);
    out.write(reinterpret_cast<const char*>(c.data()), (std::streamsize)(c.size() * sizeof(FPTYPE)));
  }
  flash::unmap_file(a);
  flash::unmap_file(ia);
  flash::unmap_file(ja);
  flash::flash_destroy();
  return rc == 0 ? 0 : 1;
}
