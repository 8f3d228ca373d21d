// TEST INFRASTRUCTURE — entry point that calls the reference's own flash::csrgemv (src/blas/csrgemv.cpp:82-97,
// linked unmodified from oracle/_ref/libfblas.a).  The reference's drivers/csrgemv.cpp never calls
// flash::flash_setup, so its main thread has no AIO context and the first synchronous read of the offsets
// (csrgemv.cpp:86) fails with EINVAL and the process hangs in exit(); this main registers the thread first.
//   ref_csrgemv <a_vals> <a_idxs> <a_offs> <x> <y> <a_nrows> <a_ncols> <trans_a>
#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

#include "flash_blas.h"
#include "lib_funcs.h"

int main(int argc, char** argv) {
  if (argc != 9) {
    std::fprintf(stderr, "usage: %s a_vals a_idxs a_offs x y a_nrows a_ncols trans_a\n", argv[0]);
    return 2;
  }
  std::string fv = argv[1], fi = argv[2], fo = argv[3], fx = argv[4], fy = argv[5];
  FBLAS_UINT m = std::stoull(argv[6]), n = std::stoull(argv[7]);
  char trans = argv[8][0];
  flash::flash_setup("./");
  auto a = flash::map_file<FPTYPE>(fv, flash::Mode::READWRITE);
  auto ja = flash::map_file<MKL_INT>(fi, flash::Mode::READWRITE);
  auto ia = flash::map_file<MKL_INT>(fo, flash::Mode::READWRITE);
  FBLAS_UINT xlen = trans == 'N' ? n : m, ylen = trans == 'N' ? m : n;
  std::vector<FPTYPE> x(xlen), y(ylen, 0.0f);
  std::ifstream in(fx, std::ios::binary);
  in.read((char*) x.data(), xlen * sizeof(FPTYPE));
  in.close();
  FBLAS_INT ret = flash::csrgemv(trans, m, n, a, ia, ja, x.data(), y.data());
  flash::unmap_file(a);
  flash::unmap_file(ja);
  flash::unmap_file(ia);
  std::ofstream out(fy, std::ios::binary);
  out.write((char*) y.data(), ylen * sizeof(FPTYPE));
  out.close();
  flash::flash_destroy();
  return (int) ret;
}
