// TEST INFRASTRUCTURE — entry point that calls the reference's own flash::csrcsc (src/blas/csrcsc.cpp:32-159,
// linked unmodified from oracle/_ref/libfblas.a).  The reference's drivers/csrcsc.cpp cannot be used as it is:
// it hard-codes the scratch directory "/raid/tmp/" (drivers/csrcsc.cpp:40), which does not exist here and
// lies outside this repository.  Same positional arguments as that driver plus the scratch directory:
//   ref_csrcsc <a_vals> <a_idxs> <a_offs> <atr_vals> <atr_idxs> <atr_offs> <n_rows> <n_cols> <scratch_dir/>
#include <cstdio>
#include <string>

#include "flash_blas.h"
#include "lib_funcs.h"

int main(int argc, char** argv) {
  if (argc != 10) {
    std::fprintf(stderr, "usage: %s a_vals a_idxs a_offs atr_vals atr_idxs atr_offs n_rows n_cols scratch_dir/\n", argv[0]);
    return 2;
  }
  std::string names[6];
  for (int i = 0; i < 6; i++) names[i] = argv[1 + i];
  FBLAS_UINT m = std::stoull(argv[7]), n = std::stoull(argv[8]);
  flash::flash_setup(argv[9]);
  auto a = flash::map_file<FPTYPE>(names[0], flash::Mode::READWRITE);
  auto ja = flash::map_file<MKL_INT>(names[1], flash::Mode::READWRITE);
  auto ia = flash::map_file<MKL_INT>(names[2], flash::Mode::READWRITE);
  auto a_tr = flash::map_file<FPTYPE>(names[3], flash::Mode::READWRITE);
  auto ja_tr = flash::map_file<MKL_INT>(names[4], flash::Mode::READWRITE);
  auto ia_tr = flash::map_file<MKL_INT>(names[5], flash::Mode::READWRITE);
  FBLAS_INT ret = flash::csrcsc(m, n, ia, ja, a, ia_tr, ja_tr, a_tr);
  flash::unmap_file(a);
  flash::unmap_file(ja);
  flash::unmap_file(ia);
  flash::unmap_file(a_tr);
  flash::unmap_file(ja_tr);
  flash::unmap_file(ia_tr);
  flash::flash_destroy();
  return (int) ret;
}
