// csrcsc driver -- CLI of the reference's drivers/csrcsc.cpp / drivers/in_mem_csrcsc.cpp:14-16:
//   <vals_a> <indices_a> <offsets_a> <vals_a_tr> <indices_a_tr> <offsets_a_tr> <n_rows> <n_cols>
// The three output files are created here with their final sizes (nnz values, nnz int64 indices,
// n_cols + 1 int64 offsets), as the in-memory driver writes them (in_mem_csrcsc.cpp:82-95).
#include "driver_common.h"

static void make_file(const char* name, size_t bytes) {
  std::string f(name);
  flash::FlashFileHandle fh;
  if (fh.open(f, flash::Mode::READWRITE, bytes == 0 ? 8 : bytes) != 0) {
    std::fprintf(stderr, "cannot create %s\n", name);
    std::exit(1);
  }
}

int main(int argc, char** argv) {
  if (argc != 9)
    drv::usage_exit("csrcsc <vals_a> <indices_a> <offsets_a> <vals_a_tr> <indices_a_tr> <offsets_a_tr> <n_rows> <n_cols>");
  flash::flash_setup("/tmp/");
  const FBLAS_UINT m = drv::to_u(argv[7]), n = drv::to_u(argv[8]);
  auto a = flash::map_file<FPTYPE>(argv[1], flash::Mode::READWRITE);
  auto ja = flash::map_file<MKL_INT>(argv[2], flash::Mode::READWRITE);
  auto ia = flash::map_file<MKL_INT>(argv[3], flash::Mode::READWRITE);
  const FBLAS_UINT nnz = (FBLAS_UINT)(ia.ptr[m] - ia.ptr[0]);
  make_file(argv[4], nnz * sizeof(FPTYPE));
  make_file(argv[5], nnz * sizeof(MKL_INT));
  make_file(argv[6], (n + 1) * sizeof(MKL_INT));
  auto a_tr = flash::map_file<FPTYPE>(argv[4], flash::Mode::READWRITE);
  auto ja_tr = flash::map_file<MKL_INT>(argv[5], flash::Mode::READWRITE);
  auto ia_tr = flash::map_file<MKL_INT>(argv[6], flash::Mode::READWRITE);
  drv::StopWatch sw;
  const FBLAS_INT rc = flash::csrcsc(m, n, ia, ja, a, ia_tr, ja_tr, a_tr);
  drv::report("csrcsc()", sw.seconds(), rc);
  flash::unmap_file(a);
  flash::unmap_file(a_tr);
  for (auto p : {ia, ja, ia_tr, ja_tr}) flash::unmap_file(p);
  flash::flash_destroy();
  return rc == 0 ? 0 : 1;
}
