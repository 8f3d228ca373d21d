// CPU-side check of the C++ drop-in layer (no GPU work): flash_ptr arithmetic and coercions, map_file /
// unmap_file / flash_malloc / flash_free, read_sync / write_sync / flash_memset / flash_memcpy on file-backed
// pointers, the mapping registry, and the error convention of the flash:: kernels (-1, never exit()).
// Prints HOST_LAYER_OK on success; any failed check aborts with a message.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <numeric>
#include <vector>

#include "flash_blas.h"
#include "lib_funcs.h"

#define CHECK(cond)                                                          \
  do {                                                                       \
    if (!(cond)) {                                                           \
      std::fprintf(stderr, "check failed at line %d: %s\n", __LINE__, #cond); \
      return 1;                                                              \
    }                                                                        \
  } while (0)

int main(int argc, char** argv) {
  if (argc != 2) return 2;
  const std::string dir = std::string(argv[1]) + "/";
  flash::mnt_dir = dir;  // flash_setup() would also create the GPU context; not wanted here

  // a file of 4096 floats: 0, 1, 2, ...
  const std::string fname = dir + "iota.bin";
  {
    std::vector<float> v(4096);
    std::iota(v.begin(), v.end(), 0.f);
    std::ofstream(fname, std::ios::binary).write(reinterpret_cast<const char*>(v.data()), v.size() * sizeof(float));
  }
  auto p = flash::map_file<FPTYPE>(fname, flash::Mode::READWRITE);
  CHECK(p.ptr != nullptr && p.foffset == 0 && p.fop != nullptr);
  CHECK(p.ptr[17] == 17.f && *(p + 100) == 100.f);

  // pointer arithmetic moves address and file offset together; coercion keeps the position
  auto q = p + 10;
  CHECK(q.ptr == p.ptr + 10 && q.foffset == 10 * sizeof(FPTYPE) && q.fop == p.fop);
  flash::flash_ptr<void> qv = q;
  CHECK(qv.get_raw_ptr() == static_cast<void*>(p.ptr + 10) && qv.foffset == q.foffset);
  flash::flash_ptr<MKL_INT> qi = q;  // reinterpretation, same position
  CHECK(reinterpret_cast<void*>(qi.ptr) == static_cast<void*>(q.ptr));
  CHECK(q == qv && !(p == q));
  CHECK(flash::FlashPtrEq()(qv, flash::flash_ptr<void>(q)) && flash::FlashPtrHasher()(qv) == flash::FlashPtrHasher()(flash::flash_ptr<void>(q)));
  bool threw = false;
  try { flash::flash_ptr<FPTYPE> bad(p.ptr); (void)bad; } catch (const std::logic_error&) { threw = true; }
  CHECK(threw);

  // the mapping is registered with the library (base address -> descriptor)
  CHECK(bof_unregister_mapping(p.ptr) == 0);             // map_file registered it
  CHECK(bof_unregister_mapping(p.ptr) != 0);             // and only once
  CHECK(bof_register_mapping(p.ptr, 4096 * sizeof(float), 0, 0) == 0);
  CHECK(bof_register_mapping(nullptr, 16, 0, 0) != 0 && bof_register_mapping(p.ptr, 0, 0, 0) != 0);

  // synchronous helpers go through the file handle
  std::vector<FPTYPE> buf(8);
  CHECK(flash::read_sync(buf.data(), p + 32, buf.size()) == 0 && buf[0] == 32.f && buf[7] == 39.f);
  for (auto& x : buf) x = -1.f;
  CHECK(flash::write_sync(p + 64, buf.data(), buf.size()) == 0 && p.ptr[64] == -1.f && p.ptr[71] == -1.f && p.ptr[72] == 72.f);
  flash::flash_memset(p + 128, 0, 16 * sizeof(FPTYPE));
  CHECK(p.ptr[128] == 0.f && p.ptr[143] == 0.f && p.ptr[144] == 144.f);

  // a temporary on "flash": file name, size rounding, contents, removal
  auto t = flash::flash_malloc<FPTYPE>(1000 * sizeof(FPTYPE), "scratch");
  const std::string tname = dir + "tmp_scratch_4096";
  CHECK(std::ifstream(tname).good());
  auto src = p + 200;
  flash::flash_memcpy(t, src, 50 * sizeof(FPTYPE));
  CHECK(t.ptr[0] == 200.f && t.ptr[49] == 249.f);
  flash::flash_free(t);
  CHECK(!std::ifstream(tname).good());
  threw = false;
  try { flash::flash_malloc<FPTYPE>(0); } catch (const std::invalid_argument&) { threw = true; }
  CHECK(threw);
  threw = false;
  try { flash::map_file<FPTYPE>(dir + "does_not_exist.bin", flash::Mode::READ); } catch (const std::runtime_error&) { threw = true; }
  CHECK(threw);

  // error convention: with no GPU the kernels report -1 (and say why on stderr); with a GPU a bad character
  // argument does.  Never an exit() from library code.
  FBLAS_INT rc = flash::gemm('X', 'N', 'N', 4, 4, 4, 1.f, 0.f, p, p + 16, p + 32);
  CHECK(rc == -1);
  rc = flash::csrgemv('Q', 4, 4, p, flash::flash_ptr<MKL_INT>(p), flash::flash_ptr<MKL_INT>(p), buf.data(), buf.data());
  CHECK(rc == -1);

  flash::unmap_file(p);
  flash::flash_destroy();
  std::printf("HOST_LAYER_OK\n");
  return 0;
}
