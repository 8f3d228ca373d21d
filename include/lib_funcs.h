// Library lifecycle and small synchronous helpers of the flash:: API
// (reference include/lib_funcs.h:24-128, src/lib_funcs.cpp:18-34).
//
// flash_setup() creates the per-process B200 context (one GPU per process: CUDA ordinal from
// BOF_DEVICE, else LOCAL_RANK, else 0) that replaces the reference's static scheduler; flash_destroy()
// tears it down.  Kernels are blocking and one-at-a-time per process, as in the reference.
#pragma once

#include <fcntl.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "bof_b200.h"
#include "pointers/allocator.h"
#include "pointers/pointer.h"

namespace flash {

extern std::string mnt_dir;

void flash_setup(std::string mntdir);
void flash_destroy();

// The process-wide context behind flash::* (created on first use if flash_setup was not called).
bof_ctx* flash_context();

template <typename T>
FBLAS_INT read_sync(T* dest, flash_ptr<T> src, size_t len) {
  return src.fop->read(src.foffset, len * sizeof(T), dest);
}
template <typename T>
FBLAS_INT write_sync(flash_ptr<T> dest, T* src, size_t len) {
  return dest.fop->write(dest.foffset, len * sizeof(T), src);
}

template <typename T>
void flash_memset(flash_ptr<T> fptr, int val, FBLAS_UINT n_bytes) {
  std::vector<char> buf(std::min<FBLAS_UINT>(n_bytes, 16u << 20), (char)val);
  for (FBLAS_UINT done = 0; done < n_bytes; done += buf.size())
    fptr.fop->write(fptr.foffset + done, std::min<FBLAS_UINT>(buf.size(), n_bytes - done), buf.data());
}

template <typename T, typename W>
void flash_memcpy(flash_ptr<T> dest, flash_ptr<W>& src, FBLAS_UINT n_bytes) {
  src.fop->copy(src.foffset, *dest.fop, dest.foffset, n_bytes);
}

template <typename T>
void flash_truncate(flash_ptr<T> fptr, uint64_t new_size) {
  auto* fh = dynamic_cast<FlashFileHandle*>(fptr.fop);
  if (fh == nullptr || ::ftruncate(fh->file_desc, (off_t)(fptr.foffset + new_size)) != 0)
    std::fprintf(stderr, "flash_truncate failed: %s\n", std::strerror(errno));
}

// A temporary on "flash": file mnt_dir/tmp_<name>_<bytes>, size rounded to 4 KiB, mapped read-write.
template <typename T>
flash_ptr<T> flash_malloc(FBLAS_UINT n_bytes, std::string opt_name = "") {
  if (n_bytes == 0) throw std::invalid_argument("flash_malloc: cannot allocate 0 bytes");
  n_bytes = ROUND_UP(n_bytes, 4096);
  std::string fname = mnt_dir + "tmp_" + (opt_name.empty() ? "" : opt_name + "_") + std::to_string(n_bytes);
  const int fd = ::open(fname.c_str(), O_RDWR | O_CREAT, 0666);
  if (fd < 0 || ::ftruncate(fd, (off_t)n_bytes) != 0) {
    if (fd >= 0) ::close(fd);
    throw std::runtime_error("flash_malloc: cannot create " + fname + ": " + std::strerror(errno));
  }
  ::close(fd);
  return map_file<T>(fname, Mode::READWRITE);
}

template <typename T>
void flash_free(flash_ptr<T> fptr) {
  auto* fh = dynamic_cast<FlashFileHandle*>(fptr.fop);
  const std::string fname = fh ? fh->get_filename() : std::string();
  unmap_file<T>(fptr);
  if (!fname.empty()) ::remove(fname.c_str());
}

}  // namespace flash
