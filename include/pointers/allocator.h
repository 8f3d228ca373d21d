// map_file / unmap_file (reference include/pointers/allocator.h:19-59): open the file, mmap it whole,
// hand out a flash_ptr at byte `foffset`.  The mapping is what the GPU pipeline reads from and writes
// to, so MAP_SHARED + (for outputs) PROT_WRITE is all the "flash" there is on this side.
#pragma once

#include <sys/mman.h>

#include <stdexcept>
#include <string>

#include "bof_b200.h"
#include "file_handles/flash_file_handle.h"
#include "pointers/pointer.h"

namespace flash {

template <typename T>
flash_ptr<T> map_file(std::string fname, Mode mode, FBLAS_UINT foffset = 0, int flags = 0) {
  auto* fh = new FlashFileHandle();
  if (fh->open(fname, mode) != 0) {
    delete fh;
    throw std::runtime_error("map_file: cannot open " + fname + ": " + std::strerror(errno));
  }
  if (fh->file_sz == 0 || foffset > fh->file_sz) {
    delete fh;
    throw std::runtime_error("map_file: " + fname + " is empty or shorter than the requested offset");
  }
  const int prot = (mode == Mode::READ) ? PROT_READ : (PROT_READ | PROT_WRITE);
  void* base = ::mmap(nullptr, fh->file_sz, prot, MAP_SHARED | flags, fh->file_desc, 0);
  if (base == MAP_FAILED) {
    delete fh;
    throw std::runtime_error("map_file: mmap of " + fname + " failed: " + std::strerror(errno));
  }
  // let the GPU pipeline stage this file with pread/pwrite instead of faulting the mapping page by page
  bof_register_mapping(base, fh->file_sz, fh->file_desc, 0);
  return flash_ptr<T>(reinterpret_cast<T*>(static_cast<char*>(base) + foffset), foffset, fh);
}

template <typename T>
void unmap_file(flash_ptr<T> fptr) {
  auto* fh = dynamic_cast<FlashFileHandle*>(fptr.fop);
  if (fh == nullptr) return;
  bof_unregister_mapping(reinterpret_cast<char*>(fptr.ptr) - fptr.foffset);
  ::munmap(reinterpret_cast<char*>(fptr.ptr) - fptr.foffset, fh->file_sz);
  delete fh;
}

}  // namespace flash
