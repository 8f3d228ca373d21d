// FlashFileHandle: a file on "flash" (reference include/file_handles/flash_file_handle.h).  The
// reference drives O_DIRECT + libaio with per-thread io contexts; here the bulk data path is the
// GPU pipeline reading the mmap of the file, so this class only needs plain pread/pwrite for the
// small synchronous accesses the API exposes (read_sync/write_sync, flash_memset, ...).
#pragma once

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstring>
#include <vector>

#include "file_handles/file_handle.h"

namespace flash {

class FlashFileHandle : public BaseFileHandle {
 public:
  int file_desc = -1;
  FBLAS_UINT file_sz = 0;

  FlashFileHandle() = default;
  ~FlashFileHandle() override { close(); }

  // The reference registers an AIO context per thread (flash_file_handle.cpp:137-190); nothing to do here.
  static void register_thread() {}
  static void deregister_thread() {}

  FBLAS_INT open(std::string& fname, Mode fmode, FBLAS_UINT size = 0) override {
    int flags = (fmode == Mode::READ) ? O_RDONLY : O_RDWR;
    if (fmode != Mode::READ && size != 0) flags |= O_CREAT;
    file_desc = ::open(fname.c_str(), flags, 0666);
    if (file_desc < 0) return -1;
    if (size != 0 && fmode != Mode::READ && ::ftruncate(file_desc, (off_t)size) != 0) return -1;
    struct stat st {};
    if (::fstat(file_desc, &st) != 0) return -1;
    file_sz = (FBLAS_UINT)st.st_size;
    filename_ = fname;
    return 0;
  }

  FBLAS_INT close() override {
    if (file_desc >= 0) {
      ::close(file_desc);
      file_desc = -1;
    }
    return 0;
  }

  const std::string& get_filename() const { return filename_; }

  FBLAS_INT read(FBLAS_UINT offset, FBLAS_UINT len, void* buf, const Callback& cb = dummy_std_func) override {
    const FBLAS_INT rc = xfer(false, offset, len, buf);
    cb();
    return rc;
  }
  FBLAS_INT write(FBLAS_UINT offset, FBLAS_UINT len, void* buf, const Callback& cb = dummy_std_func) override {
    const FBLAS_INT rc = xfer(true, offset, len, buf);
    cb();
    return rc;
  }
  FBLAS_INT copy(FBLAS_UINT self_offset, BaseFileHandle& dest, FBLAS_UINT dest_offset, FBLAS_UINT len,
                 const Callback& cb = dummy_std_func) override {
    std::vector<char> tmp(std::min<FBLAS_UINT>(len, 32u << 20));
    FBLAS_INT rc = 0;
    for (FBLAS_UINT done = 0; done < len && rc == 0; done += tmp.size()) {
      const FBLAS_UINT n = std::min<FBLAS_UINT>(tmp.size(), len - done);
      rc = read(self_offset + done, n, tmp.data());
      if (rc == 0) rc = dest.write(dest_offset + done, n, tmp.data());
    }
    cb();
    return rc;
  }
  FBLAS_INT sread(FBLAS_UINT offset, StrideInfo s, void* buf, const Callback& cb = dummy_std_func) override {
    FBLAS_INT rc = 0;
    for (FBLAS_UINT i = 0; i < s.n_strides && rc == 0; ++i)
      rc = xfer(false, offset + i * s.stride, s.len_per_stride, (char*)buf + i * s.len_per_stride);
    cb();
    return rc;
  }
  FBLAS_INT swrite(FBLAS_UINT offset, StrideInfo s, void* buf, const Callback& cb = dummy_std_func) override {
    FBLAS_INT rc = 0;
    for (FBLAS_UINT i = 0; i < s.n_strides && rc == 0; ++i)
      rc = xfer(true, offset + i * s.stride, s.len_per_stride, (char*)buf + i * s.len_per_stride);
    cb();
    return rc;
  }
  FBLAS_INT scopy(FBLAS_UINT self_offset, BaseFileHandle& dest, FBLAS_UINT dest_offset, StrideInfo s,
                  const Callback& cb = dummy_std_func) override {
    std::vector<char> tmp(s.len_per_stride);
    FBLAS_INT rc = 0;
    for (FBLAS_UINT i = 0; i < s.n_strides && rc == 0; ++i) {
      rc = xfer(false, self_offset + i * s.stride, s.len_per_stride, tmp.data());
      if (rc == 0) rc = dest.write(dest_offset + i * s.stride, s.len_per_stride, tmp.data());
    }
    cb();
    return rc;
  }

 private:
  std::string filename_;

  FBLAS_INT xfer(bool wr, FBLAS_UINT offset, FBLAS_UINT len, void* buf) {
    char* p = static_cast<char*>(buf);
    while (len > 0) {
      const ssize_t n = wr ? ::pwrite(file_desc, p, len, (off_t)offset) : ::pread(file_desc, p, len, (off_t)offset);
      if (n < 0 && errno == EINTR) continue;
      if (n <= 0) return -1;
      p += n;
      offset += (FBLAS_UINT)n;
      len -= (FBLAS_UINT)n;
    }
    return 0;
  }
};

}  // namespace flash
