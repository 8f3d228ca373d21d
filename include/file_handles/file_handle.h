// File-handle interface of the flash:: API (reference include/file_handles/file_handle.h:39-73).
// Only the blocking contiguous/strided read/write/copy surface survives: the asynchronous libaio
// machinery behind it (src/file_handles/flash_file_handle.cpp) is replaced by the CUDA-stream tile
// pipeline inside libbof_b200, which reads operands straight from the file mapping.
#pragma once

#include <functional>
#include <string>

#include "bof_types.h"

namespace flash {

enum class Mode { READ, WRITE, READWRITE };

// `n_strides` runs of `len_per_stride` bytes, `stride` bytes apart (stride >= len_per_stride).
struct StrideInfo {
  FBLAS_UINT stride = 0;
  FBLAS_UINT n_strides = 0;
  FBLAS_UINT len_per_stride = 0;

  bool operator==(const StrideInfo& o) const {
    return stride == o.stride && n_strides == o.n_strides && len_per_stride == o.len_per_stride;
  }
  operator std::string() const {
    return std::to_string(stride) + ":" + std::to_string(n_strides) + ":" + std::to_string(len_per_stride);
  }
};

using Callback = std::function<void(void)>;
extern Callback dummy_std_func;

class BaseFileHandle {
 public:
  virtual ~BaseFileHandle() = default;

  virtual FBLAS_INT open(std::string& fname, Mode fmode, FBLAS_UINT size = 0) = 0;
  virtual FBLAS_INT close() = 0;

  // contiguous; all calls block and return 0 on success, -1 on failure
  virtual FBLAS_INT read(FBLAS_UINT offset, FBLAS_UINT len, void* buf, const Callback& cb = dummy_std_func) = 0;
  virtual FBLAS_INT write(FBLAS_UINT offset, FBLAS_UINT len, void* buf, const Callback& cb = dummy_std_func) = 0;
  virtual FBLAS_INT copy(FBLAS_UINT self_offset, BaseFileHandle& dest, FBLAS_UINT dest_offset, FBLAS_UINT len,
                         const Callback& cb = dummy_std_func) = 0;

  // strided: the strides are packed back to back in `buf`
  virtual FBLAS_INT sread(FBLAS_UINT offset, StrideInfo sinfo, void* buf, const Callback& cb = dummy_std_func) = 0;
  virtual FBLAS_INT swrite(FBLAS_UINT offset, StrideInfo sinfo, void* buf, const Callback& cb = dummy_std_func) = 0;
  virtual FBLAS_INT scopy(FBLAS_UINT self_offset, BaseFileHandle& dest, FBLAS_UINT dest_offset, StrideInfo sinfo,
                          const Callback& cb = dummy_std_func) = 0;
};

}  // namespace flash
