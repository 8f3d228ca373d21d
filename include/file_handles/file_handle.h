// File-handle interface of the flash:: API (reference include/file_handles/file_handle.h:39-73).
// Only the blocking contiguous/strided read/write/copy surface survives: the asynchronous libaio
// machinery behind it (src/file_handles/flash_file_handle.cpp) is replaced by the CUDA-stream tile
// pipeline inside libbof_b200, which reads operands straight from the file mapping.  Method names,
// argument order and the trailing completion-callback argument are kept so that code written against
// the reference's handles still compiles; every call here blocks, returns 0 / -1, and then runs the
// callback.
#pragma once

#include <functional>
#include <string>

#include "bof_types.h"
#include "file_handles/stride_info.h"

namespace flash {

enum class Mode { READ, WRITE, READWRITE };

using Callback = std::function<void(void)>;
extern Callback dummy_std_func;  // the "no callback" default (a no-op)

class BaseFileHandle {
 public:
  virtual ~BaseFileHandle() = default;

  // lifetime
  virtual FBLAS_INT open(std::string& fname, Mode fmode, FBLAS_UINT size = 0) = 0;
  virtual FBLAS_INT close() = 0;

  // file -> memory: `n_bytes` from byte `pos`, or the strides of `sinfo` starting at `pos` packed into `dst`
  virtual FBLAS_INT read(FBLAS_UINT pos, FBLAS_UINT n_bytes, void* dst, const Callback& done = dummy_std_func) = 0;
  virtual FBLAS_INT sread(FBLAS_UINT pos, StrideInfo sinfo, void* dst, const Callback& done = dummy_std_func) = 0;

  // memory -> file, same addressing
  virtual FBLAS_INT write(FBLAS_UINT pos, FBLAS_UINT n_bytes, void* src, const Callback& done = dummy_std_func) = 0;
  virtual FBLAS_INT swrite(FBLAS_UINT pos, StrideInfo sinfo, void* src, const Callback& done = dummy_std_func) = 0;

  // file -> file (this handle is the source)
  virtual FBLAS_INT copy(FBLAS_UINT pos, BaseFileHandle& to, FBLAS_UINT to_pos, FBLAS_UINT n_bytes,
                         const Callback& done = dummy_std_func) = 0;
  virtual FBLAS_INT scopy(FBLAS_UINT pos, BaseFileHandle& to, FBLAS_UINT to_pos, StrideInfo sinfo,
                          const Callback& done = dummy_std_func) = 0;
};

}  // namespace flash
