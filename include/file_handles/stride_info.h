// StrideInfo: the strided-access descriptor of the flash:: file-handle interface
// (reference include/file_handles/file_handle.h:19-34) -- `n_strides` runs of `len_per_stride` bytes whose
// starts are `stride` bytes apart, stride >= len_per_stride.  Field names and the string form
// "stride:n_strides:len_per_stride" are part of the API (the reference hashes tiles by that string).
#pragma once

#include <string>

#include "bof_types.h"

namespace flash {

struct StrideInfo {
  FBLAS_UINT stride = 0;
  FBLAS_UINT n_strides = 0;
  FBLAS_UINT len_per_stride = 0;

  // bytes the description covers when packed back to back
  FBLAS_UINT packed_bytes() const { return n_strides * len_per_stride; }
  // bytes of file between the first and one past the last byte it touches
  FBLAS_UINT span_bytes() const { return n_strides == 0 ? 0 : (n_strides - 1) * stride + len_per_stride; }

  operator std::string() const;
};

inline bool operator==(const StrideInfo& a, const StrideInfo& b) {
  return a.stride == b.stride && a.n_strides == b.n_strides && a.len_per_stride == b.len_per_stride;
}
inline bool operator!=(const StrideInfo& a, const StrideInfo& b) { return !(a == b); }

inline StrideInfo::operator std::string() const {
  std::string out;
  for (FBLAS_UINT field : {stride, n_strides, len_per_stride}) {
    if (!out.empty()) out += ':';
    out += std::to_string(field);
  }
  return out;
}

}  // namespace flash
