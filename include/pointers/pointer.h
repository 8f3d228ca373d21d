// flash_ptr<T>: the fat pointer of the flash:: API -- {mapped address, byte offset in the file, file
// handle}.  Layout, operators and coercions as the reference's include/pointers/pointer.h:15-60 so
// that caller code compiles unchanged; it is a non-owning, trivially copyable value.
#pragma once

#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <type_traits>

#include "bof_types.h"
#include "file_handles/file_handle.h"

namespace flash {

template <typename T>
struct flash_ptr {
  T* ptr = nullptr;               // address inside the mmap of the file
  FBLAS_UINT foffset = 0;         // the same position as a byte offset from the start of the file
  BaseFileHandle* fop = nullptr;  // handle that owns the descriptor (see map_file / unmap_file)

  flash_ptr() = default;
  flash_ptr(T*) { throw std::logic_error("flash_ptr cannot be built from a bare pointer"); }
  flash_ptr(T* p, FBLAS_UINT off, BaseFileHandle* h) : ptr(p), foffset(off), fop(h) {}

  // element-wise advance: moves the address and the file offset together
  flash_ptr operator+(FBLAS_UINT n_vals) const {
    using Step = typename std::conditional<std::is_void<T>::value, char, T>::type;
    return flash_ptr(reinterpret_cast<T*>(reinterpret_cast<Step*>(ptr) + n_vals), foffset + n_vals * sizeof(Step), fop);
  }

  template <typename X>
  bool operator==(const flash_ptr<X>& o) const {
    return static_cast<const void*>(ptr) == static_cast<const void*>(o.ptr) && foffset == o.foffset && fop == o.fop;
  }

  T* get_raw_ptr() const { return ptr; }

  template <class Q = T>
  typename std::enable_if<!std::is_void<Q>::value, Q>::type& operator*() {
    return *ptr;
  }

  // reinterpretation keeps the position
  template <typename W>
  operator flash_ptr<W>() const {
    return flash_ptr<W>(reinterpret_cast<W*>(ptr), foffset, fop);
  }

  operator std::string() const {
    return "[" + std::to_string(reinterpret_cast<uintptr_t>(fop)) + "-" + std::to_string(foffset) + "]";
  }
};

// hashing / equality on the mapped address only (reference pointer.h:62-75)
struct FlashPtrHasher {
  size_t operator()(const flash_ptr<void>& k) const { return std::hash<void*>()(k.get_raw_ptr()); }
};
struct FlashPtrEq {
  bool operator()(const flash_ptr<void>& a, const flash_ptr<void>& b) const { return a.get_raw_ptr() == b.get_raw_ptr(); }
};

}  // namespace flash
