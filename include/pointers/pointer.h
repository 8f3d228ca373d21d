// flash_ptr<T>: the fat pointer of the flash:: API -- {mapped address, byte offset in the file, file
// handle}.  Member names, their order, the operators and the coercions are those of the reference's
// include/pointers/pointer.h:15-60 so that caller code compiles unchanged; it is a non-owning,
// trivially copyable value.  Unlike the reference (whose `ptr` is a fake address inside a reserved
// range), `ptr` here is a real address: the file is mmap'd, and that address is what the C ABI takes.
#pragma once

#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <type_traits>

#include "bof_types.h"
#include "file_handles/file_handle.h"

namespace flash {

namespace detail {
// element type used for address arithmetic; void pointers step in bytes
template <typename T>
using step_t = typename std::conditional<std::is_void<T>::value, char, T>::type;

inline std::string describe_position(const BaseFileHandle* handle, FBLAS_UINT byte_offset) {
  return "[" + std::to_string(reinterpret_cast<uintptr_t>(handle)) + "-" + std::to_string(byte_offset) + "]";
}
}  // namespace detail

template <typename T>
struct flash_ptr {
  T* ptr = nullptr;               // address inside the mmap of the file
  FBLAS_UINT foffset = 0;         // the same position as a byte offset from the start of the file
  BaseFileHandle* fop = nullptr;  // handle that owns the descriptor (see map_file / unmap_file)

  flash_ptr() = default;
  flash_ptr(T* p, FBLAS_UINT off, BaseFileHandle* h) : ptr(p), foffset(off), fop(h) {}
  // a bare address carries no file position: the reference asserts here, we throw
  flash_ptr(T*) { throw std::logic_error("flash_ptr cannot be built from a bare pointer"); }

  T* get_raw_ptr() const { return ptr; }

  // element-wise advance: moves the address and the file offset together
  flash_ptr operator+(FBLAS_UINT n_vals) const;

  // same position (address, offset and handle), whatever the element types
  template <typename X>
  bool operator==(const flash_ptr<X>& o) const {
    return fop == o.fop && foffset == o.foffset && static_cast<const void*>(ptr) == static_cast<const void*>(o.ptr);
  }

  // dereference (not for flash_ptr<void>)
  template <class Q = T>
  typename std::enable_if<!std::is_void<Q>::value, Q>::type& operator*() {
    return *ptr;
  }

  // reinterpretation keeps the position
  template <typename W>
  operator flash_ptr<W>() const {
    return flash_ptr<W>(reinterpret_cast<W*>(ptr), foffset, fop);
  }

  operator std::string() const { return detail::describe_position(fop, foffset); }
};

template <typename T>
flash_ptr<T> flash_ptr<T>::operator+(FBLAS_UINT n_vals) const {
  auto* advanced = reinterpret_cast<detail::step_t<T>*>(ptr) + n_vals;
  return flash_ptr(reinterpret_cast<T*>(advanced), foffset + n_vals * sizeof(detail::step_t<T>), fop);
}

static_assert(std::is_trivially_copyable<flash_ptr<float>>::value, "flash_ptr is passed by value everywhere");

// hashing / equality on the mapped address only (reference pointer.h:62-75)
struct FlashPtrHasher {
  size_t operator()(const flash_ptr<void>& k) const { return std::hash<void*>()(k.get_raw_ptr()); }
};
struct FlashPtrEq {
  bool operator()(const flash_ptr<void>& a, const flash_ptr<void>& b) const { return a.get_raw_ptr() == b.get_raw_ptr(); }
};

}  // namespace flash
