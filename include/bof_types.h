// Scalar types of the flash:: API.  Same names and widths as the reference's include/bof_types.h:14-20
// (64-bit integers, fp32 values).  MKL is not a dependency of this implementation: MKL_INT is kept only
// because it appears in the public signatures (ILP64 => int64_t, reference CMakeLists.txt:104).
#pragma once

#include <cfloat>
#include <cstdint>

#ifndef MKL_INT
#define MKL_INT int64_t
#endif

using FBLAS_INT = int64_t;
using FBLAS_UINT = uint64_t;
using CHAR = char;
using FPTYPE = float;       // the B200 path is fp32-only (3xTF32 for GEMM-shaped work)
using LONGFPTYPE = double;

#define FPTYPE_MAX FLT_MAX

#ifndef SECTOR_LEN
#define SECTOR_LEN 512  // reference CMakeLists.txt:44; only used to size flash_malloc'd files
#endif

#ifndef ROUND_UP
#define ROUND_UP(X, Y) (((uint64_t)(X) + (uint64_t)(Y) - 1) / (uint64_t)(Y) * (uint64_t)(Y))
#endif
