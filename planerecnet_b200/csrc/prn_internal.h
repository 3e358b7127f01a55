// prn_internal.h — shared host-side helpers for libprn_b200 (error reporting, driver entry points).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/prn_b200.h"

namespace prn {

int set_error(int code, const char* fmt, ...);

#define PRN_REQUIRE(cond, ...)                                  \
  do {                                                          \
    if (!(cond)) return ::prn::set_error(PRN_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define PRN_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return ::prn::set_error(PRN_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));    \
  } while (0)

// 2D row-major [rows][cols] 16-bit matrix -> tensor map with box {64 cols, box_rows}, 128B swizzle.
// pitch_elems: row pitch in elements (0 = cols).
int encode_tmap_2d_sw128(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                         int dtype, uint64_t pitch_elems = 0);

// General tiled tensor map over a 16-bit tensor: dims / box innermost first, strides_bytes[i] = byte stride of dimension
// i + 1 (rank - 1 entries), swizzle_bytes in {0, 32, 64, 128} (box[0] * 2 bytes must not exceed it), OOB elements read as 0.
int encode_tmap_nd(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes, int dtype);

int sm_count();

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace prn
