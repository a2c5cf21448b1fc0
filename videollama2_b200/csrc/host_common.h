// Host-side helpers shared by the C-ABI translation units: error reporting, launch counting, TMA tensor maps.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vl2.h"

namespace vl2 {

int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);
int sm_count();

// Encode a tiled bf16 tensor map with 128-byte swizzle and zero OOB fill.
// dims/strides are innermost-first; strides (bytes) are given for dims 1..rank-1.
int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);

#define VL2_CHECK_CUDA(expr)                                                                      \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return ::vl2::set_error(VL2_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                                \
  } while (0)

#define VL2_CHECK_LAUNCH(name)                                                                              \
  do {                                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                                    \
    if (_e != cudaSuccess)                                                                                  \
      return ::vl2::set_error(VL2_E_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e));         \
    ::vl2::count_launch();                                                                                  \
  } while (0)

#define VL2_REQUIRE(cond, code, ...)                           \
  do {                                                         \
    if (!(cond)) return ::vl2::set_error(code, __VA_ARGS__);   \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace vl2
