// tsb_common.h — shared declarations of libtsb (host + device).
#pragma once

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#ifdef TSB_HOST_EMU
#include "emu/cuda_emu.h"
#else
#include <cuda_runtime.h>
#include "emu/cuda_emu.h"  // only the TSB_LAUNCH / TSB_DYN_SMEM macros in this configuration
#endif

#include "../../include/tsb.h"

#define TSB_HD __host__ __device__ __forceinline__
#define TSB_D __device__ __forceinline__

namespace tsb {

void set_error(const char* fmt, ...);
const char* last_error();

struct Error {
  int code;
};

#define TSB_CHECK(cond, code, ...)            \
  do {                                        \
    if (!(cond)) {                            \
      ::tsb::set_error(__VA_ARGS__);          \
      throw ::tsb::Error{code};               \
    }                                         \
  } while (0)

#define TSB_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t e__ = (expr);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      ::tsb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),     \
                       __FILE__, __LINE__);                                         \
      throw ::tsb::Error{TSB_ERR_CUDA};                                             \
    }                                                                               \
  } while (0)

inline int dtype_size(int dt) { return dt == TSB_U8 ? 1 : dt == TSB_I32 ? 4 : 8; }
inline bool is_pow2(int64_t x) { return x > 0 && (x & (x - 1)) == 0; }
inline int ilog2(int64_t x) {
  int r = 0;
  while ((int64_t(1) << (r + 1)) <= x) r++;
  return r;
}

}  // namespace tsb
