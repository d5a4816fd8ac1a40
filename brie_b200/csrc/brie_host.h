// Host-side helpers shared by the translation units of libbrie_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/brie_b200.h"

namespace brie {

// records the thread-local message returned by brie_last_error() and returns `code`
int fail(int code, const char* fmt, ...);

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// grid-stride launches: enough CTAs to fill the 148 SMs, never more than the work
inline int grid_1d(int64_t n, int block) {
  int64_t g = ceil_div(n, block);
  if (g > 148 * 32) g = 148 * 32;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace brie

#define BRIE_CUDA(call)                                                                        \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return brie::fail(BRIE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                   \
  } while (0)
