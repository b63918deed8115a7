// Shared host-side helpers for the litridge C-ABI translation units.
#pragma once
#include <cstdio>
#include <cstdint>
#include <cstdarg>
#include <cuda_runtime.h>

namespace lit {

// Thread-local last-error string, exported through lit_last_error().
void set_error(const char* fmt, ...);
const char* get_error();

// Number of SMs of the current device (cached per device).
int sm_count();

}  // namespace lit

#define LIT_OK 0
#define LIT_ERR_INVALID (-22) /* -EINVAL  */
#define LIT_ERR_CUDA (-5)     /* -EIO     */
#define LIT_ERR_NOMEM (-12)   /* -ENOMEM  */

#define LIT_CUDA_CHECK(expr)                                                                     \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      lit::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return LIT_ERR_CUDA;                                                                       \
    }                                                                                            \
  } while (0)

#define LIT_REQUIRE(cond, ...)     \
  do {                             \
    if (!(cond)) {                 \
      lit::set_error(__VA_ARGS__); \
      return LIT_ERR_INVALID;      \
    }                              \
  } while (0)

#define LIT_LAUNCH_CHECK() LIT_CUDA_CHECK(cudaGetLastError())
