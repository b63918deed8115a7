// Error reporting and device queries shared by all litridge translation units.
#include "common.cuh"
#include "../../include/litridge.h"

#include <cstring>

namespace lit {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

int sm_count() {
  static int cached_dev = -1;
  static int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace lit

extern "C" const char* lit_last_error(void) { return lit::get_error(); }
extern "C" int lit_abi_version(void) { return LITRIDGE_ABI_VERSION; }

extern "C" int lit_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* free_bytes, size_t* total_bytes) {
  int dev = 0;
  LIT_CUDA_CHECK(cudaGetDevice(&dev));
  int v = 0;
  if (sm_count) {
    LIT_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    *sm_count = v;
  }
  if (cc_major) {
    LIT_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
    *cc_major = v;
  }
  if (cc_minor) {
    LIT_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
    *cc_minor = v;
  }
  if (free_bytes || total_bytes) {
    size_t f = 0, t = 0;
    LIT_CUDA_CHECK(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
  }
  return LIT_OK;
}

extern "C" int lit_memcpy_2d(void* dst, size_t dpitch_bytes, const void* src, size_t spitch_bytes, size_t width_bytes,
                             size_t height, int kind, void* stream) {
  LIT_REQUIRE(kind >= 1 && kind <= 3, "memcpy_2d: kind must be 1 (H2D), 2 (D2H) or 3 (D2D)");
  if (width_bytes == 0 || height == 0) return LIT_OK;
  const cudaMemcpyKind k =
      kind == 1 ? cudaMemcpyHostToDevice : (kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice);
  LIT_CUDA_CHECK(cudaMemcpy2DAsync(dst, dpitch_bytes, src, spitch_bytes, width_bytes, height, k,
                                   static_cast<cudaStream_t>(stream)));
  return LIT_OK;
}

// 0 = ordinary pageable host memory, 1 = page-locked (pinned / registered) host memory, 2 = device or managed memory.
extern "C" int lit_host_pointer_kind(const void* ptr) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
    cudaGetLastError();  // clear the sticky "invalid value" some drivers return for unregistered pointers
    return 0;
  }
  if (a.type == cudaMemoryTypeHost) return 1;
  if (a.type == cudaMemoryTypeUnregistered) return 0;
  return 2;
}
