// Symmetric eigendecomposition of the p x p Gram: the ONLY library call on the path
// (cuSOLVER Xsyevd / XsyevBatched), timed and reported separately by the host code.
// It replaces torch.linalg.svd inside svd_wrapper (encoding/models/ridge_utils.py:49-67):
// X = U S V^T  =>  X^T X = V S^2 V^T, so lam = S^2 and the eigenvectors are the rows of Vh.
#include "common.cuh"
#include "../../include/litridge.h"

#include <cusolverDn.h>
#include <mutex>

namespace lit {

struct SolverCtx {
  cusolverDnHandle_t handle = nullptr;
  cusolverDnParams_t params = nullptr;
};

// One cuSOLVER handle per calling host thread: syevd blocks its caller, so concurrent decompositions are
// issued from several worker threads, each on its own stream with its own handle and workspace.
static int get_solver(SolverCtx** out) {
  static thread_local SolverCtx ctx;
  if (!ctx.handle) {
    cusolverStatus_t st = cusolverDnCreate(&ctx.handle);
    if (st != CUSOLVER_STATUS_SUCCESS) {
      set_error("cusolverDnCreate failed (%d)", (int)st);
      ctx.handle = nullptr;
      return LIT_ERR_CUDA;
    }
    st = cusolverDnCreateParams(&ctx.params);
    if (st != CUSOLVER_STATUS_SUCCESS) {
      set_error("cusolverDnCreateParams failed (%d)", (int)st);
      return LIT_ERR_CUDA;
    }
  }
  *out = &ctx;
  return LIT_OK;
}

}  // namespace lit

using namespace lit;

extern "C" int lit_syevd_workspace(int n, int dtype, int batch, size_t* device_bytes, size_t* host_bytes) {
  LIT_REQUIRE(n > 0 && batch > 0 && device_bytes && host_bytes, "syevd_workspace: bad arguments");
  LIT_REQUIRE(dtype == 0 || dtype == 1, "syevd: dtype must be 0 (f32) or 1 (f64)");
  SolverCtx* c;
  int rc = get_solver(&c);
  if (rc) return rc;
  const cudaDataType dt = dtype == 0 ? CUDA_R_32F : CUDA_R_64F;
  cusolverStatus_t st;
  if (batch == 1)
    st = cusolverDnXsyevd_bufferSize(c->handle, c->params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, dt,
                                     nullptr, n, dt, nullptr, dt, device_bytes, host_bytes);
  else
    st = cusolverDnXsyevBatched_bufferSize(c->handle, c->params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n,
                                           dt, nullptr, n, dt, nullptr, dt, device_bytes, host_bytes, batch);
  if (st != CUSOLVER_STATUS_SUCCESS) {
    set_error("cusolver syevd bufferSize failed (%d) for n=%d batch=%d", (int)st, n, batch);
    return LIT_ERR_CUDA;
  }
  return LIT_OK;
}

extern "C" int lit_syevd(void* G, int n, long ld, int dtype, int batch, void* lam, void* work, size_t work_bytes, void* work_h,
                         size_t work_h_bytes, int* info, void* stream) {
  LIT_REQUIRE(n > 0 && batch > 0 && G && lam && info, "syevd: bad arguments");
  LIT_REQUIRE(ld >= n && (batch == 1 || ld == n), "syevd: pitch must be >= n (== n for the batched solver)");
  LIT_REQUIRE(dtype == 0 || dtype == 1, "syevd: dtype must be 0 (f32) or 1 (f64)");
  SolverCtx* c;
  int rc = get_solver(&c);
  if (rc) return rc;
  cusolverStatus_t st = cusolverDnSetStream(c->handle, (cudaStream_t)stream);
  if (st != CUSOLVER_STATUS_SUCCESS) {
    set_error("cusolverDnSetStream failed (%d)", (int)st);
    return LIT_ERR_CUDA;
  }
  const cudaDataType dt = dtype == 0 ? CUDA_R_32F : CUDA_R_64F;
  // The Gram is symmetric, so its row-major storage is also its column-major storage; cuSOLVER
  // returns eigenvector j in (column-major) column j == (row-major) row j, eigenvalues ascending.
  if (batch == 1)
    st = cusolverDnXsyevd(c->handle, c->params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, dt, G, ld, dt, lam,
                          dt, work, work_bytes, work_h, work_h_bytes, info);
  else
    st = cusolverDnXsyevBatched(c->handle, c->params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, dt, G, n, dt,
                                lam, dt, work, work_bytes, work_h, work_h_bytes, info, batch);
  if (st != CUSOLVER_STATUS_SUCCESS) {
    set_error("cusolver syevd failed (%d) for n=%d batch=%d", (int)st, n, batch);
    return LIT_ERR_CUDA;
  }
  return LIT_OK;
}
