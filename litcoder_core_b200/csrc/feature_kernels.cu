// Feature-construction kernels: FIR delay stacking and Lanczos TR-resampling.
// Both are HBM-bound (fp64 output): one read of the source, ndelays (resp. one) coalesced
// fp64 writes; threads map to the contiguous feature axis.
//   FIR.make_delayed        encoding/features/FIR_expander.py:24-43
//   lanczosinterp2D/fun     encoding/downsample/interpdata.py:45-63,87-126
#include "common.cuh"
#include "../../include/litridge.h"

namespace lit {

template <typename T>
__global__ void fir_kernel(const T* __restrict__ stim, long nt, long ndim, long ld_stim,
                           const int32_t* __restrict__ delays, int ndelays, int circpad, double* __restrict__ out,
                           long ld_out) {
  // one thread per (t, delay, pair of columns); consecutive threads walk the columns
  const long half = (ndim + 1) / 2;
  const long total = nt * ndelays * half;
  const long stride = (long)gridDim.x * blockDim.x;
  const bool vec_ok = (ndim % 2 == 0) && (ld_out % 2 == 0);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long c2 = i % half;
    const long rest = i / half;
    const int di = (int)(rest % ndelays);
    const long t = rest / ndelays;
    const long c = c2 * 2;
    const long d = delays[di];
    long ts = t - d;
    bool valid = ts >= 0 && ts < nt;
    if (!valid && circpad && nt > 0) {
      // reference slicing semantics: a shift with |d| < nt wraps around once; with |d| >= nt
      // both slice assignments cover the whole array and the block degenerates to a plain copy
      ts = (d < nt && -d < nt) ? ((ts % nt) + nt) % nt : t;
      valid = true;
    }
    double v0 = 0.0, v1 = 0.0;
    if (valid) {
      v0 = (double)stim[ts * ld_stim + c];
      if (c + 1 < ndim) v1 = (double)stim[ts * ld_stim + c + 1];
    }
    double* o = out + t * ld_out + (long)di * ndim + c;
    if (vec_ok) {
      *reinterpret_cast<double2*>(o) = make_double2(v0, v1);
    } else {
      o[0] = v0;
      if (c + 1 < ndim) o[1] = v1;
    }
  }
}

// Lanczos kernel value exactly as interpdata.lanczosfun: t already multiplied by the cutoff.
__device__ __forceinline__ double lanczos_weight(double t, double window) {
  if (t == 0.0) return 1.0;
  if (fabs(t) > window) return 0.0;
  const double pi = 3.141592653589793;
  const double pit = pi * t;
  return window * sin(pit) * sin(pit / window) / (pi * pi * (t * t));
}

// grid.x = TR index, grid.y = column tile.  The block first evaluates the weights of a chunk
// of samples cooperatively into shared memory, then every thread accumulates its column(s).
template <typename T, int CHUNK>
__global__ void lanczos_kernel(const T* __restrict__ data, long n_samples, long ndim, long ld_data,
                               const double* __restrict__ data_times, const double* __restrict__ tr_times,
                               double window, double cutoff, int rectify, const int32_t* __restrict__ lo,
                               const int32_t* __restrict__ hi, double* __restrict__ out, long ld_out) {
  __shared__ double w_sh[CHUNK];
  const long i = blockIdx.x;
  const long c = (long)blockIdx.y * blockDim.x + threadIdx.x;
  const double tr = tr_times[i];
  const long j_begin = lo ? (long)lo[i] : 0;
  const long j_end = hi ? (long)hi[i] : n_samples;
  // Without a band (lo == NULL) the kernel is the dense product of the reference: zero weights are
  // multiplied too, so that non-finite samples poison the output exactly as np.dot(sincmat, data) does.
  const bool dense = lo == nullptr;
  double acc = 0.0, acc_neg = 0.0;
  for (long j0 = j_begin; j0 < j_end; j0 += CHUNK) {
    const long cnt = (j_end - j0) < CHUNK ? (j_end - j0) : CHUNK;
    __syncthreads();
    for (long q = threadIdx.x; q < cnt; q += blockDim.x)
      w_sh[q] = lanczos_weight((tr - data_times[j0 + q]) * cutoff, window);
    __syncthreads();
    if (c < ndim) {
      for (long q = 0; q < cnt; ++q) {
        const double w = w_sh[q];
        if (w != 0.0 || dense) {
          const double x = (double)data[(j0 + q) * ld_data + c];
          if (rectify) {
            // np.clip keeps NaN; CUDA's fmin / fmax would drop it
            acc_neg = fma(w, x != x ? x : fmin(x, 0.0), acc_neg);
            acc = fma(w, x != x ? x : fmax(x, 0.0), acc);
          } else {
            acc = fma(w, x, acc);
          }
        }
      }
    }
  }
  if (c < ndim) {
    if (rectify) {
      out[i * ld_out + c] = acc_neg;
      out[i * ld_out + ndim + c] = acc;
    } else {
      out[i * ld_out + c] = acc;
    }
  }
}

}  // namespace lit

using namespace lit;

extern "C" int lit_fir_make_delayed(const void* stim, int dtype_in, long nt, long ndim, long ld_stim,
                                    const int32_t* delays, int ndelays, int circpad, double* out, long ld_out,
                                    void* stream) {
  LIT_REQUIRE(nt >= 0 && ndim >= 0 && ndelays >= 0, "fir: negative extent");
  LIT_REQUIRE(ld_stim >= ndim && ld_out >= ndim * ndelays, "fir: pitch too small");
  LIT_REQUIRE(dtype_in == 0 || dtype_in == 1, "fir: dtype_in must be 0 (f32) or 1 (f64)");
  LIT_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "fir: output must be 16-byte aligned");
  const long total = nt * ndelays * ((ndim + 1) / 2);
  if (total == 0) return LIT_OK;
  long grid = (total + 255) / 256;
  const long cap = (long)sm_count() * 64;
  if (grid > cap) grid = cap;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype_in == 0)
    fir_kernel<float><<<(int)grid, 256, 0, s>>>((const float*)stim, nt, ndim, ld_stim, delays, ndelays, circpad, out,
                                                ld_out);
  else
    fir_kernel<double><<<(int)grid, 256, 0, s>>>((const double*)stim, nt, ndim, ld_stim, delays, ndelays, circpad, out,
                                                 ld_out);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}

extern "C" int lit_lanczos_downsample(const void* data, int dtype_in, long n_samples, long ndim, long ld_data,
                                      const double* data_times, const double* tr_times, long n_tr, double window,
                                      double cutoff, int rectify, const int32_t* lo, const int32_t* hi, double* out,
                                      long ld_out, void* stream) {
  LIT_REQUIRE(n_samples >= 0 && ndim >= 0 && n_tr >= 0, "lanczos: negative extent");
  LIT_REQUIRE(ld_data >= ndim && ld_out >= (rectify ? 2 : 1) * ndim, "lanczos: pitch too small");
  LIT_REQUIRE(dtype_in == 0 || dtype_in == 1, "lanczos: dtype_in must be 0 (f32) or 1 (f64)");
  LIT_REQUIRE((lo == nullptr) == (hi == nullptr), "lanczos: lo and hi must be given together");
  LIT_REQUIRE(n_tr <= 2147483647L, "lanczos: too many TRs");
  if (n_tr == 0 || ndim == 0) return LIT_OK;
  const int block = ndim >= 256 ? 256 : (ndim >= 128 ? 128 : 64);
  dim3 grid((unsigned)n_tr, (unsigned)((ndim + block - 1) / block));
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype_in == 0)
    lanczos_kernel<float, 256><<<grid, block, 0, s>>>((const float*)data, n_samples, ndim, ld_data, data_times,
                                                      tr_times, window, cutoff, rectify, lo, hi, out, ld_out);
  else
    lanczos_kernel<double, 256><<<grid, block, 0, s>>>((const double*)data, n_samples, ndim, ld_data, data_times,
                                                       tr_times, window, cutoff, rectify, lo, hi, out, ld_out);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}
