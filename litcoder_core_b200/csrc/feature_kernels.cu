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

// float32 fast path: one thread per (t, delay, 4 columns): a 16-byte load, two 16-byte stores.
__global__ void fir_f32x4_kernel(const float* __restrict__ stim, long nt, long ndim, long ld_stim,
                                 const int32_t* __restrict__ delays, int ndelays, int circpad,
                                 double* __restrict__ out, long ld_out) {
  const long quads = ndim / 4;
  const long total = nt * ndelays * quads;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long c4 = i % quads;
    const long rest = i / quads;
    const int di = (int)(rest % ndelays);
    const long t = rest / ndelays;
    const long d = delays[di];
    long ts = t - d;
    bool valid = ts >= 0 && ts < nt;
    if (!valid && circpad) {
      ts = (d < nt && -d < nt) ? ((ts % nt) + nt) % nt : t;
      valid = true;
    }
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) v = __ldg(reinterpret_cast<const float4*>(stim + ts * ld_stim + c4 * 4));
    double* o = out + t * ld_out + (long)di * ndim + c4 * 4;
    *reinterpret_cast<double2*>(o) = make_double2((double)v.x, (double)v.y);
    *reinterpret_cast<double2*>(o + 2) = make_double2((double)v.z, (double)v.w);
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
// interpdata.sincfun (:29-42) before renormalisation: B = cutoff, t = newtime - oldtime (seconds).
__device__ __forceinline__ double sinc_weight(double t, double B, double window, int causal) {
  const double two_pi = 6.283185307179586;
  double v = 2.0 * B * sin(two_pi * B * t) / (two_pi * B * t + 1e-20);
  if (fabs(t) > window / (2.0 * B)) v = 0.0;
  if (causal && t < 0.0) v = 0.0;
  return v;
}

// grid.x = group of TRS consecutive TRs, grid.y = column tile.  The block first evaluates the weights of a
// chunk of samples for all its TRs cooperatively into shared memory, then every thread accumulates its
// column for the TRS outputs from ONE read of each sample (the bands of neighbouring TRs overlap almost
// entirely, so a sample row is fetched once per block instead of once per TR).
// KIND 0: Lanczos (optionally rectified output); KIND 1: sinc (optionally causal / renormalised by the
// row sum of the weights, which is accumulated in a first pass over the same sample range).
template <typename T, int CHUNK, int KIND, int TRS>
__global__ void resample_kernel(const T* __restrict__ data, long n_samples, long ndim, long ld_data,
                                const double* __restrict__ data_times, const double* __restrict__ tr_times, long n_tr,
                                double window, double cutoff, int flag_a, int flag_b, const int32_t* __restrict__ lo,
                                const int32_t* __restrict__ hi, double* __restrict__ out, long ld_out) {
  __shared__ double w_sh[TRS][CHUNK];
  __shared__ double red_sh[TRS][8];
  __shared__ double scale_sh[TRS];
  const long i0 = (long)blockIdx.x * TRS;
  const int n_here = (int)((n_tr - i0) < TRS ? (n_tr - i0) : TRS);
  const long c = (long)blockIdx.y * blockDim.x + threadIdx.x;
  const int rectify = KIND == 0 ? flag_a : 0;
  const int causal = KIND == 1 ? flag_a : 0;
  const int renorm = KIND == 1 ? flag_b : 0;
  // Without a band (lo == NULL) the kernel is the dense product of the reference: zero weights are
  // multiplied too, so that non-finite samples poison the output exactly as np.dot(sincmat, data) does.
  const bool dense = lo == nullptr;
  // union of the bands of this block's TRs (each TR still only uses weights inside its own band)
  long j_begin = n_samples, j_end = 0;
  for (int r = 0; r < n_here; ++r) {
    const long b = lo ? (long)lo[i0 + r] : 0, e = hi ? (long)hi[i0 + r] : n_samples;
    j_begin = b < j_begin ? b : j_begin;
    j_end = e > j_end ? e : j_end;
  }
  if (threadIdx.x < TRS) scale_sh[threadIdx.x] = 1.0;
  if (renorm) {  // val = val / np.sum(val) unless the sum is exactly 0 (interpdata.py:36-37)
    for (int r = 0; r < n_here; ++r) {
      const double tr = tr_times[i0 + r];
      const long b = lo ? (long)lo[i0 + r] : 0, e = hi ? (long)hi[i0 + r] : n_samples;
      double part = 0.0;
      for (long j = b + threadIdx.x; j < e; j += blockDim.x) part += sinc_weight(tr - data_times[j], cutoff, window, causal);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if ((threadIdx.x & 31) == 0) red_sh[r][threadIdx.x >> 5] = part;
    }
    __syncthreads();
    if (threadIdx.x < n_here) {
      double tot = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red_sh[threadIdx.x][w];
      if (tot != 0.0) scale_sh[threadIdx.x] = 1.0 / tot;
    }
  }
  __syncthreads();
  double acc[TRS], acc_neg[TRS];
#pragma unroll
  for (int r = 0; r < TRS; ++r) acc[r] = acc_neg[r] = 0.0;
  for (long j0 = j_begin; j0 < j_end; j0 += CHUNK) {
    const long cnt = (j_end - j0) < CHUNK ? (j_end - j0) : CHUNK;
    __syncthreads();
    for (long q = threadIdx.x; q < cnt * TRS; q += blockDim.x) {
      const int r = (int)(q / cnt);
      const long jj = q - (long)r * cnt;
      double w = 0.0;
      if (r < n_here) {
        const long j = j0 + jj;
        const long b = lo ? (long)lo[i0 + r] : 0, e = hi ? (long)hi[i0 + r] : n_samples;
        if (j >= b && j < e) {
          const double dt = tr_times[i0 + r] - data_times[j];
          w = KIND == 0 ? lanczos_weight(dt * cutoff, window) : sinc_weight(dt, cutoff, window, causal) * scale_sh[r];
        }
      }
      w_sh[r][jj] = w;
    }
    __syncthreads();
    if (c < ndim) {
      for (long q = 0; q < cnt; ++q) {
        const double x = (double)data[(j0 + q) * ld_data + c];
        const double xn = x != x ? x : fmin(x, 0.0);  // np.clip keeps NaN; CUDA's fmin / fmax would drop it
        const double xp = x != x ? x : fmax(x, 0.0);
#pragma unroll
        for (int r = 0; r < TRS; ++r) {
          const double w = w_sh[r][q];
          if (w != 0.0 || dense) {
            if (rectify) {
              acc_neg[r] = fma(w, xn, acc_neg[r]);
              acc[r] = fma(w, xp, acc[r]);
            } else {
              acc[r] = fma(w, x, acc[r]);
            }
          }
        }
      }
    }
  }
  if (c < ndim) {
#pragma unroll
    for (int r = 0; r < TRS; ++r) {
      if (r >= n_here) break;
      if (rectify) {
        out[(i0 + r) * ld_out + c] = acc_neg[r];
        out[(i0 + r) * ld_out + ndim + c] = acc[r];
      } else {
        out[(i0 + r) * ld_out + c] = acc[r];
      }
    }
  }
}

// Unweighted rows are accumulated in the INPUT precision, one entry after the other, and divided by the
// count in that precision: exactly what np.sum / np.mean(axis=0) do on a float32 (or float64) block, so the
// result is bit-identical to the reference's before it is widened to float64.
template <typename T>
__global__ void csr_rows_kernel(const T* __restrict__ data, long ndim, long ld_data,
                                const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                                const double* __restrict__ weights, int mean, double* __restrict__ out, long ld_out) {
  const long r = blockIdx.x;
  const long c = (long)blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= ndim) return;
  const int e0 = row_ptr[r], e1 = row_ptr[r + 1];
  double res;
  if (weights) {
    double acc = 0.0;
    for (int e = e0; e < e1; ++e) acc = fma(weights[e], (double)data[(long)col_idx[e] * ld_data + c], acc);
    if (mean && e1 > e0) acc /= (double)(e1 - e0);
    res = acc;
  } else {
    T acc = (T)0;
    for (int e = e0; e < e1; ++e) acc += data[(long)col_idx[e] * ld_data + c];
    if (mean && e1 > e0) acc /= (T)(e1 - e0);
    res = (double)acc;
  }
  out[r * ld_out + c] = res;
}

// Gabor transform magnitude (interpdata.gabor_xfm / gabor_xfm2D, :129-145, and downsampling.py:165):
//   out[i][d * n_freq + f] = | sum_j exp(-0.5 (t_j - tr_i)^2 / (2 sigma^2)) * data[j][d] * exp(i 2 pi f t_j) |
// grid.x = TR, grid.y = tile over the (d, f) pairs; the Gaussian envelope of a chunk of samples is staged
// in shared memory, sin / cos are evaluated per (sample, frequency) by the owning thread.
template <typename T, int CHUNK>
__global__ void gabor_kernel(const T* __restrict__ data, long n_samples, long ndim, long ld_data,
                             const double* __restrict__ data_times, const double* __restrict__ tr_times,
                             const double* __restrict__ freqs, int n_freq, double sigma, double* __restrict__ out,
                             long ld_out) {
  __shared__ double g_sh[CHUNK];
  __shared__ double t_sh[CHUNK];
  const long i = blockIdx.x;
  const long pair = (long)blockIdx.y * blockDim.x + threadIdx.x;
  const bool active = pair < ndim * n_freq;
  const long d = active ? pair / n_freq : 0;
  const double f = active ? freqs[pair % n_freq] : 0.0;
  const double tr = tr_times[i];
  const double two_pi = 6.283185307179586;
  double re = 0.0, im = 0.0;
  for (long j0 = 0; j0 < n_samples; j0 += CHUNK) {
    const long cnt = (n_samples - j0) < CHUNK ? (n_samples - j0) : CHUNK;
    __syncthreads();
    for (long q = threadIdx.x; q < cnt; q += blockDim.x) {
      const double tj = data_times[j0 + q];
      t_sh[q] = tj;
      g_sh[q] = exp(-0.5 * (tj - tr) * (tj - tr) / (2.0 * sigma * sigma));
    }
    __syncthreads();
    if (active) {
      for (long q = 0; q < cnt; ++q) {
        const double gx = g_sh[q] * (double)data[(j0 + q) * ld_data + d];
        double sn, cs;
        sincos(t_sh[q] * f * two_pi, &sn, &cs);
        re = fma(cs, gx, re);
        im = fma(sn, gx, im);
      }
    }
  }
  if (active) out[i * ld_out + pair] = sqrt(re * re + im * im);
}

// ---------------------------------------------------------------------------------------------
// Fused FIR delay stacking + per-story trim + column z-score + nan_to_num, written as fp32 straight into the
// rows of the design matrix that fit_predict consumes (the float64 delayed matrix never exists):
//   trainer.py:203-209 (apply_fir_delays), :236-239 / :250-253 (zs of the trimmed story, nan_to_num),
//   encoding/utils.py:23-29 (zs: population std; a zero-std column is centred only).
// Column (delay i, feature c) of the delayed matrix is feature c shifted by delays[i], so its statistics over
// the trimmed rows [row_start, row_stop) are those of stim[row_start - d .. row_stop - d) with zero fill.
// Block = 32 columns x 8 row slices; fp64 moments about the column's first value (a constant column gets an
// exact zero std), three passes over the (L2-resident) story.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ double fir_value(const T* __restrict__ stim, long nt, long ld_stim, long t, long d, long c,
                                            int circpad) {
  long ts = t - d;
  bool valid = ts >= 0 && ts < nt;
  if (!valid && circpad && nt > 0) {
    ts = (d < nt && -d < nt) ? ((ts % nt) + nt) % nt : t;
    valid = true;
  }
  return valid ? (double)stim[ts * ld_stim + c] : 0.0;
}

template <typename T>
__global__ void __launch_bounds__(256)
fir_zscore_kernel(const T* __restrict__ stim, long nt, long ndim, long ld_stim, const int32_t* __restrict__ delays,
                  int ndelays, int circpad, long row_start, long row_stop, int zscore, float* __restrict__ out,
                  long ld_out) {
  __shared__ double red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long j = (long)blockIdx.x * 32 + tx;  // output column = delay * ndim + feature
  const bool col_ok = j < ndim * ndelays;
  const long c = col_ok ? j % ndim : 0;
  const long d = col_ok ? (long)delays[j / ndim] : 0;
  const long n = row_stop - row_start;
  double mean = 0.0, inv_std = 1.0;
  if (zscore) {
    const double shift = col_ok ? fir_value(stim, nt, ld_stim, row_start, d, c, circpad) : 0.0;
    double acc = 0.0;
    if (col_ok)
      for (long t = row_start + ty; t < row_stop; t += 8) acc += fir_value(stim, nt, ld_stim, t, d, c, circpad) - shift;
    red[ty][tx] = acc;
    __syncthreads();
    acc = 0.0;
#pragma unroll
    for (int y = 0; y < 8; ++y) acc += red[y][tx];
    mean = shift + acc / (double)n;
    __syncthreads();
    double ssq = 0.0;
    if (col_ok)
      for (long t = row_start + ty; t < row_stop; t += 8) {
        const double e = fir_value(stim, nt, ld_stim, t, d, c, circpad) - mean;
        ssq += e * e;
      }
    red[ty][tx] = ssq;
    __syncthreads();
    ssq = 0.0;
#pragma unroll
    for (int y = 0; y < 8; ++y) ssq += red[y][tx];
    const double sd = sqrt(ssq / (double)n);
    inv_std = (sd != 0.0) ? 1.0 / sd : 1.0;  // NaN std: `s != 0` holds in the reference too -> NaN column -> 0
  }
  if (!col_ok) return;
  for (long t = row_start + ty; t < row_stop; t += 8) {
    double v = fir_value(stim, nt, ld_stim, t, d, c, circpad);
    if (zscore) {
      v = (v - mean) * inv_std;
      if (isnan(v)) v = 0.0;  // np.nan_to_num; +-inf become +-DBL_MAX there, i.e. +-inf again once cast to fp32
    }
    out[(t - row_start) * ld_out + j] = (float)v;
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
  if (dtype_in == 0 && ndim % 4 == 0 && ld_stim % 4 == 0 && ld_out % 2 == 0 &&
      (reinterpret_cast<uintptr_t>(stim) & 15) == 0) {
    const long items = nt * ndelays * (ndim / 4);
    long g4 = (items + 255) / 256;
    if (g4 > cap) g4 = cap;
    fir_f32x4_kernel<<<(int)g4, 256, 0, s>>>((const float*)stim, nt, ndim, ld_stim, delays, ndelays, circpad, out,
                                             ld_out);
    LIT_LAUNCH_CHECK();
    return LIT_OK;
  }
  if (dtype_in == 0)
    fir_kernel<float><<<(int)grid, 256, 0, s>>>((const float*)stim, nt, ndim, ld_stim, delays, ndelays, circpad, out,
                                                ld_out);
  else
    fir_kernel<double><<<(int)grid, 256, 0, s>>>((const double*)stim, nt, ndim, ld_stim, delays, ndelays, circpad, out,
                                                 ld_out);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}

template <int KIND>
static int launch_resample(const void* data, int dtype_in, long n_samples, long ndim, long ld_data,
                           const double* data_times, const double* tr_times, long n_tr, double window, double cutoff,
                           int flag_a, int flag_b, const int32_t* lo, const int32_t* hi, double* out, long ld_out,
                           void* stream) {
  const int block = ndim >= 256 ? 256 : (ndim >= 128 ? 128 : 64);
  constexpr int TRS = 4;
  dim3 grid((unsigned)((n_tr + TRS - 1) / TRS), (unsigned)((ndim + block - 1) / block));
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype_in == 0)
    resample_kernel<float, 128, KIND, TRS><<<grid, block, 0, s>>>((const float*)data, n_samples, ndim, ld_data,
                                                                   data_times, tr_times, n_tr, window, cutoff, flag_a,
                                                                   flag_b, lo, hi, out, ld_out);
  else
    resample_kernel<double, 128, KIND, TRS><<<grid, block, 0, s>>>((const double*)data, n_samples, ndim, ld_data,
                                                                    data_times, tr_times, n_tr, window, cutoff, flag_a,
                                                                    flag_b, lo, hi, out, ld_out);
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
  return launch_resample<0>(data, dtype_in, n_samples, ndim, ld_data, data_times, tr_times, n_tr, window, cutoff,
                            rectify, 0, lo, hi, out, ld_out, stream);
}

extern "C" int lit_sinc_downsample(const void* data, int dtype_in, long n_samples, long ndim, long ld_data,
                                   const double* data_times, const double* tr_times, long n_tr, double window,
                                   double cutoff, int causal, int renorm, const int32_t* lo, const int32_t* hi,
                                   double* out, long ld_out, void* stream) {
  LIT_REQUIRE(n_samples >= 0 && ndim >= 0 && n_tr >= 0, "sinc: negative extent");
  LIT_REQUIRE(ld_data >= ndim && ld_out >= ndim, "sinc: pitch too small");
  LIT_REQUIRE(dtype_in == 0 || dtype_in == 1, "sinc: dtype_in must be 0 (f32) or 1 (f64)");
  LIT_REQUIRE((lo == nullptr) == (hi == nullptr), "sinc: lo and hi must be given together");
  LIT_REQUIRE(n_tr <= 2147483647L, "sinc: too many TRs");
  if (n_tr == 0 || ndim == 0) return LIT_OK;
  return launch_resample<1>(data, dtype_in, n_samples, ndim, ld_data, data_times, tr_times, n_tr, window, cutoff, causal,
                            renorm, lo, hi, out, ld_out, stream);
}

extern "C" int lit_csr_rows_apply(const void* data, int dtype_in, long ndim, long ld_data, const int32_t* row_ptr,
                                  const int32_t* col_idx, const double* weights, long n_rows_out, int mean, double* out,
                                  long ld_out, void* stream) {
  LIT_REQUIRE(ndim >= 0 && n_rows_out >= 0 && ld_data >= ndim && ld_out >= ndim, "csr_rows: bad extents");
  LIT_REQUIRE(dtype_in == 0 || dtype_in == 1, "csr_rows: dtype_in must be 0 (f32) or 1 (f64)");
  LIT_REQUIRE(n_rows_out <= 2147483647L, "csr_rows: too many rows");
  if (n_rows_out == 0 || ndim == 0) return LIT_OK;
  const int block = ndim >= 256 ? 256 : (ndim >= 128 ? 128 : 64);
  dim3 grid((unsigned)n_rows_out, (unsigned)((ndim + block - 1) / block));
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype_in == 0)
    csr_rows_kernel<float><<<grid, block, 0, s>>>((const float*)data, ndim, ld_data, row_ptr, col_idx, weights, mean,
                                                  out, ld_out);
  else
    csr_rows_kernel<double><<<grid, block, 0, s>>>((const double*)data, ndim, ld_data, row_ptr, col_idx, weights, mean,
                                                   out, ld_out);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}

extern "C" int lit_gabor_downsample(const void* data, int dtype_in, long n_samples, long ndim, long ld_data,
                                    const double* data_times, const double* tr_times, long n_tr, const double* freqs,
                                    int n_freq, double sigma, double* out, long ld_out, void* stream) {
  LIT_REQUIRE(n_samples >= 0 && ndim >= 0 && n_tr >= 0 && n_freq >= 0, "gabor: negative extent");
  LIT_REQUIRE(ld_data >= ndim && ld_out >= ndim * n_freq, "gabor: pitch too small");
  LIT_REQUIRE(dtype_in == 0 || dtype_in == 1, "gabor: dtype_in must be 0 (f32) or 1 (f64)");
  LIT_REQUIRE(n_tr <= 2147483647L, "gabor: too many TRs");
  const long pairs = ndim * n_freq;
  if (n_tr == 0 || pairs == 0) return LIT_OK;
  const int block = pairs >= 256 ? 256 : (pairs >= 128 ? 128 : 64);
  dim3 grid((unsigned)n_tr, (unsigned)((pairs + block - 1) / block));
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype_in == 0)
    gabor_kernel<float, 256><<<grid, block, 0, s>>>((const float*)data, n_samples, ndim, ld_data, data_times, tr_times,
                                                    freqs, n_freq, sigma, out, ld_out);
  else
    gabor_kernel<double, 256><<<grid, block, 0, s>>>((const double*)data, n_samples, ndim, ld_data, data_times,
                                                     tr_times, freqs, n_freq, sigma, out, ld_out);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}

extern "C" int lit_fir_zscore_rows(const void* stim, int dtype_in, long nt, long ndim, long ld_stim,
                                   const int32_t* delays, int ndelays, int circpad, long row_start, long row_stop,
                                   int zscore, float* out, long ld_out, void* stream) {
  LIT_REQUIRE(nt >= 0 && ndim >= 0 && ndelays >= 0, "fir_zscore: negative extent");
  LIT_REQUIRE(ld_stim >= ndim && ld_out >= ndim * ndelays, "fir_zscore: pitch too small");
  LIT_REQUIRE(dtype_in == 0 || dtype_in == 1, "fir_zscore: dtype_in must be 0 (f32) or 1 (f64)");
  LIT_REQUIRE(row_start >= 0 && row_stop <= nt && row_start <= row_stop, "fir_zscore: trimmed rows [%ld, %ld) outside [0, %ld)",
              row_start, row_stop, nt);
  const long ncols = ndim * ndelays;
  if (ncols == 0 || row_stop == row_start) return LIT_OK;
  const unsigned grid = (unsigned)((ncols + 31) / 32);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype_in == 0)
    fir_zscore_kernel<float><<<grid, 256, 0, s>>>((const float*)stim, nt, ndim, ld_stim, delays, ndelays, circpad,
                                                  row_start, row_stop, zscore, out, ld_out);
  else
    fir_zscore_kernel<double><<<grid, 256, 0, s>>>((const double*)stim, nt, ndim, ld_stim, delays, ndelays, circpad,
                                                   row_start, row_stop, zscore, out, ld_out);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}
