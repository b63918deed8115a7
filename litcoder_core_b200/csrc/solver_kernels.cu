// Kernels of the eigendecomposition-free inner-fold solver.
//
// For an inner fold with training Gram G (p x p) and centred validation design P_c (m x p), the
// alpha-stacked prediction operator of ridge_corr_torch (ridge_regression.py:115-120)
//     pred_a = P_c (G + a^2 I)^-1 C,     a = alpha * S[0]   (S[0]^2 = lambda_max(G) when normalpha)
// only needs the m x p matrices M_a = P_c (G + a^2 I)^-1.  Instead of an eigendecomposition of G they are
// obtained with tensor-core GEMMs:
//   * lambda_max(G) by a short Lanczos recurrence (sym_gemv + lanczos_step + a Sturm bisection on the
//     tridiagonal matrix),
//   * small alphas by Chebyshev iteration on M (G + a^2 I) = P_c, spectrum in [a^2, lambda_max + a^2]:
//     one GEMM (lit_gemm_tf32x3_nt with beta/Cin) and one cheb_update kernel per step,
//   * large alphas (a^2 >> lambda_max) by the truncated Neumann series in the shared powers P_c G^q.
// The per-step scalars are computed on the host from lambda_max (control flow only).
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "../../include/litridge.h"

namespace lit {

__host__ __device__ static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Batched Lanczos: blockIdx.y is the matrix index b.  Matrix b is Gs.p[b] (all n x n, pitch ld); its work vectors
// live at vec + b * 3n (three n-vectors whose roles rotate: offsets o_v / o_prev / o_y in units of n) and its
// scalars at scal + b * scal_stride: [dot, -, alpha[steps], beta[steps + 1]].
constexpr int LANCZOS_MAX_BATCH = 32;
struct MatPtrs {
  const float* p[LANCZOS_MAX_BATCH];
};

// y = G v (G symmetric, row-major, n x n), and dot += v . y.  One warp per row.
__global__ void sym_gemv_dot_kernel(MatPtrs Gs, long ld, int n, float* __restrict__ vec, int o_v, int o_y,
                                    double* __restrict__ scal, long scal_stride) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n) return;
  const float* __restrict__ G = Gs.p[blockIdx.y];
  const float* __restrict__ v = vec + ((long)blockIdx.y * 3 + o_v) * n;
  float* __restrict__ y = vec + ((long)blockIdx.y * 3 + o_y) * n;
  double* dot = scal + (long)blockIdx.y * scal_stride;
  const float* row = G + (long)warp * ld;
  double acc = 0.0;
  const bool vec4 = (ld % 4 == 0) && aligned16(G) && aligned16(v);
  if (vec4) {
    const int n4 = n >> 2;
    for (int j = lane; j < n4; j += 32) {
      const float4 g = *reinterpret_cast<const float4*>(row + 4 * j);
      const float4 x = *reinterpret_cast<const float4*>(v + 4 * j);
      acc += (double)g.x * x.x + (double)g.y * x.y + (double)g.z * x.z + (double)g.w * x.w;
    }
    for (int j = (n4 << 2) + lane; j < n; j += 32) acc += (double)row[j] * v[j];
  } else {
    for (int j = lane; j < n; j += 32) acc += (double)row[j] * v[j];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    y[warp] = (float)acc;
    atomicAdd(dot, acc * (double)v[warp]);
  }
}

__device__ __forceinline__ double block_sum(double x, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = x;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
  return t;
}

// Deterministic start vector (unit norm).  One block per matrix.
__global__ void lanczos_init_kernel(float* __restrict__ vec, int n, double* __restrict__ scal, long scal_stride,
                                    int steps) {
  __shared__ double sh[32];
  float* __restrict__ v = vec + (long)blockIdx.x * 3 * n;
  float* __restrict__ v_prev = v + n;
  double* dot = scal + (long)blockIdx.x * scal_stride;
  double* alpha = dot + 2;
  double* beta = alpha + steps;
  double part = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + 12345u;
    h ^= h >> 15;
    h *= 2246822519u;
    h ^= h >> 13;
    const float x = (float)(h & 0xffffff) / 8388608.f - 1.f + 1e-3f;
    v[i] = x;
    v_prev[i] = 0.f;
    part += (double)x * x;
  }
  const double nrm = sqrt(block_sum(part, sh));
  for (int i = threadIdx.x; i < n; i += blockDim.x) v[i] = (float)(v[i] / nrm);
  for (int i = threadIdx.x; i <= steps; i += blockDim.x) {
    if (i < steps) alpha[i] = 0.0;
    beta[i] = 0.0;
  }
  if (threadIdx.x == 0) *dot = 0.0;
}

// One three-term Lanczos step after y = G v_j and dot = v_j . y have been formed:
//   alpha_j = dot;  w = y - alpha_j v_j - beta_j v_{j-1};  beta_{j+1} = ||w||;  v_{j+1} = w / beta_{j+1}
// v_next overwrites v_prev (it is read before it is written by the same thread).  One block per matrix.
__global__ void lanczos_step_kernel(float* __restrict__ vec, int o_y, int o_v, int o_prev, int n, int j, int steps,
                                    double* __restrict__ scal, long scal_stride) {
  __shared__ double sh[32];
  const float* __restrict__ y = vec + ((long)blockIdx.x * 3 + o_y) * n;
  const float* __restrict__ v = vec + ((long)blockIdx.x * 3 + o_v) * n;
  float* v_prev_next = vec + ((long)blockIdx.x * 3 + o_prev) * n;
  double* dot = scal + (long)blockIdx.x * scal_stride;
  double* alpha = dot + 2;
  double* beta = alpha + steps;
  const double a = *dot;
  const double b = beta[j];
  double part = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double w = (double)y[i] - a * (double)v[i] - b * (double)v_prev_next[i];
    part += w * w;
  }
  const double nrm = sqrt(block_sum(part, sh));
  const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double w = (double)y[i] - a * (double)v[i] - b * (double)v_prev_next[i];
    v_prev_next[i] = (float)(w * inv);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    alpha[j] = a;
    beta[j + 1] = nrm;
    *dot = 0.0;
  }
}

// Largest eigenvalue of the symmetric tridiagonal (alpha[0..m), beta[1..m)) by Sturm multisection (fp64):
// one warp evaluates 32 trial points per round (5 bits per round instead of 1).
__global__ void tridiag_lmax_kernel(const double* __restrict__ scal, long scal_stride, int m,
                                    float* __restrict__ out_f32, double* __restrict__ out_f64) {
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  const double* __restrict__ alpha = scal + (long)blockIdx.x * scal_stride + 2;
  const double* __restrict__ beta = alpha + m;
  if (out_f32) out_f32 += blockIdx.x;
  if (out_f64) out_f64 += blockIdx.x;
  // an (almost) invariant subspace ends the recurrence early: keep the leading block
  double scale = 0.0;
  for (int i = 0; i < m; ++i) scale = fmax(scale, fabs(alpha[i]));
  int mm = m;
  for (int i = 1; i < m; ++i)
    if (beta[i] <= 1e-12 * scale) {
      mm = i;
      break;
    }
  double lo = alpha[0], hi = alpha[0];
  for (int i = 0; i < mm; ++i) {
    const double r = (i > 0 ? fabs(beta[i]) : 0.0) + (i + 1 < mm ? fabs(beta[i + 1]) : 0.0);
    lo = fmin(lo, alpha[i] - r);
    hi = fmax(hi, alpha[i] + r);
  }
  const double tiny = 1e-300 + 1e-30 * fmax(fabs(lo), fabs(hi));
  for (int it = 0; it < 14; ++it) {  // 33^14 > 2^70 subdivisions of the Gershgorin interval
    const double x = lo + (hi - lo) * (double)(lane + 1) / 33.0;
    int below = 0;  // number of eigenvalues smaller than x
    double q = alpha[0] - x;
    if (fabs(q) < tiny) q = -tiny;
    if (q < 0.0) ++below;
    for (int i = 1; i < mm; ++i) {
      q = alpha[i] - x - beta[i] * beta[i] / q;
      if (fabs(q) < tiny) q = -tiny;
      if (q < 0.0) ++below;
    }
    // lambda_max lies between the last point with an eigenvalue at or above it and the first point above all
    const unsigned above_all = __ballot_sync(0xffffffffu, below >= mm);
    const int first = above_all ? __ffs(above_all) - 1 : 32;
    const double step = (hi - lo) / 33.0;
    const double new_lo = lo + step * (double)first;       // point index first-1 (or lo itself)
    const double new_hi = first < 32 ? lo + step * (double)(first + 1) : hi;
    lo = new_lo;
    hi = new_hi;
  }
  const double lam = 0.5 * (lo + hi);
  if (lane == 0) {
    if (out_f32) *out_f32 = (float)lam;
    if (out_f64) *out_f64 = lam;
  }
}

// Chebyshev step on row-matrices (rows x cols):
//   d = c1 * d + c2 * r        (search direction; also written as a 3xTF32 split pair for the next GEMM)
//   x = x + d
//   t = r - a2 * d             (the GEMM then forms  r = t - d G)
template <bool VEC>
__global__ void cheb_update_kernel(float* __restrict__ d, const float* __restrict__ r, float* __restrict__ x,
                                   float* __restrict__ t, float* __restrict__ d_hi, float* __restrict__ d_lo, long ld,
                                   long rows, long cols, float c1, float c2, float a2, int first) {
  constexpr int CPT = VEC ? 4 : 1;
  const long ct = (cols + CPT - 1) / CPT;
  const long c = ((long)blockIdx.x * blockDim.x + threadIdx.x) * CPT;
  if (c >= cols) return;
  (void)ct;
  for (long row = blockIdx.y; row < rows; row += gridDim.y) {
    const long o = row * ld + c;
    float dv[CPT], rv[CPT], xv[CPT];
    if (VEC) {
      const float4 r4 = *reinterpret_cast<const float4*>(r + o);
      rv[0] = r4.x, rv[1 % CPT] = r4.y, rv[2 % CPT] = r4.z, rv[3 % CPT] = r4.w;
      if (first) {
#pragma unroll
        for (int u = 0; u < CPT; ++u) dv[u] = 0.f, xv[u] = 0.f;
      } else {
        const float4 d4 = *reinterpret_cast<const float4*>(d + o);
        const float4 x4 = *reinterpret_cast<const float4*>(x + o);
        dv[0] = d4.x, dv[1 % CPT] = d4.y, dv[2 % CPT] = d4.z, dv[3 % CPT] = d4.w;
        xv[0] = x4.x, xv[1 % CPT] = x4.y, xv[2 % CPT] = x4.z, xv[3 % CPT] = x4.w;
      }
    } else {
      rv[0] = r[o];
      dv[0] = first ? 0.f : d[o];
      xv[0] = first ? 0.f : x[o];
    }
    float h[CPT], l[CPT], tv[CPT];
#pragma unroll
    for (int u = 0; u < CPT; ++u) {
      dv[u] = fmaf(c1, dv[u], c2 * rv[u]);
      xv[u] += dv[u];
      tv[u] = fmaf(-a2, dv[u], rv[u]);
      h[u] = ptx::to_tf32(dv[u]);
      l[u] = ptx::to_tf32(dv[u] - h[u]);
    }
    if (VEC) {
      *reinterpret_cast<float4*>(d + o) = make_float4(dv[0], dv[1 % CPT], dv[2 % CPT], dv[3 % CPT]);
      *reinterpret_cast<float4*>(x + o) = make_float4(xv[0], xv[1 % CPT], xv[2 % CPT], xv[3 % CPT]);
      *reinterpret_cast<float4*>(t + o) = make_float4(tv[0], tv[1 % CPT], tv[2 % CPT], tv[3 % CPT]);
      *reinterpret_cast<float4*>(d_hi + o) = make_float4(h[0], h[1 % CPT], h[2 % CPT], h[3 % CPT]);
      *reinterpret_cast<float4*>(d_lo + o) = make_float4(l[0], l[1 % CPT], l[2 % CPT], l[3 % CPT]);
    } else {
      d[o] = dv[0];
      x[o] = xv[0];
      t[o] = tv[0];
      d_hi[o] = h[0];
      d_lo[o] = l[0];
    }
  }
}

// out[(slot_g * rows_pad + t)][c] = sum_q coef[g][q] * (Q_hi[q] + Q_lo[q])[t][c]  for t < rows, 0 for pad rows,
// written as a split pair.  Up to 4 source matrices (Q_0 = P_c, Q_q = P_c G^q), n_groups target alphas.
struct PolySources {
  const float* hi[4];
  const float* lo[4];  // may be null
};
__global__ void poly_combine_kernel(PolySources src, int n_src, long ld_src, long rows, long rows_pad, long cols,
                                    const double* __restrict__ coef /* [n_groups][4] */,
                                    const int32_t* __restrict__ slots, int n_groups, float* __restrict__ out_hi,
                                    float* __restrict__ out_lo, long ld_out) {
  const long c = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= cols) return;
  for (long t = blockIdx.y; t < rows_pad; t += gridDim.y) {
    float q[4][4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (s < n_src && t < rows) {
        v = *reinterpret_cast<const float4*>(src.hi[s] + t * ld_src + c);
        if (src.lo[s]) {
          const float4 w = *reinterpret_cast<const float4*>(src.lo[s] + t * ld_src + c);
          v.x += w.x, v.y += w.y, v.z += w.z, v.w += w.w;
        }
      }
      q[s][0] = v.x, q[s][1] = v.y, q[s][2] = v.z, q[s][3] = v.w;
    }
    for (int g = 0; g < n_groups; ++g) {
      float h[4], l[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        double acc = 0.0;
#pragma unroll
        for (int s = 0; s < 4; ++s)
          if (s < n_src) acc += coef[g * 4 + s] * (double)q[s][u];
        const float o = (float)acc;
        h[u] = ptx::to_tf32(o);
        l[u] = ptx::to_tf32(o - h[u]);
      }
      const long orow = (long)slots[g] * rows_pad + t;
      *reinterpret_cast<float4*>(out_hi + orow * ld_out + c) = make_float4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<float4*>(out_lo + orow * ld_out + c) = make_float4(l[0], l[1], l[2], l[3]);
    }
  }
}

// Series tiles of the alpha stack (lit_gemm_corr_series): the four terms Q_q = scale[q] * (src_hi[q] + src_lo[q])
// (Q_0 = P_c, Q_q = P_c G^q / lambda_max^q) of the Neumann series, interleaved so that one 256-row tile holds
// 64 time points x 4 terms:  out row = tile*256 + half*128 + q*32 + i  <->  time point t = tile*64 + half*32 + i.
// Rows with t >= rows are zero.  Written as a 3xTF32 split pair.
__global__ void series_stack_kernel(PolySources src, long ld_src, long rows, long cols, double s0, double s1, double s2,
                                    double s3, long n_tiles, float* __restrict__ out_hi, float* __restrict__ out_lo,
                                    long ld_out) {
  const long c = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= cols) return;
  const double sc[4] = {s0, s1, s2, s3};
  for (long orow = blockIdx.y; orow < n_tiles * 256; orow += gridDim.y) {
    const long tile = orow >> 8;
    const int r = (int)(orow & 255);
    const int q = (r & 127) >> 5;
    const long t = tile * 64 + (r >> 7) * 32 + (r & 31);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < rows) {
      v = *reinterpret_cast<const float4*>(src.hi[q] + t * ld_src + c);
      if (src.lo[q]) {
        const float4 w = *reinterpret_cast<const float4*>(src.lo[q] + t * ld_src + c);
        v.x += w.x, v.y += w.y, v.z += w.z, v.w += w.w;
      }
      const double f = sc[q];
      v.x = (float)(f * (double)v.x), v.y = (float)(f * (double)v.y);
      v.z = (float)(f * (double)v.z), v.w = (float)(f * (double)v.w);
    }
    float4 h, l;
    h.x = ptx::to_tf32(v.x), h.y = ptx::to_tf32(v.y), h.z = ptx::to_tf32(v.z), h.w = ptx::to_tf32(v.w);
    l.x = ptx::to_tf32(v.x - h.x), l.y = ptx::to_tf32(v.y - h.y);
    l.z = ptx::to_tf32(v.z - h.z), l.w = ptx::to_tf32(v.w - h.w);
    *reinterpret_cast<float4*>(out_hi + orow * ld_out + c) = h;
    *reinterpret_cast<float4*>(out_lo + orow * ld_out + c) = l;
  }
}

}  // namespace lit

using namespace lit;

static int lanczos_batch(const float* const* G, int batch, long ld, int n, int steps, float* vec_scratch,
                         double* scal_scratch, float* lam_out_f32, double* lam_out_f64, cudaStream_t s) {
  const long scal_stride = 2L * steps + 4;
  for (int b0 = 0; b0 < batch; b0 += LANCZOS_MAX_BATCH) {
    const int nb = batch - b0 < LANCZOS_MAX_BATCH ? batch - b0 : LANCZOS_MAX_BATCH;
    MatPtrs ptrs = {};
    for (int b = 0; b < nb; ++b) {
      LIT_REQUIRE(G[b0 + b], "lanczos_lambda_max: null matrix");
      ptrs.p[b] = G[b0 + b];
    }
    float* vec = vec_scratch + (long)b0 * 3 * n;
    double* scal = scal_scratch + (long)b0 * scal_stride;
    lanczos_init_kernel<<<nb, 1024, 0, s>>>(vec, n, scal, scal_stride, steps);
    LIT_LAUNCH_CHECK();
    const int rows_per_block = 8;
    const int gblocks = (n + rows_per_block - 1) / rows_per_block;
    int o_v = 0, o_prev = 1;  // roles of the three work vectors (offset 2 holds y)
    for (int j = 0; j < steps; ++j) {
      sym_gemv_dot_kernel<<<dim3(gblocks, nb), rows_per_block * 32, 0, s>>>(ptrs, ld, n, vec, o_v, 2, scal, scal_stride);
      lanczos_step_kernel<<<nb, 1024, 0, s>>>(vec, 2, o_v, o_prev, n, j, steps, scal, scal_stride);
      const int tmp = o_v;  // v_{j+1} was written over v_{j-1}
      o_v = o_prev;
      o_prev = tmp;
    }
    LIT_LAUNCH_CHECK();
    tridiag_lmax_kernel<<<nb, 32, 0, s>>>(scal, scal_stride, steps, lam_out_f32 ? lam_out_f32 + b0 : nullptr,
                                          lam_out_f64 ? lam_out_f64 + b0 : nullptr);
    LIT_LAUNCH_CHECK();
  }
  return LIT_OK;
}

extern "C" int lit_lanczos_lambda_max(const float* G, long ld, int n, int steps, float* vec_scratch /* 3*n floats */,
                                      double* scal_scratch /* 2*steps + 4 doubles */, float* lam_out_f32,
                                      double* lam_out_f64, void* stream) {
  LIT_REQUIRE(n > 0 && ld >= n && steps > 0, "lanczos_lambda_max: bad extents");
  LIT_REQUIRE(vec_scratch && scal_scratch, "lanczos_lambda_max: scratch required");
  if (steps > n) steps = n;
  return lanczos_batch(&G, 1, ld, n, steps, vec_scratch, scal_scratch, lam_out_f32, lam_out_f64, (cudaStream_t)stream);
}

extern "C" int lit_lanczos_lambda_max_batched(const float* const* G /* host array of `batch` device pointers */,
                                              int batch, long ld, int n, int steps,
                                              float* vec_scratch /* batch * 3*n floats */,
                                              double* scal_scratch /* batch * (2*steps + 4) doubles */,
                                              double* lam_out_f64 /* batch */, void* stream) {
  LIT_REQUIRE(batch >= 0 && n > 0 && ld >= n && steps > 0, "lanczos_lambda_max_batched: bad extents");
  LIT_REQUIRE(steps <= n, "lanczos_lambda_max_batched: steps must not exceed n (the scratch layout depends on it)");
  LIT_REQUIRE(batch == 0 || (G && vec_scratch && scal_scratch && lam_out_f64), "lanczos_lambda_max_batched: null argument");
  if (batch == 0) return LIT_OK;
  return lanczos_batch(G, batch, ld, n, steps, vec_scratch, scal_scratch, nullptr, lam_out_f64, (cudaStream_t)stream);
}

extern "C" int lit_cheb_update(float* d, const float* r, float* x, float* t, float* d_hi, float* d_lo, long ld, long rows,
                               long cols, float c1, float c2, float a2, int first, void* stream) {
  LIT_REQUIRE(rows >= 0 && cols >= 0 && ld >= cols, "cheb_update: bad extents");
  if (rows == 0 || cols == 0) return LIT_OK;
  const bool vec = cols % 4 == 0 && ld % 4 == 0 && aligned16(d) && aligned16(r) && aligned16(x) && aligned16(t) &&
                   aligned16(d_hi) && aligned16(d_lo);
  const int block = 128;
  const long ct = vec ? cols / 4 : cols;
  const int gx = (int)((ct + block - 1) / block);
  long gy = rows;
  const long want = ((long)sm_count() * 16 + gx - 1) / gx;
  if (gy > want) gy = want;
  if (gy > 65535) gy = 65535;
  dim3 grid(gx, (unsigned)gy);
  cudaStream_t s = (cudaStream_t)stream;
  if (vec)
    cheb_update_kernel<true><<<grid, block, 0, s>>>(d, r, x, t, d_hi, d_lo, ld, rows, cols, c1, c2, a2, first);
  else
    cheb_update_kernel<false><<<grid, block, 0, s>>>(d, r, x, t, d_hi, d_lo, ld, rows, cols, c1, c2, a2, first);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}

extern "C" int lit_poly_combine(const float* const* src_hi, const float* const* src_lo, int n_src, long ld_src, long rows,
                                long rows_pad, long cols, const double* coef, const int32_t* slots, int n_groups,
                                float* out_hi, float* out_lo, long ld_out, void* stream) {
  LIT_REQUIRE(n_src >= 1 && n_src <= 4 && n_groups >= 0, "poly_combine: 1..4 sources");
  // columns are processed four at a time; a ragged width is rounded up into the (never read) pitch padding
  cols = (cols + 3) / 4 * 4;
  LIT_REQUIRE(rows >= 0 && rows_pad >= rows && ld_src % 4 == 0 && ld_out % 4 == 0 && cols <= ld_src && cols <= ld_out,
              "poly_combine: pitches must be multiples of 4 floats and cover the width rounded up to 4");
  if (n_groups == 0 || rows_pad == 0 || cols == 0) return LIT_OK;
  PolySources ps;
  for (int i = 0; i < 4; ++i) {
    ps.hi[i] = i < n_src ? src_hi[i] : nullptr;
    ps.lo[i] = (i < n_src && src_lo) ? src_lo[i] : nullptr;
    LIT_REQUIRE(i >= n_src || (ps.hi[i] && aligned16(ps.hi[i]) && (!ps.lo[i] || aligned16(ps.lo[i]))),
                "poly_combine: source alignment");
  }
  LIT_REQUIRE(aligned16(out_hi) && aligned16(out_lo), "poly_combine: output alignment");
  const int block = 128;
  const int gx = (int)((cols / 4 + block - 1) / block);
  long gy = rows_pad;
  const long want = ((long)sm_count() * 16 + gx - 1) / gx;
  if (gy > want) gy = want;
  if (gy > 65535) gy = 65535;
  poly_combine_kernel<<<dim3(gx, (unsigned)gy), block, 0, (cudaStream_t)stream>>>(ps, n_src, ld_src, rows, rows_pad, cols,
                                                                                   coef, slots, n_groups, out_hi, out_lo,
                                                                                   ld_out);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}

extern "C" int lit_series_stack(const float* const* src_hi, const float* const* src_lo, long ld_src, long rows, long cols,
                                const double* scale /* 4 host doubles */, long n_tiles, float* out_hi, float* out_lo,
                                long ld_out, void* stream) {
  cols = (cols + 3) / 4 * 4;
  LIT_REQUIRE(rows >= 0 && n_tiles >= 0 && rows <= n_tiles * 64, "series_stack: %ld rows do not fit %ld tiles", rows,
              n_tiles);
  LIT_REQUIRE(ld_src % 4 == 0 && ld_out % 4 == 0 && cols <= ld_src && cols <= ld_out,
              "series_stack: pitches must be multiples of 4 floats and cover the width rounded up to 4");
  if (n_tiles == 0 || cols == 0) return LIT_OK;
  PolySources ps;
  for (int i = 0; i < 4; ++i) {
    ps.hi[i] = src_hi[i];
    ps.lo[i] = src_lo ? src_lo[i] : nullptr;
    LIT_REQUIRE(ps.hi[i] && aligned16(ps.hi[i]) && (!ps.lo[i] || aligned16(ps.lo[i])), "series_stack: source alignment");
  }
  LIT_REQUIRE(aligned16(out_hi) && aligned16(out_lo), "series_stack: output alignment");
  const int block = 128;
  const int gx = (int)((cols / 4 + block - 1) / block);
  long gy = n_tiles * 256;
  const long want = ((long)sm_count() * 16 + gx - 1) / gx;
  if (gy > want) gy = want;
  if (gy > 65535) gy = 65535;
  series_stack_kernel<<<dim3(gx, (unsigned)gy), block, 0, (cudaStream_t)stream>>>(
      ps, ld_src, rows, cols, scale[0], scale[1], scale[2], scale[3], n_tiles, out_hi, out_lo, ld_out);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}
