// Ridge-specific streaming kernels: column statistics and z-scoring of the responses,
// construction of the alpha-stacked validation design in the eigenbasis, per-voxel
// shrinkage, inner-CV score finalisation and the per-voxel argmax over alphas.
// All are HBM-bound: threads map to the contiguous (voxel / feature) axis so every warp
// access is a full 128-byte line; reductions over rows are per-thread serial (no atomics).
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "../../include/litridge.h"

#include <cfloat>

namespace lit {

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline int blocks_for(long items, int block) { return (int)((items + block - 1) / block); }

// ---------------------------------------------------------------------------------------------
// Column mean / std over a gathered row set.  One thread per column, row loop split over
// blockIdx.y into ROW_SPLIT slices that are combined with fp64 atomics on (sum, sumsq) of
// values shifted by the first row (avoids cancellation when |mean| >> std).
// ---------------------------------------------------------------------------------------------
template <int CPT>  // columns per thread: 4 (float4 loads) or 1
__global__ void col_moments_kernel(const float* __restrict__ src, long ld_src, const int32_t* __restrict__ idx,
                                   long n_idx, long cols, double* __restrict__ acc /* [2][cols] */) {
  const long c = ((long)blockIdx.x * blockDim.x + threadIdx.x) * CPT;
  if (c >= cols) return;
  const long per = (n_idx + gridDim.y - 1) / gridDim.y;
  const long r_begin = (long)blockIdx.y * per;
  long r_end = r_begin + per;
  if (r_end > n_idx) r_end = n_idx;
  const long r_first = idx ? (long)idx[0] : 0;
  float shift[CPT];
  double s[CPT], q[CPT];
#pragma unroll
  for (int u = 0; u < CPT; ++u) {
    shift[u] = (c + u < cols) ? src[r_first * ld_src + c + u] : 0.f;
    s[u] = 0.0;
    q[u] = 0.0;
  }
  constexpr int UNROLL = 4;
  long r = r_begin;
  for (; r + UNROLL <= r_end; r += UNROLL) {
    float v[UNROLL][CPT];
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      const long sr = idx ? (long)idx[r + k] : r + k;
      if (CPT == 4) {
        const float4 x = *reinterpret_cast<const float4*>(src + sr * ld_src + c);
        v[k][0] = x.x;
        v[k][1 % CPT] = x.y;
        v[k][2 % CPT] = x.z;
        v[k][3 % CPT] = x.w;
      } else {
        v[k][0] = src[sr * ld_src + c];
      }
    }
#pragma unroll
    for (int k = 0; k < UNROLL; ++k)
#pragma unroll
      for (int u = 0; u < CPT; ++u) {
        const double d = (double)v[k][u] - (double)shift[u];
        s[u] += d;
        q[u] += d * d;
      }
  }
  for (; r < r_end; ++r) {
    const long sr = idx ? (long)idx[r] : r;
#pragma unroll
    for (int u = 0; u < CPT; ++u) {
      if (c + u < cols) {
        const double d = (double)src[sr * ld_src + c + u] - (double)shift[u];
        s[u] += d;
        q[u] += d * d;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < CPT; ++u) {
    if (c + u >= cols) break;
    if (gridDim.y == 1) {
      acc[c + u] = s[u];
      acc[cols + c + u] = q[u];
    } else {
      atomicAdd(&acc[c + u], s[u]);
      atomicAdd(&acc[cols + c + u], q[u]);
    }
  }
}

__global__ void col_stats_finish_kernel(const float* __restrict__ src, long ld_src, const int32_t* __restrict__ idx,
                                        long n_idx, long cols, int ddof, const double* __restrict__ acc,
                                        float* __restrict__ mean, float* __restrict__ stdv) {
  const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const long r_first = idx ? (long)idx[0] : 0;
  const double shift = (double)src[r_first * ld_src + c];
  const double s = acc[c], q = acc[cols + c];
  const double n = (double)n_idx;
  const double m = s / n;
  double var = (q - s * m) / (n - (double)ddof);
  if (var < 0.0) var = 0.0;
  if (mean) mean[c] = (float)(shift + m);
  if (stdv) stdv[c] = (float)sqrt(var);
}

// dst[i][c] = (src[idx[i]][c] - mean[c]) * scale(c)
// grid.x covers the columns (CPT per thread), grid.y groups of RPT output rows: the per-column mean / scale
// are computed once per thread and reused for its RPT rows, whose loads are all in flight together.
template <bool VEC, bool SPLIT>
__global__ void gather_normalize_kernel(const float* __restrict__ src, long ld_src, const int32_t* __restrict__ idx,
                                        long n_idx, long cols, const float* __restrict__ mean,
                                        const float* __restrict__ stdv, int mode, float eps, float* __restrict__ dst,
                                        float* __restrict__ dst_lo, long ld_dst, long n_rows_out) {
  constexpr int CPT = VEC ? 4 : 1;
  constexpr int RPT = 4;
  const long c = ((long)blockIdx.x * blockDim.x + threadIdx.x) * CPT;
  if (c >= cols) return;
  const float rs = n_idx > 1 ? rsqrtf((float)(n_idx - 1)) : 0.f;
  float m[CPT], sc[CPT];
#pragma unroll
  for (int u = 0; u < CPT; ++u) {
    m[u] = mean[c + u];
    sc[u] = 1.f;
    if (mode == 0) {
      sc[u] = 1.f / (stdv[c + u] + eps);
    } else if (mode == 1) {
      // a constant response column has no correlation: NaN here propagates through the fused
      // reduction and becomes (r, p) = (0, 1) in pearson_finalize, as SciPy's pearsonr -> NaN
      const float sd = stdv[c + u];
      sc[u] = sd > 0.f ? rs / sd : __int_as_float(0x7fc00000);
    } else if (mode == 3) {
      const float sd = stdv[c + u];
      sc[u] = sd != 0.f ? 1.f / sd : 1.f;  // zs(): zero-variance columns are centred only
    }
  }
  for (long r0 = (long)blockIdx.y * RPT; r0 < n_rows_out; r0 += (long)gridDim.y * RPT) {
    float x[RPT][CPT];
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const long r = r0 + k;
#pragma unroll
      for (int u = 0; u < CPT; ++u) x[k][u] = 0.f;
      if (r < n_idx) {
        const long sr = idx ? (long)idx[r] : r;
        if (VEC) {
          const float4 v = *reinterpret_cast<const float4*>(src + sr * ld_src + c);
          x[k][0] = v.x, x[k][1 % CPT] = v.y, x[k][2 % CPT] = v.z, x[k][3 % CPT] = v.w;
        } else {
          x[k][0] = src[sr * ld_src + c];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const long r = r0 + k;
      if (r >= n_rows_out) break;
      float h[CPT], l[CPT];
#pragma unroll
      for (int u = 0; u < CPT; ++u) {
        const float y = r < n_idx ? (x[k][u] - m[u]) * sc[u] : 0.f;
        if (SPLIT) {
          h[u] = ptx::to_tf32(y);
          l[u] = ptx::to_tf32(y - h[u]);
        } else {
          h[u] = y;
          l[u] = 0.f;
        }
      }
      if (VEC) {
        *reinterpret_cast<float4*>(dst + r * ld_dst + c) = make_float4(h[0], h[1 % CPT], h[2 % CPT], h[3 % CPT]);
        if (SPLIT)
          *reinterpret_cast<float4*>(dst_lo + r * ld_dst + c) = make_float4(l[0], l[1 % CPT], l[2 % CPT], l[3 % CPT]);
      } else {
        dst[r * ld_dst + c] = h[0];
        if (SPLIT) dst_lo[r * ld_dst + c] = l[0];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Alpha stack: out[a*rows_pad + t][j] = (L[t][j] - mean_j) * keep_j / (lam_j + (alpha_a*s)^2)
// grid: x over column tiles (128 columns), y over (alpha, row-chunk).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float norm_scale(const float* lam, int k, int normalpha) {
  // lam ascending (cuSOLVER syevd) -> largest eigenvalue is the last one.
  return normalpha ? sqrtf(fmaxf(lam[k - 1], 0.f)) : 1.f;
}

template <int CPT>  // columns per thread (4: float4 accesses)
__global__ void alpha_stack_kernel(const float* __restrict__ L, long ld_l, long n_rows, long rows_pad, int k,
                                   const float* __restrict__ lam, const double* __restrict__ alphas, int normalpha,
                                   float singcutoff, const float* __restrict__ col_mean, float* __restrict__ out_hi,
                                   float* __restrict__ out_lo, long ld_out, int row_chunk) {
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * CPT;
  const int n_chunks = (int)((rows_pad + row_chunk - 1) / row_chunk);
  const int a = blockIdx.y / n_chunks;
  const int ch = blockIdx.y - a * n_chunks;
  if (j >= k) return;
  const float s = norm_scale(lam, k, normalpha);
  const double an = alphas[a] * (double)s;  // alpha * S[0] in double, as the Python floats of the reference
  const float a2 = (float)(an * an);
  float d[CPT], m[CPT];
#pragma unroll
  for (int u = 0; u < CPT; ++u) {
    const float lj = lam[j + u];
    const bool keep = sqrtf(fmaxf(lj, 0.f)) > singcutoff;
    d[u] = keep ? 1.f / (lj + a2) : 0.f;
    m[u] = col_mean[j + u];
  }
  const long t0 = (long)ch * row_chunk;
  long t1 = t0 + row_chunk;
  if (t1 > rows_pad) t1 = rows_pad;
  for (long t = t0; t < t1; t += 2) {  // two rows in flight per thread
    float v[2][CPT];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const bool live = t + r < n_rows;
      if (CPT == 4) {
        const float4 x = live ? *reinterpret_cast<const float4*>(L + (t + r) * ld_l + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[r][0] = x.x, v[r][1 % CPT] = x.y, v[r][2 % CPT] = x.z, v[r][3 % CPT] = x.w;
      } else {
        v[r][0] = live ? L[(t + r) * ld_l + j] : 0.f;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (t + r >= t1) break;
      const bool live = t + r < n_rows;
      float h[CPT], l[CPT];
#pragma unroll
      for (int u = 0; u < CPT; ++u) {
        const float x = live ? (v[r][u] - m[u]) * d[u] : 0.f;
        h[u] = ptx::to_tf32(x);
        l[u] = ptx::to_tf32(x - h[u]);
      }
      const long o = ((long)a * rows_pad + t + r) * ld_out + j;
      if (CPT == 4) {
        *reinterpret_cast<float4*>(out_hi + o) = make_float4(h[0], h[1 % CPT], h[2 % CPT], h[3 % CPT]);
        *reinterpret_cast<float4*>(out_lo + o) = make_float4(l[0], l[1 % CPT], l[2 % CPT], l[3 % CPT]);
      } else {
        out_hi[o] = h[0];
        out_lo[o] = l[0];
      }
    }
  }
}

// out[v][j] = (Zhi+Zlo)[v][j] * keep_j / (lam_j + (alpha_v*s)^2)
template <bool VEC>
__global__ void scale_rows_kernel(const float* __restrict__ Z_hi, const float* __restrict__ Z_lo, long ld_z,
                                  long n_vox, int k, const float* __restrict__ lam, const float* __restrict__ alpha_v,
                                  int normalpha, float singcutoff, float* __restrict__ out_hi,
                                  float* __restrict__ out_lo, long ld_out) {
  constexpr int CPT = VEC ? 4 : 1;
  const float s = norm_scale(lam, k, normalpha);
  const long kt = (k + CPT - 1) / CPT;
  const long total = n_vox * kt;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long v = i / kt;
    const int j = (int)(i - v * kt) * CPT;
    const float an = alpha_v[v] * s;  // fp32 product, as nalphas = alphas * norm in ridge_torch
    const float a2 = an * an;
    float z[CPT], lj[CPT];
    if (VEC) {
      const float4 zh = *reinterpret_cast<const float4*>(Z_hi + v * ld_z + j);
      z[0] = zh.x, z[1 % CPT] = zh.y, z[2 % CPT] = zh.z, z[3 % CPT] = zh.w;
      if (Z_lo) {
        const float4 zl = *reinterpret_cast<const float4*>(Z_lo + v * ld_z + j);
        z[0] += zl.x, z[1 % CPT] += zl.y, z[2 % CPT] += zl.z, z[3 % CPT] += zl.w;
      }
      const float4 l4 = *reinterpret_cast<const float4*>(lam + j);
      lj[0] = l4.x, lj[1 % CPT] = l4.y, lj[2 % CPT] = l4.z, lj[3 % CPT] = l4.w;
    } else {
      z[0] = Z_hi[v * ld_z + j];
      if (Z_lo) z[0] += Z_lo[v * ld_z + j];
      lj[0] = lam[j];
    }
    float h[CPT], l[CPT];
#pragma unroll
    for (int u = 0; u < CPT; ++u) {
      const bool keep = sqrtf(fmaxf(lj[u], 0.f)) > singcutoff;
      const float o = keep ? z[u] / (lj[u] + a2) : 0.f;
      h[u] = ptx::to_tf32(o);
      l[u] = ptx::to_tf32(o - h[u]);
    }
    if (VEC) {
      *reinterpret_cast<float4*>(out_hi + v * ld_out + j) = make_float4(h[0], h[1 % CPT], h[2 % CPT], h[3 % CPT]);
      *reinterpret_cast<float4*>(out_lo + v * ld_out + j) = make_float4(l[0], l[1 % CPT], l[2 % CPT], l[3 % CPT]);
    } else {
      out_hi[v * ld_out + j] = h[0];
      out_lo[v * ld_out + j] = l[0];
    }
  }
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float nan_to_num(float x) {
  if (isnan(x)) return 0.f;
  if (isinf(x)) return x > 0.f ? FLT_MAX : -FLT_MAX;
  return x;
}

// metric 0: correlation (Yz z-scored).  metric 1: signed sqrt of R^2 (Yz centred only, resp_std = unbiased
// std of the validation responses): Rsq = 1 - var(Q - pred)/var(Q) with
// sum (q_c - p_c)^2 = (n-1) var(Q) - 2 dot + ssq   (ridge_regression.py:126-130).
__device__ __forceinline__ float inner_score(float d, float q, int metric, bool raw, float qvar, float inv_n,
                                             float inv_nm1, float nm1, float eps) {
  float c;
  if (metric == 0) {
    const float sd = sqrtf(q * inv_nm1);
    c = (d * inv_n) / (sd + eps);
  } else {
    const float resvar = (qvar * nm1 - 2.f * d + q) * inv_nm1;
    const float rsq = 1.f - resvar / qvar;
    c = sqrtf(fabsf(rsq)) * (rsq > 0.f ? 1.f : (rsq < 0.f ? -1.f : 0.f));
    if (isnan(rsq)) c = rsq;
  }
  return raw ? c : nan_to_num(c);
}

// inv_row[v] / inv_tile[N tile]: power-of-two operand scales of the fp16-split GEMM to undo (NULL = ones); a tile
// is 256 stacked rows = 2 parts.
__global__ void corr_finalize_kernel(const float* __restrict__ dot_part, const float* __restrict__ ssq_part,
                                     long ld_part, int tiles_per_group, int n_groups, long n_vox, long n_rows, float eps,
                                     int accumulate, int metric, const float* __restrict__ resp_std,
                                     const float* __restrict__ inv_row, const float* __restrict__ inv_tile,
                                     const int32_t* __restrict__ slots, float* __restrict__ corr, long ld_corr) {
  const long v = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vox) return;
  const float inv_n = 1.f / (float)n_rows;
  const float inv_nm1 = 1.f / (float)(n_rows - 1);
  const bool raw = (metric & 2) != 0;  // keep NaN / inf (ridge_corr_pred_torch does not scrub them)
  metric &= 1;
  float qvar = 0.f;
  if (metric == 1) {
    const float sd = resp_std[v];
    qvar = sd * sd;
  }
  const float ir = inv_row ? inv_row[v] : 1.f;
  for (int g = 0; g < n_groups; ++g) {
    float d = 0.f, q = 0.f;
    for (int t = 0; t < tiles_per_group; ++t) {
      const long part = (long)g * tiles_per_group + t;
      const long o = part * ld_part + v;
      if (inv_row || inv_tile) {  // fp16 split pairs: undo the power-of-two operand scales (exact)
        const float is = ir * (inv_tile ? inv_tile[part >> 1] : 1.f);
        d += dot_part[o] * is;
        q += ssq_part[o] * is * is;
      } else {
        d += dot_part[o];
        q += ssq_part[o];
      }
    }
    const float c = inner_score(d, q, metric, raw, qvar, inv_n, inv_nm1, (float)(n_rows - 1), eps);
    float* dst = corr + (long)(slots ? slots[g] : g) * ld_corr + v;
    *dst = accumulate ? *dst + c : c;
  }
}

// Scores of the alphas served by the Neumann series, from the 14 per-voxel sums of the series tiles
// (lit_gemm_corr_series): with T_q[t] = Q_q[t] . c_v the prediction for alpha a is  sum_q coef[a][q] T_q[t], so
//   dot_a = sum_q coef_q D_q,   ssq_a = sum_{q,q'} coef_q coef_q' S_qq'
// (D_q = sum_t T_q y, S_qq' = sum_t T_q T_q'), combined in fp64 and scored like every other alpha.
// series_part[part*14 + j]: j = 0..3 -> D_q; 4..13 -> S_00 S_01 S_02 S_03 S_11 S_12 S_13 S_22 S_23 S_33.
__global__ void corr_finalize_series_kernel(const float* __restrict__ series_part, long ld_part, int n_parts,
                                            long n_vox, long n_rows, float eps, int accumulate, int metric,
                                            const float* __restrict__ resp_std, const float* __restrict__ inv_row,
                                            const float* __restrict__ inv_tile /* of the series tiles */,
                                            const double* __restrict__ coef /* [n_alphas][4] */,
                                            const int32_t* __restrict__ slots, int n_alphas,
                                            float* __restrict__ corr, long ld_corr) {
  const long v = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vox) return;
  const float inv_n = 1.f / (float)n_rows;
  const float inv_nm1 = 1.f / (float)(n_rows - 1);
  const bool raw = (metric & 2) != 0;
  metric &= 1;
  float qvar = 0.f;
  if (metric == 1) {
    const float sd = resp_std[v];
    qvar = sd * sd;
  }
  const double ir = inv_row ? (double)inv_row[v] : 1.0;
  double sum[14];
#pragma unroll
  for (int j = 0; j < 14; ++j) sum[j] = 0.0;
  for (int part = 0; part < n_parts; ++part) {
    const double is = ir * (inv_tile ? (double)inv_tile[part >> 1] : 1.0);
    const float* src = series_part + (long)part * 14 * ld_part + v;
#pragma unroll
    for (int j = 0; j < 14; ++j) sum[j] += (double)src[(long)j * ld_part] * (j < 4 ? is : is * is);
  }
  for (int a = 0; a < n_alphas; ++a) {
    const double c0 = coef[a * 4], c1 = coef[a * 4 + 1], c2 = coef[a * 4 + 2], c3 = coef[a * 4 + 3];
    const double d = c0 * sum[0] + c1 * sum[1] + c2 * sum[2] + c3 * sum[3];
    const double q = c0 * c0 * sum[4] + c1 * c1 * sum[8] + c2 * c2 * sum[11] + c3 * c3 * sum[13] +
                     2.0 * (c0 * c1 * sum[5] + c0 * c2 * sum[6] + c0 * c3 * sum[7] + c1 * c2 * sum[9] +
                            c1 * c3 * sum[10] + c2 * c3 * sum[12]);
    const float c = inner_score((float)d, (float)q, metric, raw, qvar, inv_n, inv_nm1, (float)(n_rows - 1), eps);
    float* dst = corr + (long)slots[a] * ld_corr + v;
    *dst = accumulate ? *dst + c : c;
  }
}

__global__ void argmax_alpha_kernel(const float* __restrict__ corr_sum, long ld_corr, int n_alphas, long n_vox,
                                    int n_folds, const float* __restrict__ alphas, int32_t* __restrict__ best,
                                    float* __restrict__ alpha_out, double* __restrict__ col_sums) {
  extern __shared__ double sh[];  // [n_alphas] block partial sums
  for (int a = threadIdx.x; a < n_alphas; a += blockDim.x) sh[a] = 0.0;
  __syncthreads();
  const long v = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const float nf = (float)n_folds;
  int bi = 0;
  float bv = -INFINITY;
  for (int a = 0; a < n_alphas; ++a) {
    float m = 0.f;
    if (v < n_vox) {
      m = corr_sum[(long)a * ld_corr + v] / nf;
      if (m > bv) {  // strict: first maximum wins (torch.argmax)
        bv = m;
        bi = a;
      }
    }
    if (col_sums) {
      double x = (double)m;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) == 0) atomicAdd(&sh[a], x);
    }
  }
  if (v < n_vox) {
    if (best) best[v] = bi;
    if (alpha_out) alpha_out[v] = alphas[bi];
  }
  if (col_sums) {
    __syncthreads();
    for (int a = threadIdx.x; a < n_alphas; a += blockDim.x) atomicAdd(&col_sums[a], sh[a]);
  }
}

// Deterministic counting sort of the voxels by their selected alpha index (one block: V is ~1e5 and the keys < 32).
//   pos[v]         row of voxel v in the grouped layout: groups in index order, each padded to a multiple of `tile` rows
//   perm[s]        voxel at grouped row s, -1 for padding rows                  (s < rows_cap)
//   tile_group[t]  group of the t-th row tile, -1 past the last group           (t < rows_cap / tile)
__global__ void __launch_bounds__(1024) group_plan_kernel(const int32_t* __restrict__ idx, long n_vox, int n_groups,
                                                          int tile, long rows_cap, int32_t* __restrict__ pos,
                                                          int32_t* __restrict__ perm, int32_t* __restrict__ tile_group) {
  extern __shared__ int cnt[];  // [n_groups][1024] per-thread counts -> offsets inside the group
  __shared__ int warp_tot[32];
  __shared__ long start[33];
  const int t = threadIdx.x;
  const long chunk = (n_vox + 1023) / 1024;
  const long v0 = t * chunk, v1 = min(n_vox, v0 + chunk);
  for (int g = 0; g < n_groups; ++g) cnt[g * 1024 + t] = 0;
  for (long v = v0; v < v1; ++v) {
    const int g = min(max(idx[v], 0), n_groups - 1);
    cnt[g * 1024 + t]++;
  }
  for (long s = t; s < rows_cap; s += 1024) perm[s] = -1;
  __syncthreads();
  if (t == 0) start[0] = 0;
  for (int g = 0; g < n_groups; ++g) {  // exclusive scan of the group's counts over the threads
    const int c = cnt[g * 1024 + t];
    int x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if ((t & 31) >= o) x += y;
    }
    if ((t & 31) == 31) warp_tot[t >> 5] = x;
    __syncthreads();
    if (t < 32) {
      int w = warp_tot[t];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, o);
        if (t >= o) w += y;
      }
      warp_tot[t] = w;  // inclusive
    }
    __syncthreads();
    const int before = (t >> 5) ? warp_tot[(t >> 5) - 1] : 0;
    cnt[g * 1024 + t] = before + x - c;
    if (t == 0) start[g + 1] = start[g] + ((long)(warp_tot[31] + tile - 1) / tile) * tile;
    __syncthreads();
  }
  for (long v = v0; v < v1; ++v) {
    const int g = min(max(idx[v], 0), n_groups - 1);
    const long s = start[g] + cnt[g * 1024 + t]++;
    pos[v] = (int32_t)s;
    perm[s] = (int32_t)v;
  }
  for (long tl = t; tl < rows_cap / tile; tl += 1024) {
    const long r = tl * tile;
    int g = -1;
    for (int q = 0; q < n_groups; ++q)
      if (r >= start[q] && r < start[q + 1]) g = q;
    tile_group[tl] = g;
  }
}

}  // namespace lit

using namespace lit;

extern "C" int lit_group_plan(const int32_t* idx, long n_vox, int n_groups, int tile, long rows_cap, int32_t* pos,
                              int32_t* perm, int32_t* tile_group, void* stream) {
  LIT_REQUIRE(n_vox >= 0 && n_groups >= 1 && n_groups <= 32 && tile >= 1, "group_plan: bad extents");
  LIT_REQUIRE(rows_cap % tile == 0 && rows_cap >= ((n_vox + tile - 1) / tile + n_groups) * tile,
              "group_plan: rows_cap must be a multiple of the tile and >= round_up(n_vox, tile) + n_groups * tile");
  static bool attr_set = false;
  const int smem = n_groups * 1024 * (int)sizeof(int);
  if (!attr_set) {
    LIT_CUDA_CHECK(cudaFuncSetAttribute(group_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024 * 4));
    attr_set = true;
  }
  group_plan_kernel<<<1, 1024, smem, (cudaStream_t)stream>>>(idx, n_vox, n_groups, tile, rows_cap, pos, perm, tile_group);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}

extern "C" int lit_col_stats(const float* src, long ld_src, const int32_t* idx, long n_idx, long cols, int ddof,
                             float* mean, float* stdv, double* scratch, void* stream) {
  LIT_REQUIRE(n_idx > 0 && cols >= 0, "col_stats: need at least one row");
  LIT_REQUIRE(scratch != nullptr, "col_stats: scratch (2*cols doubles) required");
  if (cols == 0) return LIT_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int block = 128;
  const bool vec = cols % 4 == 0 && ld_src % 4 == 0 && aligned16(src);
  const int gx = blocks_for(vec ? cols / 4 : cols, block);
  // enough row slices to fill the machine when there are few columns
  int gy = 1;
  const int target = sm_count() * 8;
  if (gx < target) {
    gy = (target + gx - 1) / gx;
    const long max_gy = (n_idx + 63) / 64;
    if (gy > max_gy) gy = (int)max_gy;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
  }
  if (gy > 1) LIT_CUDA_CHECK(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * cols, s));
  if (vec)
    col_moments_kernel<4><<<dim3(gx, gy), block, 0, s>>>(src, ld_src, idx, n_idx, cols, scratch);
  else
    col_moments_kernel<1><<<dim3(gx, gy), block, 0, s>>>(src, ld_src, idx, n_idx, cols, scratch);
  LIT_LAUNCH_CHECK();
  const int gxf = blocks_for(cols, block);
  col_stats_finish_kernel<<<gxf, block, 0, s>>>(src, ld_src, idx, n_idx, cols, ddof, scratch, mean, stdv);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}

extern "C" int lit_gather_normalize_rows(const float* src, long ld_src, const int32_t* idx, long n_idx, long cols,
                                         const float* mean, const float* stdv, int mode, float eps, float* dst,
                                         float* dst_lo, long ld_dst, long n_rows_out, void* stream) {
  LIT_REQUIRE(ld_src >= cols && ld_dst >= cols && n_rows_out >= n_idx, "gather_normalize: bad extents");
  LIT_REQUIRE(mode >= 0 && mode <= 3, "gather_normalize: mode must be 0..3");
  LIT_REQUIRE(mode == 2 || stdv != nullptr, "gather_normalize: std required");
  if (n_rows_out == 0 || cols == 0) return LIT_OK;
  const bool vec = cols % 4 == 0 && ld_src % 4 == 0 && ld_dst % 4 == 0 && aligned16(src) && aligned16(dst) &&
                   (!dst_lo || aligned16(dst_lo));
  const int gx = blocks_for(vec ? cols / 4 : cols, 128);
  long gy = (n_rows_out + 3) / 4;
  const long want = ((long)sm_count() * 32 + gx - 1) / gx;  // enough blocks to fill the machine a few times over
  if (gy > want) gy = want;
  if (gy > 65535) gy = 65535;
  if (gy < 1) gy = 1;
  const dim3 grid(gx, (unsigned)gy);
  cudaStream_t s = (cudaStream_t)stream;
#define LIT_GN(V, S)                                                                                             \
  gather_normalize_kernel<V, S><<<grid, 128, 0, s>>>(src, ld_src, idx, n_idx, cols, mean, stdv, mode, eps, dst, \
                                                     dst_lo, ld_dst, n_rows_out)
  if (vec) {
    if (dst_lo)
      LIT_GN(true, true);
    else
      LIT_GN(true, false);
  } else {
    if (dst_lo)
      LIT_GN(false, true);
    else
      LIT_GN(false, false);
  }
#undef LIT_GN
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}

extern "C" int lit_build_alpha_stack(const float* L, long ld_l, long n_rows, long rows_pad, int k, const float* lam,
                                     const double* alphas, int n_alphas, int normalpha, float singcutoff,
                                     float* col_mean, double* scratch, float* out_hi, float* out_lo, long ld_out,
                                     void* stream) {
  LIT_REQUIRE(n_rows > 0 && rows_pad >= n_rows && k > 0 && n_alphas > 0, "alpha_stack: bad extents");
  LIT_REQUIRE(ld_l >= k && ld_out >= k, "alpha_stack: pitch smaller than k");
  int rc = lit_col_stats(L, ld_l, nullptr, n_rows, k, 0, col_mean, nullptr, scratch, stream);
  if (rc) return rc;
  const int block = 128;
  const int row_chunk = 64;
  const int n_chunks = (int)((rows_pad + row_chunk - 1) / row_chunk);
  LIT_REQUIRE((long)n_alphas * n_chunks <= 65535, "alpha_stack: too many (alpha, row-chunk) pairs");
  const bool vec = k % 4 == 0 && ld_l % 4 == 0 && ld_out % 4 == 0 && aligned16(L) && aligned16(out_hi) && aligned16(out_lo);
  dim3 grid(blocks_for(vec ? k / 4 : k, block), n_alphas * n_chunks);
  if (vec)
    alpha_stack_kernel<4><<<grid, block, 0, (cudaStream_t)stream>>>(L, ld_l, n_rows, rows_pad, k, lam, alphas, normalpha,
                                                                    singcutoff, col_mean, out_hi, out_lo, ld_out,
                                                                    row_chunk);
  else
    alpha_stack_kernel<1><<<grid, block, 0, (cudaStream_t)stream>>>(L, ld_l, n_rows, rows_pad, k, lam, alphas, normalpha,
                                                                    singcutoff, col_mean, out_hi, out_lo, ld_out,
                                                                    row_chunk);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}

extern "C" int lit_scale_rows_by_alpha(const float* Z_hi, const float* Z_lo, long ld_z, long n_vox, int k,
                                       const float* lam, const float* alpha_v, int normalpha, float singcutoff,
                                       float* out_hi, float* out_lo, long ld_out, void* stream) {
  LIT_REQUIRE(ld_z >= k && ld_out >= k && k > 0, "scale_rows: bad extents");
  if (n_vox == 0) return LIT_OK;
  const bool vec = k % 4 == 0 && ld_z % 4 == 0 && ld_out % 4 == 0 && aligned16(Z_hi) && (!Z_lo || aligned16(Z_lo)) &&
                   aligned16(out_hi) && aligned16(out_lo) && aligned16(lam);
  const long items = n_vox * (long)(vec ? k / 4 : k);
  long grid = (items + 255) / 256;
  const long cap = (long)sm_count() * 64;
  if (grid > cap) grid = cap;
  if (vec)
    scale_rows_kernel<true><<<(int)grid, 256, 0, (cudaStream_t)stream>>>(Z_hi, Z_lo, ld_z, n_vox, k, lam, alpha_v,
                                                                         normalpha, singcutoff, out_hi, out_lo, ld_out);
  else
    scale_rows_kernel<false><<<(int)grid, 256, 0, (cudaStream_t)stream>>>(Z_hi, Z_lo, ld_z, n_vox, k, lam, alpha_v,
                                                                          normalpha, singcutoff, out_hi, out_lo, ld_out);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}

static int corr_finalize_common(const float* dot_part, const float* ssq_part, long ld_part, int tiles_per_group,
                                int n_groups, long n_vox, long n_rows, float eps, int accumulate, int metric,
                                const float* resp_std, const float* inv_row, const float* inv_tile, const int32_t* slots,
                                float* corr, long ld_corr, void* stream) {
  LIT_REQUIRE(ld_part >= n_vox && ld_corr >= n_vox, "corr_finalize: pitch smaller than n_vox");
  LIT_REQUIRE(metric >= 0 && metric <= 3, "corr_finalize: metric must be 0..3");
  LIT_REQUIRE((metric & 1) == 0 || resp_std, "corr_finalize: the R^2 metric needs the response std");
  if (n_vox == 0 || n_groups == 0) return LIT_OK;
  corr_finalize_kernel<<<blocks_for(n_vox, 256), 256, 0, (cudaStream_t)stream>>>(
      dot_part, ssq_part, ld_part, tiles_per_group, n_groups, n_vox, n_rows, eps, accumulate, metric, resp_std, inv_row,
      inv_tile, slots, corr, ld_corr);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}

extern "C" int lit_corr_finalize(const float* dot_part, const float* ssq_part, long ld_part, int tiles_per_group,
                                 int n_groups, long n_vox, long n_rows, float eps, int accumulate, int metric,
                                 const float* resp_std, float* corr, long ld_corr, void* stream) {
  return corr_finalize_common(dot_part, ssq_part, ld_part, tiles_per_group, n_groups, n_vox, n_rows, eps, accumulate,
                              metric, resp_std, nullptr, nullptr, nullptr, corr, ld_corr, stream);
}

extern "C" int lit_corr_finalize_scaled(const float* dot_part, const float* ssq_part, long ld_part, int tiles_per_group,
                                        int n_groups, long n_vox, long n_rows, float eps, int accumulate, int metric,
                                        const float* resp_std, const float* inv_row, const float* inv_tile,
                                        const int32_t* slots, float* corr, long ld_corr, void* stream) {
  return corr_finalize_common(dot_part, ssq_part, ld_part, tiles_per_group, n_groups, n_vox, n_rows, eps, accumulate,
                              metric, resp_std, inv_row, inv_tile, slots, corr, ld_corr, stream);
}

extern "C" int lit_corr_finalize_series(const float* series_part, long ld_part, int n_parts, long n_vox, long n_rows,
                                        float eps, int accumulate, int metric, const float* resp_std,
                                        const float* inv_row, const float* inv_tile, const double* coef,
                                        const int32_t* slots, int n_alphas, float* corr, long ld_corr, void* stream) {
  LIT_REQUIRE(ld_part >= n_vox && ld_corr >= n_vox, "corr_finalize_series: pitch smaller than n_vox");
  LIT_REQUIRE(metric >= 0 && metric <= 3, "corr_finalize_series: metric must be 0..3");
  LIT_REQUIRE((metric & 1) == 0 || resp_std, "corr_finalize_series: the R^2 metric needs the response std");
  LIT_REQUIRE(n_alphas == 0 || (coef && slots), "corr_finalize_series: coefficients / slots missing");
  if (n_vox == 0 || n_alphas == 0) return LIT_OK;
  corr_finalize_series_kernel<<<blocks_for(n_vox, 128), 128, 0, (cudaStream_t)stream>>>(
      series_part, ld_part, n_parts, n_vox, n_rows, eps, accumulate, metric, resp_std, inv_row, inv_tile, coef, slots,
      n_alphas, corr, ld_corr);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}

extern "C" int lit_argmax_alpha(const float* corr_sum, long ld_corr, int n_alphas, long n_vox, int n_folds,
                                const float* alphas, int32_t* best, float* alpha_out, double* col_sums, void* stream) {
  LIT_REQUIRE(n_alphas > 0 && n_folds > 0 && ld_corr >= n_vox, "argmax_alpha: bad extents");
  cudaStream_t s = (cudaStream_t)stream;
  // an empty voxel shard (n_vox <= 128 * (world - 1)) still owes the all-reduce true zeros
  if (col_sums) LIT_CUDA_CHECK(cudaMemsetAsync(col_sums, 0, sizeof(double) * n_alphas, s));
  if (n_vox == 0) return LIT_OK;
  argmax_alpha_kernel<<<blocks_for(n_vox, 256), 256, sizeof(double) * n_alphas, s>>>(
      corr_sum, ld_corr, n_alphas, n_vox, n_folds, alphas, best, alpha_out, col_sums);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}
