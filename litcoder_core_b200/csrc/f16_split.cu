// fp16 split pairs for the kind::f16 variant of the fused prediction + correlation GEMM
// (lit_gemm_f16x3_nt_corr; replaces the same reference lines as the 3xTF32 variant:
// encoding/models/ridge_regression.py:115-133).
//
// A value x is carried as  hi = fp16(s x),  lo = fp16(s x - hi)  with a power-of-two scale s shared by a
// GROUP of consecutive rows (one voxel row of the coefficient matrix; all rows of one alpha of the stacked
// validation design), chosen so that the largest magnitude of the group lands in [2^14, 2^15).  fp16 has the
// 11-bit significand of TF32, so hi*hi + hi*lo + lo*hi is as exact as 3xTF32 (2^-22 relative) for every
// element within 2^-14 of its group's maximum; smaller elements keep an ABSOLUTE error of 2^-40 of the group
// maximum (fp16 subnormal spacing 2^-24 on a 2^15 scale), far below the fp32 rounding of the sums they enter.
// kind::f16 MMAs consume 16 values of K per instruction instead of 8: twice the tensor-core rate and half the
// operand bytes of the TF32 form.  The scales are undone on the per-voxel partial sums (lit_corr_finalize_scaled).
//
// Three bandwidth-bound passes: group |max| (warp per row, float4 loads, one atomicMax per row), the scales,
// and the conversion (8 values per thread: two float4 loads per plane, one 16-byte store per output plane).
#include "common.cuh"
#include "../../include/litridge.h"

#include <cuda_fp16.h>

namespace lit {

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

__device__ __forceinline__ float absmax4(float4 x) {
  return fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w)));
}

// gmax[r / rows_per_group] = max |hi + lo| as the bit pattern of a non-negative float (ordered like uint32).
__global__ void __launch_bounds__(256)
group_absmax_kernel(const float* __restrict__ hi, const float* __restrict__ lo, long ld, long rows, long cols,
                    long rows_per_group, unsigned* __restrict__ gmax, int vec) {
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long n_warps = ((long)gridDim.x * blockDim.x) >> 5;
  const long cols4 = vec ? (cols & ~3L) : 0;
  for (long r = warp; r < rows; r += n_warps) {
    const float* h = hi + r * ld;
    const float* l = lo ? lo + r * ld : nullptr;
    float m = 0.f;
#pragma unroll 4
    for (long c = lane * 4L; c < cols4; c += 128) {
      float4 x = *reinterpret_cast<const float4*>(h + c);
      if (l) {
        const float4 y = *reinterpret_cast<const float4*>(l + c);
        x.x += y.x;
        x.y += y.y;
        x.z += y.z;
        x.w += y.w;
      }
      m = fmaxf(m, absmax4(x));
    }
    for (long c = cols4 + lane; c < cols; c += 32) m = fmaxf(m, fabsf(l ? h[c] + l[c] : h[c]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) atomicMax(gmax + r / rows_per_group, __float_as_uint(m));
  }
}

// scale = 2^(15 - e) with max = f 2^e, f in [0.5, 1): the scaled maximum is < 2^15 (fp16 max is 65504).
// An all-zero or non-finite group keeps scale 1 (inf / NaN then propagate as they do in the TF32 form).
__global__ void f16_scales_kernel(const unsigned* __restrict__ gmax, long n, float* __restrict__ scale,
                                  float* __restrict__ inv_scale) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float m = __uint_as_float(gmax[i]);
  float s = 1.f, inv = 1.f;
  if (m > 0.f && isfinite(m)) {
    int e;
    frexpf(m, &e);
    const int se = min(max(15 - e, -100), 100);
    s = ldexpf(1.f, se);
    inv = ldexpf(1.f, -se);
  }
  scale[i] = s;
  inv_scale[i] = inv;
}

__device__ __forceinline__ void split_h(float y, __half& h, __half& l) {
  h = __float2half_rn(y);
  l = __float2half_rn(y - __half2float(h));
}

__global__ void __launch_bounds__(256)
split_f16_kernel(const float* __restrict__ hi, const float* __restrict__ lo, long ld, long rows, long cols,
                 long rows_per_group, const float* __restrict__ scale, __half* __restrict__ out_hi,
                 __half* __restrict__ out_lo, long ld_out, int vec) {
  const long cols8 = (cols + 7) / 8;
  const long total = rows * cols8;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long r = i / cols8;
    const long c = (i - r * cols8) * 8;
    const float s = scale[r / rows_per_group];
    const float* h = hi + r * ld + c;
    const float* l = lo ? lo + r * ld + c : nullptr;
    __half* oh = out_hi + r * ld_out + c;
    __half* ol = out_lo + r * ld_out + c;
    if (vec && c + 8 <= cols) {
      __align__(16) float x[8];
      *reinterpret_cast<float4*>(x) = *reinterpret_cast<const float4*>(h);
      *reinterpret_cast<float4*>(x + 4) = *reinterpret_cast<const float4*>(h + 4);
      if (l) {
        __align__(16) float y[8];
        *reinterpret_cast<float4*>(y) = *reinterpret_cast<const float4*>(l);
        *reinterpret_cast<float4*>(y + 4) = *reinterpret_cast<const float4*>(l + 4);
#pragma unroll
        for (int q = 0; q < 8; ++q) x[q] += y[q];
      }
      __align__(16) __half vh[8];
      __align__(16) __half vl[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) split_h(x[q] * s, vh[q], vl[q]);
      *reinterpret_cast<uint4*>(oh) = *reinterpret_cast<const uint4*>(vh);
      *reinterpret_cast<uint4*>(ol) = *reinterpret_cast<const uint4*>(vl);
    } else {
      for (int q = 0; q < 8 && c + q < cols; ++q) {
        const float x = l ? h[q] + l[q] : h[q];
        split_h(x * s, oh[q], ol[q]);
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Scales chosen from a BOUND instead of the data, so that a producer can write fp16 pairs directly.
//
// The cross product of an inner fold is the outer fold's minus the removed rows' contribution,
// C_i^T = C_o^T - Y_R^T X_R, so |C_i^T[v][j]| <= max_j |C_o^T[v][j]| + |y_{v,R}|_2 max_j |x_{j,R}|_2 (Cauchy-Schwarz on
// the removed rows only: typically within 2^3 of the true row maximum).  A gathered response row never exceeds the
// column maximum of |Y| over all rows.  Both bounds are known BEFORE the producing kernel runs; scaling the bound (plus
// a 2^-10 rounding margin) into [2^14, 2^15) can therefore never overflow fp16, and the pair keeps the accuracy
// stated at the top of this file as long as the true maximum is within 2^19 of the bound.

// out_sumsq[c] += sum_r src[row(r)][c]^2, out_absmax[c] = max_r |src[row(r)][c]| over a slab of rows per blockIdx.y
// (outputs zeroed by the caller; absmax as the bit pattern of a non-negative float).
__global__ void __launch_bounds__(256)
gather_col_reduce_kernel(const float* __restrict__ src, long ld, const int32_t* __restrict__ idx, long n_idx, long cols,
                         float* __restrict__ out_sumsq, unsigned* __restrict__ out_absmax) {
  const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const long per = (n_idx + gridDim.y - 1) / gridDim.y;
  const long r0 = blockIdx.y * per, r1 = min(n_idx, r0 + per);
  float ss = 0.f, m = 0.f;
  bool nan = false;  // fmaxf drops NaN; keep it so that the column falls back to scale 1 like lit_split_f16
#pragma unroll 8
  for (long r = r0; r < r1; ++r) {
    const long sr = idx ? (long)idx[r] : r;
    const float x = sr >= 0 ? src[sr * ld + c] : 0.f;
    ss = fmaf(x, x, ss);
    m = fmaxf(m, fabsf(x));
    nan |= x != x;
  }
  if (r1 > r0) {
    if (out_sumsq) atomicAdd(out_sumsq + c, ss);
    if (out_absmax) atomicMax(out_absmax + c, nan ? 0x7fc00000u : __float_as_uint(m));
  }
}

// bound[r] = absmax[r] + sqrt(row_sumsq[r] * max_j col_sumsq[j]); scale[r] = 2^k with bound * (1 + 2^-10) * scale in
// [2^14, 2^15) (1 for a zero or non-finite bound).
__global__ void __launch_bounds__(256)
bound_scales_kernel(const unsigned* __restrict__ absmax, const float* __restrict__ row_sumsq,
                    const float* __restrict__ col_sumsq, long n_cs, long rows, float* __restrict__ scale,
                    float* __restrict__ inv_scale) {
  __shared__ float red[8];
  float cmax = 0.f;
  if (row_sumsq && col_sumsq) {
    for (long j = threadIdx.x; j < n_cs; j += blockDim.x) {
      const float x = col_sumsq[j];
      cmax = x != x ? x : fmaxf(cmax, x);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float y = __shfl_xor_sync(0xffffffffu, cmax, o);
      cmax = (y != y || cmax != cmax) ? nanf("") : fmaxf(cmax, y);
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cmax;
    __syncthreads();
    cmax = red[0];
    for (int w = 1; w < 8; ++w) cmax = (red[w] != red[w] || cmax != cmax) ? nanf("") : fmaxf(cmax, red[w]);
  }
  const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float bound = absmax ? __uint_as_float(absmax[r]) : 0.f;
  if (row_sumsq && col_sumsq) bound += sqrtf(row_sumsq[r] * cmax);
  bound *= 1.f + 0x1p-10f;
  float s = 1.f, inv = 1.f;
  if (bound > 0.f && isfinite(bound)) {
    int e;
    frexpf(bound, &e);
    const int se = min(max(15 - e, -100), 100);
    s = ldexpf(1.f, se);
    inv = ldexpf(1.f, -se);
  }
  scale[r] = s;
  inv_scale[r] = inv;
}

// dst[c][r] = fp16 pair of scale[c] * src[row(r)][c]: the gathered, transposed response rows of a fold as the A
// operand of an fp16-pair GEMM (one scale per voxel = per output row).  64 x 64 tile through shared memory; columns
// r in [n_rows, n_rows_pad) are zero-filled.  The tile has no padding: 16-byte groups are XOR-swizzled with
// (row >> 3) & 7, which makes the float4 stores of the load phase and the scalar column reads of the store phase
// (8 lanes x 8 rows apart, 4 adjacent columns) both bank-conflict-free.  (A 64 x 128 tile with a padded pitch and
// scalar stores was measured slower: 0.305 vs 0.285 ms for the 1,500-row gather.)
__global__ void __launch_bounds__(256)
transpose64_f16_kernel(const float* __restrict__ src, long ld_src, const int32_t* __restrict__ idx, long n_rows,
                       long n_rows_pad, long cols, const float* __restrict__ scale, __half* __restrict__ dst_hi,
                       __half* __restrict__ dst_lo, long ld_dst, int vec) {
  __shared__ __align__(16) float tile[64 * 64];
  const long tiles_r = (n_rows_pad + 63) / 64;
  const long tiles_c = (cols + 63) / 64;
  const long total = tiles_r * tiles_c;
  for (long t = blockIdx.x; t < total; t += gridDim.x) {
    const long tr = t % tiles_r;  // consecutive blocks walk along the gathered rows
    const long tc = t / tiles_r;
    const long r0 = tr * 64, c0 = tc * 64;
    {
      const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 float4 column groups x 16 rows per pass
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const long r = r0 + ty + 16 * k;
        const long c = c0 + tx * 4;
        v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < n_rows && c < cols) {
          const long sr = idx ? (long)idx[r] : r;
          if (sr >= 0) {
            const float* sp = src + sr * ld_src + c;
            if (vec && c + 3 < cols) {
              v[k] = *reinterpret_cast<const float4*>(sp);
            } else {
              v[k].x = sp[0];
              if (c + 1 < cols) v[k].y = sp[1];
              if (c + 2 < cols) v[k].z = sp[2];
              if (c + 3 < cols) v[k].w = sp[3];
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int row = ty + 16 * k;
        *reinterpret_cast<float4*>(tile + row * 64 + ((tx ^ ((row >> 3) & 7)) << 2)) = v[k];
      }
    }
    __syncthreads();
    {
      const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;  // 8 values (16 bytes per plane) per thread, 32 rows per pass
#pragma unroll
      for (int k = 0; k < 64; k += 32) {
        const int cl = ty + k;       // column inside the tile = destination row
        const long c = c0 + cl;
        const long r = r0 + tx * 8;  // destination column group
        if (c < cols && r < n_rows_pad) {
          const float s = scale[c];
          __align__(16) __half vh[8];
          __align__(16) __half vl[8];
          // rows 8 tx + q: (row >> 3) & 7 == tx for every q
          const float* col = tile + tx * 8 * 64 + ((((cl >> 2) ^ tx) << 2) | (cl & 3));
#pragma unroll
          for (int q = 0; q < 8; ++q) split_h(col[q * 64] * s, vh[q], vl[q]);
          __half* oh = dst_hi + c * ld_dst + r;
          __half* ol = dst_lo + c * ld_dst + r;
          if (vec && r + 7 < n_rows_pad) {
            *reinterpret_cast<uint4*>(oh) = *reinterpret_cast<const uint4*>(vh);
            *reinterpret_cast<uint4*>(ol) = *reinterpret_cast<const uint4*>(vl);
          } else {
            for (int q = 0; q < 8 && r + q < n_rows_pad; ++q) {
              oh[q] = vh[q];
              ol[q] = vl[q];
            }
          }
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace lit

using namespace lit;

extern "C" int lit_split_f16(const float* src_hi, const float* src_lo, long ld_src, long rows, long cols,
                             long rows_per_group, void* out_hi, void* out_lo, long ld_out, float* inv_scale,
                             void* scratch, void* stream) {
  LIT_REQUIRE(rows >= 0 && cols >= 0 && rows_per_group > 0, "split_f16: bad extents");
  LIT_REQUIRE(ld_src >= cols && ld_out >= cols, "split_f16: pitch smaller than cols");
  LIT_REQUIRE(src_hi && out_hi && out_lo && inv_scale && scratch, "split_f16: null pointer");
  if (rows == 0 || cols == 0) return LIT_OK;
  const long n_groups = (rows + rows_per_group - 1) / rows_per_group;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned* gmax = static_cast<unsigned*>(scratch);
  float* scale = reinterpret_cast<float*>(gmax + n_groups);
  const int vec_in = aligned16(src_hi) && (!src_lo || aligned16(src_lo)) && ld_src % 4 == 0;
  const int vec_out = vec_in && aligned16(out_hi) && aligned16(out_lo) && ld_out % 8 == 0;
  LIT_CUDA_CHECK(cudaMemsetAsync(gmax, 0, (size_t)n_groups * sizeof(unsigned), s));
  const long cap = (long)sm_count() * 8;  // 8 resident 256-thread blocks per SM
  long blocks = (rows + 7) / 8;           // one warp per row
  group_absmax_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, s>>>(src_hi, src_lo, ld_src, rows, cols,
                                                                             rows_per_group, gmax, vec_in);
  LIT_LAUNCH_CHECK();
  f16_scales_kernel<<<(unsigned)((n_groups + 255) / 256), 256, 0, s>>>(gmax, n_groups, scale, inv_scale);
  LIT_LAUNCH_CHECK();
  blocks = (rows * ((cols + 7) / 8) + 255) / 256;
  split_f16_kernel<<<(unsigned)(blocks < cap * 4 ? blocks : cap * 4), 256, 0, s>>>(
      src_hi, src_lo, ld_src, rows, cols, rows_per_group, scale, static_cast<__half*>(out_hi),
      static_cast<__half*>(out_lo), ld_out, vec_out);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}

extern "C" int lit_gather_col_reduce(const float* src, long ld_src, const int32_t* idx, long n_idx, long cols,
                                     float* out_sumsq, float* out_absmax, void* stream) {
  LIT_REQUIRE(n_idx >= 0 && cols >= 0 && ld_src >= cols, "gather_col_reduce: bad extents");
  LIT_REQUIRE(out_sumsq || out_absmax, "gather_col_reduce: no output");
  if (cols == 0) return LIT_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (out_sumsq) LIT_CUDA_CHECK(cudaMemsetAsync(out_sumsq, 0, (size_t)cols * sizeof(float), s));
  if (out_absmax) LIT_CUDA_CHECK(cudaMemsetAsync(out_absmax, 0, (size_t)cols * sizeof(float), s));
  if (n_idx == 0) return LIT_OK;
  const long bx = (cols + 255) / 256;
  long by = (long)sm_count() * 8 / bx;  // enough row slabs to fill the machine when there are few columns
  by = by < 1 ? 1 : (by > 64 ? 64 : by);
  if (by > (n_idx + 31) / 32) by = (n_idx + 31) / 32;
  gather_col_reduce_kernel<<<dim3((unsigned)bx, (unsigned)by), 256, 0, s>>>(src, ld_src, idx, n_idx, cols, out_sumsq,
                                                                            reinterpret_cast<unsigned*>(out_absmax));
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}

extern "C" int lit_row_absmax(const float* src, long ld_src, long rows, long cols, float* out, void* stream) {
  LIT_REQUIRE(rows >= 0 && cols >= 0 && ld_src >= cols && out, "row_absmax: bad arguments");
  if (rows == 0) return LIT_OK;
  cudaStream_t s = (cudaStream_t)stream;
  LIT_CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t)rows * sizeof(float), s));
  if (cols == 0) return LIT_OK;
  const long cap = (long)sm_count() * 8;
  const long blocks = (rows + 7) / 8;
  group_absmax_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, s>>>(
      src, nullptr, ld_src, rows, cols, 1, reinterpret_cast<unsigned*>(out), aligned16(src) && ld_src % 4 == 0);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}

extern "C" int lit_f16_bound_scales(const float* absmax, const float* row_sumsq, const float* col_sumsq, long n_cs,
                                    long rows, float* scale, float* inv_scale, void* stream) {
  LIT_REQUIRE(rows >= 0 && n_cs >= 0 && scale && inv_scale, "f16_bound_scales: bad arguments");
  LIT_REQUIRE((row_sumsq == nullptr) == (col_sumsq == nullptr), "f16_bound_scales: the two norm vectors go together");
  LIT_REQUIRE(absmax || row_sumsq, "f16_bound_scales: no bound given");
  if (rows == 0) return LIT_OK;
  bound_scales_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const unsigned*>(absmax), row_sumsq, col_sumsq, n_cs, rows, scale, inv_scale);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}

extern "C" int lit_gather_rows_transpose_f16(const float* src, long ld_src, const int32_t* idx, long n_idx, long cols,
                                             const float* scale, void* dst_hi, void* dst_lo, long ld_dst, void* stream) {
  LIT_REQUIRE(ld_src >= cols && ld_dst >= n_idx && n_idx >= 0 && cols >= 0, "gather_rows_transpose_f16: bad extents");
  LIT_REQUIRE(scale && dst_hi && dst_lo, "gather_rows_transpose_f16: null pointer");
  if (cols == 0 || ld_dst == 0) return LIT_OK;
  const int vec = aligned16(src) && ld_src % 4 == 0 && aligned16(dst_hi) && aligned16(dst_lo) && ld_dst % 8 == 0;
  const long tiles = ((ld_dst + 63) / 64) * ((cols + 63) / 64);
  const long cap = (long)sm_count() * 16;
  transpose64_f16_kernel<<<(unsigned)(tiles < cap ? tiles : cap), 256, 0, (cudaStream_t)stream>>>(
      src, ld_src, idx, n_idx, ld_dst, cols, scale, static_cast<__half*>(dst_hi), static_cast<__half*>(dst_lo), ld_dst,
      vec);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}
