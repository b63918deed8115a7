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
