// Bandwidth-bound layout / conversion kernels: dtype conversion, TF32 hi/lo split,
// transpose, row gather, gather+transpose, axpy.  All are coalesced along the contiguous
// dimension and float4-vectorised where the pitch allows it; grids are sized from the SM count.
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "../../include/litridge.h"

namespace lit {

static inline int grid_for(size_t work_items, int block, int max_waves = 8) {
  size_t blocks = (work_items + block - 1) / block;
  size_t cap = (size_t)sm_count() * 8 * max_waves;  // persistent-ish grid-stride launch
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---------------------------------------------------------------------------------------------
__global__ void convert_f64_f32_kernel(const double* __restrict__ src, float* __restrict__ dst, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n2 = n / 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    const double2 v = reinterpret_cast<const double2*>(src)[i];
    reinterpret_cast<float2*>(dst)[i] = make_float2((float)v.x, (float)v.y);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && (n & 1)) dst[n - 1] = (float)src[n - 1];
}
__global__ void convert_f32_f64_kernel(const float* __restrict__ src, double* __restrict__ dst, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n2 = n / 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    const float2 v = reinterpret_cast<const float2*>(src)[i];
    reinterpret_cast<double2*>(dst)[i] = make_double2((double)v.x, (double)v.y);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && (n & 1)) dst[n - 1] = (double)src[n - 1];
}
__global__ void fill_kernel(float* __restrict__ dst, size_t n, float v) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = v;
}

// ---------------------------------------------------------------------------------------------
// Generic row-wise elementwise template: out(r, c) over a [rows][cols] matrix, 4 columns per
// thread when VEC (all pitches and cols multiples of 4, bases 16-byte aligned).
template <bool VEC, class F>
__global__ void rowwise_kernel(long rows, long cols, F f) {
  const long cpt = VEC ? 4 : 1;
  const long cols_t = (cols + cpt - 1) / cpt;
  const long total = rows * cols_t;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long r = i / cols_t;
    const long c = (i - r * cols_t) * cpt;
    f(r, c);
  }
}
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

struct SplitOp {
  const float* src;
  long ld_src;
  float* hi;
  float* lo;
  long ld_dst;
  bool vec;
  __device__ void operator()(long r, long c) const {
    if (vec) {
      const float4 x = *reinterpret_cast<const float4*>(src + r * ld_src + c);
      float4 h, l;
      h.x = ptx::to_tf32(x.x);
      h.y = ptx::to_tf32(x.y);
      h.z = ptx::to_tf32(x.z);
      h.w = ptx::to_tf32(x.w);
      l.x = ptx::to_tf32(x.x - h.x);
      l.y = ptx::to_tf32(x.y - h.y);
      l.z = ptx::to_tf32(x.z - h.z);
      l.w = ptx::to_tf32(x.w - h.w);
      *reinterpret_cast<float4*>(hi + r * ld_dst + c) = h;
      *reinterpret_cast<float4*>(lo + r * ld_dst + c) = l;
    } else {
      const float x = src[r * ld_src + c];
      const float h = ptx::to_tf32(x);
      hi[r * ld_dst + c] = h;
      lo[r * ld_dst + c] = ptx::to_tf32(x - h);
    }
  }
};

struct GatherOp {
  const float* src;
  long ld_src;
  const int32_t* idx;
  long n_idx;
  float* dst;
  float* dst_lo;
  long ld_dst;
  bool vec;
  __device__ void operator()(long r, long c) const {
    if (vec) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < n_idx) {
        const long sr = idx ? (long)idx[r] : r;
        if (sr >= 0) x = *reinterpret_cast<const float4*>(src + sr * ld_src + c);  // negative index: a zero row
      }
      if (dst_lo) {
        float4 h, l;
        h.x = ptx::to_tf32(x.x);
        h.y = ptx::to_tf32(x.y);
        h.z = ptx::to_tf32(x.z);
        h.w = ptx::to_tf32(x.w);
        l.x = ptx::to_tf32(x.x - h.x);
        l.y = ptx::to_tf32(x.y - h.y);
        l.z = ptx::to_tf32(x.z - h.z);
        l.w = ptx::to_tf32(x.w - h.w);
        *reinterpret_cast<float4*>(dst + r * ld_dst + c) = h;
        *reinterpret_cast<float4*>(dst_lo + r * ld_dst + c) = l;
      } else {
        *reinterpret_cast<float4*>(dst + r * ld_dst + c) = x;
      }
    } else {
      float x = 0.f;
      if (r < n_idx) {
        const long sr = idx ? (long)idx[r] : r;
        if (sr >= 0) x = src[sr * ld_src + c];
      }
      if (dst_lo) {
        const float h = ptx::to_tf32(x);
        dst[r * ld_dst + c] = h;
        dst_lo[r * ld_dst + c] = ptx::to_tf32(x - h);
      } else {
        dst[r * ld_dst + c] = x;
      }
    }
  }
};

struct AxpyOp {
  float a;
  const float* x_hi;
  const float* x_lo;
  long ld_x;
  float* y;
  long ld_y;
  bool vec;
  __device__ void operator()(long r, long c) const {
    if (vec) {
      float4 x = *reinterpret_cast<const float4*>(x_hi + r * ld_x + c);
      if (x_lo) {
        const float4 l = *reinterpret_cast<const float4*>(x_lo + r * ld_x + c);
        x.x += l.x;
        x.y += l.y;
        x.z += l.z;
        x.w += l.w;
      }
      float4 o = *reinterpret_cast<float4*>(y + r * ld_y + c);
      o.x = fmaf(a, x.x, o.x);
      o.y = fmaf(a, x.y, o.y);
      o.z = fmaf(a, x.z, o.z);
      o.w = fmaf(a, x.w, o.w);
      *reinterpret_cast<float4*>(y + r * ld_y + c) = o;
    } else {
      float x = x_hi[r * ld_x + c];
      if (x_lo) x += x_lo[r * ld_x + c];
      y[r * ld_y + c] = fmaf(a, x, y[r * ld_y + c]);
    }
  }
};

template <class F>
static int launch_rowwise(long rows, long cols, bool vec, F f, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return LIT_OK;
  const long items = rows * ((cols + (vec ? 3 : 0)) / (vec ? 4 : 1));
  const int block = 256;
  const int grid = grid_for((size_t)items, block);
  if (vec)
    rowwise_kernel<true, F><<<grid, block, 0, s>>>(rows, cols, f);
  else
    rowwise_kernel<false, F><<<grid, block, 0, s>>>(rows, cols, f);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}

// ---------------------------------------------------------------------------------------------
// Transposes through a padded shared tile.  Source rows may be gathered through idx.
// dst[c][r] = src[row(r)][c]; columns r in [n_rows, n_rows_pad) of dst are zero-filled.
//
// Fast path (VEC): 64 x 64 tile, 256 threads, float4 global loads along the source columns and float4
// global stores along the destination columns (= source rows), 16 KB in flight per block.
template <bool SPLIT>
__global__ void __launch_bounds__(256)
transpose64_kernel(const float* __restrict__ src, long ld_src, const int32_t* __restrict__ idx, long n_rows,
                   long n_rows_pad, long cols, float* __restrict__ dst, float* __restrict__ dst_lo, long ld_dst) {
  __shared__ float tile[64][65];
  const long tiles_r = (n_rows_pad + 63) / 64;
  const long tiles_c = (cols + 63) / 64;
  const long total = tiles_r * tiles_c;
  const int tx = threadIdx.x & 15;   // float4 column group inside the tile
  const int ty = threadIdx.x >> 4;   // 0..15
  for (long t = blockIdx.x; t < total; t += gridDim.x) {
    const long tr = t % tiles_r;  // consecutive blocks walk along the gathered rows
    const long tc = t / tiles_r;
    const long r0 = tr * 64, c0 = tc * 64;
#pragma unroll
    for (int k = 0; k < 64; k += 16) {
      const long r = r0 + ty + k;
      const long c = c0 + tx * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < n_rows && c < cols) {
        const long sr = idx ? (long)idx[r] : r;
        const float* sp = src + sr * ld_src + c;
        if (c + 3 < cols) {
          v = *reinterpret_cast<const float4*>(sp);
        } else {
          v.x = sp[0];
          if (c + 1 < cols) v.y = sp[1];
          if (c + 2 < cols) v.z = sp[2];
        }
      }
      tile[ty + k][tx * 4 + 0] = v.x;
      tile[ty + k][tx * 4 + 1] = v.y;
      tile[ty + k][tx * 4 + 2] = v.z;
      tile[ty + k][tx * 4 + 3] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 64; k += 16) {
      const long c = c0 + ty + k;       // destination row
      const long r = r0 + tx * 4;       // destination column group
      if (c < cols && r < n_rows_pad) {
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = tile[tx * 4 + q][ty + k];
        float* dp = dst + c * ld_dst + r;
        if (r + 3 < n_rows_pad) {
          if (SPLIT) {
            float h[4], l[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              h[q] = ptx::to_tf32(v[q]);
              l[q] = ptx::to_tf32(v[q] - h[q]);
            }
            *reinterpret_cast<float4*>(dp) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(dst_lo + c * ld_dst + r) = make_float4(l[0], l[1], l[2], l[3]);
          } else {
            *reinterpret_cast<float4*>(dp) = make_float4(v[0], v[1], v[2], v[3]);
          }
        } else {
          for (int q = 0; q < 4 && r + q < n_rows_pad; ++q) {
            if (SPLIT) {
              const float h = ptx::to_tf32(v[q]);
              dp[q] = h;
              dst_lo[c * ld_dst + r + q] = ptx::to_tf32(v[q] - h);
            } else {
              dp[q] = v[q];
            }
          }
        }
      }
    }
    __syncthreads();
  }
}

// Generic path (any alignment): 32 x 32 tile, scalar accesses.
template <bool SPLIT>
__global__ void transpose_kernel(const float* __restrict__ src, long ld_src, const int32_t* __restrict__ idx,
                                 long n_rows, long n_rows_pad, long cols, float* __restrict__ dst,
                                 float* __restrict__ dst_lo, long ld_dst) {
  __shared__ float tile[32][33];
  const long tiles_r = (n_rows_pad + 31) / 32;
  const long tiles_c = (cols + 31) / 32;
  const long total = tiles_r * tiles_c;
  for (long t = blockIdx.x; t < total; t += gridDim.x) {
    const long tr = t % tiles_r;  // consecutive blocks walk along the gathered rows
    const long tc = t / tiles_r;
    const long r0 = tr * 32, c0 = tc * 32;
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
      const long r = r0 + threadIdx.y + k;
      const long c = c0 + threadIdx.x;
      float v = 0.f;
      if (r < n_rows && c < cols) {
        const long sr = idx ? (long)idx[r] : r;
        v = src[sr * ld_src + c];
      }
      tile[threadIdx.y + k][threadIdx.x] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
      const long c = c0 + threadIdx.y + k;
      const long r = r0 + threadIdx.x;
      if (c < cols && r < n_rows_pad) {
        const float v = tile[threadIdx.x][threadIdx.y + k];
        if (SPLIT) {
          const float h = ptx::to_tf32(v);
          dst[c * ld_dst + r] = h;
          dst_lo[c * ld_dst + r] = ptx::to_tf32(v - h);
        } else {
          dst[c * ld_dst + r] = v;
        }
      }
    }
    __syncthreads();
  }
}

static int launch_transpose(const float* src, long ld_src, const int32_t* idx, long n_rows, long n_rows_pad, long cols,
                            float* dst, float* dst_lo, long ld_dst, cudaStream_t s) {
  if (n_rows_pad <= 0 || cols <= 0) return LIT_OK;
  const bool vec = ld_src % 4 == 0 && ld_dst % 4 == 0 && aligned16(src) && aligned16(dst) && (!dst_lo || aligned16(dst_lo));
  if (vec) {
    const long tiles = ((n_rows_pad + 63) / 64) * ((cols + 63) / 64);
    long grid = tiles;
    const long cap = (long)sm_count() * 16;
    if (grid > cap) grid = cap;
    if (dst_lo)
      transpose64_kernel<true><<<(int)grid, 256, 0, s>>>(src, ld_src, idx, n_rows, n_rows_pad, cols, dst, dst_lo, ld_dst);
    else
      transpose64_kernel<false><<<(int)grid, 256, 0, s>>>(src, ld_src, idx, n_rows, n_rows_pad, cols, dst, dst_lo,
                                                          ld_dst);
    LIT_LAUNCH_CHECK();
    return LIT_OK;
  }
  const long tiles = ((n_rows_pad + 31) / 32) * ((cols + 31) / 32);
  long grid = tiles;
  const long cap = (long)sm_count() * 32;
  if (grid > cap) grid = cap;
  dim3 block(32, 8);
  if (dst_lo)
    transpose_kernel<true><<<(int)grid, block, 0, s>>>(src, ld_src, idx, n_rows, n_rows_pad, cols, dst, dst_lo, ld_dst);
  else
    transpose_kernel<false><<<(int)grid, block, 0, s>>>(src, ld_src, idx, n_rows, n_rows_pad, cols, dst, dst_lo,
                                                         ld_dst);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}

}  // namespace lit

using namespace lit;

extern "C" int lit_convert_f64_to_f32(const double* src, float* dst, size_t n, void* stream) {
  if (n == 0) return LIT_OK;
  LIT_REQUIRE(aligned16(src) && (reinterpret_cast<uintptr_t>(dst) & 7) == 0, "convert: misaligned buffers");
  convert_f64_f32_kernel<<<grid_for(n / 2 + 1, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, n);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}
extern "C" int lit_convert_f32_to_f64(const float* src, double* dst, size_t n, void* stream) {
  if (n == 0) return LIT_OK;
  LIT_REQUIRE(aligned16(dst) && (reinterpret_cast<uintptr_t>(src) & 7) == 0, "convert: misaligned buffers");
  convert_f32_f64_kernel<<<grid_for(n / 2 + 1, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, n);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}
extern "C" int lit_fill_f32(float* dst, size_t n, float value, void* stream) {
  if (n == 0) return LIT_OK;
  fill_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(dst, n, value);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}

extern "C" int lit_split_tf32(const float* src, long rows, long cols, long ld_src, float* hi, float* lo, long ld_dst,
                              void* stream) {
  LIT_REQUIRE(ld_src >= cols && ld_dst >= cols, "split: pitch smaller than cols");
  const bool vec = cols % 4 == 0 && ld_src % 4 == 0 && ld_dst % 4 == 0 && aligned16(src) && aligned16(hi) && aligned16(lo);
  SplitOp op{src, ld_src, hi, lo, ld_dst, vec};
  return launch_rowwise(rows, cols, vec, op, (cudaStream_t)stream);
}

extern "C" int lit_transpose_f32(const float* src, long rows, long cols, long ld_src, float* dst, float* dst_lo,
                                 long ld_dst, void* stream) {
  LIT_REQUIRE(ld_src >= cols && ld_dst >= rows, "transpose: pitch too small");
  return launch_transpose(src, ld_src, nullptr, rows, rows, cols, dst, dst_lo, ld_dst, (cudaStream_t)stream);
}

extern "C" int lit_gather_rows_f32(const float* src, long ld_src, const int32_t* idx, long n_idx, long cols, float* dst,
                                   float* dst_lo, long ld_dst, long n_rows_out, void* stream) {
  LIT_REQUIRE(ld_src >= cols && ld_dst >= cols && n_rows_out >= n_idx, "gather_rows: bad extents");
  const bool vec = cols % 4 == 0 && ld_src % 4 == 0 && ld_dst % 4 == 0 && aligned16(src) && aligned16(dst) &&
                   (!dst_lo || aligned16(dst_lo));
  GatherOp op{src, ld_src, idx, n_idx, dst, dst_lo, ld_dst, vec};
  return launch_rowwise(n_rows_out, cols, vec, op, (cudaStream_t)stream);
}

extern "C" int lit_gather_rows_transpose_split(const float* src, long ld_src, const int32_t* idx, long n_idx, long cols,
                                               float* dst_hi, float* dst_lo, long ld_dst, void* stream) {
  LIT_REQUIRE(ld_src >= cols && ld_dst >= n_idx, "gather_rows_transpose: bad extents");
  LIT_REQUIRE(dst_hi, "gather_rows_transpose: destination missing");  // dst_lo == NULL: plain fp32 plane
  return launch_transpose(src, ld_src, idx, n_idx, ld_dst, cols, dst_hi, dst_lo, ld_dst, (cudaStream_t)stream);
}

extern "C" int lit_axpy_f32(float a, const float* x_hi, const float* x_lo, long ld_x, float* y, long ld_y, long rows,
                            long cols, void* stream) {
  LIT_REQUIRE(ld_x >= cols && ld_y >= cols, "axpy: pitch smaller than cols");
  const bool vec = cols % 4 == 0 && ld_x % 4 == 0 && ld_y % 4 == 0 && aligned16(x_hi) && aligned16(y) &&
                   (!x_lo || aligned16(x_lo));
  AxpyOp op{a, x_hi, x_lo, ld_x, y, ld_y, vec};
  return launch_rowwise(rows, cols, vec, op, (cudaStream_t)stream);
}
