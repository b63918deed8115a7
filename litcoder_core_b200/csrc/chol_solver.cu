// Batched direct solver for the small-alpha systems of the inner folds.
//
// For an inner fold with training Gram G (p x p) and centred validation design P_c (m x p), ridge_corr_torch
// (encoding/models/ridge_regression.py:115-120) needs, per alpha, the m x p matrix
//     M_a = P_c (G + a^2 I)^-1,          a = alpha * S[0]
// (pred_a = M_a X^T Y).  The large alphas ride a Neumann series (solver_kernels.cu); the few small ones were
// solved by ~66 single-wave Chebyshev GEMM steps per fold in round 1 (30 % of a config-2 fit).  Here ALL
// (fold, alpha) systems of an outer fold are solved together by a blocked right-looking Cholesky factorisation
// whose trailing updates are batched tcgen05 GEMMs, applied to an AUGMENTED matrix so that every product is an
// "NT" GEMM (both operands K-major) and no triangular solve or transposed factor is ever needed:
//
//            [ A = G + a^2 I ]  n rows            [ L          ]
//     F  =   [ R = P_c       ]  mp rows   --->    [ Y = R L^-T ]      (the elimination that turns A into L applies
//            [ I             ]  n rows            [ W = L^-T   ]       L^-T from the right to every row below)
//
//     M = R A^-1 = R L^-T L^-1 = Y W^T              one more NT GEMM (K loop starts at the tile's first row of W:
//                                                    W is upper triangular)
//
// Per 128-column panel j (o = 128 j):
//   chol_split_kernel        the panel F[o+128:, o:o+128] as TF32 hi/lo planes (S)
//   chol_diag_kernel         Cholesky of the 128 x 128 diagonal block F[o:o+128, o:o+128] in shared memory and its
//                            triangular inverse Linv (split pair out); one CTA per system
//   panel GEMM               S[o+128:, o:o+128] <- S[o+128:, o:o+128] Linv^T          (M = all rows below, N = K = 128)
//   trailing GEMM            F[o+128:act, o+128:n] -= S[o+128:act, panel] S[o+128:n, panel]^T     (K = 128)
// where act = n + mp + o + 128 bounds the rows that are non-zero so far (identity rows become active panel by panel).
// Cost per system: ~ (n + mp) n^2 flops for the elimination + m n^2 for the final product, against ~66 x 2 m n k
// for the Chebyshev chains; everything is launched batched (one launch per step for all systems).
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "../../include/litridge.h"

namespace lit {

int gemm_nt_batched(const float* A_hi, const float* A_lo, long lda, long bs_a, const float* B_hi, const float* B_lo,
                    long ldb, long bs_b, int M, int N, int K, float alpha, const float* Cin, long ldc, long bs_c,
                    float beta, float* D, float* D_lo, long ldd, long bs_d, int batch, int tri_k, cudaStream_t s);

constexpr int CHOL_NB = 128;         // panel width
constexpr int CHOL_MAX_BATCH = 128;  // systems per launch (descriptors travel as kernel parameters: 3 KB)

struct SpdSystems {
  const float* G[CHOL_MAX_BATCH];  // n x n symmetric, pitch ldg
  const float* R[CHOL_MAX_BATCH];  // m x n right-hand side rows, pitch ldr
  int m[CHOL_MAX_BATCH];
  float a2[CHOL_MAX_BATCH];
};

// F[b] <- [G + a2 I (padded with an identity block to n_pad); R (zero rows up to mp); I]   (fp32, pitch ldw)
__global__ void chol_init_kernel(SpdSystems sys, long ldg, long ldr, int n, int n_pad, int mp, float* __restrict__ F,
                                 long ldw, long sys_stride) {
  const int b = blockIdx.z;
  const long row = blockIdx.y;
  const int c4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c4 >= n_pad) return;
  float* out = F + b * sys_stride + row * ldw + c4;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (row < n_pad) {
    if (row < n) {
      const float* g = sys.G[b] + row * ldg;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (c4 + q < n) v[q] = g[c4 + q] + ((c4 + q) == row ? sys.a2[b] : 0.f);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = (c4 + q) == row ? 1.f : 0.f;
    }
  } else if (row < n_pad + mp) {
    const long r = row - n_pad;
    if (r < sys.m[b]) {
      const float* src = sys.R[b] + r * ldr;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (c4 + q < n) v[q] = src[c4 + q];
    }
  } else {
    const long i = row - n_pad - mp;
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = (c4 + q) == i ? 1.f : 0.f;
  }
  *reinterpret_cast<float4*>(out) = make_float4(v[0], v[1], v[2], v[3]);
}

// Panel split: S[rows o+NB.., cols o..o+NB) <- (hi, lo) TF32 planes of F there.  64 rows per block.
__global__ void __launch_bounds__(256) chol_split_kernel(const float* __restrict__ F, long ldw, long sys_stride, int o,
                                                         long rows_total, float* __restrict__ S_hi,
                                                         float* __restrict__ S_lo) {
  constexpr int NB = CHOL_NB;
  const long base = (long)blockIdx.y * sys_stride;
  const long r0 = (long)o + NB + 64L * blockIdx.x;
  const int c4 = (threadIdx.x & 31) * 4;
#pragma unroll
  for (int rr = threadIdx.x >> 5; rr < 64; rr += 8) {
    const long row = r0 + rr;
    if (row >= rows_total) break;
    const long off = base + row * ldw + o + c4;
    const float4 v = *reinterpret_cast<const float4*>(F + off);
    const float x[4] = {v.x, v.y, v.z, v.w};
    float h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      h[q] = ptx::to_tf32(x[q]);
      l[q] = ptx::to_tf32(x[q] - h[q]);
    }
    *reinterpret_cast<float4*>(S_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(S_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
  }
}

// Cholesky of the 128 x 128 diagonal block F[o:o+NB, o:o+NB] in shared memory and its triangular inverse
// Linv = L^-1, written as a TF32 split pair (the B operand of the panel GEMM).  One CTA per system.
__global__ void __launch_bounds__(256) chol_diag_kernel(const float* __restrict__ F, long ldw, long sys_stride, int o,
                                                        float* __restrict__ Dg_hi, float* __restrict__ Dg_lo,
                                                        int* __restrict__ info) {
  constexpr int NB = CHOL_NB;
  const int b = blockIdx.x;
  const float* Fb = F + b * sys_stride;
  extern __shared__ float sm[];
  constexpr int LD = NB + 1;
  float* A = sm;            // [NB][LD]  the block, then L (lower)
  float* X = sm + NB * LD;  // [NB][LD]  L^-1 (lower)
  __shared__ int bad;
  const int tid = threadIdx.x;
  if (tid == 0) bad = 0;
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int i = idx / NB, j = idx % NB;
    A[i * LD + j] = Fb[(long)(o + i) * ldw + o + j];
    X[i * LD + j] = 0.f;
  }
  __syncthreads();
  // Blocked right-looking Cholesky with 32-wide sub-panels: 3 block barriers per sub-panel instead of 3 per column
  // (the first version of this kernel spent most of its 156 us in 384 __syncthreads).
  const int warp = tid >> 5, lane = tid & 31;
  for (int c0 = 0; c0 < NB; c0 += 32) {
    // (1) 32 x 32 diagonal sub-block by one warp, left-looking: at step j lane i forms
    //     L[i][j] = (a[i][j] - sum_{k<j} L[i][k] L[j][k]) / L[j][j]  from rows i and j in shared memory
    if (warp == 0) {
      const int ri = (c0 + lane) * LD + c0;
      for (int j = 0; j < 32; ++j) {
        const int rj = (c0 + j) * LD + c0;
        float sacc = A[ri + j];
        for (int k = 0; k < j; ++k) sacc = fmaf(-A[ri + k], A[rj + k], sacc);
        float dk = __shfl_sync(0xffffffffu, sacc, j);
        if (!(dk > 0.f)) {  // not positive definite (or NaN): flag it, keep going on a harmless pivot
          if (lane == 0) bad = c0 + j + 1;
          dk = 1.f;
        }
        const float piv = sqrtf(dk);
        if (lane >= j) A[ri + j] = (lane == j) ? piv : sacc / piv;
        __syncwarp();
      }
    }
    __syncthreads();
    const int below = NB - c0 - 32;  // rows under the sub-block
    if (below > 0) {
      // (2) sub-panel: one thread per row r solves x L_D^T = A[r][c0 : c0+32] in place
      if (tid < below) {
        const int rr = (c0 + 32 + tid) * LD + c0;
        for (int j = 0; j < 32; ++j) {
          const int rj = (c0 + j) * LD + c0;
          float sacc = A[rr + j];
          for (int k = 0; k < j; ++k) sacc = fmaf(-A[rr + k], A[rj + k], sacc);
          A[rr + j] = sacc / A[rj + j];
        }
      }
      __syncthreads();
      // (3) trailing update of the lower triangle below / right of the sub-panel: rank-32, 16 x 16 threads, each
      //     thread owns the elements (i0 + ty + 16 m, i0 + tx + 16 n)
      const int i0 = c0 + 32, ty = tid >> 4, tx = tid & 15;
      for (int m = 0; m * 16 < below; ++m) {
        const int i = i0 + ty + 16 * m;
        for (int n2 = 0; n2 <= m; ++n2) {
          const int j = i0 + tx + 16 * n2;
          if (i < NB && j <= i) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 32; ++k) acc = fmaf(A[i * LD + c0 + k], A[j * LD + c0 + k], acc);
            A[i * LD + j] -= acc;
          }
        }
      }
      __syncthreads();
    }
  }
  // L^-1 in 32 x 32 blocks.  (a) the four diagonal blocks by forward substitution, one column per thread:
  //   X[i][c] = (delta_ic - sum_{k=c}^{i-1} L[i][k] X[k][c]) / L[i][i]
  if (tid < NB) {
    const int base = tid & ~31, c = tid & 31;
    for (int i = c; i < 32; ++i) {
      float s = (i == c) ? 1.f : 0.f;
      for (int k = c; k < i; ++k) s = fmaf(-A[(base + i) * LD + base + k], X[(base + k) * LD + base + c], s);
      X[(base + i) * LD + base + c] = s / A[(base + i) * LD + base + i];
    }
  }
  __syncthreads();
  // (b) block rows j = 1..3:  X_ji = -X_jj (sum_{k=i}^{j-1} L_jk X_ki)  for all i < j at once
  for (int j = 1; j < NB / 32; ++j) {
    const int n_out = j * 32 * 32;  // elements of block row j left of its diagonal block
    float t[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) {
      const int idx = tid + e * 256;
      t[e] = 0.f;
      if (idx < n_out) {
        const int r = idx / (j * 32), c = idx % (j * 32);  // row inside block row j, absolute column c
        float acc = 0.f;
        for (int k = c & ~31; k < j * 32; ++k) acc = fmaf(A[(j * 32 + r) * LD + k], X[k * LD + c], acc);
        t[e] = acc;
      }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 12; ++e) {
      const int idx = tid + e * 256;
      if (idx < n_out) X[(j * 32 + idx / (j * 32)) * LD + idx % (j * 32)] = t[e];  // T parked in X_ji
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 12; ++e) {
      const int idx = tid + e * 256;
      t[e] = 0.f;
      if (idx < n_out) {
        const int r = idx / (j * 32), c = idx % (j * 32);
        float acc = 0.f;
        for (int q = 0; q <= r; ++q) acc = fmaf(X[(j * 32 + r) * LD + j * 32 + q], X[(j * 32 + q) * LD + c], acc);
        t[e] = -acc;
      }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 12; ++e) {
      const int idx = tid + e * 256;
      if (idx < n_out) X[(j * 32 + idx / (j * 32)) * LD + idx % (j * 32)] = t[e];
    }
    __syncthreads();
  }
  float* oh = Dg_hi + (long)b * NB * NB;
  float* ol = Dg_lo + (long)b * NB * NB;
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int i = idx / NB, j = idx % NB;
    const float x = (j <= i) ? X[i * LD + j] : 0.f;
    const float h = ptx::to_tf32(x);
    oh[idx] = h;
    ol[idx] = ptx::to_tf32(x - h);
  }
  if (tid == 0 && bad && info) atomicMax(info + b, o + bad);
}

// Randomised a-posteriori check of M (G + a2 I) = R on one probe vector x (deterministic +-1 pattern):
//   num = || M (G x + a2 x) - R x ||^2,  den = || R x ||^2   per system (the host takes sqrt(num / den)).
__device__ __forceinline__ float probe_x(int j) { return ((j * 2654435761u) >> 13) & 1u ? 1.f : -1.f; }

// y_b = (G_b + a2_b I) x.  One warp per row; grid (ceil(n / 8), nb).
__global__ void __launch_bounds__(256) spd_probe_y_kernel(SpdSystems sys, long ldg, int n, float* __restrict__ y) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  const float* __restrict__ g = sys.G[b] + (long)i * ldg;
  double acc = 0.0;
  for (int j = lane; j < n; j += 32) acc += (double)g[j] * probe_x(j);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) y[(long)b * n + i] = (float)(acc + (double)sys.a2[b] * probe_x(i));
}

// One warp per right-hand-side row; grid (ceil(max m / 8), nb); numden[2 b], numden[2 b + 1] accumulate num, den.
__global__ void __launch_bounds__(256) spd_probe_err_kernel(SpdSystems sys, long ldr, int n, const float* __restrict__ Msol,
                                                            long ldm, long m_stride, const float* __restrict__ y,
                                                            double* __restrict__ numden) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= sys.m[b]) return;
  const float* __restrict__ mrow = Msol + b * m_stride + (long)i * ldm;
  const float* __restrict__ rrow = sys.R[b] + (long)i * ldr;
  const float* __restrict__ yb = y + (long)b * n;
  double my = 0.0, rx = 0.0;
  for (int j = lane; j < n; j += 32) {
    my += (double)mrow[j] * yb[j];
    rx += (double)rrow[j] * probe_x(j);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    my += __shfl_xor_sync(0xffffffffu, my, o);
    rx += __shfl_xor_sync(0xffffffffu, rx, o);
  }
  if (lane == 0) {
    atomicAdd(numden + 2 * b, (my - rx) * (my - rx));
    atomicAdd(numden + 2 * b + 1, rx * rx);
  }
}

static int fill_systems(SpdSystems& sys, int nb, const void* const* G, const void* const* R, const int* m,
                        const float* a2) {
  LIT_REQUIRE(nb >= 0 && nb <= CHOL_MAX_BATCH, "spd_solve: at most %d systems per call (got %d)", CHOL_MAX_BATCH, nb);
  for (int b = 0; b < nb; ++b) {
    sys.G[b] = static_cast<const float*>(G[b]);
    sys.R[b] = static_cast<const float*>(R[b]);
    sys.m[b] = m[b];
    sys.a2[b] = a2[b];
    LIT_REQUIRE(sys.G[b] && sys.R[b] && m[b] >= 0, "spd_solve: null operand / negative row count");
  }
  return LIT_OK;
}

}  // namespace lit

using namespace lit;

extern "C" int lit_spd_solve_workspace(int nb, int n, int mp, size_t* f_floats, size_t* s_floats, size_t* dg_floats,
                                       long* ldw, long* rows_total) {
  LIT_REQUIRE(nb >= 0 && n >= 1 && mp >= 0, "spd_solve_workspace: bad extents");
  const long n_pad = (n + CHOL_NB - 1) / CHOL_NB * CHOL_NB;
  const long rows = 2 * n_pad + mp;
  *ldw = n_pad;
  *rows_total = rows;
  *f_floats = (size_t)nb * rows * n_pad;
  *s_floats = (size_t)nb * rows * n_pad;  // per plane
  *dg_floats = (size_t)nb * CHOL_NB * CHOL_NB;  // per plane
  return LIT_OK;
}

// Factor and eliminate (see the header comment).  On return, for system b (stride rows_total * ldw floats):
//   S rows [n_pad, n_pad + mp)      : Y = R L^-T        (split pair)
//   S rows [n_pad + mp, 2 n_pad + mp): W = L^-T          (split pair, upper triangular)
// so that M_b = Y_b W_b^T, e.g. by lit_gemm_tf32x3_nt_batched(..., tri_k = 1).  info[b] != 0: system b was not
// positive definite (1-based index of the first bad pivot).
extern "C" int lit_spd_solve_batched(int nb, int n, int mp, const void* const* G_h, long ldg, const void* const* R_h,
                                     long ldr, const int* m_h, const float* a2_h, float* F, float* S_hi, float* S_lo,
                                     float* Dg_hi, float* Dg_lo, int* info, void* stream) {
  LIT_REQUIRE(n >= 1 && mp >= 0 && ldg >= n && ldr >= n, "spd_solve: bad extents");
  if (nb == 0) return LIT_OK;
  SpdSystems sys = {};
  int rc = fill_systems(sys, nb, G_h, R_h, m_h, a2_h);
  if (rc) return rc;
  for (int b = 0; b < nb; ++b) LIT_REQUIRE(m_h[b] <= mp, "spd_solve: right-hand side has more rows than mp");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int n_pad = (n + CHOL_NB - 1) / CHOL_NB * CHOL_NB;
  const long ldw = n_pad;
  const long rows = 2L * n_pad + mp;
  const long sys_stride = rows * ldw;
  LIT_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int) * nb, s));
  {
    dim3 grid((n_pad / 4 + 127) / 128, (unsigned)rows, nb);
    chol_init_kernel<<<grid, 128, 0, s>>>(sys, ldg, ldr, n, n_pad, mp, F, ldw, sys_stride);
    LIT_LAUNCH_CHECK();
  }
  static bool attr_set = false;
  const int diag_smem = 2 * CHOL_NB * (CHOL_NB + 1) * (int)sizeof(float);
  if (!attr_set) {
    LIT_CUDA_CHECK(cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, diag_smem));
    attr_set = true;
  }
  const int steps = n_pad / CHOL_NB;
  for (int j = 0; j < steps; ++j) {
    const int o = j * CHOL_NB;
    const long below = rows - (o + CHOL_NB);  // rows under the diagonal block (L21, R, identity rows)
    if (below > 0) {
      dim3 grid((unsigned)((below + 63) / 64), nb);
      chol_split_kernel<<<grid, 256, 0, s>>>(F, ldw, sys_stride, o, rows, S_hi, S_lo);
      LIT_LAUNCH_CHECK();
    }
    chol_diag_kernel<<<nb, 256, diag_smem, s>>>(F, ldw, sys_stride, o, Dg_hi, Dg_lo, info);
    LIT_LAUNCH_CHECK();
    if (below <= 0) break;
    const long off_panel = (long)(o + CHOL_NB) * ldw + o;
    // panel: S[o+NB:, o:o+NB] <- S[o+NB:, o:o+NB] Linv^T   (in place: a tile reads exactly the rows it writes)
    rc = gemm_nt_batched(S_hi + off_panel, S_lo + off_panel, ldw, sys_stride, Dg_hi, Dg_lo, CHOL_NB,
                         (long)CHOL_NB * CHOL_NB, (int)below, CHOL_NB, CHOL_NB, 1.f, nullptr, 0, 0, 0.f,
                         S_hi + off_panel, S_lo + off_panel, ldw, sys_stride, nb, 0, s);
    if (rc) return rc;
    const int r = n_pad - (o + CHOL_NB);  // columns (and L rows) still to eliminate
    if (r > 0) {
      long act = (long)n_pad + mp + o + CHOL_NB;  // rows that can be non-zero in this panel
      if (act > rows) act = rows;
      const long off_trail = (long)(o + CHOL_NB) * ldw + (o + CHOL_NB);
      const int m_upd = (int)(act - (o + CHOL_NB));
      if (j % 2 == 0 && r > CHOL_NB) {
        // first panel of a pair: only the NEXT 128 columns need it now (they are the second panel); the rest of the
        // trailing matrix gets both panels at once below -- half as many K = 128 passes over the big matrix
        rc = gemm_nt_batched(S_hi + off_panel, S_lo + off_panel, ldw, sys_stride, S_hi + off_panel, S_lo + off_panel, ldw,
                             sys_stride, m_upd, CHOL_NB, CHOL_NB, -1.f, F + off_trail, ldw, sys_stride, 1.f,
                             F + off_trail, nullptr, ldw, sys_stride, nb, 0, s);
      } else if (j % 2 == 1) {
        // second panel of a pair: rank-256 update with both panels (adjacent columns o-128 .. o+128 of S)
        const long off_pair = off_panel - CHOL_NB;
        rc = gemm_nt_batched(S_hi + off_pair, S_lo + off_pair, ldw, sys_stride, S_hi + off_pair, S_lo + off_pair, ldw,
                             sys_stride, m_upd, r, 2 * CHOL_NB, -1.f, F + off_trail, ldw, sys_stride, 1.f, F + off_trail,
                             nullptr, ldw, sys_stride, nb, 0, s);
      } else {
        // an unpaired first panel whose trailing matrix is a single 128-column block
        rc = gemm_nt_batched(S_hi + off_panel, S_lo + off_panel, ldw, sys_stride, S_hi + off_panel, S_lo + off_panel, ldw,
                             sys_stride, m_upd, r, CHOL_NB, -1.f, F + off_trail, ldw, sys_stride, 1.f, F + off_trail,
                             nullptr, ldw, sys_stride, nb, 0, s);
      }
      if (rc) return rc;
    }
  }
  return LIT_OK;
}

extern "C" int lit_spd_probe_residual(int nb, int n, const void* const* G_h, long ldg, const void* const* R_h, long ldr,
                                      const int* m_h, const float* a2_h, const float* M, long ldm, long m_stride,
                                      float* scratch, double* numden, void* stream) {
  if (nb == 0) return LIT_OK;
  SpdSystems sys = {};
  int rc = fill_systems(sys, nb, G_h, R_h, m_h, a2_h);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int m_max = 0;
  for (int b = 0; b < nb; ++b) m_max = m_h[b] > m_max ? m_h[b] : m_max;
  LIT_CUDA_CHECK(cudaMemsetAsync(numden, 0, sizeof(double) * 2 * nb, s));
  spd_probe_y_kernel<<<dim3((n + 7) / 8, nb), 256, 0, s>>>(sys, ldg, n, scratch);
  LIT_LAUNCH_CHECK();
  if (m_max > 0) {
    spd_probe_err_kernel<<<dim3((m_max + 7) / 8, nb), 256, 0, s>>>(sys, ldr, n, M, ldm, m_stride, scratch, numden);
    LIT_LAUNCH_CHECK();
  }
  return LIT_OK;
}
