// 3xTF32 split-precision GEMM on the sm_100a tensor cores (tcgen05 + TMEM + TMA).
//
//   D[M,N] = alpha * A[M,K] * B[N,K]^T (+ beta * Cin)        ("NT": both operands K-major)
//
// Every dense contraction of the nested-CV ridge path is expressed in this one form
// (see DESIGN.md): the Gram X^T X, the cross product Y^T X, the rotation into the
// eigenbasis, the per-alpha validation predictions and the weight solve.  They replace
// the torch.matmul / torch.linalg.svd call sites of the reference
// (encoding/models/ridge_regression.py:32,59-61,104-105,120; nested_cv.py:151,251).
//
// Precision.  Operands arrive as two fp32 planes hi = rna_tf32(x), lo = rna_tf32(x - hi)
// (both exactly representable in TF32), and each k-step issues three kind::tf32 MMAs
//   lo*hi + hi*lo + hi*hi
// into the same fp32 TMEM accumulator.  The products are then exact to ~2^-22, but the
// tensor core's fp32 accumulator TRUNCATES on every accumulation (measured on B200:
// relative bias -6e-9 * K, i.e. -2e-5 at K = 3072, 20x the error of an fp32 FMA chain;
// profiles/r1_gemm_accuracy.md).  The kernel therefore accumulates only a short K chunk
// (kc_blocks * 32 values of K, 128 by default) inside TMEM and sums the chunks in
// registers with round-to-nearest fp32 adds -- the accumulate-outside-the-tensor-core
// scheme of Ootomo & Yokota, mapped onto TMEM double buffering: while the tensor core
// fills one 128 x BN accumulator buffer, the epilogue warps drain the other one.
//
// Kernel structure (persistent, warp-specialised, one CTA or one CTA pair per SM, 384 threads):
//   warp 0    : TMA producer  (4 tiled loads per k-block: A_hi, A_lo, B_hi, B_lo; SWIZZLE_128B)
//   warp 1    : MMA issuer    (one thread issues tcgen05.mma; commits free the smem stage and
//                              publish a finished K chunk)
//   warp 2    : TMEM allocator
//   warps 4-11: accumulate + epilogue.  Warp w owns TMEM lanes 32*(w%4).. and the column half
//               (w-4)/4 of the tile: 128 x BN fp32 running sums live in registers
//               (BN/2 per thread; setmaxnreg moves registers from warps 0-3 to these warps).
// Two epilogues:
//   EPI_STORE : D = alpha*acc + beta*Cin, optionally written as a (hi, lo) TF32 split pair
//   EPI_CORR  : fused column reduction for per-voxel correlation.  Accumulator rows are
//               voxels, columns are (alpha group, time) pairs; each thread reduces its row
//               against the z-scored responses Yz[t][v] and emits per-half-tile partial sums
//               sum(pred*yz) and sum(pred^2).  Predictions never reach HBM
//               (the reference materialises them per alpha: ridge_regression.py:120-125).
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "../../include/litridge.h"

#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <cstdlib>
#include <mutex>

namespace lit {

enum { EPI_STORE = 0, EPI_CORR = 1 };

struct GemmParams {
  int M, N, K;
  int num_m_tiles;  // in units of BM*CG rows
  int num_n_tiles;
  int num_k_blocks;
  int kc_blocks;  // k-blocks accumulated inside TMEM before the chunk is drained into registers
  int group_m;    // rasterisation group (tiles along M that share a B tile wave)
  // EPI_STORE
  float* D;
  float* D_lo;  // optional: when non-null D receives hi and D_lo receives lo
  long ldd;
  const float* Cin;
  long ldc;
  float alpha, beta;
  // EPI_STORE on fp16 split pairs: power-of-two operand scales to undo, D = alpha * scale_m[row] * scale_n[col] * acc
  // (either may be NULL = ones; lit_split_f16 writes them as inv_scale)
  const float* scale_m;
  const float* scale_n;
  // EPI_STORE, fp16-pair output (optional, H_hi != NULL): the result is ALSO (or, with D == NULL, only) written as
  // the scaled fp16 split pair the next fp16-pair GEMM consumes: hi = fp16(s d), lo = fp16(s d - hi) with
  // s = out_scale[row] (a power of two chosen by the caller from a bound on |d|; see lit_f16_bound_scales).
  __half* H_hi;
  __half* H_lo;
  long ldh;
  const float* out_scale;
  // EPI_CORR
  const float* Yz;  // [parts_per_group*BN/2 rows][>= M cols], row pitch ldy
  long ldy;
  int tiles_per_group;  // N tiles (of BN columns) per alpha group
  float* dot_part;      // [2*num_n_tiles][ld_part]
  float* ssq_part;
  long ld_part;
  // EPI_CORR, series tiles (N tiles >= series_tile0): the 256 columns of a tile are 64 time points x 4 series
  // terms, laid out [half][q][32 time points]; every epilogue thread then holds T_q[t] for q = 0..3 and 32 time
  // points and emits the 4 sums T_q y and the 10 sums T_q T_q' (q <= q') into series_part[part*14 + j][ld_part].
  int series_tile0;
  float* series_part;
  // Batched launches (EPI_STORE): `batch` equally shaped problems; operand b lives at base + b * stride (the
  // strides of A and B are part of their 3-D tensor maps), D / D_lo at + b * bs_d floats, Cin at + b * bs_c.
  int batch;
  long bs_d, bs_c;
  // tri_k: B is upper triangular in the sense B[n][k] == 0 for k < n (rows of an inverse Cholesky factor):
  // the K loop of a tile starts at its first B row instead of 0.
  int tri_k;
  // Grouped launches (EPI_STORE): A and D are single matrices whose row tiles (of BM * CG rows) belong to groups;
  // row tile mt multiplies B operand number tile_group[mt] (B is a stack of n_groups matrices whose stride is part of
  // its 3-D tensor map); tile_group[mt] < 0: nothing to do for that row tile.
  const int* tile_group;
};

// F16_ = 0: operands are fp32 planes holding TF32 values (kind::tf32, 8 values of K per MMA);
// F16_ = 1: operands are fp16 planes (kind::f16, 16 values of K per MMA, twice the TF32 rate).  A k-block is one
// 128-byte swizzle span in both cases, so tiles, descriptors and the K offsets inside a span are byte-identical.
template <int BN_, int CG_, int F16_ = 0>
struct GemmShape {
  static constexpr int BM = 128;  // accumulator rows per CTA (= TMEM lanes)
  static constexpr int BN = BN_;  // accumulator columns
  static constexpr int ELEM = F16_ ? 2 : 4;
  static constexpr int BK = 128 / ELEM;  // operand values per k-block = one 128-byte swizzle span
  static constexpr int CG = CG_;  // CTAs cooperating on one MMA (cta_group)
  static constexpr int UMMA_K = 32 / ELEM;
  static constexpr int B_ROWS = BN / CG;  // rows of B staged by each CTA
  static constexpr int A_BYTES = BM * 128;
  static constexpr int B_BYTES = B_ROWS * 128;
  static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);
  static constexpr int SMEM_BUDGET = 227 * 1024 - 2048;  // tiles; barriers + alignment slack live in the rest
  static constexpr int STAGES_RAW = SMEM_BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int TMEM_COLS = 2 * BN;  // double-buffered chunk accumulator
  static constexpr int STAGING_BYTES = 8 * 32 * 32 * 4;  // store epilogue: one 32 x 32 fp32 block per epilogue warp
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + STAGING_BYTES;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static constexpr int EPI_WARPS = 8;
  static constexpr int THREADS = 128 + EPI_WARPS * 32;
  static constexpr int COLS = BN / 2;  // accumulator columns owned by one epilogue thread
  static_assert(STAGES >= 2, "need at least a double-buffered smem ring");
  static_assert(TMEM_COLS == 64 || TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM cols pow2");
  static_assert(BN % 64 == 0 && BN <= 256, "BN");
};

__device__ __forceinline__ void tile_coords(const GemmParams& p, int tile, int& mt, int& nt) {
  const int gsz = p.group_m * p.num_n_tiles;
  const int g = tile / gsz;
  const int first_m = g * p.group_m;
  const int gm = min(p.num_m_tiles - first_m, p.group_m);
  const int r = tile - g * gsz;
  mt = first_m + r % gm;
  nt = r / gm;
}

// Batched tile index -> (batch, m tile, n tile); first k-block of the tile (tri_k: see GemmParams).
// Returns the index of the B operand (= b for batched launches, the row tile's group for grouped ones; < 0: skip).
template <int BN, int BK>
__device__ __forceinline__ int batch_tile_coords(const GemmParams& p, int tile, int& b, int& mt, int& nt, int& kb_begin) {
  const int per = p.num_m_tiles * p.num_n_tiles;
  b = tile / per;
  tile_coords(p, tile - b * per, mt, nt);
  kb_begin = p.tri_k ? min((nt * BN) / BK, p.num_k_blocks - 1) : 0;
  return p.tile_group ? __ldg(p.tile_group + mt) : b;
}

template <int BN, int CG, int EPI, int F16 = 0>
__global__ void __launch_bounds__(GemmShape<BN, CG, F16>::THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                   const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                   const GemmParams p) {
  using S = GemmShape<BN, CG, F16>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment (same offset in both CTAs of a pair).
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::STAGES * S::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + S::STAGES;
  uint64_t* tmem_full_bar = empty_bar + S::STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? ptx::cluster_ctarank() : 0u;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmAh);
    ptx::prefetch_tensormap(&tmAl);
    ptx::prefetch_tensormap(&tmBh);
    ptx::prefetch_tensormap(&tmBl);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S::STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full_bar[a], 1);
      ptx::mbar_init(&tmem_empty_bar[a], CG * S::EPI_WARPS);  // one arrive per epilogue warp of every CTA
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<CG>(tmem_ptr_smem, S::TMEM_COLS);
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int num_tiles = p.num_m_tiles * p.num_n_tiles * p.batch;
  const int worker = blockIdx.x / CG;
  const int num_workers = gridDim.x / CG;

  if (warp < 4) {
    // The data-movement / issue warps need few registers; hand the rest to the accumulating warps.
    ptx::setmaxnreg_dec<40>();
    if (warp == 0) {
      // ======================= TMA producer =======================
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = worker; tile < num_tiles; tile += num_workers) {
          int b, mt, nt, kb_begin;
          const int bb = batch_tile_coords<BN, S::BK>(p, tile, b, mt, nt, kb_begin);
          if (bb < 0) continue;
          const int m_row = (mt * CG + (int)cta_rank) * S::BM;
          const int n_row = nt * BN + (int)cta_rank * S::B_ROWS;
          for (int kb = kb_begin; kb < p.num_k_blocks; ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* st = smem + stage * S::STAGE_BYTES;
            const int k0 = kb * S::BK;
            if constexpr (CG == 1) {
              ptx::mbar_arrive_expect_tx(&full_bar[stage], S::STAGE_BYTES);
              ptx::tma_load_3d(st, &tmAh, &full_bar[stage], k0, m_row, b);
              ptx::tma_load_3d(st + S::A_BYTES, &tmAl, &full_bar[stage], k0, m_row, b);
              ptx::tma_load_3d(st + 2 * S::A_BYTES, &tmBh, &full_bar[stage], k0, n_row, bb);
              ptx::tma_load_3d(st + 2 * S::A_BYTES + S::B_BYTES, &tmBl, &full_bar[stage], k0, n_row, bb);
            } else {
              // Both CTAs load their halves; all bytes are accounted on the leader's barrier.
              if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * S::STAGE_BYTES);
              ptx::tma_load_3d_2sm(st, &tmAh, &full_bar[stage], k0, m_row, b);
              ptx::tma_load_3d_2sm(st + S::A_BYTES, &tmAl, &full_bar[stage], k0, m_row, b);
              ptx::tma_load_3d_2sm(st + 2 * S::A_BYTES, &tmBh, &full_bar[stage], k0, n_row, bb);
              ptx::tma_load_3d_2sm(st + 2 * S::A_BYTES + S::B_BYTES, &tmBl, &full_bar[stage], k0, n_row, bb);
            }
            if (++stage == S::STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    } else if (warp == 1) {
      // ======================= MMA issuer =======================
      if (lane == 0 && cta_rank == 0) {
        constexpr uint32_t idesc = F16 ? ptx::umma_idesc_f16(S::BM * CG, BN) : ptx::umma_idesc_tf32(S::BM * CG, BN);
        int stage = 0;
        uint32_t phase = 0;
        uint32_t cc = 0;  // running chunk counter: TMEM buffer = cc & 1
        for (int tile = worker; tile < num_tiles; tile += num_workers) {
          int b, mt, nt, kb_begin;
          if (batch_tile_coords<BN, S::BK>(p, tile, b, mt, nt, kb_begin) < 0) continue;
          for (int kb0 = kb_begin; kb0 < p.num_k_blocks; kb0 += p.kc_blocks, ++cc) {
            const uint32_t buf = cc & 1u;
            ptx::mbar_wait(&tmem_empty_bar[buf], ((cc >> 1) & 1u) ^ 1u);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * BN;
            const int kb1 = min(kb0 + p.kc_blocks, p.num_k_blocks);
            for (int kb = kb0; kb < kb1; ++kb) {
              ptx::mbar_wait(&full_bar[stage], phase);
              ptx::tc_fence_after();
              const uint32_t st = ptx::smem_u32(smem + stage * S::STAGE_BYTES);
              const uint32_t a_hi = st, a_lo = st + S::A_BYTES;
              const uint32_t b_hi = st + 2 * S::A_BYTES, b_lo = b_hi + S::B_BYTES;
#pragma unroll
              for (int ks = 0; ks < S::BK / S::UMMA_K; ++ks) {
                const uint32_t koff = ks * S::UMMA_K * S::ELEM;  // bytes along K inside the swizzle span (32 per step)
                const uint64_t dah = ptx::umma_desc_k_sw128(a_hi + koff);
                const uint64_t dal = ptx::umma_desc_k_sw128(a_lo + koff);
                const uint64_t dbh = ptx::umma_desc_k_sw128(b_hi + koff);
                const uint64_t dbl = ptx::umma_desc_k_sw128(b_lo + koff);
                ptx::umma_split<CG, F16>(d_tmem, dal, dbh, idesc, (uint32_t)((kb != kb0) | (ks != 0)));
                ptx::umma_split<CG, F16>(d_tmem, dah, dbl, idesc, 1u);
                ptx::umma_split<CG, F16>(d_tmem, dah, dbh, idesc, 1u);
              }
              if constexpr (CG == 1)
                ptx::umma_commit(&empty_bar[stage]);
              else
                ptx::umma_commit_2sm_mc(&empty_bar[stage], 0b11);
              if (++stage == S::STAGES) {
                stage = 0;
                phase ^= 1;
              }
            }
            if constexpr (CG == 1)
              ptx::umma_commit(&tmem_full_bar[buf]);
            else
              ptx::umma_commit_2sm_mc(&tmem_full_bar[buf], 0b11);
          }
        }
      }
    }
  } else {
    // ======================= accumulate + epilogue =======================
    ptx::setmaxnreg_inc<216>();
    const int ew = warp - 4;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access (warp id % 4)
    const int half = ew >> 2;   // column half of the tile owned by this warp
    constexpr int COLS = S::COLS;
    uint32_t cc = 0;
    float acc[COLS];
    for (int tile = worker; tile < num_tiles; tile += num_workers) {
      int b, mt, nt, kb_begin;
      if (batch_tile_coords<BN, S::BK>(p, tile, b, mt, nt, kb_begin) < 0) continue;
      const int num_chunks = (p.num_k_blocks - kb_begin + p.kc_blocks - 1) / p.kc_blocks;
      if constexpr (EPI == EPI_STORE) {
        // Cin of this tile towards L2 now: it is read only after the tile's MMAs, several microseconds from here
        if (p.Cin) {
          const long prow = (long)(mt * CG + (int)cta_rank) * S::BM + quad * 32 + lane;
          const long pcol = (long)nt * BN + half * COLS;
          if (prow < p.M) {
            const float* pc = p.Cin + b * p.bs_c + prow * p.ldc + pcol;
#pragma unroll
            for (int q = 0; q < COLS / 32; ++q)
              if (pcol + 32 * q < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(pc + 32 * q));
          }
        }
      }
#pragma unroll
      for (int i = 0; i < COLS; ++i) acc[i] = 0.f;
      for (int c = 0; c < num_chunks; ++c, ++cc) {
        const uint32_t buf = cc & 1u;
        ptx::mbar_wait(&tmem_full_bar[buf], (cc >> 1) & 1u);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * BN + half * COLS;
#pragma unroll
        for (int j = 0; j < COLS / 64; ++j) {
          float v0[32], v1[32];
          ptx::tmem_ld_32x32(taddr + j * 64, v0);
          ptx::tmem_ld_32x32(taddr + j * 64 + 32, v1);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            acc[j * 64 + i] += v0[i];  // round-to-nearest fp32 adds across K chunks
            acc[j * 64 + 32 + i] += v1[i];
          }
        }
        // Release this chunk buffer back to the MMA issuer (leader CTA).
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 1)
            ptx::mbar_arrive(&tmem_empty_bar[buf]);
          else
            ptx::mbar_arrive_cluster(&tmem_empty_bar[buf], 0);
        }
      }

      const long row = (long)(mt * CG + (int)cta_rank) * S::BM + quad * 32 + lane;
      const bool row_ok = row < p.M;
      if constexpr (EPI == EPI_STORE) {
        // The accumulator arrives one ROW per thread (TMEM lane = row).  Written that way, a warp store touches 32
        // rows with 16 bytes each: measured 1.4 TB/s of partial-sector writes, which bounded every small-K launch
        // (the trailing updates of the Cholesky solver).  So every 32 x 32 block of the warp's tile goes through a
        // 4 KB shared-memory staging buffer (XOR-swizzled 16-byte chunks: conflict-free both ways) and comes back
        // with 8 lanes per row: a warp load of Cin / store of D then moves four complete 128-byte row segments.
        // The loops over column blocks and row groups are deliberately NOT unrolled: the round-1 epilogue was
        // ~3,600 straight-line instructions per thread and tile, and ncu showed the small-K launches stalled on
        // instruction fetch (icache hit rate 68 %); only the staging stores need compile-time register indices.
        float* stg = reinterpret_cast<float*>(smem + S::STAGES * S::STAGE_BYTES + 256) + ew * 1024;
        const long row0 = (long)(mt * CG + (int)cta_rank) * S::BM + quad * 32;
        const int rsub = lane >> 3, chunk = lane & 7;
        const float sm_own = (p.scale_m && row_ok) ? __ldg(p.scale_m + row) : 1.f;
        const bool has_c = p.Cin != nullptr, has_lo = p.D_lo != nullptr, has_d = p.D != nullptr, has_h = p.H_hi != nullptr;
        const float os_own = (has_h && row_ok) ? __ldg(p.out_scale + row) : 1.f;
        const float* cb = has_c ? p.Cin + b * p.bs_c : nullptr;
        float* db = has_d ? p.D + b * p.bs_d : nullptr;
        float* lb = has_lo ? p.D_lo + b * p.bs_d : nullptr;
        const uint32_t my_sts = ptx::smem_u32(stg + lane * 32);
#pragma unroll 1
        for (int cc = 0; cc < COLS / 32; ++cc) {
#pragma unroll
          for (int c2 = 0; c2 < COLS / 32; ++c2) {
            if (c2 == cc) {
#pragma unroll
              for (int c = 0; c < 8; ++c)
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(my_sts + (((c ^ (lane & 7)) << 4))),
                             "f"(acc[c2 * 32 + 4 * c]), "f"(acc[c2 * 32 + 4 * c + 1]), "f"(acc[c2 * 32 + 4 * c + 2]),
                             "f"(acc[c2 * 32 + 4 * c + 3])
                             : "memory");
            }
          }
          __syncwarp();
          const long col = (long)nt * BN + half * COLS + cc * 32 + chunk * 4;
          const bool col_ok = col < p.N, vec_ok = col + 3 < p.N;
          float sn[4] = {1.f, 1.f, 1.f, 1.f};
          if (p.scale_n) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (col + q < p.N) sn[q] = __ldg(p.scale_n + col + q);
          }
          {
            constexpr int i0 = 0;
            // all Cin rows of the block are loaded before its first store (D may alias Cin: in-place updates;
            // 8 independent 128-byte row segments in flight per quarter warp)
            float4 cv[8];
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) {
              const long rw = row0 + 4 * (i0 + ii) + rsub;
              cv[ii] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (has_c && rw < p.M && col_ok) {
                const float* cp = cb + rw * p.ldc + col;
                if (vec_ok) {
                  cv[ii] = *reinterpret_cast<const float4*>(cp);
                } else {
                  cv[ii].x = cp[0];
                  if (col + 1 < p.N) cv[ii].y = cp[1];
                  if (col + 2 < p.N) cv[ii].z = cp[2];
                }
              }
            }
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) {
              const int r = 4 * (i0 + ii) + rsub;
              const long rw = row0 + r;
              const float am = p.alpha * __shfl_sync(0xffffffffu, sm_own, r);
              const float os = __shfl_sync(0xffffffffu, os_own, r);
              const float4 v = *reinterpret_cast<const float4*>(stg + r * 32 + ((chunk ^ (r & 7)) << 2));
              if (rw >= p.M || !col_ok) continue;
              float h[4] = {am * sn[0] * v.x + p.beta * cv[ii].x, am * sn[1] * v.y + p.beta * cv[ii].y,
                            am * sn[2] * v.z + p.beta * cv[ii].z, am * sn[3] * v.w + p.beta * cv[ii].w};
              if (has_h) {
                __align__(8) __half ph[4];
                __align__(8) __half pl[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float y = h[q] * os;
                  ph[q] = __float2half_rn(y);
                  pl[q] = __float2half_rn(y - __half2float(ph[q]));
                }
                __half* hp = p.H_hi + rw * p.ldh + col;
                __half* lp = p.H_lo + rw * p.ldh + col;
                if (vec_ok) {
                  *reinterpret_cast<uint2*>(hp) = *reinterpret_cast<const uint2*>(ph);
                  *reinterpret_cast<uint2*>(lp) = *reinterpret_cast<const uint2*>(pl);
                } else {
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    if (col + q < p.N) {
                      hp[q] = ph[q];
                      lp[q] = pl[q];
                    }
                  }
                }
                if (!has_d) continue;
              }
              float l[4] = {0.f, 0.f, 0.f, 0.f};
              if (has_lo) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float o = h[q];
                  h[q] = ptx::to_tf32(o);
                  l[q] = ptx::to_tf32(o - h[q]);
                }
              }
              float* dp = db + rw * p.ldd + col;
              if (vec_ok) {
                *reinterpret_cast<float4*>(dp) = make_float4(h[0], h[1], h[2], h[3]);
                if (has_lo) *reinterpret_cast<float4*>(lb + rw * p.ldd + col) = make_float4(l[0], l[1], l[2], l[3]);
              } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  if (col + q < p.N) {
                    dp[q] = h[q];
                    if (has_lo) lb[rw * p.ldd + col + q] = l[q];
                  }
                }
              }
            }
          }
          __syncwarp();  // the staging buffer is rewritten by the next column block
        }
      } else {
        // Fused per-voxel reduction: rows = voxels, columns = time points of one alpha group.
        if (row_ok && nt >= p.series_tile0) {
          if constexpr (COLS == 128) {
            const long part = (long)(nt - p.series_tile0) * 2 + half;
            const float* ycol = p.Yz + row + part * 32 * p.ldy;  // time points part*32 .. part*32 + 31
            float y[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) y[i] = __ldg(ycol + (long)i * p.ldy);
            float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
            float s00 = 0.f, s01 = 0.f, s02 = 0.f, s03 = 0.f, s11 = 0.f, s12 = 0.f, s13 = 0.f, s22 = 0.f, s23 = 0.f,
                  s33 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float a0 = acc[i], a1 = acc[32 + i], a2 = acc[64 + i], a3 = acc[96 + i];
              d0 = fmaf(a0, y[i], d0);
              d1 = fmaf(a1, y[i], d1);
              d2 = fmaf(a2, y[i], d2);
              d3 = fmaf(a3, y[i], d3);
              s00 = fmaf(a0, a0, s00);
              s01 = fmaf(a0, a1, s01);
              s02 = fmaf(a0, a2, s02);
              s03 = fmaf(a0, a3, s03);
              s11 = fmaf(a1, a1, s11);
              s12 = fmaf(a1, a2, s12);
              s13 = fmaf(a1, a3, s13);
              s22 = fmaf(a2, a2, s22);
              s23 = fmaf(a2, a3, s23);
              s33 = fmaf(a3, a3, s33);
            }
            float* o = p.series_part + part * 14 * p.ld_part + row;
            const float vals[14] = {d0, d1, d2, d3, s00, s01, s02, s03, s11, s12, s13, s22, s23, s33};
#pragma unroll
            for (int j = 0; j < 14; ++j) o[(long)j * p.ld_part] = vals[j];
          }
        } else if (row_ok) {
          float dot = 0.f, ssq = 0.f;
          const long t_base = (long)(nt % p.tiles_per_group) * BN + half * COLS;
          const float* ycol = p.Yz + row + t_base * p.ldy;
#pragma unroll
          for (int j = 0; j < COLS; j += 32) {
            float y[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) y[i] = __ldg(ycol + (long)(j + i) * p.ldy);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              dot = fmaf(acc[j + i], y[i], dot);
              ssq = fmaf(acc[j + i], acc[j + i], ssq);
            }
          }
          const long part = (long)nt * 2 + half;
          p.dot_part[part * p.ld_part + row] = dot;
          p.ssq_part[part * p.ld_part + row] = ssq;
        }
      }
    }
  }

  // ======================= teardown =======================
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) ptx::cluster_sync_all();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<CG>(tmem_base, S::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &ptr, 12000, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  });
  return fn;
}

// Tensor map of a row-major [rows][k] fp32 matrix (row pitch ld floats); box = box_rows x 32 floats.
static int make_operand_map(CUtensorMap* tm, const void* base, long rows, long k, long ld, int box_rows, int f16 = 0,
                            int batch = 1, long batch_stride = 0) {
  auto enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    return LIT_ERR_CUDA;
  }
  LIT_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "GEMM operand base must be 16-byte aligned");
  const int elem = f16 ? 2 : 4;
  LIT_REQUIRE((ld * elem) % 16 == 0 && ld >= k, "GEMM operand pitch must be a multiple of 16 bytes and >= K (ld=%ld K=%ld)",
              ld, k);
  if (batch <= 1 || batch_stride <= 0) batch_stride = (rows > 0 ? rows : 1) * ld;  // a single matrix: any valid stride
  LIT_REQUIRE((batch_stride * elem) % 16 == 0, "GEMM batch stride must be a multiple of 16 bytes");
  cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)rows, (cuuint64_t)(batch > 0 ? batch : 1)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * elem, (cuuint64_t)batch_stride * elem};
  cuuint32_t box[3] = {(cuuint32_t)(128 / elem), (cuuint32_t)box_rows, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = enc(tm, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base),
                   dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%ld k=%ld ld=%ld box_rows=%d)", (int)r, rows, k, ld,
              box_rows);
    return LIT_ERR_CUDA;
  }
  return LIT_OK;
}

// K chunk (in 32-wide k-blocks) accumulated inside TMEM between register drains.  4 (K = 128) keeps the
// truncation bias of the tensor-core accumulator below fp32 rounding noise; LIT_GEMM_KC_BLOCKS overrides it
// (development knob: a huge value reproduces plain in-TMEM accumulation).
static int kc_blocks_default() {
  static int v = [] {
    const char* e = getenv("LIT_GEMM_KC_BLOCKS");
    int x = e ? atoi(e) : 4;
    return x > 0 ? x : 4;
  }();
  return v;
}

// Rasterisation: consecutive tiles walk `group_m` M tiles (of BM*CG rows) before advancing along N, so that a
// wave of concurrently running tiles shares few A and B strips.  LIT_GEMM_GROUP_M overrides (development knob).
static int group_m_default(int cg) {
  static int v = [] {
    const char* e = getenv("LIT_GEMM_GROUP_M");
    return e ? atoi(e) : 0;
  }();
  return v > 0 ? v : 24 / cg;  // 12 CTA pairs: least DRAM traffic (19.6 GB vs 27.4 at 8, 22.5 at 16) and 1-3 % faster on the
                               // config-2 fused GEMM (profiles/r2_group_m_sweep.md)
}

static int g_sm_limit = [] {  // 0 = use every SM; LIT_GEMM_SM_LIMIT presets it (development knob)
  const char* e = getenv("LIT_GEMM_SM_LIMIT");
  return e ? atoi(e) : 0;
}();

template <int BN, int CG, int EPI, int F16 = 0>
static int launch_gemm(const void* A_hi, const void* A_lo, long lda, const void* B_hi, const void* B_lo, long ldb,
                       GemmParams p, cudaStream_t stream, long bs_a = 0, long bs_b = 0, int n_groups = 0) {
  using S = GemmShape<BN, CG, F16>;
  CUtensorMap tmAh, tmAl, tmBh, tmBl;
  int rc;
  if (p.batch < 1) p.batch = 1;
  const int b_count = n_groups > 0 ? n_groups : p.batch;  // grouped: B is a stack of n_groups matrices, A / D are single
  if ((rc = make_operand_map(&tmAh, A_hi, p.M, p.K, lda, S::BM, F16, p.batch, bs_a))) return rc;
  if ((rc = make_operand_map(&tmAl, A_lo, p.M, p.K, lda, S::BM, F16, p.batch, bs_a))) return rc;
  if ((rc = make_operand_map(&tmBh, B_hi, p.N, p.K, ldb, S::B_ROWS, F16, b_count, bs_b))) return rc;
  if ((rc = make_operand_map(&tmBl, B_lo, p.N, p.K, ldb, S::B_ROWS, F16, b_count, bs_b))) return rc;

  p.num_m_tiles = (p.M + S::BM * CG - 1) / (S::BM * CG);
  p.num_n_tiles = (p.N + BN - 1) / BN;
  p.num_k_blocks = (p.K + S::BK - 1) / S::BK;
  if (p.num_k_blocks < 1) p.num_k_blocks = 1;  // K == 0 still zero-initialises the accumulator via OOB fill
  if (p.group_m <= 0) p.group_m = group_m_default(CG);
  if (p.kc_blocks <= 0) p.kc_blocks = kc_blocks_default();  // 4 k-blocks: K = 128 (tf32) / 256 (fp16), 48 accumulations either way
  const int tiles = p.num_m_tiles * p.num_n_tiles * p.batch;
  if (tiles == 0) return LIT_OK;

  auto kfn = gemm_tf32x3_kernel<BN, CG, EPI, F16>;
  static bool attr_set = false;  // per template instantiation
  if (!attr_set) {
    LIT_CUDA_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM_BYTES));
    attr_set = true;
  }
  int sms = sm_count();
  if (g_sm_limit > 0 && g_sm_limit < sms) sms = g_sm_limit;
  int workers = sms / CG;
  if (workers < 1) workers = 1;
  if (workers > tiles) workers = tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(workers * CG);
  cfg.blockDim = dim3(S::THREADS);
  cfg.dynamicSmemBytes = S::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LIT_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kfn, tmAh, tmAl, tmBh, tmBl, p));
  return LIT_OK;
}

}  // namespace lit

using namespace lit;

extern "C" int lit_gemm_set_sm_limit(int n_sms) {
  LIT_REQUIRE(n_sms >= 0, "gemm_set_sm_limit: negative SM count");
  g_sm_limit = n_sms;
  return LIT_OK;
}

extern "C" int lit_gemm_tf32x3_nt(const float* A_hi, const float* A_lo, long lda, const float* B_hi,
                                  const float* B_lo, long ldb, int M, int N, int K, float alpha, const float* Cin,
                                  long ldc, float beta, float* D, float* D_lo, long ldd, int variant, void* stream) {
  LIT_REQUIRE(M >= 0 && N >= 0 && K >= 0, "negative GEMM extent");
  LIT_REQUIRE(ldd % 4 == 0 && ldd >= N, "output pitch must be a multiple of 4 floats and >= N");
  LIT_REQUIRE((reinterpret_cast<uintptr_t>(D) & 15) == 0, "output must be 16-byte aligned");
  LIT_REQUIRE(!Cin || (ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(Cin) & 15) == 0), "Cin alignment");
  LIT_REQUIRE(!D_lo || (reinterpret_cast<uintptr_t>(D_lo) & 15) == 0, "D_lo alignment");
  if (M == 0 || N == 0) return LIT_OK;
  GemmParams p = {};
  p.M = M;
  p.N = N;
  p.K = K;
  p.D = D;
  p.D_lo = D_lo;
  p.ldd = ldd;
  p.Cin = Cin;
  p.ldc = ldc;
  p.alpha = alpha;
  p.beta = beta;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // AUTO: the CTA-pair kernel (256 x 256 tile, B halves shared through the pair) is ~10 % faster
  // whenever there is more than one 128-row block of A to pair up.
  if (variant == LIT_GEMM_AUTO) variant = M > 128 ? LIT_GEMM_2CTA_N256 : LIT_GEMM_1CTA_N256;
  switch (variant) {
    case LIT_GEMM_1CTA_N256:
      return launch_gemm<256, 1, EPI_STORE>(A_hi, A_lo, lda, B_hi, B_lo, ldb, p, s);
    case LIT_GEMM_1CTA_N128:
      return launch_gemm<128, 1, EPI_STORE>(A_hi, A_lo, lda, B_hi, B_lo, ldb, p, s);
    case LIT_GEMM_2CTA_N256:
      return launch_gemm<256, 2, EPI_STORE>(A_hi, A_lo, lda, B_hi, B_lo, ldb, p, s);
    default:
      set_error("unknown GEMM variant %d", variant);
      return LIT_ERR_INVALID;
  }
}

// `batch` equally shaped products D_b = alpha * A_b B_b^T + beta * Cin_b in ONE launch (3xTF32): operand b lives at
// base + b * bs_x floats.  Used by the batched Cholesky inner solver (chol_solver.cu) and the stacked Neumann powers.
namespace lit {
int gemm_nt_batched(const float* A_hi, const float* A_lo, long lda, long bs_a, const float* B_hi, const float* B_lo,
                    long ldb, long bs_b, int M, int N, int K, float alpha, const float* Cin, long ldc, long bs_c,
                    float beta, float* D, float* D_lo, long ldd, long bs_d, int batch, int tri_k, cudaStream_t s) {
  LIT_REQUIRE(M >= 0 && N >= 0 && K >= 0 && batch >= 0, "negative GEMM extent");
  LIT_REQUIRE(ldd % 4 == 0 && ldd >= N && bs_d % 4 == 0, "output pitch / batch stride must be multiples of 4 floats");
  LIT_REQUIRE((reinterpret_cast<uintptr_t>(D) & 15) == 0, "output must be 16-byte aligned");
  LIT_REQUIRE(!Cin || (ldc % 4 == 0 && bs_c % 4 == 0 && (reinterpret_cast<uintptr_t>(Cin) & 15) == 0), "Cin alignment");
  LIT_REQUIRE(!D_lo || (reinterpret_cast<uintptr_t>(D_lo) & 15) == 0, "D_lo alignment");
  LIT_REQUIRE(!tri_k || N <= K + 255, "tri_k needs a square-ish B (N <= K)");
  if (M == 0 || N == 0 || batch == 0) return LIT_OK;
  GemmParams p = {};
  p.M = M;
  p.N = N;
  p.K = K;
  p.D = D;
  p.D_lo = D_lo;
  p.ldd = ldd;
  p.Cin = Cin;
  p.ldc = ldc;
  p.alpha = alpha;
  p.beta = beta;
  p.batch = batch;
  p.bs_d = bs_d;
  p.bs_c = bs_c;
  p.tri_k = tri_k;
  if (N <= 128) return launch_gemm<128, 1, EPI_STORE>(A_hi, A_lo, lda, B_hi, B_lo, ldb, p, s, bs_a, bs_b);  // panels
  if (M > 128) return launch_gemm<256, 2, EPI_STORE>(A_hi, A_lo, lda, B_hi, B_lo, ldb, p, s, bs_a, bs_b);
  return launch_gemm<256, 1, EPI_STORE>(A_hi, A_lo, lda, B_hi, B_lo, ldb, p, s, bs_a, bs_b);
}
}  // namespace lit

// Grouped product: D[rows of tile t] = A[rows of tile t] * B_{tile_group[t]}^T for the 256-row tiles of A (the
// voxels of the outer fit, sorted by their selected alpha; B_g = (G + a_g^2 I)^-1).  tile_group lives on the DEVICE,
// so no host synchronisation is needed to learn the group sizes; tiles with a negative group are skipped.
extern "C" int lit_gemm_tf32x3_nt_grouped(const float* A_hi, const float* A_lo, long lda, const float* B_hi,
                                          const float* B_lo, long ldb, long bs_b, int n_groups, int M, int N, int K,
                                          const int32_t* tile_group, float* D, float* D_lo, long ldd, void* stream) {
  LIT_REQUIRE(M >= 0 && N >= 0 && K >= 0 && n_groups >= 1, "gemm_grouped: bad extents");
  LIT_REQUIRE(M % 256 == 0, "gemm_grouped: the row count must be padded to a multiple of the 256-row tile");
  LIT_REQUIRE(tile_group != nullptr, "gemm_grouped: tile_group missing");
  LIT_REQUIRE(ldd % 4 == 0 && ldd >= N && (reinterpret_cast<uintptr_t>(D) & 15) == 0, "gemm_grouped: output alignment");
  LIT_REQUIRE(!D_lo || (reinterpret_cast<uintptr_t>(D_lo) & 15) == 0, "D_lo alignment");
  if (M == 0 || N == 0) return LIT_OK;
  GemmParams p = {};
  p.M = M;
  p.N = N;
  p.K = K;
  p.D = D;
  p.D_lo = D_lo;
  p.ldd = ldd;
  p.alpha = 1.f;
  p.batch = 1;
  p.tile_group = tile_group;
  return launch_gemm<256, 2, EPI_STORE>(A_hi, A_lo, lda, B_hi, B_lo, ldb, p, static_cast<cudaStream_t>(stream), 0, bs_b,
                                        n_groups);
}

extern "C" int lit_gemm_tf32x3_nt_batched(const float* A_hi, const float* A_lo, long lda, long bs_a, const float* B_hi,
                                          const float* B_lo, long ldb, long bs_b, int M, int N, int K, float alpha,
                                          const float* Cin, long ldc, long bs_c, float beta, float* D, float* D_lo,
                                          long ldd, long bs_d, int batch, int tri_k, void* stream) {
  return lit::gemm_nt_batched(A_hi, A_lo, lda, bs_a, B_hi, B_lo, ldb, bs_b, M, N, K, alpha, Cin, ldc, bs_c, beta, D, D_lo,
                              ldd, bs_d, batch, tri_k, static_cast<cudaStream_t>(stream));
}

// D = alpha * A B^T + beta * Cin on fp16 split pairs (lit_split_f16; one scale per row of A and per row of B, whose
// inverses inv_a / inv_b the epilogue multiplies back in): three kind::f16 MMAs per k-step, twice the rate of the
// 3xTF32 form at the same product accuracy.
static int gemm_f16x3(const void* A_hi, const void* A_lo, long lda, const void* B_hi, const void* B_lo, long ldb, int M,
                      int N, int K, float alpha, const float* Cin, long ldc, float beta, float* D, float* D_lo, long ldd,
                      const float* inv_a, const float* inv_b, const float* out_scale, void* H_hi, void* H_lo, long ldh,
                      int variant, void* stream) {
  LIT_REQUIRE(M >= 0 && N >= 0 && K >= 0, "negative GEMM extent");
  LIT_REQUIRE(D || H_hi, "GEMM without an output");
  LIT_REQUIRE(!D || (ldd % 4 == 0 && ldd >= N), "output pitch must be a multiple of 4 floats and >= N");
  LIT_REQUIRE((reinterpret_cast<uintptr_t>(D) & 15) == 0, "output must be 16-byte aligned");
  LIT_REQUIRE(!Cin || (ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(Cin) & 15) == 0), "Cin alignment");
  LIT_REQUIRE(!D_lo || (D && (reinterpret_cast<uintptr_t>(D_lo) & 15) == 0), "D_lo alignment");
  LIT_REQUIRE(!inv_b || (reinterpret_cast<uintptr_t>(inv_b) & 15) == 0, "inv_b must be 16-byte aligned");
  LIT_REQUIRE(!H_hi || (H_lo && out_scale && ldh % 4 == 0 && ldh >= N && (reinterpret_cast<uintptr_t>(H_hi) & 7) == 0 &&
                        (reinterpret_cast<uintptr_t>(H_lo) & 7) == 0),
              "fp16-pair output: both planes, the row scales, 8-byte alignment and a pitch that is a multiple of 4");
  if (M == 0 || N == 0) return LIT_OK;
  GemmParams p = {};
  p.M = M;
  p.N = N;
  p.K = K;
  p.D = D;
  p.D_lo = D_lo;
  p.ldd = ldd;
  p.Cin = Cin;
  p.ldc = ldc;
  p.alpha = alpha;
  p.beta = beta;
  p.scale_m = inv_a;
  p.scale_n = inv_b;
  p.H_hi = static_cast<__half*>(H_hi);
  p.H_lo = static_cast<__half*>(H_lo);
  p.ldh = ldh;
  p.out_scale = out_scale;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (variant == LIT_GEMM_AUTO) variant = M > 128 ? LIT_GEMM_2CTA_N256 : LIT_GEMM_1CTA_N256;
  switch (variant) {
    case LIT_GEMM_1CTA_N256:
      return launch_gemm<256, 1, EPI_STORE, 1>(A_hi, A_lo, lda, B_hi, B_lo, ldb, p, s);
    case LIT_GEMM_2CTA_N256:
      return launch_gemm<256, 2, EPI_STORE, 1>(A_hi, A_lo, lda, B_hi, B_lo, ldb, p, s);
    default:
      set_error("unknown f16x3 GEMM variant %d", variant);
      return LIT_ERR_INVALID;
  }
}

extern "C" int lit_gemm_f16x3_nt(const void* A_hi, const void* A_lo, long lda, const void* B_hi, const void* B_lo, long ldb,
                                 int M, int N, int K, float alpha, const float* Cin, long ldc, float beta, float* D,
                                 float* D_lo, long ldd, const float* inv_a, const float* inv_b, int variant,
                                 void* stream) {
  LIT_REQUIRE(D, "output missing");
  return gemm_f16x3(A_hi, A_lo, lda, B_hi, B_lo, ldb, M, N, K, alpha, Cin, ldc, beta, D, D_lo, ldd, inv_a, inv_b, nullptr,
                    nullptr, nullptr, 0, variant, stream);
}

// The same product with the result written as the scaled fp16 split pair (H_hi, H_lo; row scales out_scale) that the
// next fp16-pair GEMM consumes, and optionally (D != NULL) as fp32 too: the downdated cross product of an inner fold
// goes straight from this epilogue into the fused prediction + correlation GEMM without a lit_split_f16 pass.
extern "C" int lit_gemm_f16x3_nt_pairout(const void* A_hi, const void* A_lo, long lda, const void* B_hi, const void* B_lo,
                                         long ldb, int M, int N, int K, float alpha, const float* Cin, long ldc,
                                         float beta, float* D, long ldd, const float* inv_a, const float* inv_b,
                                         const float* out_scale, void* H_hi, void* H_lo, long ldh, int variant,
                                         void* stream) {
  LIT_REQUIRE(H_hi && H_lo && out_scale, "fp16-pair output missing");
  return gemm_f16x3(A_hi, A_lo, lda, B_hi, B_lo, ldb, M, N, K, alpha, Cin, ldc, beta, D, nullptr, ldd, inv_a, inv_b,
                    out_scale, H_hi, H_lo, ldh, variant, stream);
}

// Shared body of the fused prediction + correlation entry points.  f16 = 0: 3xTF32 split pairs; 1: fp16 split pairs.
static int corr_gemm(int f16, const void* A_hi, const void* A_lo, long lda, const void* B_hi, const void* B_lo, long ldb,
                     int M, int n_groups, int rows_per_group, int n_series_tiles, int K, const float* Yz, long ldy,
                     float* dot_part, float* ssq_part, float* series_part, long ld_part, int variant, void* stream) {
  LIT_REQUIRE(M >= 0 && n_groups >= 0 && rows_per_group >= 0 && K >= 0 && n_series_tiles >= 0, "negative extent");
  LIT_REQUIRE(rows_per_group % 256 == 0, "rows_per_group must be padded to a multiple of 256 (got %d)",
              rows_per_group);
  LIT_REQUIRE(ld_part >= M && ldy >= M, "partial / response pitch smaller than M");
  LIT_REQUIRE(n_series_tiles == 0 || series_part, "series tiles need the series_part output");
  const long n_plain = (long)n_groups * rows_per_group;
  if (M == 0 || n_plain + n_series_tiles == 0) return LIT_OK;
  GemmParams p = {};
  p.M = M;
  p.N = (int)(n_plain + 256L * n_series_tiles);
  p.K = K;
  p.Yz = Yz;
  p.ldy = ldy;
  p.tiles_per_group = rows_per_group >= 256 ? rows_per_group / 256 : 1;
  p.dot_part = dot_part;
  p.ssq_part = ssq_part;
  p.ld_part = ld_part;
  p.series_tile0 = (int)(n_plain / 256);
  p.series_part = series_part;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (variant == LIT_GEMM_AUTO) variant = M > 128 ? LIT_GEMM_2CTA_N256 : LIT_GEMM_1CTA_N256;
  switch (variant) {
    case LIT_GEMM_1CTA_N256:
      return f16 ? launch_gemm<256, 1, EPI_CORR, 1>(A_hi, A_lo, lda, B_hi, B_lo, ldb, p, s)
                 : launch_gemm<256, 1, EPI_CORR, 0>(A_hi, A_lo, lda, B_hi, B_lo, ldb, p, s);
    case LIT_GEMM_2CTA_N256:
      return f16 ? launch_gemm<256, 2, EPI_CORR, 1>(A_hi, A_lo, lda, B_hi, B_lo, ldb, p, s)
                 : launch_gemm<256, 2, EPI_CORR, 0>(A_hi, A_lo, lda, B_hi, B_lo, ldb, p, s);
    default:
      set_error("unknown corr-GEMM variant %d", variant);
      return LIT_ERR_INVALID;
  }
}

extern "C" int lit_gemm_tf32x3_nt_corr(const float* A_hi, const float* A_lo, long lda, const float* B_hi,
                                       const float* B_lo, long ldb, int M, int n_groups, int rows_per_group, int K,
                                       const float* Yz, long ldy, float* dot_part, float* ssq_part, long ld_part,
                                       int variant, void* stream) {
  return corr_gemm(0, A_hi, A_lo, lda, B_hi, B_lo, ldb, M, n_groups, rows_per_group, 0, K, Yz, ldy, dot_part, ssq_part,
                   nullptr, ld_part, variant, stream);
}

// Same fused prediction + correlation GEMM with fp16 split pairs (see lit_split_f16): hi = fp16(s x),
// lo = fp16(s x - hi) with a power-of-two scale s per row (A) / per 256-row tile (B), three kind::f16 MMAs per
// k-step.  The 11-bit significands make the products as exact as 3xTF32, at twice the tensor-core rate and half
// the operand bytes.  The partial sums are those of the SCALED predictions; lit_corr_finalize_scaled /
// lit_corr_finalize_series take the scale vectors to undo it.
extern "C" int lit_gemm_f16x3_nt_corr(const void* A_hi, const void* A_lo, long lda, const void* B_hi, const void* B_lo,
                                      long ldb, int M, int n_groups, int rows_per_group, int K, const float* Yz, long ldy,
                                      float* dot_part, float* ssq_part, long ld_part, int variant, void* stream) {
  return corr_gemm(1, A_hi, A_lo, lda, B_hi, B_lo, ldb, M, n_groups, rows_per_group, 0, K, Yz, ldy, dot_part, ssq_part,
                   nullptr, ld_part, variant, stream);
}

// The fused GEMM over a stack whose last n_series_tiles tiles carry the four terms of the Neumann series
// (see GemmParams::series_tile0 and lit_series_stack): precision 0 = 3xTF32 pairs, 1 = fp16 pairs.
extern "C" int lit_gemm_corr_series(int precision, const void* A_hi, const void* A_lo, long lda, const void* B_hi,
                                    const void* B_lo, long ldb, int M, int n_groups, int rows_per_group,
                                    int n_series_tiles, int K, const float* Yz, long ldy, float* dot_part,
                                    float* ssq_part, float* series_part, long ld_part, int variant, void* stream) {
  LIT_REQUIRE(precision == 0 || precision == 1, "corr_series: precision must be 0 (tf32x3) or 1 (f16x3)");
  return corr_gemm(precision, A_hi, A_lo, lda, B_hi, B_lo, ldb, M, n_groups, rows_per_group, n_series_tiles, K, Yz, ldy,
                   dot_part, ssq_part, series_part, ld_part, variant, stream);
}
