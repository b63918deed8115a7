/* litridge.h -- C ABI of the B200 (sm_100a) hot path of LITcoder's encoding models.
 *
 * The reference (GT-LIT-Lab/litcoder_core) is pure Python and has no FFI of its own; its
 * boundary for this path is three Python call signatures
 *     NestedCVModel.fit_predict      encoding/models/nested_cv.py:18-42
 *     Downsampler.downsample         encoding/downsample/downsampling.py:395-424
 *     FIR.make_delayed               encoding/features/FIR_expander.py:24-43
 * The Python package `litcoder_core_b200` mirrors those signatures and drives the entry
 * points below through ctypes.  Each entry point replaces one (group of) torch / NumPy /
 * SciPy call site(s) of the reference, cited next to it.
 *
 * Conventions
 *   - every function returns 0 on success or a negative errno-style code; the message is
 *     available from lit_last_error() (thread-local);
 *   - pointers are DEVICE pointers unless the parameter name ends in `_h`;
 *   - matrices are row-major fp32 with an explicit leading dimension (pitch, in elements);
 *   - `stream` is a cudaStream_t passed as void*; all work is asynchronous on that stream
 *     unless stated otherwise; nothing here allocates device memory behind the caller's
 *     back except lit_syevd's workspace query helper (sizes are reported, caller allocates);
 *   - a "split pair" (hi, lo) is the 3xTF32 representation of an fp32 matrix:
 *     hi = rna_tf32(x), lo = rna_tf32(x - hi); the two planes share one pitch.
 */
#ifndef LITRIDGE_H_
#define LITRIDGE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LITRIDGE_ABI_VERSION 6

const char* lit_last_error(void);
int lit_abi_version(void);
/* SM count, compute capability and memory of the current device. */
int lit_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* free_bytes, size_t* total_bytes);

/* ------------------------------------------------------------------------------------------
 * Dense contractions: 3xTF32 tcgen05 GEMMs
 * ---------------------------------------------------------------------------------------- */
enum lit_gemm_variant {
  LIT_GEMM_AUTO = 0,
  LIT_GEMM_1CTA_N256 = 1, /* 128x256 tile per CTA, cta_group::1 */
  LIT_GEMM_1CTA_N128 = 2, /* 128x128 tile per CTA, cta_group::1 */
  LIT_GEMM_2CTA_N256 = 3  /* 256x256 tile per CTA pair, cta_group::2 */
};

/* The GEMMs are persistent kernels with one CTA per SM.  n_sms > 0 restricts their grids to that many
 * SMs so that other work (the cuSOLVER eigendecompositions on a side stream) can run beside them;
 * 0 restores the default (every SM).  Process-wide setting. */
int lit_gemm_set_sm_limit(int n_sms);

/* D[M,N] = alpha * A[M,K] * B[N,K]^T + beta * Cin[M,N]   (Cin may be NULL).
 * A and B are split pairs.  If D_lo is non-NULL the result is written as a split pair
 * (D = hi, D_lo = lo) so it can feed the next GEMM directly.  K is accumulated in chunks of 128
 * inside the tensor core (whose fp32 accumulator truncates) and across chunks with round-to-nearest
 * fp32 adds in registers, which keeps the result at fp32-FMA accuracy for any K.
 * Replaces torch.matmul at ridge_regression.py:32,59-61,104,105 and nested_cv.py:151,251
 * (after the Gram/eigen reformulation described in DESIGN.md). */
int lit_gemm_tf32x3_nt(const float* A_hi, const float* A_lo, long lda, const float* B_hi, const float* B_lo, long ldb,
                       int M, int N, int K, float alpha, const float* Cin, long ldc, float beta, float* D, float* D_lo,
                       long ldd, int variant, void* stream);

/* Prediction GEMM with the correlation reduction fused into the epilogue.
 *   acc[v, g*R + t] = sum_k A[v,k] * B[g*R + t, k]       v < M (voxels), g < n_groups, t < R
 *   dot_part[part][v] = sum_{t in part} acc * Yz[t][v]      ssq_part[part][v] = sum_{t in part} acc^2
 * with R = rows_per_group (multiple of 256; pad rows of B and Yz must be zero) and
 * part = g*(R/128) + t/128 (each epilogue thread reduces 128 consecutive time points, so the part
 * arrays have n_groups * R / 128 rows of pitch ld_part).  Yz is [R][ldy] (time-major, voxel contiguous).
 * Replaces the per-alpha loop body of ridge_corr_torch (ridge_regression.py:115-125) and the
 * outer-test prediction + per-voxel Pearson loop (nested_cv.py:151,251,418-438). */
int lit_gemm_tf32x3_nt_corr(const float* A_hi, const float* A_lo, long lda, const float* B_hi, const float* B_lo,
                            long ldb, int M, int n_groups, int rows_per_group, int K, const float* Yz, long ldy,
                            float* dot_part, float* ssq_part, long ld_part, int variant, void* stream);

/* The same fused GEMM on fp16 split pairs (kind::f16 MMAs: 16 values of K per instruction, twice the TF32 rate,
 * half the operand bytes).  A_* / B_* are fp16 planes written by lit_split_f16 (pitches lda / ldb in fp16
 * elements, multiples of 8); the partial sums are those of the SCALED predictions
 * sA[v] * sB[g] * pred -- lit_corr_finalize_scaled undoes the scales.  Same reference lines as above. */
int lit_gemm_f16x3_nt_corr(const void* A_hi, const void* A_lo, long lda, const void* B_hi, const void* B_lo, long ldb,
                           int M, int n_groups, int rows_per_group, int K, const float* Yz, long ldy, float* dot_part,
                           float* ssq_part, long ld_part, int variant, void* stream);
/* D = alpha * A B^T + beta * Cin with the operands as fp16 split pairs (lit_split_f16 with rows_per_group = 1 for
 * both): inv_a[M] / inv_b[N, readable up to the next multiple of 4] are the inverse operand scales the epilogue
 * multiplies back in (NULL = ones).  Same output options as lit_gemm_tf32x3_nt; variants 1CTA_N256 / 2CTA_N256.
 * lda / ldb in fp16 elements (multiples of 8).  Used for the voxel-side products: U^T Rresp
 * (ridge_regression.py:32,104) as C^T = Y^T X and its fold downdates, the rotation into the eigenbasis and the
 * weights (ridge_regression.py:59-61). */
int lit_gemm_f16x3_nt(const void* A_hi, const void* A_lo, long lda, const void* B_hi, const void* B_lo, long ldb, int M,
                      int N, int K, float alpha, const float* Cin, long ldc, float beta, float* D, float* D_lo, long ldd,
                      const float* inv_a, const float* inv_b, int variant, void* stream);
/* The fused GEMM over a stack whose last n_series_tiles tiles (256 rows each) are SERIES tiles written by
 * lit_series_stack: 64 time points x the 4 terms Q_q of the Neumann series of (G + a^2 I)^-1.  For those tiles the
 * epilogue emits, per part of 32 time points and voxel, 14 sums into series_part[part*14 + j][ld_part]:
 * j = 0..3: sum_t T_q[t] Yz[t][v];  j = 4..13: sum_t T_q[t] T_q'[t] for (q,q') = 00 01 02 03 11 12 13 22 23 33,
 * T_q = Q_q A[v]^T.  Every alpha of the series is then a 4-term combination (lit_corr_finalize_series) instead of
 * its own block of stacked rows: 16 of the 20 BASELINE alphas cost 4 row blocks.  The first n_groups *
 * rows_per_group rows are ordinary alpha groups (dot_part / ssq_part as in lit_gemm_tf32x3_nt_corr).
 * precision: 0 = 3xTF32 split pairs (float planes), 1 = fp16 split pairs (lit_split_f16).
 * Replaces the per-alpha loop of ridge_corr_torch (ridge_regression.py:115-133) for all alphas of a fold at once. */
int lit_gemm_corr_series(int precision, const void* A_hi, const void* A_lo, long lda, const void* B_hi, const void* B_lo,
                         long ldb, int M, int n_groups, int rows_per_group, int n_series_tiles, int K, const float* Yz,
                         long ldy, float* dot_part, float* ssq_part, float* series_part, long ld_part, int variant,
                         void* stream);
/* fp16 split pair of x = src_hi (+ src_lo when non-NULL):  out_hi = fp16(s x), out_lo = fp16(s x - out_hi) with one
 * power-of-two scale s per group of rows_per_group consecutive rows (1 for the voxel rows of A, 256 = one N tile
 * for the stacked design B), chosen so that the group's largest magnitude lands in [2^14, 2^15).
 * inv_scale[g] = 1/s.  scratch: 8 bytes per group.  ld_out in fp16 elements.  (Operand format only: no reference
 * counterpart; the reference multiplies fp32 tensors, ridge_regression.py:32,104,120.) */
int lit_split_f16(const float* src_hi, const float* src_lo, long ld_src, long rows, long cols, long rows_per_group,
                  void* out_hi, void* out_lo, long ld_out, float* inv_scale, void* scratch, void* stream);

/* fp16 pairs written by the PRODUCER (no lit_split_f16 pass): the scale of a row comes from a bound known beforehand.
 *   lit_gather_col_reduce: out_sumsq[c] = sum_r src[idx[r]][c]^2, out_absmax[c] = max_r |src[idx[r]][c]| (either
 *     may be NULL; idx == NULL: rows 0..n_idx-1; negative index = zero row).
 *   lit_row_absmax: out[r] = max_c |src[r][c]|.
 *   lit_f16_bound_scales: bound[r] = absmax[r] + sqrt(row_sumsq[r] * max_j col_sumsq[j]) (absmax or the norm pair may
 *     be NULL); scale[r] = 2^k with bound * (1 + 2^-10) * scale in [2^14, 2^15), inv_scale[r] = 1 / scale[r]
 *     (both 1 for a zero or non-finite bound).  For the downdated cross product C_i^T = C_o^T - Y_R^T X_R the bound
 *     is max_j |C_o^T[v][j]| + |y_(v,R)|_2 max_j |x_(j,R)|_2 (Cauchy-Schwarz on the removed rows).
 *   lit_gather_rows_transpose_f16: dst[c][r] = fp16 pair of scale[c] * src[idx[r]][c], columns r in [n_idx, ld_dst)
 *     zero-filled (the gathered response rows of a fold as GEMM operand, scale[c] from the column maximum of |Y|).
 *   lit_gemm_f16x3_nt_pairout: lit_gemm_f16x3_nt whose epilogue writes H = fp16 pair of out_scale[row] * D (pitch ldh
 *     in fp16 elements, multiple of 4) and, when D != NULL, D itself.
 * (Operand format only, as lit_split_f16: the reference multiplies fp32 tensors, ridge_regression.py:32,104,120.) */
int lit_gather_col_reduce(const float* src, long ld_src, const int32_t* idx, long n_idx, long cols, float* out_sumsq,
                          float* out_absmax, void* stream);
int lit_row_absmax(const float* src, long ld_src, long rows, long cols, float* out, void* stream);
int lit_f16_bound_scales(const float* absmax, const float* row_sumsq, const float* col_sumsq, long n_cs, long rows,
                         float* scale, float* inv_scale, void* stream);
int lit_gather_rows_transpose_f16(const float* src, long ld_src, const int32_t* idx, long n_idx, long cols,
                                  const float* scale, void* dst_hi, void* dst_lo, long ld_dst, void* stream);
int lit_gemm_f16x3_nt_pairout(const void* A_hi, const void* A_lo, long lda, const void* B_hi, const void* B_lo, long ldb,
                              int M, int N, int K, float alpha, const float* Cin, long ldc, float beta, float* D,
                              long ldd, const float* inv_a, const float* inv_b, const float* out_scale, void* H_hi,
                              void* H_lo, long ldh, int variant, void* stream);

/* ------------------------------------------------------------------------------------------
 * Layout / conversion (bandwidth-bound streaming kernels)
 * ---------------------------------------------------------------------------------------- */
/* dst[i] = (float)src[i]   -- the torch.tensor(..., dtype=float32) of nested_cv.py:99-100. */
int lit_convert_f64_to_f32(const double* src, float* dst, size_t n, void* stream);
/* dst[i] = (double)src[i]. */
int lit_convert_f32_to_f64(const float* src, double* dst, size_t n, void* stream);
/* hi/lo TF32 split of a [rows][cols] matrix (hi or lo may alias src). */
int lit_split_tf32(const float* src, long rows, long cols, long ld_src, float* hi, float* lo, long ld_dst,
                   void* stream);
/* dst[c][r] = src[r][c]; if dst_lo != NULL the result is written as a split pair. */
int lit_transpose_f32(const float* src, long rows, long cols, long ld_src, float* dst, float* dst_lo, long ld_dst,
                      void* stream);
/* dst[i][:] = src[idx[i]][:] for i < n_idx, zero rows for n_idx <= i < n_rows_out.
 * idx == NULL means the identity (a padded copy); a negative idx[i] gives a zero row.  Optional split output.
 * Replaces the fancy-index gathers at nested_cv.py:200-201,371-374. */
int lit_gather_rows_f32(const float* src, long ld_src, const int32_t* idx, long n_idx, long cols, float* dst,
                        float* dst_lo, long ld_dst, long n_rows_out, void* stream);
/* dst[c][i] = src[idx[i]][c] for i < n_idx (zero for n_idx <= i < ld_dst): gather + transpose,
 * written as a split pair (or, with dst_lo == NULL, as one fp32 plane: operands that are re-split into fp16
 * pairs anyway); this is how the K-major (time-contiguous) operands X^T and Y^T of a fold's training rows
 * are produced. */
int lit_gather_rows_transpose_split(const float* src, long ld_src, const int32_t* idx, long n_idx, long cols,
                                    float* dst_hi, float* dst_lo, long ld_dst, void* stream);
/* y[i] += a * (x_hi[i] (+ x_lo[i])) over a [rows][cols] matrix (x_lo may be NULL). */
int lit_axpy_f32(float a, const float* x_hi, const float* x_lo, long ld_x, float* y, long ld_y, long rows, long cols,
                 void* stream);
int lit_fill_f32(float* dst, size_t n, float value, void* stream);
/* Pitched copy between host and device (cudaMemcpy2DAsync): kind 1 = host->device, 2 = device->host,
 * 3 = device->device.  Used to upload a column block (voxel shard) of a host matrix without a
 * host-side repack.  The host side of the copy may be pageable or pinned. */
int lit_memcpy_2d(void* dst, size_t dpitch_bytes, const void* src, size_t spitch_bytes, size_t width_bytes,
                  size_t height, int kind, void* stream);
/* 0 = pageable host memory, 1 = page-locked host memory, 2 = device / managed memory.  Pageable inputs of fit_predict
 * (what np.vstack hands a drop-in caller; the reference copies them with torch.tensor at nested_cv.py:99-100) are
 * staged through page-locked buffers by a background thread so that their upload overlaps the design-side work. */
int lit_host_pointer_kind(const void* ptr);

/* ------------------------------------------------------------------------------------------
 * Column statistics / normalisation  (ridge_utils.py:6-15 z_score, :70-180 DataNormalizer)
 * ---------------------------------------------------------------------------------------- */
/* Per-column mean and std over the gathered rows idx (NULL = all n rows).  ddof = 1 matches
 * torch.std (unbiased), ddof = 0 NumPy.  Accumulation is in fp64; outputs fp32. */
int lit_col_stats(const float* src, long ld_src, const int32_t* idx, long n_idx, long cols, int ddof, float* mean,
                  float* std, double* scratch /* 2*cols doubles */, void* stream);
/* dst[i][c] = (src[idx[i]][c] - mean[c]) * scale(c), zero rows up to n_rows_out, where
 *   mode 0: scale = 1 / (std[c] + eps)                  (z_score, eps = 1e-8)
 *   mode 1: scale = 1 / (std[c] * sqrt(n_idx - 1))      (unit-norm centred column; NaN if std == 0)
 *   mode 2: scale = 1                                   (centre only)
 *   mode 3: scale = 1 / std[c] if std[c] != 0 else 1    (zs / zscore of the trainers, encoding/utils.py:23-34)
 * Optional split output. */
int lit_gather_normalize_rows(const float* src, long ld_src, const int32_t* idx, long n_idx, long cols,
                              const float* mean, const float* std, int mode, float eps, float* dst, float* dst_lo,
                              long ld_dst, long n_rows_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Eigen-decomposition of the p x p Gram (the only library call: cuSOLVER syevd)
 * replaces torch.linalg.svd in svd_wrapper (ridge_utils.py:49-67).
 * ---------------------------------------------------------------------------------------- */
/* Workspace size in bytes (device and host) for lit_syevd at order n.
 * dtype: 0 = f32, 1 = f64.  batch > 1 selects the batched solver (matrices n*n apart). */
int lit_syevd_workspace(int n, int dtype, int batch, size_t* device_bytes, size_t* host_bytes);
/* In place: on exit row j of G (pitch ld >= n) is the j-th eigenvector, lam ascending.
 * info is a device int per matrix (0 on success). work / work_h (host) from lit_syevd_workspace. */
int lit_syevd(void* G, int n, long ld, int dtype, int batch, void* lam, void* work, size_t work_bytes, void* work_h,
              size_t work_h_bytes, int* info, void* stream);

/* ------------------------------------------------------------------------------------------
 * Ridge-specific elementwise kernels
 * ---------------------------------------------------------------------------------------- */
/* Stacked, column-centred, alpha-weighted validation design in the eigenbasis:
 *   Lst[a*rows_pad + t][j] = (L[t][j] - mean_t L[:,j]) * keep_j / (lam_j + (alpha_a * s)^2)
 * s = sqrt(max(lam)) if normalpha else 1;  keep_j = sqrt(max(lam_j,0)) > singcutoff;
 * rows t >= n_rows are zero.  Output is a split pair of shape [n_alphas*rows_pad][k].
 * (ridge_regression.py:97-101,117-120 with D = S/(S^2+a^2) folded into the Gram form.) */
int lit_build_alpha_stack(const float* L, long ld_l, long n_rows, long rows_pad, int k, const float* lam,
                          const double* alphas, int n_alphas, int normalpha, float singcutoff,
                          float* col_mean /* k floats, scratch */, double* scratch /* 2*k doubles */, float* out_hi,
                          float* out_lo, long ld_out, void* stream);
/* Per-voxel shrinkage in the eigenbasis: out[v][j] = (Z_hi+Z_lo)[v][j] * keep_j / (lam_j + (alpha_v * s)^2)
 * written as a split pair (ridge_regression.py:56-61 for every voxel at once). */
int lit_scale_rows_by_alpha(const float* Z_hi, const float* Z_lo, long ld_z, long n_vox, int k, const float* lam,
                            const float* alpha_v, int normalpha, float singcutoff, float* out_hi, float* out_lo,
                            long ld_out, void* stream);
/* Inner-CV score from the fused partials (ridge_regression.py:122-133):
 *   metric 0: corr[a][v] = nan_to_num( (sum_tiles dot / n_rows) / (sqrt(sum_tiles ssq / (n_rows-1)) + eps) )
 *             (Yz z-scored with the unbiased std + eps)
 *   metric 1: signed sqrt of R^2 = 1 - var(Q - pred)/var(Q) (Yz centred only; resp_std = unbiased std of Q)
 *   metric | 2: the same without the nan_to_num scrub (ridge_corr_pred_torch, ridge_regression.py:203-214)
 * accumulate != 0 adds into corr (fold sum for nested_cv.py:391-393). */
int lit_corr_finalize(const float* dot_part, const float* ssq_part, long ld_part, int parts_per_group, int n_groups,
                      long n_vox, long n_rows, float eps, int accumulate, int metric, const float* resp_std,
                      float* corr, long ld_corr, void* stream);
/* lit_corr_finalize with the operand scales of lit_gemm_f16x3_nt_corr undone and an optional slot map: the partial
 * sums of part P (128 stacked rows) are multiplied by inv_row[v] * inv_tile[P / 2] (and its square) before they are
 * added (either vector may be NULL = ones; lit_split_f16 with rows_per_group = 256 writes inv_tile); group g is
 * written to corr row slots[g] (NULL: row g). */
int lit_corr_finalize_scaled(const float* dot_part, const float* ssq_part, long ld_part, int parts_per_group,
                             int n_groups, long n_vox, long n_rows, float eps, int accumulate, int metric,
                             const float* resp_std, const float* inv_row, const float* inv_tile, const int32_t* slots,
                             float* corr, long ld_corr, void* stream);
/* Scores of the alphas served by the truncated Neumann series from the 14 per-voxel sums of the series tiles of
 * lit_gemm_corr_series: prediction_a[t] = sum_q coef[a*4+q] T_q[t], so dot_a = sum_q coef_q D_q and
 * ssq_a = sum_{q,q'} coef_q coef_q' S_qq' (fp64), then the lit_corr_finalize formula; alpha a goes to corr row
 * slots[a].  n_parts = 2 * n_series_tiles; inv_tile indexes the SERIES tiles (NULL = ones). */
int lit_corr_finalize_series(const float* series_part, long ld_part, int n_parts, long n_vox, long n_rows, float eps,
                             int accumulate, int metric, const float* resp_std, const float* inv_row,
                             const float* inv_tile, const double* coef, const int32_t* slots, int n_alphas, float* corr,
                             long ld_corr, void* stream);
/* best[v] = first argmax_a mean[a][v], mean = corr_sum / n_folds (nested_cv.py:391-393,408-411);
 * alpha_out[v] = (float)alphas[best[v]].  col_sums (n_alphas doubles, may be NULL) receives
 * sum_v mean[a][v] for the single_alpha rule (nested_cv.py:396-400). */
int lit_argmax_alpha(const float* corr_sum, long ld_corr, int n_alphas, long n_vox, int n_folds, const float* alphas,
                     int32_t* best, float* alpha_out, double* col_sums, void* stream);

/* ------------------------------------------------------------------------------------------
 * Eigendecomposition-free inner-fold solver: M_a = P_c (G + a^2 I)^-1 with GEMMs only
 * (same quantity as PVh * D of ridge_regression.py:105,117-120, rotated back out of the eigenbasis)
 * ---------------------------------------------------------------------------------------- */
/* lambda_max of the symmetric positive semi-definite G (n x n) by `steps` steps of the three-term Lanczos
 * recurrence and a Sturm bisection on the resulting tridiagonal matrix (all on the device).
 * vec_scratch: 3*n floats; scal_scratch: 2*steps + 4 doubles.  Either output may be NULL.
 * Replaces S[0] = largest singular value used by normalpha (ridge_regression.py:39,97). */
int lit_lanczos_lambda_max(const float* G, long ld, int n, int steps, float* vec_scratch, double* scal_scratch,
                           float* lam_out_f32, double* lam_out_f64, void* stream);
/* The same for `batch` matrices of one size at once (one launch per Lanczos step for all of them): G is a HOST
 * array of device pointers (all n x n, pitch ld); vec_scratch: batch * 3*n floats; scal_scratch:
 * batch * (2*steps + 4) doubles; lam_out_f64: batch doubles.  steps <= n. */
int lit_lanczos_lambda_max_batched(const float* const* G, int batch, long ld, int n, int steps, float* vec_scratch,
                                   double* scal_scratch, double* lam_out_f64, void* stream);
/* `batch` equally shaped products D_b = alpha * A_b B_b^T + beta * Cin_b in ONE launch (3xTF32 split pairs):
 * operand b lives at base + b * bs_x floats (bs_a, bs_b, bs_c, bs_d; 16-byte multiples).  tri_k != 0: B is upper
 * triangular in the sense B[n][k] == 0 for k < n (the rows of an inverse Cholesky factor), so the K loop of a tile
 * starts at its first row of B.  Serves the batched direct solver below; same arithmetic as lit_gemm_tf32x3_nt. */
int lit_gemm_tf32x3_nt_batched(const float* A_hi, const float* A_lo, long lda, long bs_a, const float* B_hi,
                               const float* B_lo, long ldb, long bs_b, int M, int N, int K, float alpha, const float* Cin,
                               long ldc, long bs_c, float beta, float* D, float* D_lo, long ldd, long bs_d, int batch,
                               int tri_k, void* stream);
/* Grouped product for the eigendecomposition-free outer fit:  D[rows of tile t] = A[rows of tile t] * B_g^T  with
 * g = tile_group[t] (DEVICE int32 per 256-row tile of A; negative: skip), B a stack of n_groups (N x K) split pairs at
 * stride bs_b floats.  A holds the cross products X^T y of the voxels SORTED by their selected alpha (lit_group_plan),
 * B_g = (X^T X + a_g^2 I)^-1: the weights of ridge_torch (ridge_regression.py:52-61: one dense product per unique
 * alpha over the voxels that selected it) without an eigendecomposition and without a host round trip for the
 * group sizes.  M must be a multiple of 256. */
int lit_gemm_tf32x3_nt_grouped(const float* A_hi, const float* A_lo, long lda, const float* B_hi, const float* B_lo,
                               long ldb, long bs_b, int n_groups, int M, int N, int K, const int32_t* tile_group,
                               float* D, float* D_lo, long ldd, void* stream);
/* Deterministic counting sort of n_vox voxels by idx[v] in [0, n_groups <= 32) into a layout whose groups are padded to
 * multiples of `tile` rows: pos[v] = grouped row of voxel v; perm[s] = voxel at grouped row s or -1 (s < rows_cap);
 * tile_group[t] = group of row tile t or -1 (t < rows_cap / tile).  rows_cap >= round_up(n_vox, tile) + n_groups * tile.
 * Replaces the torch.unique / boolean-mask selection of ridge_regression.py:52-58. */
int lit_group_plan(const int32_t* idx, long n_vox, int n_groups, int tile, long rows_cap, int32_t* pos, int32_t* perm,
                   int32_t* tile_group, void* stream);
/* Batched direct solve of the small-alpha systems of the inner folds:  M_b = R_b (G_b + a2_b I)^-1  for up to 128
 * systems of one size at once (G_b: n x n symmetric, pitch ldg; R_b: m_b x n, pitch ldr, m_b <= mp), by a blocked
 * Cholesky factorisation of the augmented matrix [G + a2 I; R; I] whose panel and trailing updates are batched
 * tcgen05 GEMMs (chol_solver.cu).  G_h, R_h, m_h, a2_h are HOST arrays of length nb.
 * lit_spd_solve_workspace reports the sizes: F (fp32 work matrix) f_floats, S_hi / S_lo (its TF32 planes) s_floats
 * each, Dg_hi / Dg_lo dg_floats each; pitch ldw (= n rounded up to 128) and rows_total (= 2 ldw + mp) per system.
 * On return system b (at + b * rows_total * ldw floats) holds  Y_b = R_b L^-T  in rows [ldw, ldw + mp)  and
 * W_b = L^-T (upper triangular)  in rows [ldw + mp, rows_total)  of S, so that
 *     M_b = Y_b W_b^T = lit_gemm_tf32x3_nt_batched(A = Y, B = W, K = N = n, tri_k = 1).
 * info[b] (device ints) != 0: system b was not positive definite.
 * Replaces the per-alpha `(PVh * D) @ UR` factors of ridge_regression.py:115-120 for the alphas that do not ride
 * the Neumann series (round 1 solved them with ~66 Chebyshev GEMM steps per fold). */
int lit_spd_solve_workspace(int nb, int n, int mp, size_t* f_floats, size_t* s_floats, size_t* dg_floats, long* ldw,
                            long* rows_total);
int lit_spd_solve_batched(int nb, int n, int mp, const void* const* G_h, long ldg, const void* const* R_h, long ldr,
                          const int* m_h, const float* a2_h, float* F, float* S_hi, float* S_lo, float* Dg_hi,
                          float* Dg_lo, int* info, void* stream);
/* A-posteriori check of such solutions on one deterministic +-1 probe vector x per system:
 *   numden[2 b] = || M_b (G_b x + a2_b x) - R_b x ||^2,  numden[2 b + 1] = || R_b x ||^2   (M_b at M + b * m_stride,
 * pitch ldm); scratch: nb * n floats.  Read back with the fit's final synchronisation (DeviceOps.check_solver). */
int lit_spd_probe_residual(int nb, int n, const void* const* G_h, long ldg, const void* const* R_h, long ldr,
                           const int* m_h, const float* a2_h, const float* M, long ldm, long m_stride, float* scratch,
                           double* numden, void* stream);
/* One Chebyshev step on (rows x cols) row-matrices of pitch ld:
 *   d = c1*d + c2*r (also written as the split pair d_hi/d_lo),  x += d,  t = r - a2*d;
 * `first` != 0 treats d and x as zero on input.  The caller then forms r = t - d G with lit_gemm_tf32x3_nt. */
int lit_cheb_update(float* d, const float* r, float* x, float* t, float* d_hi, float* d_lo, long ld, long rows, long cols,
                    float c1, float c2, float a2, int first, void* stream);
/* out[slots[g]*rows_pad + t][:] = sum_{q<n_src} coef[g*4+q] * (src_hi[q] + src_lo[q])[t][:]  (t < rows; pad rows 0),
 * written as a split pair: the truncated Neumann series of (G + a^2 I)^-1 for alphas far above the spectrum.
 * src_hi / src_lo are HOST arrays of n_src (<= 4) device pointers (src_lo may be NULL or hold NULLs). */
int lit_poly_combine(const float* const* src_hi, const float* const* src_lo, int n_src, long ld_src, long rows,
                     long rows_pad, long cols, const double* coef, const int32_t* slots, int n_groups, float* out_hi,
                     float* out_lo, long ld_out, void* stream);
/* Series tiles of the alpha stack: out row tile*256 + half*128 + q*32 + i = scale[q] * (src_hi[q] + src_lo[q])[t],
 * t = tile*64 + half*32 + i (zero for t >= rows), as a split pair; q = 0..3.  src_hi / src_lo: HOST arrays of 4
 * device pointers (src_lo may be NULL or hold NULLs); scale: 4 HOST doubles (lambda_max^-q). */
int lit_series_stack(const float* const* src_hi, const float* const* src_lo, long ld_src, long rows, long cols,
                     const double* scale, long n_tiles, float* out_hi, float* out_lo, long ld_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Test statistics (nested_cv.py:418-477, statsmodels fdrcorrection)
 * ---------------------------------------------------------------------------------------- */
/* r[v] = clip(sum_tiles dot / sqrt(sum_tiles ssq), -1, 1) (Yz must hold unit-norm centred columns),
 * NaN -> r = 0, p = 1; p = two-sided Student-t / Beta(n/2-1, n/2-1) p-value of r with n samples,
 * evaluated in fp64; if p_round_f32 != 0 it is rounded through fp32 (SciPy >= 1.14 keeps the
 * input dtype, so the reference run in this image produces fp32 p-values). */
int lit_pearson_finalize(const float* dot_part, const float* ssq_part, long ld_part, int n_parts, long n_vox,
                         long n_samples, int p_round_f32, float* r, double* p, void* stream);
/* Benjamini-Hochberg: reject[v] (uint8), p_adj[v], *count_out (device int).  Scratch: see
 * lit_bh_workspace.  Sorting is a hand-written bitonic network (keys padded to a power of two). */
int lit_bh_workspace(long n, size_t* bytes);
int lit_bh_fdr(const double* p, long n, double alpha_fdr, uint8_t* reject, double* p_adj, int* count_out, void* work,
               size_t work_bytes, void* stream);
/* Fisher's method across n_folds p-value vectors (p is [n_folds][ld_p]); all-ones -> 1.0;
 * p_round_f32 as above. */
int lit_fisher_combine(const double* p, long ld_p, int n_folds, long n_vox, int p_round_f32, double* p_out,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * Feature construction
 * ---------------------------------------------------------------------------------------- */
/* FIR delay stacking (FIR_expander.py:24-43): out[t][i*ndim + c] = stim[t - delays[i]][c] or 0
 * (circpad: index mod nt).  dtype_in: 0 = f32, 1 = f64.  out is [nt][ndelays*ndim] f64, pitch ld_out. */
int lit_fir_make_delayed(const void* stim, int dtype_in, long nt, long ndim, long ld_stim, const int32_t* delays,
                         int ndelays, int circpad, double* out, long ld_out, void* stream);
/* FIR delay stacking fused with the trainers' per-story structuring (trainer.py:203-209,236-239,250-253;
 * encoding/utils.py:23-29): rows [row_start, row_stop) of the delayed matrix of one story, column z-scored with the
 * population std over those rows (a zero-std column is centred only), NaN -> 0, written as fp32 to out
 * [row_stop - row_start][ndelays*ndim] (pitch ld_out) -- typically a row block of the design matrix X.
 * zscore = 0 copies the trimmed delayed rows unchanged (concatenated mode, trainer.py:264-282). */
int lit_fir_zscore_rows(const void* stim, int dtype_in, long nt, long ndim, long ld_stim, const int32_t* delays,
                        int ndelays, int circpad, long row_start, long row_stop, int zscore, float* out, long ld_out,
                        void* stream);
/* Lanczos resampling (interpdata.py:45-63,87-126): out = W * data, W[i][j] = lanczos((tr_i - t_j) * cutoff),
 * cutoff = cutoff_mult / mean(diff(tr_times)) computed by the caller.  rectify -> out is [n_tr][2*ndim]
 * (negative part | positive part).  lo/hi (n_tr + 1 int32, or NULL) bound the contributing samples of
 * each TR when data_times is sorted: TR i only reads samples [lo[i], hi[i]).  Without them the kernel
 * is the dense product (every sample is multiplied, zero weights included, so non-finite samples
 * propagate as in np.dot). */
int lit_lanczos_downsample(const void* data, int dtype_in, long n_samples, long ndim, long ld_data,
                           const double* data_times, const double* tr_times, long n_tr, double window, double cutoff,
                           int rectify, const int32_t* lo, const int32_t* hi, double* out, long ld_out, void* stream);

/* Sinc resampling (interpdata.sincfun / sincinterp2D, :29-84): W[i][j] = 2B sin(2 pi B t)/(2 pi B t + 1e-20),
 * t = tr_i - t_j, B = cutoff; zero for |t| > window / (2B) and, if causal, for t < 0; if renorm each row is
 * divided by its sum unless that sum is exactly 0.  lo / hi as for Lanczos. */
int lit_sinc_downsample(const void* data, int dtype_in, long n_samples, long ndim, long ld_data,
                        const double* data_times, const double* tr_times, long n_tr, double window, double cutoff,
                        int causal, int renorm, const int32_t* lo, const int32_t* hi, double* out, long ld_out,
                        void* stream);
/* Membership (CSR) row reduction: out[r][:] = sum_{e in [row_ptr[r], row_ptr[r+1])} w[e] * data[col_idx[e]][:]
 * (w == NULL: 1; mean != 0: divided by the entry count); rows without entries are zero.  Device form of the
 * rect / average / sum / last / legacy_* downsamplers (downsampling.py:24-319): the host builds only the
 * integer membership lists. */
int lit_csr_rows_apply(const void* data, int dtype_in, long ndim, long ld_data, const int32_t* row_ptr,
                       const int32_t* col_idx, const double* weights, long n_rows_out, int mean, double* out,
                       long ld_out, void* stream);
/* Gabor transform magnitude (interpdata.gabor_xfm2D, :129-145; downsampling.py:159-166):
 * out[i][d * n_freq + f] = | sum_j exp(-0.5 (t_j - tr_i)^2 / (2 sigma^2)) data[j][d] exp(i 2 pi f t_j) |. */
int lit_gabor_downsample(const void* data, int dtype_in, long n_samples, long ndim, long ld_data,
                         const double* data_times, const double* tr_times, long n_tr, const double* freqs, int n_freq,
                         double sigma, double* out, long ld_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LITRIDGE_H_ */
