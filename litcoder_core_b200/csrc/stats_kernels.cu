// Per-voxel test statistics on device: Pearson r finalisation + two-sided p-value
// (regularised incomplete beta in fp64), Benjamini-Hochberg FDR (hand-written bitonic sort),
// and Fisher's combined probability test.  Replaces the SciPy / statsmodels Python loops of
// nested_cv.py:418-438 (pearsonr), :441-477 (combine_pvalues) and the fdrcorrection calls at
// :158,263,282.
#include "common.cuh"
#include "../../include/litridge.h"

#include <cfloat>

namespace lit {

static inline int blocks_for(long items, int block) { return (int)((items + block - 1) / block); }

// ---------------------------------------------------------------------------------------------
// Regularised incomplete beta I_x(a, b), modified Lentz continued fraction (fp64).
// ---------------------------------------------------------------------------------------------
__device__ double betacf(double a, double b, double x) {
  const double TINY = 1e-300, EPS = 1e-16;
  const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < TINY) d = TINY;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m <= 2000; ++m) {
    const double m2 = 2.0 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < TINY) d = TINY;
    c = 1.0 + aa / c;
    if (fabs(c) < TINY) c = TINY;
    d = 1.0 / d;
    h *= d * c;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < TINY) d = TINY;
    c = 1.0 + aa / c;
    if (fabs(c) < TINY) c = TINY;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < EPS) break;
  }
  return h;
}

__device__ double betainc(double a, double b, double x) {
  if (x <= 0.0) return 0.0;
  if (x >= 1.0) return 1.0;
  const double lbt = lgamma(a + b) - lgamma(a) - lgamma(b) + a * log(x) + b * log1p(-x);
  const double bt = exp(lbt);
  if (x < (a + 1.0) / (a + b + 2.0)) return bt * betacf(a, b, x) / a;
  return 1.0 - bt * betacf(b, a, 1.0 - x) / b;
}

// Two-sided p-value of a Pearson r from n samples: 2 * sf of Beta(n/2-1, n/2-1) on [-1, 1]
// at |r| (scipy.stats.pearsonr), i.e. 2 * I_{(1-|r|)/2}(n/2-1, n/2-1).
__device__ double pearson_pvalue(double r, long n) {
  if (n < 3) return 1.0;
  const double ab = 0.5 * (double)n - 1.0;
  double ar = fabs(r);
  if (ar >= 1.0) return 0.0;
  double p = 2.0 * betainc(ab, ab, 0.5 * (1.0 - ar));
  if (p > 1.0) p = 1.0;
  return p;
}

__global__ void pearson_finalize_kernel(const float* __restrict__ dot_part, const float* __restrict__ ssq_part,
                                        long ld_part, int n_tiles, long n_vox, long n_samples, int p_round_f32,
                                        float* __restrict__ r_out, double* __restrict__ p_out) {
  const long v = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vox) return;
  float d = 0.f, q = 0.f;
  for (int t = 0; t < n_tiles; ++t) {
    d += dot_part[(long)t * ld_part + v];
    q += ssq_part[(long)t * ld_part + v];
  }
  float r = d / sqrtf(q);
  double p;
  if (isnan(r)) {  // constant prediction (or response): pearsonr -> NaN -> (0.0, 1.0)
    r = 0.f;
    p = 1.0;
  } else {
    r = fminf(1.f, fmaxf(-1.f, r));
    p = pearson_pvalue((double)r, n_samples);
    if (p_round_f32) p = (double)(float)p;
  }
  r_out[v] = r;
  p_out[v] = p;
}

// ---------------------------------------------------------------------------------------------
// Bitonic sort of (key, index) pairs, ascending by key then index.  n_pad is a power of two.
// ---------------------------------------------------------------------------------------------
struct KV {
  double key;
  int idx;
  int pad;
};
__device__ __forceinline__ bool kv_less(const KV& a, const KV& b) {
  return a.key < b.key || (a.key == b.key && a.idx < b.idx);
}

__global__ void bh_init_kernel(const double* __restrict__ p, long n, long n_pad, KV* __restrict__ kv) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  KV e;
  e.pad = 0;
  if (i < n) {
    double x = p[i];
    if (isnan(x)) x = INFINITY;  // NaNs sort last
    e.key = x;
    e.idx = (int)i;
  } else {
    e.key = INFINITY;
    e.idx = 0x7fffffff;
  }
  kv[i] = e;
}

// All compare-exchange steps with partner distance < 2*blockDim inside shared memory.
// mode 0: full sort of each 2*blockDim chunk (k = 2 .. chunk); mode 1: finish stage k (j = blockDim .. 1).
__global__ void bitonic_local_kernel(KV* __restrict__ kv, long n_pad, long k_stage, int mode) {
  extern __shared__ unsigned char sh_raw[];
  KV* sh = reinterpret_cast<KV*>(sh_raw);
  const int chunk = 2 * blockDim.x;
  const long base = (long)blockIdx.x * chunk;
  sh[threadIdx.x] = kv[base + threadIdx.x];
  sh[threadIdx.x + blockDim.x] = kv[base + threadIdx.x + blockDim.x];
  __syncthreads();
  const long k_lo = mode == 0 ? 2 : k_stage;
  const long k_hi = mode == 0 ? chunk : k_stage;
  for (long k = k_lo; k <= k_hi; k <<= 1) {
    for (int j = (int)((k >> 1) < blockDim.x ? (k >> 1) : blockDim.x); j > 0; j >>= 1) {
      const int t = threadIdx.x;
      const int lo = ((t / j) * 2 * j) + (t % j);
      const int hi = lo + j;
      const bool up = (((base + lo) & k) == 0);
      KV a = sh[lo], b = sh[hi];
      if (kv_less(b, a) == up) {
        sh[lo] = b;
        sh[hi] = a;
      }
      __syncthreads();
    }
  }
  kv[base + threadIdx.x] = sh[threadIdx.x];
  kv[base + threadIdx.x + blockDim.x] = sh[threadIdx.x + blockDim.x];
}

__global__ void bitonic_global_kernel(KV* __restrict__ kv, long n_pad, long k, long j) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pad / 2) return;
  const long lo = ((t / j) * 2 * j) + (t % j);
  const long hi = lo + j;
  const bool up = ((lo & k) == 0);
  KV a = kv[lo], b = kv[hi];
  if (kv_less(b, a) == up) {
    kv[lo] = b;
    kv[hi] = a;
  }
}

// Single block: BH step-up on the sorted p-values, adjusted p (reverse running min), scatter.
__global__ void bh_apply_kernel(const KV* __restrict__ kv, long n, double alpha, uint8_t* __restrict__ reject,
                                double* __restrict__ p_adj, int* __restrict__ count_out) {
  __shared__ double smin[1024];
  __shared__ long smax[1024];
  const int T = blockDim.x;
  const int t = threadIdx.x;
  const long per = (n + T - 1) / T;
  const long b = (long)t * per;
  long e = b + per;
  if (e > n) e = n;
  const double dn = (double)n;
  // pass 1: per-thread suffix minimum of p_(i) / (i/n) and the largest rejecting rank
  double lmin = INFINITY;
  long lmax = -1;
  for (long i = e - 1; i >= b; --i) {
    const double ecdf = (double)(i + 1) / dn;
    const double pv = kv[i].key;
    const double raw = pv / ecdf;
    if (raw < lmin) lmin = raw;
    if (pv <= ecdf * alpha && i > lmax) lmax = i;
  }
  smin[t] = lmin;
  smax[t] = lmax;
  __syncthreads();
  if (t == 0) {
    // exclusive suffix-min over thread chunks and global max (T <= 1024: trivial serial pass)
    double run = INFINITY;
    long gmax = -1;
    for (int q = T - 1; q >= 0; --q) {
      const double mine = smin[q];
      smin[q] = run;
      if (mine < run) run = mine;
      if (smax[q] > gmax) gmax = smax[q];
    }
    smax[0] = gmax;
    if (count_out) *count_out = (int)(gmax + 1);
  }
  __syncthreads();
  const long kmax = smax[0];
  double run = smin[t];
  for (long i = e - 1; i >= b; --i) {
    const double ecdf = (double)(i + 1) / dn;
    const double raw = kv[i].key / ecdf;
    if (raw < run) run = raw;
    const int o = kv[i].idx;
    p_adj[o] = run > 1.0 ? 1.0 : run;
    reject[o] = (i <= kmax) ? 1 : 0;
  }
}

__global__ void fisher_kernel(const double* __restrict__ p, long ld_p, int n_folds, long n_vox, int p_round_f32,
                              double* __restrict__ out) {
  const long v = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vox) return;
  bool all_one = true, any_zero = false;
  double x = 0.0;  // x = -sum log p = statistic / 2
  for (int f = 0; f < n_folds; ++f) {
    const double pv = p[(long)f * ld_p + v];
    if (pv != 1.0) all_one = false;
    if (pv <= 0.0) any_zero = true;
    x -= log(pv);
  }
  double r;
  if (all_one) {
    r = 1.0;
  } else if (any_zero) {
    r = 0.0;  // log(0) = -inf -> statistic = inf -> sf = 0
  } else {
    // chi2.sf(2x, 2K) = exp(-x) * sum_{j<K} x^j / j!
    double term = 1.0, sum = 1.0;
    for (int j = 1; j < n_folds; ++j) {
      term *= x / (double)j;
      sum += term;
    }
    r = exp(-x) * sum;
    if (r > 1.0) r = 1.0;
    if (p_round_f32) r = (double)(float)r;
  }
  out[v] = r;
}

static long next_pow2(long n) {
  long p = 1;
  while (p < n) p <<= 1;
  return p;
}

}  // namespace lit

using namespace lit;

extern "C" int lit_pearson_finalize(const float* dot_part, const float* ssq_part, long ld_part, int n_tiles, long n_vox,
                                    long n_samples, int p_round_f32, float* r, double* p, void* stream) {
  LIT_REQUIRE(ld_part >= n_vox && n_tiles > 0, "pearson_finalize: bad extents");
  if (n_vox == 0) return LIT_OK;
  pearson_finalize_kernel<<<blocks_for(n_vox, 128), 128, 0, (cudaStream_t)stream>>>(
      dot_part, ssq_part, ld_part, n_tiles, n_vox, n_samples, p_round_f32, r, p);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}

extern "C" int lit_bh_workspace(long n, size_t* bytes) {
  LIT_REQUIRE(n >= 0 && bytes, "bh_workspace: bad arguments");
  long np2 = next_pow2(n < 2048 ? 2048 : n);
  *bytes = (size_t)np2 * sizeof(KV);
  return LIT_OK;
}

extern "C" int lit_bh_fdr(const double* p, long n, double alpha_fdr, uint8_t* reject, double* p_adj, int* count_out,
                          void* work, size_t work_bytes, void* stream) {
  LIT_REQUIRE(n > 0, "bh_fdr: empty input");
  const long n_pad = next_pow2(n < 2048 ? 2048 : n);
  LIT_REQUIRE(work && work_bytes >= (size_t)n_pad * sizeof(KV), "bh_fdr: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  KV* kv = reinterpret_cast<KV*>(work);
  bh_init_kernel<<<blocks_for(n_pad, 256), 256, 0, s>>>(p, n, n_pad, kv);
  LIT_LAUNCH_CHECK();
  const int lt = 1024;  // threads per local block -> 2048-element chunks (32 KB of smem)
  const size_t lsmem = 2 * lt * sizeof(KV);
  const int lblocks = (int)(n_pad / (2 * lt));
  bitonic_local_kernel<<<lblocks, lt, lsmem, s>>>(kv, n_pad, 0, 0);
  LIT_LAUNCH_CHECK();
  for (long k = 4 * lt; k <= n_pad; k <<= 1) {
    for (long j = k >> 1; j > lt; j >>= 1) {
      bitonic_global_kernel<<<blocks_for(n_pad / 2, 256), 256, 0, s>>>(kv, n_pad, k, j);
      LIT_LAUNCH_CHECK();
    }
    bitonic_local_kernel<<<lblocks, lt, lsmem, s>>>(kv, n_pad, k, 1);
    LIT_LAUNCH_CHECK();
  }
  bh_apply_kernel<<<1, 1024, 0, s>>>(kv, n, alpha_fdr, reject, p_adj, count_out);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}

extern "C" int lit_fisher_combine(const double* p, long ld_p, int n_folds, long n_vox, int p_round_f32, double* p_out,
                                  void* stream) {
  LIT_REQUIRE(n_folds > 0 && ld_p >= n_vox, "fisher_combine: bad extents");
  if (n_vox == 0) return LIT_OK;
  fisher_kernel<<<blocks_for(n_vox, 256), 256, 0, (cudaStream_t)stream>>>(p, ld_p, n_folds, n_vox, p_round_f32, p_out);
  LIT_LAUNCH_CHECK();
  return LIT_OK;
}
