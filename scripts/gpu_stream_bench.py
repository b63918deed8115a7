"""Achieved HBM bandwidth of the streaming kernels at BASELINE config-2 sizes (CUDA events, inputs larger
than L2 or L2 flushed between repetitions).  Prints one JSON line per kernel: algorithmic bytes, ms, GB/s and
the fraction of the measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from litcoder_core_b200.device import DeviceOps, Mat, _vp, check

    ops = DeviceOps()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    peak = peaks["hbm_gbs"]
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device="cuda")  # 256 MB > 126 MB L2

    def timeit(name, nbytes, fn, reps=5):
        fn()
        ms = []
        for _ in range(reps):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        best = min(ms)
        print(json.dumps({"kernel": name, "algorithmic_MB": round(nbytes / 1e6, 1), "ms": round(best, 4),
                          "GBps": round(nbytes / best / 1e6, 1), "frac_of_measured_copy": round(nbytes / best / 1e6 / peak, 3)}),
              flush=True)

    N, p, V, n_o, n_v, A = 9400, 3072, 95000, 7520, 1500, 20
    Y = Mat(torch.randn((N, V), device="cuda"), None, N, V)
    X = Mat(torch.randn((N, p), device="cuda"), None, N, p)
    rng = np.random.default_rng(0)
    tr = ops.upload_index(np.sort(rng.permutation(N)[:n_o]))
    va = ops.upload_index(np.sort(rng.permutation(N)[:n_v]))
    timeit("gather_rows_transpose_split Y[7520 rows] -> (V x 7520) hi/lo", n_o * V * 4 * 3, lambda: ops.gather_rows_T_split(Y, tr, n_o))
    timeit("gather_rows_transpose_split Y[1500 rows] -> (V x 1500) hi/lo", n_v * V * 4 * 3, lambda: ops.gather_rows_T_split(Y, va, n_v))
    mean, std = ops.col_stats(Y, va, n_v, 1)
    timeit("col_stats Y[1500 rows] (mean, unbiased std)", n_v * V * 4, lambda: ops.col_stats(Y, va, n_v, 1))
    timeit("gather_normalize z-score Y[1500 rows] -> 1536 padded rows", n_v * V * 4 + 1536 * V * 4, lambda: ops.gather_normalize(Y, va, n_v, mean, std, 0, 1e-8, rows_out=1536))
    Ct = ops.empty(V, p)
    Ct.hi.normal_()
    timeit("split_tf32 C^T (V x 3072)", V * p * 4 * 3, lambda: ops.split(Ct))
    lam = torch.rand(p, device="cuda") + 0.1
    av = torch.full((V,), 10.0, device="cuda")
    Zs = ops.split(Ct)
    timeit("scale_rows_by_alpha Z^T (V x 3072) hi/lo -> hi/lo", V * p * 4 * 4, lambda: ops.scale_rows_by_alpha(Zs, lam, av, True, 1e-10))
    Wm = ops.zeros(V, p)
    timeit("axpy mean-weights += W^T/5 (V x 3072, split in)", V * p * 4 * 4, lambda: ops.axpy(0.2, Zs, Wm))
    timeit("transpose W^T (V x 3072) -> (3072 x V)", V * p * 4 * 2, lambda: ops.transpose(Wm))
    L = Mat(torch.randn((n_v, p), device="cuda"), None, n_v, p)
    al = ops.upload_vector(np.logspace(-1, 8, A), "f64")
    timeit("build_alpha_stack (1500 x 3072) -> (20 x 1536 x 3072) hi/lo", n_v * p * 4 + A * 1536 * p * 8, lambda: ops.build_alpha_stack(L, n_v, 1536, lam, al, A, True, 1e-10))
    corr = ops.empty(A, V)
    corr.hi.normal_()
    timeit("argmax_alpha (20 x V)", A * V * 4 + V * 8, lambda: ops.argmax_alpha(corr, 5, al.float(), False), reps=5)

    # round 2: operand re-split into fp16 pairs, the grouped outer fit's sort / gathers
    timeit("split_f16 C^T (V x 3072) tf32 pair -> fp16 pair + row scales", V * p * (8 * 2 + 4), lambda: ops.split_f16(Zs, 1))
    idx = ops.upload_vector(rng.integers(0, A, V), "i32")
    pos, perm, tg, cap = ops.group_plan(idx, V, A)
    timeit("group_plan: counting sort of V voxels by alpha index (1 block)", V * 4 * 3, lambda: ops.group_plan(idx, V, A))
    timeit("gather_rows C^T by alpha group (V x 3072 f32 -> hi/lo, padded)", V * p * 4 * 3, lambda: ops.gather_rows(Ct, perm, cap, split=True))

    # round 2, later: fp16 pairs written by their producers (bounds instead of passes over the data)
    timeit("gather_col_reduce |max| of Y over all 9400 rows (once per fit)", N * V * 4,
           lambda: ops.col_reduce(Y, None, N, sumsq=False, absmax=True))
    ysc = ops.f16_bound_scales(V, absmax=ops.col_reduce(Y, None, N, sumsq=False, absmax=True)[1])
    timeit("gather_col_reduce sum of squares of Y[1500 rows]", n_v * V * 4, lambda: ops.col_reduce(Y, va, n_v))
    timeit("gather_col_reduce sum of squares of X[1500 rows] (3072 columns)", n_v * p * 4, lambda: ops.col_reduce(X, va, n_v))
    timeit("row_absmax C^T (V x 3072)", V * p * 4, lambda: ops.row_absmax(Ct))
    timeit("gather_rows_transpose_f16 Y[1500 rows] -> (V x 1500) fp16 hi/lo", n_v * V * (4 + 4), lambda: ops.gather_rows_T_f16(Y, va, n_v, ysc))
    timeit("gather_rows_transpose_f16 Y[9400 rows] -> (V x 9400) fp16 hi/lo", N * V * (4 + 4), lambda: ops.gather_rows_T_f16(Y, ops.upload_index(np.arange(N)), N, ysc), reps=3)

    # feature construction: device-resident buffers through the C ABI (the API-level calls add H2D/D2H)
    nt, D = 9400, 768
    stim = torch.randn((nt, D), device="cuda")
    delays = ops.upload_vector(np.array([1, 2, 3, 4]), "i32")
    out = torch.empty((nt, 4 * D), dtype=torch.float64, device="cuda")
    timeit("fir_make_delayed 9400 x 768 f32 -> 9400 x 3072 f64", nt * D * 4 + nt * 4 * D * 8,
           lambda: check(ops.lib.lit_fir_make_delayed(_vp(stim.data_ptr()), 0, nt, D, D, _vp(delays.data_ptr()), 4, 0,
                                                      _vp(out.data_ptr()), 4 * D, _vp(ops.stream)), "fir"))
    n_s, n_tr = 28000, 9400  # all 25 stories' words in one call (the API runs one story at a time)
    words = torch.randn((n_s, D), device="cuda")
    wt = np.sort(rng.uniform(0, 2.0 * n_tr, n_s))
    trt = np.arange(n_tr) * 2.0 + 1.0
    from litcoder_core_b200.downsample import lanczos_band

    lo, hi = lanczos_band(wt, trt, 3, 0.5)
    d_wt, d_tr = ops.upload_vector(wt, "f64"), ops.upload_vector(trt, "f64")
    d_lo, d_hi = ops.upload_vector(lo, "i32"), ops.upload_vector(hi, "i32")
    lout = torch.empty((n_tr, D), dtype=torch.float64, device="cuda")
    timeit("lanczos_downsample 28000 words x 768 f32 -> 9400 TRs x 768 f64", n_s * D * 4 + n_tr * D * 8,
           lambda: check(ops.lib.lit_lanczos_downsample(_vp(words.data_ptr()), 0, n_s, D, D, _vp(d_wt.data_ptr()),
                                                        _vp(d_tr.data_ptr()), n_tr, 3.0, 0.5, 0, _vp(d_lo.data_ptr()),
                                                        _vp(d_hi.data_ptr()), _vp(lout.data_ptr()), D, _vp(ops.stream)),
                         "lanczos"))


if __name__ == "__main__":
    main()
