"""Launch only the fused prediction + correlation GEMM at the BASELINE config-2 inner-fold shape, for
`ncu --set full` captures.  Prints the CUDA-event time per launch.

    python scripts/gpu_corr_gemm_only.py [reps] [--form compact|full] [--precision f16x3|tf32x3]

form "compact" (what a fit launches): 95,000 voxels x (4 solved alphas x 1,536 padded validation TRs + 24 series
tiles of 256 rows = 12,288 stacked rows), K = 3,072, through lit_gemm_corr_series.
form "full": one block of 1,536 rows for each of the 20 alphas (30,720 stacked rows).
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reps", nargs="?", type=int, default=3)
    ap.add_argument("--form", default="compact", choices=["compact", "full"])
    ap.add_argument("--precision", default="f16x3", choices=["f16x3", "tf32x3"])
    args = ap.parse_args()
    import torch

    from litcoder_core_b200.device import DeviceOps, Mat, SeriesStack

    ops = DeviceOps()
    M, G, R, K, n_va = 95000, 20, 1536, 3072, 1500
    A = ops.split(Mat(torch.randn((M, K), device="cuda"), None, M, K))
    if args.form == "compact":
        n_cheb, n_tiles = 4, -(-n_va // 64)
        rows = n_cheb * R + n_tiles * 256
        stack = SeriesStack(ops.split(Mat(torch.randn((rows, K), device="cuda"), None, rows, K)), n_cheb, R, n_tiles,
                            np.arange(n_cheb, dtype=np.int32), np.arange(n_cheb, G, dtype=np.int32),
                            np.ones((G - n_cheb, 4)))
    else:
        rows = G * R
        stack = ops.split(Mat(torch.randn((rows, K), device="cuda"), None, rows, K))
    Yz = Mat(torch.randn((R, M), device="cuda"), None, R, M)
    times = []
    for _ in range(args.reps):
        ops.reset_counters()
        ops.gemm_corr(A, stack, G, R, Yz, precision=args.precision)
        torch.cuda.synchronize()
        times.append(ops.corr_launches()[-1][0])  # the GEMM alone (the fp16 re-split of the operands is outside)
    flops = 2.0 * M * rows * K
    print(json.dumps({"form": args.form, "precision": args.precision, "M": M, "stacked_rows": rows, "K": K, "ms": times,
                      "algorithmic_tflops": [flops / t / 1e9 for t in times]}))


if __name__ == "__main__":
    main()
