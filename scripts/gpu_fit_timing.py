"""Phase timing of one nested-CV ridge fit at a BASELINE config on a B200 (development aid).

    python scripts/gpu_fit_timing.py --config 2 [--voxels 95000] [--no-overlap] [--no-downdate]

Inputs are synthetic and resident in HBM (generated with torch -- data generation is not part of
the product path).  Prints one JSON object with per-category CUDA-event milliseconds.
"""
import argparse
import json
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CONFIGS = {1: (9400, 4, 95000), 2: (9400, 3072, 95000), 3: (9400, 5120, 95000), 4: (2226, 3072, 81924),
           5: (9400, 16384, 95000), 6: (9400, 10240, 95000)}  # 6 = the "~2.5k x 4 delays" reading of config 3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--voxels", type=int, default=None)
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--no-downdate", action="store_true")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--train-test", action="store_true")
    args = ap.parse_args()
    import torch

    import litcoder_core_b200 as L
    from litcoder_core_b200 import engine as E

    N, p, V = CONFIGS[args.config]
    V = args.voxels or V
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn((N, p), device="cuda", generator=g)
    for j in range(1, min(p, 8)):
        X[:, j] = 0.6 * X[:, j - 1] + 0.8 * X[:, j]
    pw = min(p, 3072)  # signal lives in the first pw features (keeps the generator's scratch small)
    W = torch.randn((pw, V), device="cuda", generator=g) / pw ** 0.5
    W *= (torch.rand((1, V), device="cuda", generator=g) < 0.3)
    Y = X[:, :pw] @ W
    del W
    Y += 3.0 * torch.randn((N, V), device="cuda", generator=g)
    torch.cuda.synchronize()

    if args.no_overlap or args.no_downdate:
        orig = E.RidgeConfig.__init__

        def patched(self, *a, **k):
            orig(self, *a, **k)
            self.overlap_eig = not args.no_overlap
            self.downdate = not args.no_downdate

        E.RidgeConfig.__init__ = patched

    model = L.NestedCVModel("ridge_regression")
    alphas = np.logspace(-1, 8, 20)
    out = []
    for rep in range(args.reps):
        random.seed(0)
        t0 = time.perf_counter()
        if args.train_test:
            m, w, a = model.fit_predict(X[:7520], Y[:7520], X_test=X[7520:], y_test=Y[7520:], alphas=alphas)
        else:
            m, w, a = model.fit_predict(X, Y, alphas=alphas)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        out.append({"rep": rep, "wall_s": wall, "timings_ms": model.last_timings, "stats": model.last_stats,
                    "median_r": m["median_score"], "n_significant": m["n_significant"],
                    "mem_peak_gb": torch.cuda.max_memory_allocated() / 1e9})
        print(json.dumps(out[-1]), flush=True)


if __name__ == "__main__":
    main()
