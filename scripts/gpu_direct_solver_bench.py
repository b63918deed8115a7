"""Times DeviceOps.solve_blocks_many on the BASELINE config-2 shape (p = 3072, 1500 validation rows, 4 solved alphas
per fold, 5 folds per outer fold).  Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel split."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from litcoder_core_b200.device import default_ops

    ops = default_ops()
    folds = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    p, m, n = 3072, 1500, 6000
    g = torch.Generator(device="cuda").manual_seed(0)
    jobs = []
    for f in range(folds):
        X = torch.randn((n, p), device="cuda", generator=g)
        X[1:] = 0.6 * X[:-1] + 0.8 * X[1:]
        G = (X.T @ X).contiguous()
        lmax = float(torch.linalg.eigvalsh(G.double())[-1])
        Pc = torch.randn((m, p), device="cuda", generator=g)
        Pc -= Pc.mean(0)
        a2 = [(a ** 2) * lmax for a in np.logspace(-1, 8, 20)]
        jobs.append(dict(G=ops.wrap(G), Pc=ops.wrap(Pc.contiguous()), n_rows=m, lam_max=lmax, a2=a2))
    for r in range(reps):
        ops.reset_counters()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ops.solve_blocks_many(jobs)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ops.check_solver()
        print(f"rep {r}: {dt * 1e3:.1f} ms for {folds} folds x 4 alphas; timed: "
              f"{ {k: round(v, 2) for k, v in ops.timings().items()} } residual {ops.last_solver_residual:.2e}", flush=True)


if __name__ == "__main__":
    main()
