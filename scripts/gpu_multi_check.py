"""Multi-GPU consistency check (run under torchrun on a multi-GPU box):
every rank runs the sharded fit; rank 0 also runs the same fit unsharded and compares.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/gpu_multi_check.py
"""
import json
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist

    import litcoder_core_b200 as L
    from litcoder_core_b200.engine import SingleProcess

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(3)
    N, p, V = 2400, 512, 5000
    X = rng.standard_normal((N, p)).astype(np.float32)
    for j in range(1, p):
        X[:, j] = 0.6 * X[:, j - 1] + 0.8 * X[:, j]
    W = (rng.standard_normal((p, V)) / np.sqrt(p)).astype(np.float32) * (rng.random(V) < 0.3)
    Y = (X @ W + 3.0 * rng.standard_normal((N, V))).astype(np.float32)
    out = {}
    # --row-shard-gram: only the run with the outer Gram formed over 1/world of the TRs per rank + one all-reduce
    # (agreement with the unsharded fit then holds to fp32 rounding of that sum, not bit for bit)
    row_shard = "--row-shard-gram" in sys.argv
    for single_alpha in ((False,) if row_shard else (False, True)):
        kw = dict(n_outer_folds=3, n_inner_folds=3, chunk_length=20, alphas=np.logspace(-1, 6, 12), single_alpha=single_alpha)
        random.seed(1)
        m, w, a = L.NestedCVModel("ridge_regression").fit_predict(X, Y, row_shard_gram=row_shard, **kw)
        if rank == 0:
            random.seed(1)
            m1, w1, a1 = L.NestedCVModel("ridge_regression", comm=SingleProcess()).fit_predict(X, Y, **kw)
            same = np.isclose(a, a1)
            r, r1 = np.asarray(m["correlations"]), np.asarray(m1["correlations"])
            out[f"single_alpha={single_alpha}" + (",row_shard_gram" if row_shard else "")] = {
                "world": world, "alpha_agreement": float(same.mean()), "max_dr_same_alpha": float(np.abs(r - r1)[same].max()),
                "max_dw_rel": float(np.abs(w[:, same] - w1[:, same]).max() / np.abs(w1).max()),
                "n_significant": [m["n_significant"], m1["n_significant"]], "w_shape": list(w.shape)}
            assert same.mean() > (0.97 if row_shard else 0.999) and np.abs(r - r1)[same].max() < 1e-5, out
            assert w.shape == w1.shape and abs(m["n_significant"] - m1["n_significant"]) <= 1
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
