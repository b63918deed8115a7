"""One small config where BOTH arms run the whole nested-CV fit end to end, on identical inputs, fully measured
("add one fully measured small config where both arms run end to end").

    python scripts/gpu_small_both_arms.py [--out gpurun_out/small_both_arms.json]

Workload `dev_small_2000x256x4096` (2,000 TRs x 256 delayed features x 4,096 voxels, 20 alphas, 5 x 5 chunked folds,
SURVEY 8d pipeline data).  Arm A: the UNMODIFIED reference (`baseline/_ref` through oracle/ref_shim.py; the NumPy oracle
port if it is absent), `NestedCVModel.fit_predict(..., use_gpu=False)` on all host threads, SciPy per-voxel loops and
all.  Arm B: this package through the same public call with the same host arrays (H2D and D2H inside the timed
region).  The two results are then put through the parity proof of tests/parity.py.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import logging
import os
import random
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scripts"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "small_both_arms.json"))
    args = ap.parse_args()

    import torch

    import litcoder_core_b200 as L
    import synth8d
    from bench import WORKLOADS, use_all_host_threads
    from oracle import ref_shim
    from oracle import ridge_oracle as O
    from parity import prove_fit_parity

    cores = use_all_host_threads()
    workload = "dev_small_2000x256x4096"
    N, p, V, A, Ko, Ki, chunk = WORKLOADS[workload]
    X = synth8d.design_host(synth8d.make_stories(workload, 0))
    Y = synth8d.responses_host(X, V, seed=0, true_r=0.2)
    alphas = np.logspace(-1, 8, A)
    kw = dict(alphas=alphas, n_outer_folds=Ko, n_inner_folds=Ki, chunk_length=chunk, folding_type="chunked")

    # ---- arm A: the reference, whole fit on the CPU
    ref = ref_shim.load_reference()
    kind = "reference" if ref is not None else "port"
    t_ref = []
    for rep in range(2):
        random.seed(7)
        np.random.seed(7)
        t0 = time.perf_counter()
        if ref is not None:
            logging.disable(logging.CRITICAL)
            with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
                warnings.simplefilter("ignore")
                m_ref, w_ref, a_ref = ref.NestedCVModel("ridge_regression").fit_predict(X, Y, use_gpu=False, **kw)
            logging.disable(logging.NOTSET)
        else:
            m_ref, w_ref, a_ref = O.fit_predict(X, Y, **kw)
        t_ref.append(time.perf_counter() - t0)

    # ---- arm B: the product, same host arrays through the public API
    model = L.NestedCVModel("ridge_regression")
    t_gpu = []
    for rep in range(5):
        random.seed(7)
        np.random.seed(7)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m, w, a = model.fit_predict(X, Y, **kw)
        torch.cuda.synchronize()
        t_gpu.append(time.perf_counter() - t0)

    # ---- parity: the product against the oracle (proof), the oracle's headline numbers against the reference's
    info = prove_fit_parity(model.last_fold_results, m, w, X, Y, 7, max_ambiguous=6, **kw)
    r, r_ref = np.asarray(m["correlations"]), np.asarray(m_ref["correlations"], dtype=np.float64)
    same = np.isclose(a, np.asarray(a_ref), rtol=1e-6)
    units = V * A * Ko * Ki
    rec = {
        "workload": workload, "TRs": N, "features": p, "voxels": V, "alphas": A, "folds": f"{Ko}x{Ki} chunked({chunk})",
        "reference_arm": {"kind": kind, "cores": cores, "fit_seconds": min(t_ref), "fit_seconds_all": t_ref,
                          "value": units / min(t_ref), "unit": "voxel*alpha*fold/s",
                          "what": "NestedCVModel.fit_predict(use_gpu=False), whole fit incl. SciPy per-voxel loops"},
        "b200_arm": {"fit_seconds": min(t_gpu[1:]), "fit_seconds_all": t_gpu, "value": units / min(t_gpu[1:]),
                     "unit": "voxel*alpha*fold/s", "what": "host arrays in, host arrays out (H2D / D2H timed)"},
        "speedup_e2e": min(t_ref) / min(t_gpu[1:]),
        "parity": {"alphas_differing_all_proven_near_ties": info["disagreeing_alphas"], "voxel_folds": info["voxel_folds"],
                   "max_abs_dr_vs_reference_at_product_alphas": info["max_dr"], "weights_max_rel_err": info["weights_rel_err"],
                   "n_significant_product": int(m["n_significant"]), "n_significant_reference": int(m_ref["n_significant"]),
                   "bh_ambiguous": info["ambiguous_bh"],
                   "max_abs_dr_vs_reference_run_same_mean_alpha": float(np.abs(r - r_ref)[same].max()),
                   "fraction_same_mean_alpha": float(same.mean())},
    }
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
