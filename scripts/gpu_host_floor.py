"""Host-side floor of one config-2 fit: the same design (9,400 x 3,072, 5 x 5 folds, 20 alphas) on few voxels, so that
the voxel-side GEMMs vanish and what remains is the design side + the host's launch / bookkeeping time."""
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def main():
    import torch

    import litcoder_core_b200 as L
    import synth8d
    from litcoder_core_b200.device import default_ops

    V = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    ops = default_ops()
    X = synth8d.design_device(synth8d.make_stories("config2_gpt2_9400x3072x95000", 0), ops).contiguous()
    Y = synth8d.responses_device(torch, X, V, 0)
    model = L.NestedCVModel("ridge_regression")
    kw = dict(alphas=np.logspace(-1, 8, 20), n_outer_folds=5, n_inner_folds=5, chunk_length=20)
    for rep in range(4):
        random.seed(rep)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.fit_predict(X, Y, device_outputs=True, **kw)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        t = model.last_timings
        print(f"V={V} rep {rep}: wall {dt:.1f} ms; fit(events) {t.get('fit', 0):.1f} design {t.get('phase_design', 0):.1f} "
              f"inner_cv {t.get('phase_inner_cv', 0):.1f} outer {t.get('phase_outer_fit', 0):.1f} "
              f"stats+metrics(host) {t.get('host_stats_metrics_ms', 0):.1f} launches {model.last_stats['launches']}", flush=True)


if __name__ == "__main__":
    main()
