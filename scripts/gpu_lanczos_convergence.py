"""Ritz-value convergence of the batched Lanczos on the config-2 design (device-built, SURVEY 8d pipeline):
lambda_max of 6 training-row subsets after 40..96 steps, relative to the 96-step value."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scripts")):
    sys.path.insert(0, p)


def main():
    import torch

    import synth8d
    from litcoder_core_b200.device import default_ops

    ops = default_ops()
    X = synth8d.design_device(synth8d.make_stories("config2_gpt2_9400x3072x95000", 0), ops).contiguous()
    N = X.shape[0]
    mats = []
    for f in range(6):
        keep = np.ones(N, dtype=bool)
        keep[f * 1500:(f + 1) * 1500 + 380 * (f % 2)] = False
        Xs = X[torch.from_numpy(np.nonzero(keep)[0]).cuda()].double()
        mats.append(ops.wrap((Xs.T @ Xs).float().contiguous()))
    ref = ops.lambda_max_batched(mats, steps=96).cpu().numpy()
    rec = {"lambda_max_96": ref.tolist()}
    for steps in (40, 48, 56, 64, 72, 80, 88):
        got = ops.lambda_max_batched(mats, steps=steps).cpu().numpy()
        rec[f"rel_diff_{steps}"] = [float(abs(a - b) / b) for a, b in zip(got, ref)]
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
