"""Host <-> device copy rates on the box, next to what the upload paths of DeviceOps reach (e2e accounting).

    python scripts/gpu_pcie_bench.py

Prints one JSON record: raw pinned H2D / D2H rates (torch copy_, CUDA events), DeviceOps.upload_matrix from pinned
memory, DeviceOps.upload_matrix_bg from pageable memory, both for the config-2 response block (9,400 x 95,000 fp32).
"""
from __future__ import annotations

import json
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
    N, V = 9400, 95000
    host = torch.empty((N, V), dtype=torch.float32, pin_memory=True)
    host.normal_()
    dev = torch.empty((N, V), dtype=torch.float32, device="cuda")
    gb = N * V * 4 / 1e9
    rec = {"bytes": N * V * 4}

    def ev_time(fn, reps=3):
        out = []
        for _ in range(reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            out.append(e0.elapsed_time(e1))
        return out

    t = ev_time(lambda: dev.copy_(host, non_blocking=True))
    rec["raw_pinned_h2d_GBps"] = [round(gb / (x / 1e3), 1) for x in t]
    t = ev_time(lambda: host.copy_(dev, non_blocking=True))
    rec["raw_pinned_d2h_GBps"] = [round(gb / (x / 1e3), 1) for x in t]
    del dev
    torch.cuda.empty_cache()

    # DeviceOps.upload_matrix from the pinned array (what the e2e leg of bench.py passes)
    Yh = host.numpy()
    out = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with ops.copy_stream() as ticket:
            Y = ops.upload_matrix(Yh, 0, V)
        t_issue = time.perf_counter() - t0
        ops.wait_copy(ticket)
        torch.cuda.synchronize()
        out.append((round(t_issue * 1e3, 1), round((time.perf_counter() - t0) * 1e3, 1)))
        del Y
    rec["upload_matrix_pinned_ms(issue,total)"] = out
    rec["upload_matrix_pinned_GBps"] = round(gb / (min(o[1] for o in out) / 1e3), 1)

    # DeviceOps.upload_matrix_bg from a pageable copy
    Yp = np.array(Yh)
    out = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        Y, ticket = ops.upload_matrix_bg(Yp, 0, V)
        t_issue = time.perf_counter() - t0
        ops.wait_copy(ticket)
        torch.cuda.synchronize()
        out.append((round(t_issue * 1e3, 1), round((time.perf_counter() - t0) * 1e3, 1)))
        del Y
    rec["upload_matrix_bg_pageable_ms(issue,total)"] = out
    rec["upload_matrix_bg_pageable_GBps"] = round(gb / (min(o[1] for o in out) / 1e3), 1)
    print(json.dumps(rec))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "pcie_bench.json"), "w") as f:
        json.dump(rec, f, indent=1)


if __name__ == "__main__":
    main()
