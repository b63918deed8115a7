"""fp16-split (f16x3) vs 3xTF32 form of the fused prediction + correlation GEMM on a B200 (development aid).

1. accuracy against fp64 as a function of K: signed relative bias and rms of the per-voxel sums
   (does the kind::f16 accumulation truncate like kind::tf32?), random-sign and all-positive operands;
2. CUDA-event time per launch at the BASELINE config-2 inner-fold shape, both forms, plus the
   lit_split_f16 conversions the f16 form pays for.
Prints one JSON object per line."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def accuracy(ops, torch, Mat):
    M, G, R = 512, 2, 256
    for K in (256, 1024, 3072, 16384):
        for kind in ("randn", "positive"):
            g = torch.Generator(device="cuda").manual_seed(K)
            A = torch.randn((M, K), device="cuda", generator=g)
            B = torch.randn((G * R, K), device="cuda", generator=g)
            if kind == "positive":
                A, B = A.abs(), B.abs()
            Yz = Mat(torch.ones((R, M), device="cuda"), None, R, M)
            acc = A.double() @ B.double().T  # [M][G*R]
            want_d = acc.reshape(M, G, R).sum(2).T  # [G][M]
            want_q = (acc * acc).reshape(M, G, R).sum(2).T
            out = {"K": K, "kind": kind}
            for prec in ("tf32x3", "f16x3"):
                parts = ops.gemm_corr(ops.split(Mat(A, None, M, K)), ops.split(Mat(B, None, G * R, K)), G, R, Yz,
                                      precision=prec)
                d = parts.dot[:, :M].double().reshape(G, R // 128, M).sum(1)
                q = parts.ssq[:, :M].double().reshape(G, R // 128, M).sum(1)
                if parts.inv_row is not None:
                    sc = parts.inv_group[:G].double()[:, None] * parts.inv_row[:M].double()[None, :]
                    d, q = d * sc, q * sc * sc
                denom = want_d.abs().mean()
                e = d - want_d
                out[prec] = {"dot_bias_rel": float((e * want_d.sign()).mean() / denom),
                             "dot_rms_rel": float(e.pow(2).mean().sqrt() / denom),
                             "ssq_bias_rel": float((q / want_q - 1).mean()),
                             "ssq_max_rel": float((q / want_q - 1).abs().max())}
            print(json.dumps(out), flush=True)


def timing(ops, torch, Mat, reps):
    M, G, R, K = 95000, 20, 1536, 3072
    A = ops.split(Mat(torch.randn((M, K), device="cuda"), None, M, K))
    B = ops.split(Mat(torch.randn((G * R, K), device="cuda"), None, G * R, K))
    Yz = Mat(torch.randn((R, M), device="cuda"), None, R, M)
    flops = 2.0 * M * G * R * K
    for prec in ("tf32x3", "f16x3", "tf32x3", "f16x3"):
        ops.reset_counters()
        for _ in range(reps):
            ops.gemm_corr(A, B, G, R, Yz, precision=prec)
        ms = [t for t, _ in ops.corr_launches()]
        tm = ops.timings()
        print(json.dumps({"precision": prec, "gemm_ms": ms, "algorithmic_tflops": [flops / t / 1e9 for t in ms],
                          "split_f16_ms_total": tm.get("split_f16", 0.0), "reps": reps}), flush=True)
    # conversion alone: A (95,000 x 3,072 split pair, one scale per row), B (20 groups of 1,536 rows)
    for name, src, rpg in (("A", A, 1), ("B", B, R)):
        ops.reset_counters()
        for _ in range(reps):
            ops.split_f16(src, rpg)
        t = ops.timings()["split_f16"] / reps
        gb = src.rows * src.cols * (8 + 4) / 1e9  # read two fp32 planes (twice: max pass + convert), write two fp16
        print(json.dumps({"split_f16": name, "ms": t, "GBps_read2x_write": (gb + src.rows * src.cols * 8 / 1e9) / t * 1e3}),
              flush=True)


def main():
    import torch

    from litcoder_core_b200.device import DeviceOps, Mat

    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    ops = DeviceOps()
    accuracy(ops, torch, Mat)
    timing(ops, torch, Mat, reps)


if __name__ == "__main__":
    main()
