"""Accuracy of the 3xTF32 tcgen05 GEMM against fp64 as a function of K (development aid):
signed relative bias (does the tensor-core accumulator truncate?) and rms / max error, for
random-sign and all-positive operands, next to torch's fp32 matmul (TF32 disabled)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    from litcoder_core_b200.device import DeviceOps, Mat

    torch.backends.cuda.matmul.allow_tf32 = False
    ops = DeviceOps()
    M, N = 512, 512
    for K in (256, 1024, 3072, 7520, 16384):
        for kind in ("randn", "positive"):
            g = torch.Generator(device="cuda").manual_seed(K)
            A = torch.randn((M, K), device="cuda", generator=g)
            B = torch.randn((N, K), device="cuda", generator=g)
            if kind == "positive":
                A, B = A.abs(), B.abs()
            ref = A.double() @ B.double().T
            D = ops.gemm(ops.split(Mat(A, None, M, K)), ops.split(Mat(B, None, N, K)))
            got = D.hi[:, :N].double()
            f32 = (A @ B.T).double()
            denom = ref.abs().mean()
            out = {"K": K, "kind": kind}
            for name, x in (("tf32x3", got), ("fp32_matmul", f32)):
                e = (x - ref)
                out[name] = {"bias_rel": float((e * ref.sign()).mean() / denom), "rms_rel": float(e.pow(2).mean().sqrt() / denom),
                             "max_rel": float(e.abs().max() / denom)}
            print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
