"""Measured dense TF32 (and fp16) tensor-core rate of this box through cuBLAS: torch.matmul 8192^3, best of 10
(burst) and back to back for 3 s (sustained) -- the denominators for the 3xTF32 / fp16-pair GEMM rooflines next to
MEASURED_PEAKS.json's bf16 figure."""
import json
import os
import time

import torch


def rate(dtype, tf32):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    n = 8192
    a = torch.randn((n, n), device="cuda", dtype=dtype)
    b = torch.randn((n, n), device="cuda", dtype=dtype)
    fl = 2.0 * n ** 3
    for _ in range(3):
        a @ b
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    k = 0
    e0.record()
    while time.perf_counter() - t0 < 3.0:
        for _ in range(20):
            a @ b
        k += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return {"burst_tflops": round(fl / best / 1e9, 1), "sustained_tflops": round(fl * k / e0.elapsed_time(e1) / 1e9, 1)}


def main():
    rec = {"how": "torch.matmul 8192^3 through cuBLAS: best of 10 (burst), back to back for 3 s (sustained)",
           "tf32": rate(torch.float32, True), "fp16": rate(torch.float16, False), "bf16": rate(torch.bfloat16, False),
           "fp32_no_tf32": rate(torch.float32, False)}
    print(json.dumps(rec))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    json.dump(rec, open(os.path.join(root, "gpurun_out", "tf32_peak.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
