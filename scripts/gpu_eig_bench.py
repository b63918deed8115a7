"""Micro-benchmark of the one library call (cuSOLVER syevd) on the B200: host-call time vs device
time (is the call host-blocking?), batched vs single, fp32 vs fp64, and interference with a
concurrently running 3xTF32 GEMM.  Development aid; prints JSON lines."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    from litcoder_core_b200 import _lib
    from litcoder_core_b200.device import DeviceOps, Mat

    ops = DeviceOps()
    lib = ops.lib
    vp = C.c_void_p
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3072
    X = torch.randn((6016, n), device="cuda")

    def gram():
        return Mat((X.T @ X).contiguous(), None, n, n)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    # ---- single syevd, fp32: host time of the call vs device time
    for rep in range(4):
        G = gram()
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        t0 = time.perf_counter()
        lam = ops.syevd(G)
        host_ms = (time.perf_counter() - t0) * 1e3
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"what": "syevd_f32", "n": n, "rep": rep, "host_call_ms": host_ms, "device_ms": e0.elapsed_time(e1)}),
              flush=True)
    # check the decomposition
    G0 = gram()
    ref = G0.hi.clone()
    lam = ops.syevd(G0)
    V = G0.hi  # rows = eigenvectors
    rec = (V.T * lam[None, :n]) @ V
    print(json.dumps({"what": "recon_err", "rel": float((rec - ref).abs().max() / ref.abs().max())}), flush=True)

    # ---- torch.linalg.eigh for comparison
    for rep in range(2):
        A = gram().hi
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        torch.linalg.eigh(A)
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"what": "torch_eigh_f32", "n": n, "device_ms": e0.elapsed_time(e1)}), flush=True)

    # ---- batched solver (6 matrices)
    try:
        batch = 6
        Gb = torch.stack([gram().hi for _ in range(batch)]).contiguous()
        dev_b, host_b = C.c_size_t(0), C.c_size_t(0)
        _lib.check(lib.lit_syevd_workspace(n, 0, batch, C.byref(dev_b), C.byref(host_b)), "ws")
        work = torch.empty((max(dev_b.value, 16),), dtype=torch.uint8, device="cuda")
        work_h = (C.c_uint8 * max(host_b.value, 16))()
        lam = torch.empty((batch, n), device="cuda")
        info = torch.zeros((batch,), dtype=torch.int32, device="cuda")
        for rep in range(2):
            torch.cuda.synchronize()
            e0, e1 = ev(), ev()
            e0.record()
            t0 = time.perf_counter()
            _lib.check(lib.lit_syevd(vp(Gb.data_ptr()), n, n, 0, batch, vp(lam.data_ptr()), vp(work.data_ptr()), dev_b.value,
                                     C.cast(work_h, vp), host_b.value, vp(info.data_ptr()),
                                     vp(torch.cuda.current_stream().cuda_stream)), "syevd batched")
            host_ms = (time.perf_counter() - t0) * 1e3
            e1.record()
            torch.cuda.synchronize()
            print(json.dumps({"what": "syevBatched_f32", "batch": batch, "host_call_ms": host_ms,
                              "device_ms": e0.elapsed_time(e1), "ws_mb": dev_b.value / 1e6}), flush=True)
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"what": "syevBatched_f32", "error": str(e)[:300]}), flush=True)

    # ---- interference: 6 eigs (worker thread, side stream) while 6 fused GEMMs run on the main stream,
    # with the GEMM grid restricted to fewer SMs so that the cuSOLVER kernels find room
    M, N, K = 95000, 30720, 3072
    A = ops.split(Mat(torch.randn((M, K), device="cuda"), None, M, K))
    B = ops.split(Mat(torch.randn((N, K), device="cuda"), None, N, K))
    Yz = Mat(torch.randn((1536, M), device="cuda"), None, 1536, M)
    flops6 = 6 * 2.0 * M * N * K
    for limit in (0, 140, 132, 124, 116):
        ops.set_gemm_sm_limit(limit)
        for _ in range(2):
            ops.gemm_corr(A, B, 20, 1536, Yz)
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(6):
            ops.gemm_corr(A, B, 20, 1536, Yz)
        e1.record()
        torch.cuda.synchronize()
        gemm_alone = e0.elapsed_time(e1)
        Gs = [gram() for _ in range(6)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        tickets = [ops.syevd_async(G) for G in Gs]
        t_queue = (time.perf_counter() - t0) * 1e3
        for _ in range(6):
            ops.gemm_corr(A, B, 20, 1536, Yz)
        e1.record()
        for _, tk in tickets:
            ops.wait(tk)
        e2.record()
        torch.cuda.synchronize()
        print(json.dumps({"what": "interference", "gemm_sm_limit": limit, "gemm6_alone_ms": gemm_alone,
                          "gemm6_with_eigs_ms": e0.elapsed_time(e1), "both_done_ms": e0.elapsed_time(e2),
                          "host_queue_6_eigs_ms": t_queue, "corr_tflops_alone": flops6 / gemm_alone / 1e9}), flush=True)
    ops.set_gemm_sm_limit(0)
    # ---- store-epilogue GEMM of the same shape, 6 back to back (sustained clocks)
    for _ in range(2):
        D = ops.gemm(A, B)
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(6):
        D = ops.gemm(A, B, out=D)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"what": "store_gemm_sustained", "ms6": e0.elapsed_time(e1),
                      "tflops": flops6 / e0.elapsed_time(e1) / 1e9}), flush=True)
    ops.close()


if __name__ == "__main__":
    main()
