"""Micro-benchmark of the batched 3xTF32 store GEMM at the small-K shapes of the Cholesky solver's trailing
updates (epilogue-bound): variants with / without Cin, in place, split output, K = 128 / 256 / 512."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from litcoder_core_b200.device import default_ops

    ops = default_ops()
    vp = C.c_void_p
    nb, M, N, ld = 20, 4592, 2944, 3072
    rows = 7664
    A = torch.randn((nb, rows, ld), device="cuda")
    Al = torch.zeros_like(A) + 1e-4
    F = torch.randn((nb, rows, ld), device="cuda")
    F2 = torch.randn((nb, rows, ld), device="cuda")
    Dl = torch.empty_like(F2)

    def run(K, cin, out, out_lo, label, Mx=M, Nx=N):
        def call():
            rc = ops.lib.lit_gemm_tf32x3_nt_batched(
                vp(A.data_ptr()), vp(Al.data_ptr()), ld, rows * ld, vp(A.data_ptr() + 128 * 4), vp(Al.data_ptr() + 128 * 4), ld,
                rows * ld, Mx, Nx, K, -1.0, vp(cin.data_ptr() if cin is not None else 0), ld, rows * ld, 1.0 if cin is not None else 0.0,
                vp(out.data_ptr()), vp(out_lo.data_ptr() if out_lo is not None else 0), ld, rows * ld, nb, 0, vp(ops.stream))
            assert rc == 0, ops.lib.lit_last_error()
        for _ in range(2):
            call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        gb = nb * Mx * Nx * 4 * ((1 if cin is not None else 0) + (2 if out_lo is not None else 1)) / 1e9
        print(f"{label:44s} K={K:4d} M={Mx} N={Nx}: {ms * 1e3:8.1f} us  {2.0 * nb * Mx * Nx * K / ms / 1e9:7.1f} TFLOP/s alg  "
              f"{gb / ms * 1e3:7.1f} GB/s epilogue traffic", flush=True)

    if len(sys.argv) > 1 and sys.argv[1] == "one":  # the solver's trailing update, for an ncu capture
        run(128, F, F, None, "Cin in place, fp32 out")
        return
    for K in (128, 256, 512, 1024):
        run(K, None, F2, None, "no Cin, fp32 out")
        run(K, F, F2, None, "Cin separate, fp32 out")
        run(K, F, F, None, "Cin in place, fp32 out")
        run(K, None, F2, Dl, "no Cin, split out")
    run(128, None, F2, Dl, "panel shape, split out", Mx=7536, Nx=128)
    run(128, None, F2, None, "panel shape, fp32 out", Mx=7536, Nx=128)


if __name__ == "__main__":
    main()
