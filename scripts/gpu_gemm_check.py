"""GPU bring-up check for the tcgen05 3xTF32 GEMM (run under gpurun).

Each variant runs in its own subprocess so that a device trap (mbarrier timeout) in one
variant does not poison the CUDA context of the others.  Results land in gpurun_out/.

    python scripts/gpu_gemm_check.py            # driver: all variants
    python scripts/gpu_gemm_check.py --variant 1 --out gpurun_out/gemm_v1.json
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load():
    lib = ctypes.CDLL(os.path.join(ROOT, "litcoder_core_b200", "liblitridge.so"))
    lib.lit_last_error.restype = ctypes.c_char_p
    return lib


def run_variant(variant, out_path, quick=False):
    import torch

    lib = load()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    stream = torch.cuda.current_stream().cuda_stream
    c_f = ctypes.c_float
    vp = ctypes.c_void_p
    L = ctypes.c_long
    I = ctypes.c_int

    def chk(rc, what):
        if rc != 0:
            raise RuntimeError(f"{what}: rc={rc} {lib.lit_last_error().decode()}")

    def split(x):
        rows, cols = x.shape
        hi = torch.empty_like(x)
        lo = torch.empty_like(x)
        chk(lib.lit_split_tf32(vp(x.data_ptr()), L(rows), L(cols), L(x.stride(0)), vp(hi.data_ptr()),
                               vp(lo.data_ptr()), L(hi.stride(0)), vp(stream)), "split")
        return hi, lo

    def gemm(Ah, Al, Bh, Bl, M, N, K, alpha=1.0, Cin=None, beta=0.0, split_out=False):
        ldd = (N + 3) // 4 * 4
        D = torch.full((M, ldd), float("nan"), device=dev, dtype=torch.float32)
        Dl = torch.full((M, ldd), float("nan"), device=dev, dtype=torch.float32) if split_out else None
        chk(lib.lit_gemm_tf32x3_nt(vp(Ah.data_ptr()), vp(Al.data_ptr()), L(Ah.stride(0)), vp(Bh.data_ptr()),
                                   vp(Bl.data_ptr()), L(Bh.stride(0)), I(M), I(N), I(K), c_f(alpha),
                                   vp(Cin.data_ptr() if Cin is not None else 0),
                                   L(Cin.stride(0) if Cin is not None else 0), c_f(beta), vp(D.data_ptr()),
                                   vp(Dl.data_ptr() if Dl is not None else 0), L(ldd), I(variant), vp(stream)), "gemm")
        return D, Dl

    results = {"variant": variant, "cases": [], "perf": []}
    shapes = [(128, 256, 32), (128, 256, 64), (256, 512, 96), (200, 300, 100), (1000, 777, 515), (130, 36, 4),
              (4096, 3072, 1504), (777, 4096, 3072)]
    for (M, N, K) in shapes:
        ldk = (K + 3) // 4 * 4
        A = torch.zeros((M, ldk), device=dev)
        B = torch.zeros((N, ldk), device=dev)
        A[:, :K] = torch.randn(M, K, device=dev) * torch.exp(torch.randn(M, 1, device=dev))
        B[:, :K] = torch.randn(N, K, device=dev) + 0.5
        Ah, Al = split(A)
        Bh, Bl = split(B)
        torch.cuda.synchronize()
        # split exactness: hi + lo == x to ~2^-22
        serr = ((Ah.double() + Al.double() - A.double()).abs().max() / A.abs().max()).item()
        D, _ = gemm(Ah, Al, Bh, Bl, M, N, K)
        torch.cuda.synchronize()
        ref = A[:, :K].double() @ B[:, :K].double().T
        scale = (A[:, :K].double().abs() @ B[:, :K].double().abs().T) + 1e-30
        got = D[:, :N].double()
        err = ((got - ref).abs() / scale).max().item()
        nan = bool(torch.isnan(D[:, :N]).any().item())
        # compare with what plain fp32 matmul achieves
        fp32 = (A[:, :K] @ B[:, :K].T).double()
        err32 = ((fp32 - ref).abs() / scale).max().item()
        case = {"M": M, "N": N, "K": K, "max_scaled_err": err, "fp32_matmul_scaled_err": err32, "split_err": serr,
                "nan": nan}
        # alpha/beta/Cin + split output on one mid-size shape
        if (M, N, K) == (200, 300, 100):
            Cin = torch.randn(M, 300, device=dev)
            D2, D2l = gemm(Ah, Al, Bh, Bl, M, N, K, alpha=-0.5, Cin=Cin, beta=2.0, split_out=True)
            torch.cuda.synchronize()
            ref2 = -0.5 * ref + 2.0 * Cin.double()
            got2 = D2[:, :N].double() + D2l[:, :N].double()
            case["axpby_split_err"] = ((got2 - ref2).abs() / (scale + Cin.double().abs())).max().item()
        results["cases"].append(case)
        print(case, flush=True)

    # fused correlation epilogue
    for (Mv, groups, R, K, nreal) in [(300, 3, 256, 64, 200), (1000, 4, 512, 128, 500)]:
        ldk = (K + 3) // 4 * 4
        A = torch.randn(Mv, ldk, device=dev)
        B = torch.zeros(groups * R, ldk, device=dev)
        for g in range(groups):
            B[g * R:g * R + nreal] = torch.randn(nreal, ldk, device=dev) * (g + 1)
        Yz = torch.zeros(R, Mv, device=dev)
        Yz[:nreal] = torch.randn(nreal, Mv, device=dev)
        Ah, Al = split(A)
        Bh, Bl = split(B)
        ntile = groups * R // 256
        dot = torch.full((ntile, Mv), float("nan"), device=dev)
        ssq = torch.full((ntile, Mv), float("nan"), device=dev)
        cvar = variant if variant in (1, 3) else 1
        chk(lib.lit_gemm_tf32x3_nt_corr(vp(Ah.data_ptr()), vp(Al.data_ptr()), L(ldk), vp(Bh.data_ptr()),
                                        vp(Bl.data_ptr()), L(ldk), I(Mv), I(groups), I(R), I(K), vp(Yz.data_ptr()),
                                        L(Yz.stride(0)), vp(dot.data_ptr()), vp(ssq.data_ptr()), L(Mv), I(cvar),
                                        vp(stream)), "gemm_corr")
        torch.cuda.synchronize()
        pred = (A[:, :K].double() @ B[:, :K].double().T)  # [Mv, groups*R]
        pred = pred.view(Mv, groups, R)
        dref = torch.einsum("vgt,tv->gv", pred, Yz.double())
        sref = (pred ** 2).sum(-1).T
        tpg = R // 256
        dgot = dot.double().view(groups, tpg, Mv).sum(1)
        sgot = ssq.double().view(groups, tpg, Mv).sum(1)
        e1 = ((dgot - dref).abs() / (dref.abs().max())).max().item()
        e2 = ((sgot - sref).abs() / sref.abs()).max().item()
        case = {"corr_case": [Mv, groups, R, K], "dot_err": e1, "ssq_rel_err": e2}
        results["cases"].append(case)
        print(case, flush=True)

    if not quick:
        # throughput on the shapes of BASELINE config 2 (scaled M) -- inputs > L2
        for (M, N, K) in [(16384, 30720, 3072), (32768, 3072, 6016), (8192, 8192, 8192)]:
            A = torch.randn(M, K, device=dev)
            B = torch.randn(N, K, device=dev)
            Ah, Al = split(A)
            Bh, Bl = split(B)
            del A, B
            D = torch.empty(M, N, device=dev)
            args = (vp(Ah.data_ptr()), vp(Al.data_ptr()), L(K), vp(Bh.data_ptr()), vp(Bl.data_ptr()), L(K), I(M), I(N),
                    I(K), c_f(1.0), vp(0), L(0), c_f(0.0), vp(D.data_ptr()), vp(0), L(N), I(variant), vp(stream))
            for _ in range(2):
                chk(lib.lit_gemm_tf32x3_nt(*args), "gemm")
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            reps = 3
            for _ in range(reps):
                chk(lib.lit_gemm_tf32x3_nt(*args), "gemm")
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            tf = 2.0 * M * N * K / ms / 1e9
            perf = {"M": M, "N": N, "K": K, "ms": ms, "fp32_equiv_tflops": tf, "tf32_mma_tflops": 3 * tf}
            results["perf"].append(perf)
            print(perf, flush=True)
            del Ah, Al, Bh, Bl, D
            torch.cuda.empty_cache()

    with open(out_path, "w") as f:
        json.dump(results, f, indent=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", type=int, default=None)
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    if a.variant is not None:
        run_variant(a.variant, a.out or os.path.join(ROOT, "gpurun_out", f"gemm_v{a.variant}.json"), a.quick)
        return
    for v in (1, 2, 3):
        t0 = time.time()
        out = os.path.join(ROOT, "gpurun_out", f"gemm_v{v}.json")
        log = os.path.join(ROOT, "gpurun_out", f"gemm_v{v}.log")
        with open(log, "w") as lf:
            try:
                rc = subprocess.run([sys.executable, __file__, "--variant", str(v), "--out", out] +
                                    (["--quick"] if a.quick else []), stdout=lf, stderr=subprocess.STDOUT,
                                    timeout=420).returncode
            except subprocess.TimeoutExpired:
                rc = "timeout"
        print(f"variant {v}: rc={rc} ({time.time() - t0:.1f}s)")
        with open(log) as lf:
            print(lf.read()[-3000:])


if __name__ == "__main__":
    main()
