"""Launch only the fp16-pair store GEMM at the shape of an inner fold's cross-product downdate
(C_i^T = C_o^T - Y_R^T X_R: 95,000 x 3,072, K = 1,500, Cin) for `ncu --set full` captures; prints CUDA-event times."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    from litcoder_core_b200.device import DeviceOps, Mat

    ops = DeviceOps()
    M, N, K = 95000, 3072, 1500
    A = ops.split_f16(Mat(torch.randn((M, K), device="cuda"), None, M, K), 1)
    B = ops.split_f16(Mat(torch.randn((N, K), device="cuda"), None, N, K), 1)
    Cin = ops.empty(M, N)
    Cin.hi.normal_()
    out = ops.empty(M, N)
    import ctypes as C

    from litcoder_core_b200.device import _vp, check

    times = []
    pairout = len(sys.argv) > 2 and sys.argv[2] == "pairout"
    if pairout:
        # the epilogue writes the scaled fp16 pair (row scales from the bound |Cin| row max + |A row| |B col|) and no fp32
        rs = (A.hi.float() * A.inv_scale[:M, None]).square().sum(1)
        cs = (B.hi.float() * B.inv_scale[:N, None]).square().sum(1)
        scales = ops.f16_bound_scales(M, absmax=ops.row_absmax(Cin), row_sumsq=rs.contiguous(), col_sumsq=cs.contiguous())
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if pairout:
            H = ops.gemm(A, B, alpha=-1.0, Cin=Cin, beta=1.0, precision="f16x3", pair_out=scales)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
            continue
        check(ops.lib.lit_gemm_f16x3_nt(_vp(A.hi.data_ptr()), _vp(A.lo.data_ptr()), A.ld, _vp(B.hi.data_ptr()),
                                        _vp(B.lo.data_ptr()), B.ld, M, N, K, -1.0, _vp(Cin.hi.data_ptr()), Cin.ld, 1.0,
                                        _vp(out.hi.data_ptr()), _vp(0), out.ld, _vp(A.inv_scale.data_ptr()),
                                        _vp(B.inv_scale.data_ptr()), 0, _vp(ops.stream)), "gemm_f16x3_nt")
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    fl = 2.0 * M * N * K
    print(json.dumps({"M": M, "N": N, "K": K, "output": "fp16 pair" if pairout else "fp32", "ms": times, "algorithmic_tflops": [fl / t / 1e9 for t in times]}))


if __name__ == "__main__":
    main()
