"""Launch only the dominant kernel -- the alpha-stacked prediction GEMM with the fused correlation
epilogue -- at the BASELINE config-2 inner-fold shape (95,000 voxels x 20 alphas x 1,536 padded
validation TRs, K = 3,072), for `ncu --set full` captures.  Prints the CUDA-event time per launch."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    from litcoder_core_b200.device import DeviceOps, Mat

    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    ops = DeviceOps()
    M, G, R, K = 95000, 20, 1536, 3072
    A = ops.split(Mat(torch.randn((M, K), device="cuda"), None, M, K))
    B = ops.split(Mat(torch.randn((G * R, K), device="cuda"), None, G * R, K))
    Yz = Mat(torch.randn((R, M), device="cuda"), None, R, M)
    times = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm_corr(A, B, G, R, Yz)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    flops = 2.0 * M * G * R * K
    print(json.dumps({"ms": times, "algorithmic_tflops": [flops / t / 1e9 for t in times]}))


if __name__ == "__main__":
    main()
