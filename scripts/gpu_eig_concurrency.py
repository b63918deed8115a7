"""How well do several cuSOLVER syevd calls overlap on one B200 when nothing else runs?  (development aid)
Times 12 decompositions of order n with 1..4 worker threads / streams."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    from litcoder_core_b200.device import DeviceOps, Mat

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3072
    X = torch.randn((6016, n), device="cuda")
    for workers in (1, 2, 3, 4, 6):
        os.environ["LIT_EIG_WORKERS"] = str(workers)
        ops = DeviceOps()
        ops.overlap_sms = 0
        Gs = [Mat((X.T @ X).contiguous(), None, n, n) for _ in range(12)]
        for G in Gs[:workers]:  # warm up handles / workspaces of every worker
            ops.wait(ops.syevd_async(G)[1])
        Gs = [Mat((X.T @ X).contiguous(), None, n, n) for _ in range(12)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tickets = [ops.syevd_async(G)[1] for G in Gs]
        for tk in tickets:
            ops.wait(tk)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        print(json.dumps({"n": n, "workers": workers, "ms_total_12": round(dt, 1), "ms_per_eig": round(dt / 12, 2)}), flush=True)
        ops.close()


if __name__ == "__main__":
    main()
