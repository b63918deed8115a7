"""The fp16 gather-transpose at config-2 sizes: CUDA-event time, GB/s and fraction of the measured copy bandwidth,
plus a bit-exact check against the NumPy restatement on a ragged case."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def main():
    import torch

    from fake_ops import FakeOps
    from litcoder_core_b200.device import DeviceOps, Mat

    ops = DeviceOps()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6454.6
    rng = np.random.default_rng(0)
    # correctness (ragged widths, unaligned row counts)
    for N, V, n in [(700, 333, 150), (400, 1030, 301), (90, 64, 1)]:
        Y = (rng.standard_normal((N, V)) * np.exp(rng.uniform(-6, 6, (1, V)))).astype(np.float32)
        idx = np.sort(rng.choice(N, n, replace=False))
        sc = FakeOps().f16_bound_scales(V, absmax=np.abs(Y).max(0))
        Yd = ops.upload_matrix(Y)
        T = ops.gather_rows_T_f16(Yd, ops.upload_index(idx), n, (ops.upload_vector(sc[0], "f32"), ops.upload_vector(sc[1], "f32")))
        hi, lo = T.hi.cpu().numpy().astype(np.float64), T.lo.cpu().numpy().astype(np.float64)
        assert not hi[:V, n:].any() and not lo[:V, n:].any()
        got = ((hi + lo)[:V, :n] * sc[1][:, None].astype(np.float64)).astype(np.float32)
        assert np.array_equal(got, FakeOps._pair_with_scales(Y[idx].T, sc[0])), (N, V, n)
    print("bit-exact ok")
    N, V = 9400, 95000
    Y = Mat(torch.randn((N, V), device="cuda"), None, N, V)
    ysc = ops.f16_bound_scales(V, absmax=ops.col_reduce(Y, None, N, sumsq=False, absmax=True)[1])
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device="cuda")
    for n in (1500, 1900, 9400):
        idx = ops.upload_index(np.sort(rng.permutation(N)[:n]))
        ms = []
        for _ in range(6):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.gather_rows_T_f16(Y, idx, n, ysc)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        b = n * V * 8
        print(json.dumps({"rows": n, "ms": round(min(ms[1:]), 4), "GBps": round(b / min(ms[1:]) / 1e6, 1),
                          "frac_of_measured_copy": round(b / min(ms[1:]) / 1e6 / peak, 3)}))


if __name__ == "__main__":
    main()
