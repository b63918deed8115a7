"""NumPy stand-in for litcoder_core_b200.device.DeviceOps -- TEST INFRASTRUCTURE ONLY.

It restates, operation by operation, what the C-ABI entry points of include/litridge.h compute
(in float32, GEMMs accumulated in float64 to play the role of 3xTF32), so that the host logic of
the engine (fold plans, downdates, stream tickets, accumulation over folds, sharding, metrics) can
be exercised by `pytest -m "not gpu"` on a machine without a GPU.  It is never imported by the
product package; the GPU tests run the same engine code against the real DeviceOps.
"""
from __future__ import annotations

import contextlib
import math

import numpy as np

F32 = np.float32


class FMat:
    def __init__(self, a: np.ndarray, split: bool = False):
        self.a = np.ascontiguousarray(a, dtype=F32)
        self.is_split = split

    @property
    def rows(self):
        return self.a.shape[0]

    @property
    def cols(self):
        return self.a.shape[1]

    @property
    def ld(self):
        return self.a.shape[1]


class FPair(FMat):
    """Mirror of device.MatF16 with one scale per row: `a` holds what the pair keeps, (hi + lo) / s."""

    def __init__(self, a, scales):
        super().__init__(a, split=False)
        self.rows_per_group, self.scales = 1, scales


class FakeSeriesStack:
    """Mirror of device.SeriesStack."""

    def __init__(self, mat, n_cheb, rows_pad, n_tiles, slot_cheb, slot_series, coef):
        self.mat, self.n_cheb, self.rows_pad, self.n_tiles = mat, n_cheb, rows_pad, n_tiles
        self.slot_cheb, self.slot_series, self.coef = slot_cheb, slot_series, coef


class FakeOps:
    TILE_N = 256
    PART_N = 128

    def __init__(self):
        self.reset_counters()
        self.eig_calls = 0
        self.eig_sizes = []
        self.pending = 0

    def reset_counters(self):
        self.launches = 0
        self.gemm_flops = 0.0
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    @contextlib.contextmanager
    def timed(self, name):
        yield

    def timings(self):
        return {}

    def synchronize(self):
        pass

    @contextlib.contextmanager
    def copy_stream(self):
        yield None

    @contextlib.contextmanager
    def side_stream(self):
        yield None

    def adopt(self, tensors):
        pass

    def wait_copy(self, ticket):
        pass

    def check_eig(self):
        assert self.pending == 0, "an eigendecomposition ticket was never waited on"

    # ------------------------------------------------------------------ memory
    def empty(self, rows, cols, split=False, ld=None):
        return FMat(np.full((rows, cols), np.nan, dtype=F32), split)

    def zeros(self, rows, cols):
        return FMat(np.zeros((rows, cols), dtype=F32))

    def upload_vector(self, arr, dtype):
        dt = {"f32": np.float32, "f64": np.float64, "i32": np.int32}[dtype]
        return np.array(arr, dtype=dt).reshape(-1)

    def stage_indices(self, arrays):
        return [np.asarray(a, dtype=np.int64).astype(np.int32) for a in arrays]

    def upload_index(self, idx):
        return np.asarray(idx, dtype=np.int64).astype(np.int32)

    def upload_matrix(self, host, col_start=0, col_stop=None, **kw):
        host = np.asarray(host)
        assert host.ndim == 2
        self.h2d_bytes += host[:, col_start:col_stop].nbytes
        return FMat(host[:, col_start:col_stop].astype(F32))

    def col_view(self, m, c0, c1):
        assert c0 % 4 == 0
        return FMat(m.a[:, c0:c1], m.is_split)

    def raw(self, x):
        return x.a if isinstance(x, FMat) else x

    def planes(self, m):
        return [m.a]

    def vec(self, n, dtype="f32"):
        return np.full(n, np.nan, dtype={"f32": np.float32, "f64": np.float64, "i32": np.int32, "u8": np.uint8}[dtype])

    def download(self, t):
        return np.array(t)

    def download_matrix(self, m):
        self.d2h_bytes += m.a.nbytes
        return m.a.copy()

    # ------------------------------------------------------------------ layout
    def gather_rows_T_split(self, src, idx, n, split=True):
        return FMat(src.a[np.asarray(idx[:n], dtype=np.int64)].T, split=split)

    # ---- fp16 pairs written by the producer (lit_gather_col_reduce / lit_row_absmax / lit_f16_bound_scales / ...)
    def col_reduce(self, src, idx, n, sumsq=True, absmax=False):
        rows = src.a[:n] if idx is None else src.a[np.asarray(idx[:n], dtype=np.int64)]
        self.launches += 1
        ss = (rows.astype(np.float64) ** 2).sum(0).astype(F32) if sumsq else None
        am = (np.abs(rows).max(0) if n else np.zeros(src.cols, dtype=F32)) if absmax else None
        return ss, am

    def row_absmax(self, src):
        self.launches += 1
        return np.abs(src.a).max(1) if src.cols else np.zeros(src.rows, dtype=F32)

    def f16_bound_scales(self, n, absmax=None, row_sumsq=None, col_sumsq=None):
        assert (row_sumsq is None) == (col_sumsq is None) and (absmax is not None or row_sumsq is not None)
        bound = np.zeros(n, dtype=F32) if absmax is None else np.asarray(absmax[:n], dtype=F32).copy()
        if row_sumsq is not None:
            bound = bound + np.sqrt(np.asarray(row_sumsq[:n], dtype=F32) * F32(np.max(col_sumsq) if len(col_sumsq) else 0))
        bound = (bound * F32(1 + 2.0 ** -10)).astype(F32)
        scale = np.ones(n, dtype=F32)
        ok = np.isfinite(bound) & (bound > 0)
        scale[ok] = np.ldexp(F32(1.0), np.clip(15 - np.frexp(bound[ok])[1], -100, 100)).astype(F32)
        self.launches += 1
        return scale, (F32(1.0) / scale).astype(F32)

    @staticmethod
    def _pair_with_scales(a, scale):
        """(hi + lo) / s of the fp16 pair of s * a with the given row scales; overflow would show up as inf."""
        y = (np.asarray(a, dtype=F32) * scale[:, None]).astype(F32)
        assert not y.size or np.nanmax(np.abs(y[np.isfinite(y)]), initial=0.0) < 65504.0, "fp16 overflow: the scale bound failed"
        hi = y.astype(np.float16)
        lo = (y - hi.astype(F32)).astype(np.float16)
        return ((hi.astype(np.float64) + lo.astype(np.float64)) / scale[:, None].astype(np.float64)).astype(F32)

    def gather_rows_T_f16(self, src, idx, n, scales):
        self.launches += 1
        return FPair(self._pair_with_scales(src.a[np.asarray(idx[:n], dtype=np.int64)].T, scales[0]), scales)

    def gather_rows(self, src, idx, n, rows_out=None, split=False):
        rows_out = n if rows_out is None else rows_out
        out = np.zeros((rows_out, src.cols), dtype=F32)
        if idx is None:
            out[:n] = src.a[:n]
        else:
            ix = np.asarray(idx[:n], dtype=np.int64)
            out[:n][ix >= 0] = src.a[ix[ix >= 0]]  # a negative index gives a zero row
        return FMat(out, split)

    def transpose(self, src, split=False):
        return FMat(src.a.T, split)

    def split(self, src):
        return FMat(src.a.copy(), split=True)

    def copy(self, src):
        return FMat(src.a.copy(), src.is_split)

    def axpy(self, a, x, y):
        y.a += F32(a) * x.a

    # ------------------------------------------------------------------ statistics
    def col_stats(self, src, idx, n, ddof):
        rows = src.a[:n] if idx is None else src.a[np.asarray(idx[:n], dtype=np.int64)]
        r64 = rows.astype(np.float64)
        mean = r64.mean(0)
        var = ((r64 - mean) ** 2).sum(0) / (n - ddof) if n - ddof > 0 else np.full(src.cols, np.nan)
        return mean.astype(F32), np.sqrt(var).astype(F32)

    def row_view(self, m, r0, rows):
        assert 0 <= r0 and r0 + rows <= m.rows
        v = FMat.__new__(FMat)
        v.a, v.is_split = m.a[r0:r0 + rows], m.is_split
        return v

    def upload_into(self, host, out):
        host = np.asarray(host)
        assert host.shape == (out.rows, out.cols)
        self.h2d_bytes += host.nbytes
        out.a[...] = host.astype(F32)

    def fir_zscore_rows(self, stim, delays, circpad, row_start, row_stop, zscore, out):
        """NumPy restatement of lit_fir_zscore_rows (fp64 moments, fp32 output)."""
        stim = np.asarray(stim, dtype=np.float64)
        nt, ndim = stim.shape
        blocks = []
        for d in delays:
            b = np.zeros((nt, ndim))
            if circpad and abs(d) < nt:
                b = np.roll(stim, d, axis=0)
            elif circpad:
                b = stim.copy()
            elif d >= 0:
                b[d:] = stim[:nt - d] if d < nt else 0
            else:
                b[:d] = stim[-d:] if -d < nt else 0
            blocks.append(b)
        x = np.hstack(blocks)[row_start:row_stop]
        assert out.rows == x.shape[0] and out.cols == x.shape[1]
        if zscore:
            with np.errstate(invalid="ignore", divide="ignore"):
                shift = x[:1]
                mean = shift + (x - shift).mean(0)
                sd = np.sqrt(((x - mean) ** 2).mean(0))
                x = (x - mean) * np.where(sd != 0, 1.0 / sd, 1.0)
            x = np.where(np.isnan(x), 0.0, x)
        with np.errstate(over="ignore"):
            out.a[...] = x.astype(F32)
        self.launches += 1

    def as_tensor(self, m):
        return m.a

    def gather_normalize(self, src, idx, n, mean, std, mode, eps, rows_out=None, split=False, out=None):
        rows_out = n if rows_out is None else rows_out
        rows = src.a[:n] if idx is None else src.a[np.asarray(idx[:n], dtype=np.int64)]
        x = rows - mean[None, :].astype(F32)
        with np.errstate(divide="ignore", invalid="ignore"):
            if mode == 0:
                x = x * (F32(1) / (std + F32(eps)))[None, :]
            elif mode == 1:
                sc = np.where(std > 0, F32(1.0 / math.sqrt(n - 1)) / std, np.nan).astype(F32)
                x = x * sc[None, :]
            elif mode == 3:
                x = x * np.where(std != 0, F32(1) / std, F32(1)).astype(F32)[None, :]
        if out is not None:
            assert out.rows == rows_out and out.cols == src.cols
            out.a[...] = 0
            out.a[:n] = x
            return out
        out = np.zeros((rows_out, src.cols), dtype=F32)
        out[:n] = x
        return FMat(out, split)

    # ------------------------------------------------------------------ GEMMs
    def gemm(self, A, B, alpha=1.0, Cin=None, beta=0.0, split_out=False, out=None, ld_out=None, precision="tf32x3",
             pair_out=None):
        assert precision == "f16x3" or (A.is_split and B.is_split), "3xTF32 GEMM operands must be split pairs"
        assert A.cols == B.cols and precision in ("tf32x3", "f16x3")
        assert precision == "f16x3" or not (isinstance(A, FPair) or isinstance(B, FPair))
        assert pair_out is None or (precision == "f16x3" and not split_out and out is None)
        if precision == "f16x3":  # lit_gemm_f16x3_nt: what the scaled fp16 pairs keep of the operands
            va = A.a.astype(np.float64) if isinstance(A, FPair) else self._f16_pair_value(A.a, 1)
            vb = B.a.astype(np.float64) if isinstance(B, FPair) else self._f16_pair_value(B.a, 1)
            d = alpha * (va @ vb.T)
            self.f16_gemms = getattr(self, "f16_gemms", 0) + 1
        else:
            d = alpha * (A.a.astype(np.float64) @ B.a.astype(np.float64).T)
        if Cin is not None:
            d = d + beta * Cin.a.astype(np.float64)
        self.launches += 1
        self.gemm_flops += 2.0 * A.rows * B.rows * A.cols
        if pair_out is not None:  # lit_gemm_f16x3_nt_pairout
            self.pair_out_gemms = getattr(self, "pair_out_gemms", 0) + 1
            return FPair(self._pair_with_scales(d.astype(F32), pair_out[0]), pair_out)
        return FMat(d.astype(F32), split_out)

    @staticmethod
    def _f16_pair_value(a, rows_per_group):
        """What lit_split_f16 keeps of `a`: hi = fp16(s a), lo = fp16(s a - hi), one power-of-two scale per row
        group with the group maximum in [2^14, 2^15); returns (hi + lo) / s as float64."""
        a = np.asarray(a, dtype=F32)
        n_groups = -(-a.shape[0] // rows_per_group)
        out = np.empty(a.shape, dtype=np.float64)
        for g in range(n_groups):
            blk = a[g * rows_per_group:(g + 1) * rows_per_group]
            m = float(np.abs(blk).max()) if blk.size else 0.0
            s = F32(1.0)
            if m > 0 and np.isfinite(m):
                s = F32(np.ldexp(1.0, int(np.clip(15 - np.frexp(F32(m))[1], -100, 100))))
            y = (blk * s).astype(F32)
            with np.errstate(over="ignore"):
                hi = y.astype(np.float16)
                lo = (y - hi.astype(F32)).astype(np.float16)
            out[g * rows_per_group:(g + 1) * rows_per_group] = (hi.astype(np.float64) + lo.astype(np.float64)) / float(s)
        return out

    def gemm_corr(self, A, B, n_groups, rows_per_group, Yz, precision="tf32x3"):
        """lit_gemm_corr_series: B is a split FMat of n_groups alpha groups, or a FakeSeriesStack (compact form)."""
        stack = B if isinstance(B, FakeSeriesStack) else None
        n_plain, n_st = n_groups, 0
        if stack is not None:
            assert stack.rows_pad == rows_per_group and stack.n_cheb + len(stack.slot_series) == n_groups
            B, n_plain, n_st = stack.mat, stack.n_cheb, stack.n_tiles
        assert precision in ("tf32x3", "f16x3")
        assert B.is_split and (A.is_split or precision == "f16x3")  # fp16 pairs are re-split from either form
        assert rows_per_group % self.TILE_N == 0 and B.rows == n_plain * rows_per_group + n_st * self.TILE_N
        assert Yz.rows == rows_per_group and Yz.cols == A.rows and A.cols == B.cols
        assert precision in ("tf32x3", "f16x3")
        assert not isinstance(A, FPair) or precision == "f16x3"
        if precision == "f16x3":
            va = A.a.astype(np.float64) if isinstance(A, FPair) else self._f16_pair_value(A.a, 1)
            acc = (va @ self._f16_pair_value(B.a, self.TILE_N).T).astype(F32)
        else:
            acc = (A.a.astype(np.float64) @ B.a.astype(np.float64).T).astype(F32)  # [voxel][stacked row]
        tpg = rows_per_group // self.PART_N
        dot = np.zeros((max(n_plain * tpg, 1), A.rows), dtype=F32)
        ssq = np.zeros_like(dot)
        for g in range(n_plain):
            for t in range(tpg):
                c0 = g * rows_per_group + t * self.PART_N
                blk = acc[:, c0:c0 + self.PART_N].astype(np.float64)
                yz = Yz.a[t * self.PART_N:(t + 1) * self.PART_N].astype(np.float64).T
                dot[g * tpg + t] = (blk * yz).sum(1)
                ssq[g * tpg + t] = (blk * blk).sum(1)
        series = None
        if n_st:
            # series tiles: part = 32 time points; columns [q][32] of the 128-wide half; 14 sums per (part, voxel)
            series = np.zeros((2 * n_st * 14, A.rows), dtype=F32)
            pairs = [(q, r) for q in range(4) for r in range(q, 4)]
            for part in range(2 * n_st):
                c0 = n_plain * rows_per_group + part * self.PART_N
                T = [acc[:, c0 + q * 32:c0 + (q + 1) * 32].astype(np.float64) for q in range(4)]
                yz = Yz.a[part * 32:(part + 1) * 32].astype(np.float64).T
                for q in range(4):
                    series[part * 14 + q] = (T[q] * yz).sum(1)
                for j, (q, r) in enumerate(pairs):
                    series[part * 14 + 4 + j] = (T[q] * T[r]).sum(1)
        self.launches += 1
        self.gemm_flops += 2.0 * A.rows * B.rows * A.cols
        return {"dot": dot, "ssq": ssq, "n_tiles": n_plain * tpg, "series": series, "stack": stack}

    # ------------------------------------------------------------------ eig
    def syevd(self, G, lam=None):
        a = G.a.astype(np.float64)
        a = np.tril(a) + np.tril(a, -1).T  # cuSOLVER reads the lower triangle
        w, v = np.linalg.eigh(a)
        G.a[...] = v.T.astype(F32)
        self.eig_calls += 1
        self.eig_sizes.append(G.rows)
        return w.astype(F32)

    def syevd_async(self, G):
        self.pending += 1
        snapshot = G.a.copy()
        lam = self.syevd(G)
        # emulate asynchrony: results become visible only after wait()
        result = (G.a.copy(), lam.copy())
        G.a[...] = np.nan
        lam_out = np.full_like(lam, np.nan)
        return lam_out, (G, lam_out, result, snapshot)

    def wait(self, ticket):
        G, lam_out, (vt, lam), _ = ticket
        G.a[...] = vt
        lam_out[...] = lam
        self.pending -= 1

    # ------------------------------------------------------------------ eigendecomposition-free inner solver
    def lambda_max(self, G, steps=96):
        a = G.a.astype(np.float64)
        return np.array([np.linalg.eigvalsh(0.5 * (a + a.T))[-1]])

    @staticmethod
    def solver_partition(lam_max, a2_list, series_ratio=60.0):
        series = [j for j, a2 in enumerate(a2_list) if a2 >= series_ratio * lam_max]
        cheb = [j for j in range(len(a2_list)) if j not in series]
        return cheb, series

    def solver_block_rows(self, n_rows, lam_max, a2_list, series_ratio=60.0):
        cheb, series = self.solver_partition(lam_max, a2_list, series_ratio)
        return (len(cheb) + (3 if series else 0)) * n_rows

    def lbo_prepare(self, Pv, Vt, lam_o, a2_min, steps=48, lanczos=True):
        """DeviceOps.lbo_prepare: B = X_R V, E = B diag(1/(lam_o + a2)), H = E B^T, lambda_max(H)."""
        assert Pv.is_split and Vt.is_split
        B = (Pv.a.astype(np.float64) @ Vt.a.astype(np.float64).T).astype(F32)
        E = self._lbo_scaled(B, lam_o, a2_min)
        H = (E.astype(np.float64) @ B.astype(np.float64).T).astype(F32)
        self.lbo_prepared = getattr(self, "lbo_prepared", 0) + 1
        return {"B": FMat(B, split=True), "E": FMat(E, split=True), "H": FMat(H, split=True), "a2": float(a2_min),
                "hmax_dev": self.lambda_max(FMat(H)) if lanczos else None}

    @staticmethod
    def _lbo_scaled(B, lam_o, a2):
        lam = np.asarray(lam_o, dtype=F32)
        a = F32(np.sqrt(np.float64(a2)))
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.where(lam[None, :] > 0, B / (lam[None, :] + a * a), F32(0)).astype(F32)

    def lambda_max_batched(self, mats, steps=96):
        assert len({(m.rows, m.cols) for m in mats}) <= 1, "one size per batch"
        self.lanczos_batches = getattr(self, "lanczos_batches", 0) + 1
        return np.array([self.lambda_max(m)[0] for m in mats])

    def solve_blocks(self, Gs, Pc, n_rows, lam_max, a2_list, series_ratio=60.0, lbo=None):
        """Compact per-fold solution block in the layout of DeviceOps.solve_blocks: the solutions of the small
        alphas (exact here; Chebyshev iteration on the device) followed by P_c G^q, q = 1..3.  With `lbo` the small
        alphas follow the leave-block-out identity on the outer eigendecomposition (DeviceOps._lbo_solve)."""
        assert Gs.is_split
        p = Gs.rows
        G = Gs.a.astype(np.float64)
        cheb, series = self.solver_partition(lam_max, a2_list, series_ratio)
        out = np.zeros(((len(cheb) + (3 if series else 0)) * n_rows, p), dtype=F32)
        P = Pc.a[:n_rows].astype(np.float64)
        for i, j in enumerate(cheb):
            if lbo is not None:
                assert 0.0 <= lbo["h0"] < 1.0 and lbo["lam_top"] > 0.0
                B = lbo["prep"]["B"].a
                E = self._lbo_scaled(B, lbo["lam"], float(a2_list[j])).astype(np.float64)
                H = E @ B.astype(np.float64).T
                assert np.linalg.eigvalsh(0.5 * (H + H.T))[-1] <= 1.0 - self._lbo_lo(lbo, float(a2_list[j])) + 1e-6
                Z = np.linalg.solve(np.eye(n_rows) - H, E)
                M = Z @ lbo["V"].a.astype(np.float64).T
                M = M - M.mean(0)
                self.lbo_solved = getattr(self, "lbo_solved", 0) + 1
            else:
                M = np.linalg.solve(G + float(a2_list[j]) * np.eye(p), P.T).T
            out[i * n_rows:(i + 1) * n_rows] = M.astype(F32)
        Q = P
        for q in range(3 if series else 0):
            Q = (Q @ G).astype(F32).astype(np.float64)
            out[(len(cheb) + q) * n_rows:(len(cheb) + q + 1) * n_rows] = Q
        self.solver_calls = getattr(self, "solver_calls", 0) + 1
        return FMat(out)

    def solve_blocks_many(self, jobs, series_ratio=60.0):
        """DeviceOps.solve_blocks_many: the batched direct (Cholesky) solver -- exact solves here."""
        out = []
        for job in jobs:
            assert not job["G"].is_split and not job["Pc"].is_split
            out.append(self.solve_blocks(self.split(job["G"]), job["Pc"], job["n_rows"], job["lam_max"], job["a2"],
                                         series_ratio))
            self.direct_solved = getattr(self, "direct_solved", 0) + len(
                self.solver_partition(job["lam_max"], job["a2"], series_ratio)[0])
        return out

    def check_solver(self):
        pass

    GROUP_TILE = 256

    def outer_inverses(self, G, lam_max, a2_list, series_ratio=60.0):
        """DeviceOps.outer_inverses: (G + a^2 I)^-1 per alpha (exact / the 4-term Neumann polynomial for the series)."""
        Gd = G.a.astype(np.float64)
        p = Gd.shape[0]
        cheb, series = self.solver_partition(lam_max, a2_list, series_ratio)
        out = np.zeros((len(a2_list), p, p), dtype=F32)
        for j in cheb:
            out[j] = np.linalg.inv(Gd + float(a2_list[j]) * np.eye(p)).astype(F32)
        if series:
            powers = [np.eye(p), Gd, (Gd @ Gd).astype(F32).astype(np.float64)]
            powers.append((powers[2] @ Gd).astype(F32).astype(np.float64))
            for j in series:
                a2 = float(a2_list[j])
                out[j] = sum((-1.0) ** q / a2 ** (q + 1) * powers[q] for q in range(4)).astype(F32)
        self.outer_direct = getattr(self, "outer_direct", 0) + 1
        return out

    def outer_inverse_systems(self, lam_maxs, a2_lists, series_ratio=60.0):
        return [(i, j) for i, (lm, a2) in enumerate(zip(lam_maxs, a2_lists))
                for j in self.solver_partition(lm, a2, series_ratio)[0]]

    def inverse_slot(self, inv, j):
        return [inv[j]]

    def outer_inverses_many(self, Gs, lam_maxs, a2_lists, series_ratio=60.0, owned=None):
        out = [self.outer_inverses(G, lm, a2, series_ratio) for G, lm, a2 in zip(Gs, lam_maxs, a2_lists)]
        if owned is not None:  # the slots of other ranks arrive by broadcast: poison them here
            for k, (i, j) in enumerate(self.outer_inverse_systems(lam_maxs, a2_lists, series_ratio)):
                if k not in owned:
                    out[i][j] = np.nan
        return out

    def group_plan(self, idx, n_vox, n_groups):
        tile = self.GROUP_TILE
        idx = np.asarray(idx)[:n_vox].astype(np.int64)
        cap = (-(-max(n_vox, 1) // tile) + n_groups) * tile
        pos, perm, tg = np.zeros(n_vox, np.int32), np.full(cap, -1, np.int32), np.full(cap // tile, -1, np.int32)
        start = 0
        for g in range(n_groups):
            members = np.nonzero(idx == g)[0]
            pos[members] = start + np.arange(len(members))
            perm[start:start + len(members)] = members
            n_tiles = -(-len(members) // tile)
            tg[start // tile:start // tile + n_tiles] = g
            start += n_tiles * tile
        return pos, perm, tg, cap

    def gemm_grouped(self, A, inv, tile_group, split_out=True):
        tile = self.GROUP_TILE
        out = np.zeros((A.rows, A.cols), dtype=F32)
        for t, g in enumerate(np.asarray(tile_group)):
            if g >= 0:
                out[t * tile:(t + 1) * tile] = (A.a[t * tile:(t + 1) * tile].astype(np.float64)
                                                @ inv[g].astype(np.float64).T).astype(F32)
        return FMat(out, split=split_out)

    @staticmethod
    def _lbo_lo(lbo, a2):
        """Lower end of the spectral interval the device's Chebyshev iteration assumes for I - H_a."""
        from litcoder_core_b200.device import DeviceOps
        return DeviceOps.lbo_bounds(lbo["h0"], lbo["prep"]["a2"], a2, lbo["lam_top"])[0]

    SERIES_MIN_ALPHAS = 5

    def assemble_stack(self, block, Pc, n_rows, rows_pad, lam_max, a2_list, series_ratio=60.0, series_moments=False):
        cheb, series = self.solver_partition(lam_max, a2_list, series_ratio)
        nc = len(cheb)
        Q = [Pc.a[:n_rows].astype(np.float64)] + [block.a[(nc + q) * n_rows:(nc + q + 1) * n_rows].astype(np.float64)
                                                  for q in range(3 if series else 0)]
        if series_moments and len(series) >= self.SERIES_MIN_ALPHAS:
            # lit_series_stack: row tile*256 + half*128 + q*32 + i <-> time point tile*64 + half*32 + i
            n_tiles = -(-n_rows // 64)
            out = np.zeros((nc * rows_pad + n_tiles * self.TILE_N, block.cols), dtype=F32)
            for i in range(nc):
                out[i * rows_pad:i * rows_pad + n_rows] = block.a[i * n_rows:(i + 1) * n_rows]
            for t in range(n_rows):
                tile, half, i = t // 64, (t % 64) // 32, t % 32
                for q in range(4):
                    out[nc * rows_pad + tile * 256 + half * 128 + q * 32 + i] = Q[q][t] * float(lam_max) ** -q
            coef = np.array([[(-1.0) ** q * float(lam_max) ** q / float(a2_list[j]) ** (q + 1) for q in range(4)]
                             for j in series], dtype=np.float64)
            return FakeSeriesStack(FMat(out, split=True), nc, rows_pad, n_tiles, np.asarray(cheb, dtype=np.int32),
                                   np.asarray(series, dtype=np.int32), coef)
        out = np.zeros((len(a2_list) * rows_pad, block.cols), dtype=F32)
        for i, j in enumerate(cheb):
            out[j * rows_pad:j * rows_pad + n_rows] = block.a[i * n_rows:(i + 1) * n_rows]
        for j in series:
            a2 = float(a2_list[j])
            out[j * rows_pad:j * rows_pad + n_rows] = sum((-1.0) ** q * a2 ** -(q + 1) * Q[q] for q in range(4))
        return FMat(out, split=True)

    def inverse_stack(self, Gs, Pc, n_rows, rows_pad, lam_max, a2_list, series_ratio=60.0):
        return self.assemble_stack(self.solve_blocks(Gs, Pc, n_rows, lam_max, a2_list), Pc, n_rows, rows_pad, lam_max,
                                   a2_list)

    # ------------------------------------------------------------------ ridge kernels
    @staticmethod
    def _shrink(lam, alpha_scaled_sq, singcutoff):
        keep = np.sqrt(np.maximum(lam, 0)) > singcutoff
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.where(keep, F32(1) / (lam + alpha_scaled_sq), F32(0)).astype(F32)

    def build_alpha_stack(self, L, n_rows, rows_pad, lam, alphas_dev, n_alphas, normalpha, singcutoff):
        s = math.sqrt(max(float(lam[-1]), 0.0)) if normalpha else 1.0
        Lc = L.a[:n_rows] - L.a[:n_rows].astype(np.float64).mean(0).astype(F32)[None, :]
        out = np.zeros((n_alphas * rows_pad, L.cols), dtype=F32)
        for a in range(n_alphas):
            an = float(alphas_dev[a]) * float(F32(s))
            d = self._shrink(lam, F32(an * an), singcutoff)
            out[a * rows_pad:a * rows_pad + n_rows] = Lc * d[None, :]
        return FMat(out, split=True)

    def scale_rows_by_alpha(self, Z, lam, alpha_v, normalpha, singcutoff):
        s = F32(math.sqrt(max(float(lam[-1]), 0.0))) if normalpha else F32(1)
        an = (alpha_v[:Z.rows].astype(F32) * s).astype(F32)
        a2 = (an * an).astype(F32)
        keep = np.sqrt(np.maximum(lam, 0)) > singcutoff
        with np.errstate(divide="ignore", invalid="ignore"):
            out = np.where(keep[None, :], Z.a / (lam[None, :] + a2[:, None]), F32(0))
        return FMat(out.astype(F32), split=True)

    @staticmethod
    def _inner_score(d, q, metric, raw, n_rows, eps, resp_std):
        with np.errstate(divide="ignore", invalid="ignore"):
            if metric == 0:
                c = (d / F32(n_rows)) / (np.sqrt(q / F32(n_rows - 1)) + F32(eps))
            else:
                qvar = (resp_std * resp_std).astype(F32)
                resvar = (qvar * F32(n_rows - 1) - 2 * d + q) / F32(n_rows - 1)
                rsq = 1 - resvar / qvar
                c = np.sqrt(np.abs(rsq)) * np.sign(rsq)
        return c.astype(F32) if raw else np.nan_to_num(c.astype(F32))

    def corr_finalize(self, parts, tiles_per_group, n_groups, n_vox, n_rows, eps, corr, accumulate, metric=0,
                      resp_std=None):
        raw, metric = bool(metric & 2), metric & 1
        st = parts.get("stack")
        n_plain = n_groups if st is None else st.n_cheb
        for g in range(n_plain):
            d = parts["dot"][g * tiles_per_group:(g + 1) * tiles_per_group].sum(0, dtype=F32)
            q = parts["ssq"][g * tiles_per_group:(g + 1) * tiles_per_group].sum(0, dtype=F32)
            c = self._inner_score(d, q, metric, raw, n_rows, eps, resp_std)
            slot = g if st is None else int(st.slot_cheb[g])
            corr.a[slot] = corr.a[slot] + c if accumulate else c
        if st is not None and st.n_tiles:
            # lit_corr_finalize_series: 4-term combinations of the 14 sums, in float64
            S = parts["series"].astype(np.float64).reshape(2 * st.n_tiles, 14, -1).sum(0)
            pairs = [(q, r) for q in range(4) for r in range(q, 4)]
            for a, slot in enumerate(st.slot_series):
                cf = st.coef[a]
                d = sum(cf[q] * S[q] for q in range(4))
                q2 = sum((1.0 if q == r else 2.0) * cf[q] * cf[r] * S[4 + j] for j, (q, r) in enumerate(pairs))
                c = self._inner_score(d.astype(F32), q2.astype(F32), metric, raw, n_rows, eps, resp_std)
                corr.a[slot] = corr.a[slot] + c if accumulate else c

    def argmax_alpha(self, corr_sum, n_folds, alphas_dev, want_sums):
        mean = (corr_sum.a / F32(n_folds)).astype(F32)
        best = np.argmax(mean, axis=0).astype(np.int32)
        alpha_v = np.asarray(alphas_dev, dtype=F32)[best]
        sums = mean.astype(np.float64).sum(1) if want_sums else None
        return best, alpha_v, sums

    # ------------------------------------------------------------------ test statistics
    def pearson_finalize(self, parts, n_vox, n_samples, p_round_f32):
        from scipy.special import betainc

        d = parts["dot"].sum(0, dtype=F32)
        q = parts["ssq"].sum(0, dtype=F32)
        with np.errstate(divide="ignore", invalid="ignore"):
            r = d / np.sqrt(q)
        bad = np.isnan(r)
        r = np.clip(np.where(bad, F32(0), r), -1, 1).astype(F32)
        ab = n_samples / 2.0 - 1.0
        p = np.minimum(2.0 * betainc(ab, ab, 0.5 * (1.0 - np.abs(r.astype(np.float64)))), 1.0) if n_samples >= 3 \
            else np.ones(len(r))
        p[np.abs(r) >= 1.0] = 0.0
        if p_round_f32:
            p = p.astype(F32).astype(np.float64)
        p[bad] = 1.0
        return r, p

    def bh_fdr(self, p, n, alpha):
        p = np.asarray(p[:n], dtype=np.float64)
        key = np.where(np.isnan(p), np.inf, p)
        order = np.lexsort((np.arange(n), key))
        ps = key[order]
        ecdf = np.arange(1, n + 1) / float(n)
        rej_sorted = ps <= ecdf * alpha
        kmax = np.max(np.nonzero(rej_sorted)[0]) if rej_sorted.any() else -1
        adj_sorted = np.minimum(np.minimum.accumulate((ps / ecdf)[::-1])[::-1], 1.0)
        reject = np.zeros(n, dtype=np.uint8)
        padj = np.empty(n)
        reject[order] = (np.arange(n) <= kmax).astype(np.uint8)
        padj[order] = adj_sorted
        return reject, padj, np.array([kmax + 1], dtype=np.int32)

    def stack_vectors(self, vecs, n, dtype="f64"):
        return np.stack([np.asarray(v[:n]) for v in vecs])

    def fisher(self, p_stack, n_folds, n_vox, p_round_f32):
        P = np.asarray(p_stack, dtype=np.float64)[:n_folds, :n_vox]
        with np.errstate(divide="ignore"):
            x = -np.log(P).sum(0)
        term = np.ones_like(x)
        s = np.ones_like(x)
        with np.errstate(invalid="ignore", over="ignore"):
            for j in range(1, n_folds):
                term = term * x / j
                s = s + term
            out = np.minimum(np.exp(-x) * s, 1.0)
        if p_round_f32:
            out = out.astype(F32).astype(np.float64)
        out[(P <= 0).any(0)] = 0.0
        out[(P == 1.0).all(0)] = 1.0
        return out

    # ------------------------------------------------------------------ feature construction
    def fir_make_delayed(self, stim, delays, circpad):
        stim = np.asarray(stim)
        nt, ndim = stim.shape
        out = np.zeros((nt, ndim * len(delays)))
        for i, d in enumerate(delays):
            for t in range(nt):
                ts = t - d
                if 0 <= ts < nt:
                    out[t, i * ndim:(i + 1) * ndim] = stim[ts]
                elif circpad:
                    src = ts % nt if abs(d) < nt else t
                    out[t, i * ndim:(i + 1) * ndim] = stim[src]
        return out

    def resample(self, kind, data, data_times, tr_times, window, cutoff, flag_a, flag_b, lo, hi):
        data = np.asarray(data, dtype=np.float64)
        n_tr = len(tr_times)
        rectify = kind == "lanczos" and flag_a
        parts = [np.clip(data, -np.inf, 0), np.clip(data, 0, np.inf)] if rectify else [data]
        outs = []
        for part in parts:
            out = np.zeros((n_tr, data.shape[1]))
            for i in range(n_tr):
                j0, j1 = (0, len(data_times)) if lo is None else (int(lo[i]), int(hi[i]))
                dt = tr_times[i] - data_times[j0:j1]
                with np.errstate(divide="ignore", invalid="ignore"):
                    if kind == "lanczos":
                        t = dt * cutoff
                        w = window * np.sin(np.pi * t) * np.sin(np.pi * t / window) / (np.pi ** 2 * t ** 2)
                        w[t == 0] = 1.0
                        w[np.abs(t) > window] = 0.0
                    else:
                        w = 2 * cutoff * np.sin(2 * np.pi * cutoff * dt) / (2 * np.pi * cutoff * dt + 1e-20)
                        w[np.abs(dt) > window / (2 * cutoff)] = 0
                        if flag_a:
                            w[dt < 0] = 0
                        if flag_b and w.sum() != 0.0:
                            w = w / w.sum()
                out[i] = w @ part[j0:j1]
            outs.append(out)
        return np.hstack(outs)

    def lanczos_downsample(self, data, data_times, tr_times, window, cutoff, rectify, lo, hi):
        return self.resample("lanczos", data, data_times, tr_times, window, cutoff, rectify, False, lo, hi)

    def csr_rows_apply(self, data, row_ptr, col_idx, weights, mean):
        data = np.asarray(data)
        if data.dtype not in (np.float32, np.float64) or weights is not None:
            data = data.astype(np.float64)
        out = np.zeros((len(row_ptr) - 1, data.shape[1]))
        for r in range(len(row_ptr) - 1):
            e = slice(int(row_ptr[r]), int(row_ptr[r + 1]))
            if e.stop > e.start:
                rows = data[np.asarray(col_idx[e], dtype=np.int64)]
                if weights is not None:
                    rows = rows * np.asarray(weights[e])[:, None]
                out[r] = rows.mean(0) if mean else rows.sum(0)
        return out

    def gabor_downsample(self, data, data_times, tr_times, freqs, sigma):
        data = np.asarray(data, dtype=np.float64)
        out = np.zeros((len(tr_times), data.shape[1] * len(freqs)))
        phase = np.exp(1j * 2 * np.pi * np.outer(np.asarray(freqs, dtype=np.float64), data_times))  # (F, n_s)
        for i, t in enumerate(tr_times):
            g = np.exp(-0.5 * (data_times - t) ** 2 / (2 * sigma ** 2))
            out[i] = np.abs((phase * g[None, :]) @ data).T.reshape(-1)  # (F, D) -> d-major, f-minor
        return out
