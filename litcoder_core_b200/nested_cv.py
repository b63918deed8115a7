"""Drop-in for the reference's nested-CV ridge model (encoding/models/nested_cv.py).

    NestedCVModel(model_name="ridge_regression").fit_predict(features, targets, ...)   nested_cv.py:14-331
    fit_nested_cv(features=X, targets=Y, ...)                                          README.md:137,212-226

Same keyword arguments, same return triple ``(metrics, weights, best_alphas)``, same metrics keys
(nested_cv.py:501-528, 558-614), same fold semantics (folding.py) -- but the arithmetic runs on
a B200 through the C ABI of liblitridge.so (engine.py / device.py).  Host code here only builds
fold index lists, moves arrays across the PCIe boundary and assembles the metrics dictionary.

Multi-GPU: when ``torch.distributed`` is initialised (one process per GPU, NCCL), the voxels
(columns of ``targets``) are split into contiguous blocks, one per rank; ranks exchange only
(i) the per-alpha score sums for ``single_alpha=True`` (all-reduce of A doubles per outer fold)
and (ii) the per-voxel r / p / alpha vectors and, optionally, the weight blocks (all-gather at
the end).  Every rank returns the full result.

There is no CPU fallback: ``use_gpu=False`` is accepted for signature compatibility and ignored
(with a log line); without a CUDA device or the built library the call raises.
"""
from __future__ import annotations

import logging
import os
import time
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from .engine import FoldPlan, RidgeConfig, RidgeCVEngine, SingleProcess, SolverAccuracyError
from .folding import create_folds

logger = logging.getLogger(__name__)


# ----------------------------------------------------------------------------------------------
# communicator over torch.distributed (NCCL on GPUs, gloo in the CPU tests)
# ----------------------------------------------------------------------------------------------
class TorchDistComm:
    """Host-array collectives for the engine.  NCCL needs device tensors, gloo host tensors."""

    def __init__(self, group=None, device=None):
        import torch
        import torch.distributed as dist

        self._torch, self._dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        backend = dist.get_backend(group)
        self._dev = device if (backend == "nccl" and device is not None) else (
            torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu"))
        self.device_collectives = backend == "nccl"  # tensors can be exchanged without leaving the GPU

    def all_gather_cols_device(self, block, counts: Sequence[int], ld: int):
        """All-gather of per-rank column blocks (rows x counts[r], CUDA tensors) into one (rows x ld) device
        matrix whose first sum(counts) columns are the concatenation -- NCCL over NVLink, no host hop."""
        t = self._torch
        rows, width = block.shape[0], max(counts)
        pad = t.zeros((rows, width), dtype=block.dtype, device=block.device)
        pad[:, : block.shape[1]] = block
        out = t.empty((self.world, rows, width), dtype=block.dtype, device=block.device)
        self._dist.all_gather_into_tensor(out, pad, group=self.group)
        full = t.empty((rows, ld), dtype=block.dtype, device=block.device)
        c0 = 0
        for r, c in enumerate(counts):
            full[:, c0:c0 + c] = out[r][:, :c]
            c0 += c
        return full

    def all_reduce_sum(self, arr: np.ndarray) -> np.ndarray:
        t = self._torch.from_numpy(np.ascontiguousarray(arr)).to(self._dev)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def all_reduce_sum_inplace(self, buffers) -> None:
        """Sum the ranks' buffers in place (torch tensors on the communicator's device, or NumPy arrays with gloo)."""
        for b in buffers:
            t = b if self._torch.is_tensor(b) else self._torch.from_numpy(b)
            self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)

    def broadcast_inplace(self, buffers, src: int) -> None:
        """Broadcast rank `src`'s buffers into everybody's (torch tensors on the communicator's device, or
        NumPy arrays with the gloo backend), in place."""
        for b in buffers:
            t = b if self._torch.is_tensor(b) else self._torch.from_numpy(b)
            self._dist.broadcast(t, src=self._dist.get_global_rank(self.group, src) if self.group is not None else src,
                                 group=self.group)

    def broadcast_async(self, buffers, src: int) -> list:
        """broadcast_inplace without waiting: returns the work handles; `wait_all` makes the current stream (NCCL) or
        the host (gloo) wait for them.  Lets the solution blocks of all inner folds travel while the GEMMs run."""
        works = []
        for b in buffers:
            t = b if self._torch.is_tensor(b) else self._torch.from_numpy(b)
            works.append(self._dist.broadcast(
                t, src=self._dist.get_global_rank(self.group, src) if self.group is not None else src, group=self.group,
                async_op=True))
        return works

    @staticmethod
    def wait_all(works) -> None:
        for w in works:
            w.wait()

    def all_gather_concat(self, arr: np.ndarray, counts: Optional[Sequence[int]] = None) -> np.ndarray:
        """Concatenate the ranks' arrays along the LAST axis; counts[r] = last-axis length on rank r."""
        arr = np.ascontiguousarray(arr)
        if counts is None:
            counts = [arr.shape[-1]] * self.world
        width = max(counts)
        pad = np.zeros(arr.shape[:-1] + (width,), dtype=arr.dtype)
        pad[..., : arr.shape[-1]] = arr
        t = self._torch.from_numpy(pad).to(self._dev)
        outs = [self._torch.empty_like(t) for _ in range(self.world)]
        self._dist.all_gather(outs, t, group=self.group)
        return np.concatenate([o.cpu().numpy()[..., : counts[r]] for r, o in enumerate(outs)], axis=-1)


def default_comm():
    """TorchDistComm when a process group with more than one rank is initialised, else single-process."""
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return SingleProcess()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return TorchDistComm()
    return SingleProcess()


def shard_bounds(n_vox: int, world: int) -> List[int]:
    """Contiguous voxel blocks: boundaries b[0..world]; block sizes are multiples of 128 (one GEMM
    M tile) except the last, so that every rank's shard starts on a tile boundary."""
    per = -(-n_vox // world)
    per = -(-per // 128) * 128
    return [min(r * per, n_vox) for r in range(world + 1)]


# ----------------------------------------------------------------------------------------------
# metrics dictionaries (nested_cv.py:480-616)
# ----------------------------------------------------------------------------------------------
def _summary(values: np.ndarray, infix: str = "") -> Dict[str, float]:
    return {
        f"median_{infix}score": float(np.median(values)),
        f"mean_{infix}score": float(np.mean(values)),
        f"min_{infix}score": float(np.min(values)),
        f"max_{infix}score": float(np.max(values)),
    }


def _metrics(corr: np.ndarray, pvals: np.ndarray, padj: np.ndarray, sig: np.ndarray, best_alphas: np.ndarray,
             majority: Optional[np.ndarray] = None) -> Dict[str, Any]:
    n = len(corr)
    n_sig = int(np.sum(sig))
    m: Dict[str, Any] = {
        "median_score": float(np.median(corr)),
        "mean_score": float(np.mean(corr)),
        "std_score": float(np.std(corr)),
        "min_score": float(np.min(corr)),
        "max_score": float(np.max(corr)),
        "best_alphas": best_alphas.tolist(),
        "correlations": corr.tolist(),
        "p_values": pvals.tolist(),
        "corrected_p_values": padj.tolist(),
        "significant_mask": sig.tolist(),
    }
    if majority is not None:
        m["majority_significant_mask"] = majority.tolist()
    m["n_significant"] = n_sig
    if majority is not None:
        m["n_majority_significant"] = int(np.sum(majority))
    m["percent_significant"] = float(n_sig / n * 100)
    if majority is not None:
        m["percent_majority_significant"] = float(m["n_majority_significant"] / n * 100)
    if n_sig > 0:
        m.update(_summary(corr[sig], "significant_"))
    if majority is not None and m["n_majority_significant"] > 0:
        m.update(_summary(corr[majority], "majority_significant_"))
    return m


# ----------------------------------------------------------------------------------------------
class NestedCVModel:
    """Nested cross-validated ridge regression from (delayed) stimulus features to voxels."""

    def __init__(self, model_name: str = "ridge_regression", ops=None, comm=None):
        self.model_name = model_name
        self._ops = ops
        self._comm = comm
        self.last_timings: Dict[str, float] = {}  # per-phase milliseconds of the most recent fit
        self.last_stats: Dict[str, Any] = {}
        self.last_fold_results: Dict[str, np.ndarray] = {}
        self.record_inner_scores = False  # True: last_fold_results["inner_scores"] = (folds x alphas x V) score curves

    # -- plumbing --------------------------------------------------------------------------------
    @staticmethod
    def _single_alpha_values(a_f: np.ndarray, alphas, dtype) -> np.ndarray:
        """single_alpha=True: the per-fold alpha vectors (n_folds x V; the device holds them as float32) with the grid
        value itself, in the dtype the reference returns (see fit_predict)."""
        grid = np.asarray(alphas, dtype=np.float64)
        idx = np.argmin(np.abs(grid[None, :] - a_f[:, :1].astype(np.float64)), axis=1)
        return np.repeat(grid[idx].astype(dtype)[:, None], a_f.shape[1], axis=1)

    def _get_ops(self):
        if self._ops is None:
            from .device import default_ops

            self._ops = default_ops()
        return self._ops

    @staticmethod
    def _is_torch(x) -> bool:
        return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")

    def _to_device(self, ops, arr, c0: int = 0, c1: Optional[int] = None):
        """Host ndarray (any float dtype) or resident CUDA float32 tensor -> device Mat of columns [c0, c1)."""
        if self._is_torch(arr):
            if arr.is_cuda:
                c1 = arr.shape[1] if c1 is None else c1
                if c0 == 0 and c1 == arr.shape[1]:
                    return ops.wrap(arr)
                return ops.wrap_view(arr, c0, c1)
            arr = arr.detach().numpy()
        return ops.upload_matrix(np.asarray(arr), c0, c1)

    # -- public API --------------------------------------------------------------------------------
    def fit_predict(
        self,
        features: np.ndarray,
        targets: np.ndarray,
        X_test: Optional[np.ndarray] = None,
        y_test: Optional[np.ndarray] = None,
        groups: Optional[np.ndarray] = None,
        folding_type: str = "chunked",
        n_outer_folds: int = 5,
        n_inner_folds: int = 5,
        chunk_length: int = 20,
        alphas: Optional[List[float]] = None,
        alpha_fdr: float = 0.05,
        use_gpu: bool = True,
        single_alpha: bool = False,
        normalpha: bool = True,
        use_corr: bool = True,
        normalize_features: bool = False,
        normalize_targets: bool = False,
        singcutoff: float = 1e-10,
        gather_weights: bool = True,
        device_outputs: bool = False,
        inner_solver: str = "auto",
        corr_precision: str = "auto",
        row_shard_gram: bool = False,
    ) -> Tuple[Dict[str, Union[float, List[float], List[bool]]], np.ndarray, np.ndarray]:
        """Fit with nested CV (or inner CV + a given test set), per-voxel or single alpha, FDR correction.

        Arguments and return value as the reference (nested_cv.py:18-70).  ``gather_weights`` (extension,
        multi-GPU only): False returns this rank's (p x V_rank) weight block instead of the full matrix.
        ``device_outputs`` (extension): leave the weights on the device and return this rank's (p x V_rank)
        block as a torch CUDA tensor (no D2H of the 1.2 GB weight matrix).
        ``inner_solver`` (extension): "eig" decomposes every inner-fold Gram with cuSOLVER syevd; "chebyshev"
        solves the inner folds with GEMMs only (Lanczos lambda_max + Chebyshev iteration / Neumann series);
        "auto" picks chebyshev when alphas are normalised and >= 0.05, else eig.
        ``corr_precision`` (extension): operand format of the fused inner-CV prediction + correlation GEMM:
        "tf32x3" (3xTF32 split pairs) or "f16x3" (scaled fp16 split pairs: same 2^-22 product accuracy, twice the
        tensor-core rate); "auto" = "f16x3".
        ``row_shard_gram`` (extension, multi-GPU only): every rank forms the outer-fold Gram (or kernel matrix) over
        its own 1/G of the contraction axis and the ranks all-reduce the partial sums (NCCL) instead of each forming
        the whole Gram -- for wide designs (BASELINE config 5); results then agree across rank counts to fp32
        rounding of that sum instead of bit for bit.
        """
        t_start = time.perf_counter()
        if alphas is None:
            alphas = np.logspace(-1, 8, 10)  # nested_cv.py:80-81
        # single_alpha: the reference builds torch.tensor([alphas[j]] * V) (nested_cv.py:399-401), whose dtype follows
        # the element -- float64 for an element of a float64 ndarray, torch's default float32 for a Python float
        single_dtype = np.float64 if (len(alphas) and isinstance(alphas[0], np.float64)) else np.float32
        alphas = [float(a) for a in alphas]
        if not use_gpu:
            logger.info("use_gpu=False ignored: litcoder_core_b200 always runs on the B200 (no CPU path)")
        n_samples, n_vox = features.shape[0], targets.shape[1]
        if targets.shape[0] != n_samples:
            raise ValueError("features and targets must have the same number of rows")
        train_test_mode = X_test is not None and y_test is not None  # nested_cv.py:103
        logger.info("Folding type: %s", folding_type)

        # ---- fold index lists (host; reference semantics incl. the positional `groups` -> trim_size slot) ----
        plans: List[FoldPlan] = []
        if train_test_mode:
            inner = create_folds(n_samples, folding_type, n_inner_folds, chunk_length, groups)  # nested_cv.py:130-132
            plans.append(FoldPlan(np.arange(n_samples, dtype=np.int64), np.arange(X_test.shape[0], dtype=np.int64),
                                  [(np.asarray(a, dtype=np.int64), np.asarray(b, dtype=np.int64)) for a, b in inner]))
        else:
            if groups is not None and folding_type == "group":
                outer = create_folds(n_samples, "group", n_outer_folds, groups=groups)
            else:
                outer = create_folds(n_samples, folding_type, n_outer_folds, chunk_length, groups)
            for tr, te in outer:
                tr = np.asarray(tr, dtype=np.int64)
                te = np.asarray(te, dtype=np.int64)
                if groups is not None and folding_type == "group":
                    inner = create_folds(len(tr), "group", n_inner_folds, groups=[groups[i] for i in tr])
                else:
                    inner = create_folds(len(tr), folding_type, n_inner_folds, chunk_length)
                # inner indices are positions within the outer training rows (nested_cv.py:200-201,371-374)
                plans.append(FoldPlan(tr, te, [(tr[np.asarray(a, dtype=np.int64)], tr[np.asarray(b, dtype=np.int64)])
                                               for a, b in inner]))

        cfg = RidgeConfig(alphas=alphas, alpha_fdr=alpha_fdr, single_alpha=single_alpha, normalpha=normalpha,
                          use_corr=use_corr, normalize_features=normalize_features,
                          normalize_targets=normalize_targets, singcutoff=singcutoff, n_outer_folds=n_outer_folds,
                          inner_solver=inner_solver)
        if inner_solver == "auto":
            inner_solver = cfg.inner_solver = os.environ.get("LIT_INNER_SOLVER", "auto")  # development override
        if inner_solver not in ("auto", "eig", "chebyshev"):
            raise ValueError(f"Unknown inner_solver: {inner_solver}")
        if corr_precision == "auto":
            corr_precision = os.environ.get("LIT_CORR_PRECISION", "f16x3")  # development override
        if corr_precision not in ("tf32x3", "f16x3"):
            raise ValueError(f"Unknown corr_precision: {corr_precision}")
        cfg.corr_precision = corr_precision
        cfg.row_shard_gram = bool(row_shard_gram)
        cfg.record_scores = bool(self.record_inner_scores)
        cfg.voxel_gemm_precision = os.environ.get("LIT_VOXEL_GEMM", corr_precision)  # development override
        cfg.series_moments = os.environ.get("LIT_SERIES_MOMENTS", "1") != "0"  # development override
        cfg.leave_block_out = os.environ.get("LIT_LEAVE_BLOCK_OUT", "1") != "0"  # development override

        # ---- H2D: X replicated, this rank's voxel block of Y (nested_cv.py:99-100) ----
        ops = self._get_ops()
        comm = self._comm if self._comm is not None else default_comm()
        bounds = shard_bounds(n_vox, comm.world)
        c0, c1 = bounds[comm.rank], bounds[comm.rank + 1]
        counts = [bounds[r + 1] - bounds[r] for r in range(comm.world)]
        ops.reset_counters()
        with ops.timed("h2d"):
            X = self._to_device(ops, features)
            Xt = self._to_device(ops, X_test) if train_test_mode else None
        # The design side (Grams, lambda_max, eigendecompositions) needs X only: the responses -- 97 % of the
        # input bytes -- travel on a copy stream meanwhile and are awaited right before their first use.
        if isinstance(targets, np.ndarray) and not train_test_mode and hasattr(ops, "upload_matrix_bg"):
            Y, y_ready = ops.upload_matrix_bg(targets, c0, c1)  # pageable arrays: staged by a helper thread
            Yt = None
        else:
            with ops.copy_stream() as y_ready:
                Y = self._to_device(ops, targets, c0, c1)
                Yt = self._to_device(ops, y_test, c0, c1) if train_test_mode else None

        cfg.direct_solver = os.environ.get("LIT_DIRECT_SOLVER", "1") != "0"  # development override
        cfg.direct_outer = os.environ.get("LIT_DIRECT_OUTER", "1") != "0"  # development override
        cfg.producer_pairs = os.environ.get("LIT_PRODUCER_PAIRS", "1") != "0"  # development override
        engine = RidgeCVEngine(ops, comm)
        with ops.timed("fit"):
            res = engine.fit_shard(X, Y, plans, cfg, X_test=Xt, Y_test=Yt, n_vox_total=n_vox, y_ready=y_ready)
        ops.check_eig()
        # a-posteriori check of the GEMM-only inner solves (probe residuals + pivot flags, read back here): a rejected
        # solve on ANY rank sends every rank back through the eigendecomposition route
        failed = 0.0
        try:
            ops.check_solver()
        except SolverAccuracyError as e:
            logger.warning("%s -- refitting with inner_solver='eig'", e)
            failed = 1.0
        if comm.world > 1:
            failed = float(comm.all_reduce_sum(np.array([failed]))[0])
        if failed:
            del res, engine
            cfg.inner_solver = "eig"
            engine = RidgeCVEngine(ops, comm)
            res = engine.fit_shard(X, Y, plans, cfg, X_test=Xt, Y_test=Yt, n_vox_total=n_vox)
            ops.check_eig()

        # ---- per-voxel vectors back to the host, gathered over ranks ----
        with ops.timed("d2h"):
            r_f = np.stack([ops.download(v)[: c1 - c0] for v in res.r]).astype(np.float32)
            p_f = np.stack([ops.download(v)[: c1 - c0] for v in res.p]).astype(np.float64)
            a_f = np.stack([ops.download(v)[: c1 - c0] for v in res.alpha]).astype(np.float32)
            scores = np.stack(res.scores) if res.scores else None
            if comm.world > 1:
                r_f = comm.all_gather_concat(r_f, counts)
                p_f = comm.all_gather_concat(p_f, counts)
                a_f = comm.all_gather_concat(a_f, counts)
                if scores is not None:
                    scores = comm.all_gather_concat(scores, counts)
            t_w0 = time.perf_counter()
            Wd = engine.weights_matrix(res)  # (p x V_rank) float32 on the device
            if device_outputs:
                W = Wd.hi[:, : Wd.cols]
            elif comm.world > 1 and gather_weights and getattr(comm, "device_collectives", False):
                ld = -(-n_vox // 32) * 32
                full = comm.all_gather_cols_device(Wd.hi[:, : Wd.cols], counts, ld)
                W = ops.download_matrix(type(Wd)(full, None, Wd.rows, n_vox))
            elif comm.world > 1 and gather_weights:
                W = comm.all_gather_concat(ops.download_matrix(Wd), counts)
            elif hasattr(ops, "download_matrix_start"):
                W = None  # this rank's block: copied while the host builds the metrics (below)
            else:
                W = ops.download_matrix(Wd)
        del res
        t_w = (time.perf_counter() - t_w0) * 1e3

        t_s0 = time.perf_counter()
        with ops.timed("stats"):
            masks, comb_p, sig, padj = engine.significance(p_f, cfg)
        w_finish = None
        if W is None:
            # after the device statistics (their small copies must not queue behind 1.2 GB of weights), before the
            # host-only part: the D2H of the weights runs while the metrics dictionary is built
            with ops.timed("d2h"):
                w_finish = ops.download_matrix_start(Wd)

        if train_test_mode:
            best = self._single_alpha_values(a_f, alphas, single_dtype)[0] if single_alpha else a_f[0]
            metrics = _metrics(r_f[0].astype(np.float64), p_f[0], padj, sig, best)
        else:
            if comb_p is None:  # a single outer fold: Fisher's method on one p-value is the identity
                comb_p = p_f[0]
            corr = np.mean(r_f, axis=0)  # nested_cv.py:276
            majority = np.sum(np.stack(masks), axis=0) >= (n_outer_folds // 2 + 1)  # nested_cv.py:288-290
            best = np.mean(self._single_alpha_values(a_f, alphas, single_dtype) if single_alpha else a_f,
                           axis=0)  # nested_cv.py:293
            metrics = _metrics(corr.astype(np.float64), comb_p, padj, sig, best, majority)

        # per-outer-fold vectors of the most recent fit (extension: what the parity proofs in tests/ compare fold by
        # fold; the reference only returns their means)
        self.last_fold_results = {"alphas": a_f, "correlations": r_f, "p_values": p_f, "masks": np.stack(masks)}
        if scores is not None:
            self.last_fold_results["inner_scores"] = scores
        t_stats_end = time.perf_counter()
        if w_finish is not None:
            t_w0 = time.perf_counter()
            W = w_finish()
            t_w += (time.perf_counter() - t_w0) * 1e3
        del Wd
        self.last_timings = ops.timings()
        self.last_timings["wall_ms"] = (time.perf_counter() - t_start) * 1e3
        self.last_timings["host_weights_ms"] = t_w
        self.last_timings["host_stats_metrics_ms"] = (t_stats_end - t_s0) * 1e3
        self.last_stats = {"corr_precision": cfg.corr_precision, "launches": ops.launches, "gemm_flops": ops.gemm_flops, "rank": comm.rank,
                           "world": comm.world, "voxels_this_rank": c1 - c0,
                           "h2d_bytes": getattr(ops, "h2d_bytes", 0), "d2h_bytes": getattr(ops, "d2h_bytes", 0),
                           "store_gemm_launches": getattr(ops, "store_gemms", 0),
                           "store_gemm_f16_launches": getattr(ops, "store_gemms_f16", 0),
                           "store_gemm_f16_flops": getattr(ops, "gemm_flops_f16", 0.0),
                           "voxel_gemm_precision": cfg.voxel_gemm_precision,
                           "compact_stacks": getattr(ops, "compact_stacks", 0)}
        if hasattr(ops, "corr_launches"):
            log = ops.corr_launches()
            self.last_stats["corr_launch_ms"] = [ms for ms, _ in log]
            self.last_stats["corr_launch_flops"] = [fl for _, fl in log]
        logger.info("Median correlation: %.3f", metrics["median_score"])
        logger.info("Significant voxels: %d/%d (%.1f%%)", metrics["n_significant"], n_vox,
                    metrics["percent_significant"])
        return metrics, W, best


def fit_nested_cv(features: np.ndarray, targets: np.ndarray, **kwargs):
    """Function form documented by the reference (README.md:137,212-226); same kwargs as fit_predict."""
    return NestedCVModel(model_name="ridge_regression").fit_predict(features=features, targets=targets, **kwargs)
