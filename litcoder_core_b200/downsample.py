"""Drop-in for the reference's Downsampler facade (encoding/downsample/downsampling.py:322-445).

``Downsampler().downsample(data, data_times, tr_times, method="lanczos", window=3, cutoff_mult=1.0)``
resamples word- or frame-rate features onto the fMRI TR grid.  The Lanczos method
(interpdata.py:45-63,87-126) is the one on the hot path: the reference builds a dense
(n_TR x n_samples) float64 weight matrix in a Python loop (98 % zeros) and multiplies; here the
weights are evaluated on the fly on the B200 and only the samples inside the +-window/cutoff band
of each TR are read (lit_lanczos_downsample).

Parameter validation follows the reference: required / optional keyword table per method,
unknown keywords are dropped silently (downsampling.py:361-393), unknown methods and missing
required parameters raise ValueError.  The other nine methods run on the device too:
`sinc` shares the resampling kernel (lit_sinc_downsample), `gabor` has its own (lit_gabor_downsample), and
the membership-based ones (`rect`, `average`, `sum`, `last`, `legacy_*`) reduce rows through a CSR kernel
(lit_csr_rows_apply) -- for those the host only builds the integer membership lists.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np


def lanczos_band(data_times: np.ndarray, tr_times: np.ndarray, window: int, cutoff: float):
    """Conservative [lo, hi) sample range per TR for sorted sample times, or (None, None) when the
    samples are not sorted / the cutoff is degenerate (then every sample is visited)."""
    if len(data_times) == 0 or not np.isfinite(cutoff) or cutoff == 0:
        return None, None
    if not np.all(np.isfinite(data_times)) or not np.all(np.isfinite(tr_times)):
        return None, None
    if np.any(np.diff(data_times) < 0):
        return None, None
    half = abs(window / cutoff)
    half += 1e-9 * max(1.0, half) + 1e-12 * float(np.max(np.abs(tr_times)) if len(tr_times) else 0.0)
    lo = np.searchsorted(data_times, tr_times - half, side="left").astype(np.int32)
    hi = np.searchsorted(data_times, tr_times + half, side="right").astype(np.int32)
    return lo, hi


class Downsampler:
    """Unified interface to the temporal downsampling methods (reference: downsampling.py:322)."""

    METHOD_PARAMS: Dict[str, Dict[str, List[str]]] = {
        "lanczos": {"required": ["window", "cutoff_mult"], "optional": ["rectify"]},
        "sinc": {"required": ["window", "cutoff_mult"], "optional": ["causal", "renorm"]},
        "average": {"required": ["split_indices"], "optional": []},
        "sum": {"required": ["split_indices"], "optional": []},
        "last": {"required": ["split_indices"], "optional": []},
        "legacy_average": {"required": ["split_indices"], "optional": []},
        "legacy_sum": {"required": ["split_indices"], "optional": []},
        "legacy_last": {"required": ["split_indices"], "optional": []},
        "rect": {"required": [], "optional": []},
        "gabor": {"required": ["freqs", "sigma"], "optional": []},
    }
    # registration order of the reference (downsampling.py:348-359) -> available_methods
    _ORDER = ["rect", "average", "sinc", "lanczos", "last", "gabor", "legacy_average", "legacy_last", "sum",
              "legacy_sum"]

    def __init__(self, ops=None):
        self._ops = ops
        self._methods = {name: getattr(self, f"_{name}", None) for name in self._ORDER}

    def _get_ops(self):
        if self._ops is None:
            from .device import default_ops

            self._ops = default_ops()
        return self._ops

    def _validate_method_params(self, method: str, **kwargs) -> dict:
        if method not in self._methods:
            raise ValueError(f"Unsupported downsampling method: {method}")
        spec = self.METHOD_PARAMS.get(method, {"required": [], "optional": []})
        filtered = {}
        for name in spec["required"]:
            if name not in kwargs:
                raise ValueError(f"Required parameter '{name}' missing for method '{method}'")
            filtered[name] = kwargs[name]
        for name in spec["optional"]:
            if name in kwargs:
                filtered[name] = kwargs[name]
        return filtered

    def downsample(self, data: np.ndarray, data_times: np.ndarray, tr_times: np.ndarray, method: str = "rect",
                   **kwargs) -> np.ndarray:
        """Downsample `data` (n_samples, n_features) sampled at `data_times` onto `tr_times`."""
        params = self._validate_method_params(method, **kwargs)
        return self._methods[method](data, data_times, tr_times, **params)

    @property
    def available_methods(self) -> List[str]:
        return list(self._methods.keys())

    def get_method_params(self, method: str) -> dict:
        if method not in self._methods:
            raise ValueError(f"Unsupported downsampling method: {method}")
        return self.METHOD_PARAMS.get(method, {"required": [], "optional": []})

    # ------------------------------------------------------------------------------------------
    def _lanczos(self, data, data_times, tr_times, window=3, cutoff_mult=1.0, rectify=False) -> np.ndarray:
        """interpdata.lanczosinterp2D: out = W @ data, W[i, j] = lanczos((tr_i - t_j) * cutoff)."""
        data = np.asarray(data)
        if data.ndim != 2:
            raise ValueError("data must be 2-D (n_samples, n_features)")
        data_times = np.ascontiguousarray(np.asarray(data_times, dtype=np.float64))
        tr_times = np.ascontiguousarray(np.asarray(tr_times, dtype=np.float64))
        if len(data_times) != data.shape[0]:
            # sincmat (n_TR x len(oldtime)) @ data raises in the reference
            raise ValueError(f"shapes ({len(tr_times)},{len(data_times)}) and {data.shape} not aligned")
        with np.errstate(invalid="ignore", divide="ignore"):
            cutoff = float(1 / np.mean(np.diff(tr_times)) * cutoff_mult) if len(tr_times) > 1 else float("nan")
        lo, hi = lanczos_band(data_times, tr_times, window, cutoff)
        if lo is not None and data.dtype.kind == "f" and not np.isfinite(data).all():
            lo = hi = None  # 0 * inf / 0 * nan must poison the output exactly as the dense product does
        return self._get_ops().lanczos_downsample(data, data_times, tr_times, float(window), cutoff, bool(rectify), lo,
                                                  hi)

    def _sinc(self, data, data_times, tr_times, window=1, cutoff_mult=1.0, causal=False, renorm=True) -> np.ndarray:
        """interpdata.sincinterp2D (:66-84): rows of sinc weights, optionally causal / renormalised."""
        data = np.asarray(data)
        data_times = np.ascontiguousarray(np.asarray(data_times, dtype=np.float64))
        tr_times = np.ascontiguousarray(np.asarray(tr_times, dtype=np.float64))
        if len(data_times) != data.shape[0]:
            raise ValueError(f"shapes ({len(tr_times)},{len(data_times)}) and {data.shape} not aligned")
        with np.errstate(invalid="ignore", divide="ignore"):
            cutoff = float(1 / np.mean(np.diff(tr_times)) * cutoff_mult) if len(tr_times) > 1 else float("nan")
        lo = hi = None
        if np.isfinite(cutoff) and cutoff != 0 and np.isfinite(window):
            lo, hi = lanczos_band(data_times, tr_times, float(window) / 2.0, cutoff)  # |t| <= window / (2 B)
        if lo is not None and data.dtype.kind == "f" and not np.isfinite(data).all():
            lo = hi = None
        return self._get_ops().resample("sinc", data, data_times, tr_times, float(window), cutoff, bool(causal),
                                        bool(renorm), lo, hi)

    def _gabor(self, data, data_times, tr_times, freqs=None, sigma=None) -> np.ndarray:
        """np.abs(interpdata.gabor_xfm2D(data.T, ...)).T (downsampling.py:159-166): (n_TR, n_features * n_freqs)."""
        data = np.asarray(data)
        return self._get_ops().gabor_downsample(data, np.ascontiguousarray(np.asarray(data_times, dtype=np.float64)),
                                                np.ascontiguousarray(np.asarray(tr_times, dtype=np.float64)),
                                                np.ascontiguousarray(np.asarray(freqs, dtype=np.float64)), float(sigma))

    # membership-based methods: the host builds integer (row_ptr, col_idx) lists, the device reduces the rows
    def _csr(self, data, groups: List[np.ndarray], mean: bool) -> np.ndarray:
        data = np.asarray(data)
        counts = np.fromiter((len(g) for g in groups), dtype=np.int64, count=len(groups))
        row_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        col_idx = np.concatenate(groups).astype(np.int64) if len(groups) and row_ptr[-1] else np.zeros(0, np.int64)
        return self._get_ops().csr_rows_apply(data, row_ptr, col_idx, None, mean)

    def _rect(self, data, data_times, tr_times) -> np.ndarray:
        """RectangularDownsampler (downsampling.py:24-38): mean of the samples in [t - TR/2, t + TR/2)."""
        data_times = np.asarray(data_times, dtype=np.float64)
        tr_times = np.asarray(tr_times, dtype=np.float64)
        with np.errstate(invalid="ignore"):
            tr = np.mean(np.diff(tr_times)) if len(tr_times) > 1 else np.nan
        if len(data_times) and np.all(np.diff(data_times) >= 0) and np.isfinite(tr):
            lo = np.searchsorted(data_times, tr_times - tr / 2, side="left")
            hi = np.searchsorted(data_times, tr_times + tr / 2, side="left")
            groups = [np.arange(a, max(a, b)) for a, b in zip(lo, hi)]
        else:
            groups = [np.flatnonzero((data_times >= t - tr / 2) & (data_times < t + tr / 2)) for t in tr_times]
        return self._csr(data, groups, mean=True)

    @staticmethod
    def _split_groups(split_indices, what: str) -> List[np.ndarray]:
        """Words of each TR for the split-index methods (downsampling.py:41-135, 232-279)."""
        if split_indices is None:
            raise ValueError(f"split_indices must be provided for {what} downsampling")
        arr = np.asarray(split_indices, dtype=np.int64)
        n_trs = int(arr.max()) + 1 if len(arr) else 0
        order = np.argsort(arr, kind="stable")
        order = order[arr[order] >= 0]
        bounds = np.searchsorted(arr[order], np.arange(n_trs + 1), side="left")
        return [order[bounds[i]:bounds[i + 1]] for i in range(n_trs)]

    def _average(self, data, data_times=None, tr_times=None, split_indices=None) -> np.ndarray:
        return self._csr(data, self._split_groups(split_indices, "average"), mean=True)

    def _sum(self, data, data_times=None, tr_times=None, split_indices=None) -> np.ndarray:
        return self._csr(data, self._split_groups(split_indices, "sum"), mean=False)

    def _last(self, data, data_times=None, tr_times=None, split_indices=None) -> np.ndarray:
        groups = self._split_groups(split_indices, "last point")
        return self._csr(data, [g[-1:] for g in groups], mean=False)

    @staticmethod
    def _legacy_groups(n_rows: int, split_indices) -> List[np.ndarray]:
        """np.split(data, split_indices) as index lists (downsampling.py:169-230, 282-319)."""
        if split_indices is None:
            raise ValueError("split_indices must be provided for Legacy downsampling")
        cuts = [0] + [int(i) for i in split_indices] + [n_rows]
        rows = np.arange(n_rows, dtype=np.int64)
        return [rows[cuts[i]:cuts[i + 1]] for i in range(len(cuts) - 1)]

    def _legacy_average(self, data, data_times, tr_times, split_indices=None) -> np.ndarray:
        return self._csr(data, self._legacy_groups(np.asarray(data).shape[0], split_indices), mean=True)

    def _legacy_sum(self, data, data_times, tr_times, split_indices=None) -> np.ndarray:
        return self._csr(data, self._legacy_groups(np.asarray(data).shape[0], split_indices), mean=False)

    def _legacy_last(self, data, data_times, tr_times, split_indices=None) -> np.ndarray:
        groups = self._legacy_groups(np.asarray(data).shape[0], split_indices)
        return self._csr(data, [g[-1:] for g in groups], mean=False)
