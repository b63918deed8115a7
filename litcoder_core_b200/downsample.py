"""Drop-in for the reference's Downsampler facade (encoding/downsample/downsampling.py:322-445).

``Downsampler().downsample(data, data_times, tr_times, method="lanczos", window=3, cutoff_mult=1.0)``
resamples word- or frame-rate features onto the fMRI TR grid.  The Lanczos method
(interpdata.py:45-63,87-126) is the one on the hot path: the reference builds a dense
(n_TR x n_samples) float64 weight matrix in a Python loop (98 % zeros) and multiplies; here the
weights are evaluated on the fly on the B200 and only the samples inside the +-window/cutoff band
of each TR are read (lit_lanczos_downsample).

Parameter validation follows the reference: required / optional keyword table per method,
unknown keywords are dropped silently (downsampling.py:361-393), unknown methods and missing
required parameters raise ValueError.  The other nine methods of the reference are registered so
that `available_methods` / `get_method_params` answer identically, but they are not on the B200 path
yet and raise NotImplementedError instead of silently running on the CPU.
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
        fn = self._methods[method]
        if fn is None:
            raise NotImplementedError(
                f"downsampling method '{method}' is not implemented on the B200 path (only 'lanczos' is); "
                "litcoder_core_b200 has no CPU fallback")
        return fn(data, data_times, tr_times, **params)

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
