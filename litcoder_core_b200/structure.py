"""Data structuring between FIR and fit_predict, on the device.

Drop-ins for the three steps `AbstractTrainer` runs on the host between the downsampler and the model
(encoding/trainer.py:203-282):

    apply_fir_delays          FIR.make_delayed per story                              (:203-209)
    create_train_test_split   LeBel style: last story = test; per story trim -> zs ->  (:223-262)
                              vstack, nan_to_num on the stimulus side
    create_concatenated_data  LPP / Narratives style: concatenate, trim once           (:264-282)

The reference builds a float64 delayed matrix per story, z-scores it on the host (three passes), stacks
everything into float64 arrays and hands them to fit_predict, which converts to float32 and uploads.  Here the
undelayed per-story features (n_TR x D, small) are uploaded once and ONE kernel per story
(lit_fir_zscore_rows) writes the trimmed, z-scored, NaN-scrubbed fp32 rows straight into the design matrix in
HBM; the responses are uploaded story by story and z-scored into place (lit_col_stats +
lit_gather_normalize_rows).  With ``device_outputs=True`` the results stay on the GPU as torch CUDA tensors
that `fit_predict` adopts without a copy, so X and Y never round-trip through host float64.

Outputs are float32 -- the dtype fit_predict computes in (nested_cv.py:99-100) -- where the reference returns
float64 / the brain data's dtype; values agree to float32 rounding (tests/test_host_logic.py,
tests/test_gpu_parity.py against golden outputs of the unmodified trainer).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np

from .fir import FIR

EPS = 1e-8


def _ops(ops):
    if ops is None:
        from .device import default_ops

        ops = default_ops()
    return ops


def _span(n: int, start, stop):
    """Python slice semantics of arr[start:stop] on n rows -> (first, last_exclusive) with last >= first."""
    a, b, _ = slice(start, stop).indices(n)
    return a, max(a, b)


def apply_fir_delays(features: Dict[str, np.ndarray], fir_delays: Sequence[int], ops=None) -> Dict[str, np.ndarray]:
    """trainer.py:203-209: {story: FIR.make_delayed(features[story], fir_delays)} (host float64, as the reference)."""
    return {story: FIR.make_delayed(feat, fir_delays, ops=ops) for story, feat in features.items()}


def _finish(ops, mats: Dict[str, object], device_outputs: bool):
    if device_outputs:
        return {k: ops.as_tensor(m) for k, m in mats.items()}
    return {k: ops.download_matrix(m) for k, m in mats.items()}


def _stim_side(ops, features, names, fs, fe, delays, circpad):
    """vstack of nan_to_num(zs(delayed[story][fs:fe])) for the given stories -> device matrix."""
    spans = [_span(np.asarray(features[s]).shape[0], fs, fe) for s in names]
    widths = {np.asarray(features[s]).shape[1] for s in names}
    if len(widths) != 1:
        raise ValueError("all stories must have the same number of features")  # np.vstack would raise
    X = ops.empty(sum(b - a for a, b in spans), widths.pop() * len(delays))
    r0 = 0
    for s, (a, b) in zip(names, spans):
        ops.fir_zscore_rows(np.asarray(features[s]), delays, circpad, a, b, True, ops.row_view(X, r0, b - a))
        r0 += b - a
    return X


def _resp_side(ops, brain_data, names, ts, te):
    """vstack of zs(brain_data[story][ts:te]) -> device matrix (no nan_to_num on the responses, trainer.py:240-243)."""
    blocks = []
    for s in names:
        arr = np.asarray(brain_data[s])
        a, b = _span(arr.shape[0], ts, te)
        blocks.append(arr[a:b])
    widths = {blk.shape[1] for blk in blocks}
    if len(widths) != 1:
        raise ValueError("all stories must have the same number of voxels")
    Y = ops.empty(sum(blk.shape[0] for blk in blocks), widths.pop())
    r0 = 0
    for blk in blocks:
        n = blk.shape[0]
        if n:
            M = ops.upload_matrix(blk)
            mean, std = ops.col_stats(M, None, n, ddof=0)
            ops.gather_normalize(M, None, n, mean, std, 3, EPS, out=ops.row_view(Y, r0, n))
        r0 += n
    return Y


def create_train_test_split(features: Dict[str, np.ndarray], brain_data: Dict[str, np.ndarray], trimming_config: dict,
                            fir_delays: Optional[Sequence[int]] = None, circpad: bool = False,
                            device_outputs: bool = False, ops=None) -> dict:
    """trainer.py:223-262.  `features` holds the per-story features in story order (the last story is the test
    set): the UNDELAYED downsampled features when `fir_delays` is given (FIR is fused into the structuring
    kernel), else matrices that are already delayed.  Returns {"Rstim", "Rresp", "Pstim", "Presp"} as float32
    NumPy arrays, or torch CUDA tensors with `device_outputs=True`."""
    ops = _ops(ops)
    stories = list(features.keys())
    if len(stories) < 2:
        raise ValueError("need at least one training story and one test story")  # np.vstack([]) in the reference
    delays = [int(d) for d in fir_delays] if fir_delays is not None else [0]
    g = trimming_config.get
    out = {}
    for names, prefix, kx, ky in ((stories[:-1], "train", "Rstim", "Rresp"), (stories[-1:], "test", "Pstim", "Presp")):
        out[kx] = _stim_side(ops, features, names, g(f"{prefix}_features_start", 0), g(f"{prefix}_features_end", None),
                             delays, circpad)
        out[ky] = _resp_side(ops, brain_data, names, g(f"{prefix}_targets_start", 0), g(f"{prefix}_targets_end", None))
    return _finish(ops, out, device_outputs)


def create_concatenated_data(features: Dict[str, np.ndarray], brain_data: Dict[str, np.ndarray],
                             story_order: Sequence[str], trimming_config: dict,
                             fir_delays: Optional[Sequence[int]] = None, circpad: bool = False,
                             device_outputs: bool = False, ops=None) -> dict:
    """trainer.py:264-282: concatenate the stories in `story_order`, then trim the concatenation once; no
    z-scoring.  Returns {"X", "Y"} (float32; torch CUDA tensors with `device_outputs=True`)."""
    ops = _ops(ops)
    delays = [int(d) for d in fir_delays] if fir_delays is not None else [0]
    g = trimming_config.get
    lens = [np.asarray(features[s]).shape[0] for s in story_order]
    fa, fb = _span(sum(lens), g("features_start", 0), g("features_end", None))
    ndim = np.asarray(features[story_order[0]]).shape[1]
    X = ops.empty(fb - fa, ndim * len(delays))
    base = r0 = 0
    for s, n in zip(story_order, lens):
        a, b = max(fa, base) - base, min(fb, base + n) - base  # this story's rows inside the trimmed window
        if b > a:
            ops.fir_zscore_rows(np.asarray(features[s]), delays, circpad, a, b, False, ops.row_view(X, r0, b - a))
            r0 += b - a
        base += n
    lens = [np.asarray(brain_data[s]).shape[0] for s in story_order]
    ta, tb = _span(sum(lens), g("targets_start", 0), g("targets_end", None))
    Y = ops.empty(tb - ta, np.asarray(brain_data[story_order[0]]).shape[1])
    base = r0 = 0
    for s, n in zip(story_order, lens):
        a, b = max(ta, base) - base, min(tb, base + n) - base
        if b > a:
            ops.upload_into(np.asarray(brain_data[s])[a:b], ops.row_view(Y, r0, b - a))
            r0 += b - a
        base += n
    return _finish(ops, {"X": X, "Y": Y}, device_outputs)
