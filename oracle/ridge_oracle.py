"""CPU oracle for the nested-CV ridge hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A NumPy / SciPy restatement of the reference's algorithm (GT-LIT-Lab/litcoder_core), written
from its behaviour; every function cites the reference lines it follows.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import this module; the product
package `litcoder_core_b200` never does (it fails loudly without its CUDA library).

Parity status: PINNED.  `scripts/make_golden.py` runs the unmodified reference (imported from
/root/reference through a stub shim for its absent optional dependencies) on seeded inputs and
stores the outputs under `tests/golden/`; `tests/test_oracle_golden.py` checks every function here
against those vectors.  The reference itself has no tests or golden vectors.

Arithmetic follows the reference: float32 for the ridge algebra (the reference converts to
torch.float32 at nested_cv.py:99-100 and runs LAPACK/BLAS in single precision on CPU), SciPy for
the per-voxel statistics, float64 for Lanczos / FIR.  Third-party pieces restated here because the
package is absent from the image: statsmodels 0.14.4 `fdrcorrection` (method "indep").
"""
from __future__ import annotations

import math
import random
from typing import List, Optional, Sequence, Tuple

import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------------------------
# FIR delays  (encoding/features/FIR_expander.py:24-43)
# ----------------------------------------------------------------------------------------------
def fir_make_delayed(stim: np.ndarray, delays: Sequence[int], circpad: bool = False) -> np.ndarray:
    nt, ndim = stim.shape
    blocks = []
    for d in delays:
        d = int(d)
        if d == 0:
            blocks.append(stim.copy())  # keeps the input dtype (FIR_expander.py:41)
            continue
        blk = np.zeros((nt, ndim))  # float64 (FIR_expander.py:31)
        # NumPy slice semantics of the reference: dst[d:] = src[:-d] (d > 0) / dst[:d] = src[-d:] (d < 0)
        if d > 0:
            blk[d:, :] = stim[:-d, :]
            if circpad:
                blk[:d, :] = stim[-d:, :]
        else:
            blk[:d, :] = stim[-d:, :]
            if circpad:
                blk[d:, :] = stim[:-d, :]
        blocks.append(blk)
    return np.hstack(blocks)


# ----------------------------------------------------------------------------------------------
# Lanczos resampling  (encoding/downsample/interpdata.py:45-63, 87-126)
# ----------------------------------------------------------------------------------------------
def lanczos_kernel(cutoff: float, t: np.ndarray, window: int = 3) -> np.ndarray:
    t = np.asarray(t, dtype=np.float64) * cutoff
    with np.errstate(divide="ignore", invalid="ignore"):
        val = window * np.sin(np.pi * t) * np.sin(np.pi * t / window) / (np.pi ** 2 * t ** 2)
    val[t == 0] = 1.0
    val[np.abs(t) > window] = 0.0
    return val


def lanczos_interp2d(data, oldtime, newtime, window=3, cutoff_mult=1.0, rectify=False) -> np.ndarray:
    oldtime = np.asarray(oldtime, dtype=np.float64)
    newtime = np.asarray(newtime, dtype=np.float64)
    cutoff = 1 / np.mean(np.diff(newtime)) * cutoff_mult  # interpdata.py:107
    W = np.zeros((len(newtime), len(oldtime)))
    for i in range(len(newtime)):  # interpdata.py:112-113
        W[i, :] = lanczos_kernel(cutoff, newtime[i] - oldtime, window)
    if rectify:  # interpdata.py:115-121
        return np.hstack([W @ np.clip(data, -np.inf, 0), W @ np.clip(data, 0, np.inf)])
    return W @ data


# ----------------------------------------------------------------------------------------------
# The other downsamplers of the facade  (encoding/downsample/downsampling.py:24-319, interpdata.py:29-145)
# ----------------------------------------------------------------------------------------------
def downsample_rect(data, data_times, tr_times) -> np.ndarray:
    """RectangularDownsampler (:24-38)."""
    out = np.zeros((len(tr_times), data.shape[1]))
    tr = np.mean(np.diff(tr_times))
    for i, t in enumerate(tr_times):
        mask = (data_times >= t - tr / 2) & (data_times < t + tr / 2)
        if np.any(mask):
            out[i] = np.mean(data[mask], axis=0)
    return out


def downsample_by_tr(data, split_indices, how: str) -> np.ndarray:
    """Average / Sum / LastPoint downsamplers (:41-135, 232-279): words grouped by their TR index."""
    arr = np.asarray(split_indices)
    n_trs = int(arr.max()) + 1
    out = np.zeros((n_trs, data.shape[1]))
    for tr in range(n_trs):
        idx = np.flatnonzero(arr == tr)
        if len(idx):
            out[tr] = {"average": lambda r: np.mean(r, axis=0), "sum": lambda r: np.sum(r, axis=0),
                       "last": lambda r: r[-1]}[how](data[idx])
    return out


def downsample_legacy(data, split_indices, how: str) -> np.ndarray:
    """Legacy average / sum / last (:169-230, 282-319): chunks of np.split(data, split_indices)."""
    out = np.zeros((len(split_indices) + 1, data.shape[1]))
    for ci, chunk in enumerate(np.split(data, split_indices)):
        if len(chunk):
            out[ci] = {"average": lambda r: np.mean(r, axis=0), "sum": lambda r: np.sum(r, axis=0),
                       "last": lambda r: r[-1]}[how](chunk)
    return out


def sinc_interp2d(data, oldtime, newtime, cutoff_mult=1.0, window=1, causal=False, renorm=True) -> np.ndarray:
    """interpdata.sincfun / sincinterp2D (:29-42, 66-84)."""
    B = 1 / np.mean(np.diff(newtime)) * cutoff_mult
    W = np.zeros((len(newtime), len(oldtime)))
    for i in range(len(newtime)):
        t = newtime[i] - oldtime
        val = 2 * B * np.sin(2 * np.pi * B * t) / (2 * np.pi * B * t + 1e-20)
        val[np.abs(t) > window / (2 * B)] = 0
        if causal:
            val[t < 0] = 0
        if not np.sum(val) == 0.0 and renorm:
            val = val / np.sum(val)
        W[i] = val
    return W @ data


def gabor_downsample(data, oldtimes, newtimes, freqs, sigma) -> np.ndarray:
    """np.abs(interpdata.gabor_xfm2D(data.T, ...)).T (:129-145; downsampling.py:159-166)."""
    sinv = np.vstack([np.sin(oldtimes * f * 2 * np.pi) for f in freqs])
    cosv = np.vstack([np.cos(oldtimes * f * 2 * np.pi) for f in freqs])
    blocks = []
    for d in data.T:
        out = np.zeros((len(newtimes), len(freqs)), dtype=np.complex128)
        for ti, t in enumerate(newtimes):
            g = np.exp(-0.5 * (oldtimes - t) ** 2 / (2 * sigma ** 2)) * d
            out[ti] = cosv @ g + 1j * (sinv @ g)
        blocks.append(out.T)
    return np.abs(np.vstack(blocks)).T


# ----------------------------------------------------------------------------------------------
# Fold construction  (encoding/models/folding.py:8-255)
# ----------------------------------------------------------------------------------------------
def _kfold_contiguous(n: int, k: int):
    """sklearn KFold(shuffle=False): first n % k folds get one extra sample."""
    sizes = np.full(k, n // k, dtype=int)
    sizes[: n % k] += 1
    idx = np.arange(n)
    out, start = [], 0
    for s in sizes:
        test = idx[start:start + s]
        train = np.concatenate([idx[:start], idx[start + s:]])
        out.append((train, test))
        start += s
    return out


def _kfold_shuffled(n: int, k: int):
    """sklearn KFold(shuffle=True, random_state=None): permutes with NumPy's global RNG."""
    idx = np.arange(n)
    np.random.shuffle(idx)  # check_random_state(None).shuffle
    sizes = np.full(k, n // k, dtype=int)
    sizes[: n % k] += 1
    out, start = [], 0
    for s in sizes:
        test_mask = np.zeros(n, dtype=bool)
        test_mask[idx[start:start + s]] = True
        out.append((np.arange(n)[~test_mask], np.arange(n)[test_mask]))
        start += s
    return out


def _chunked(n: int, k: int, chunk: int, shuffle: bool, trim: Optional[int] = None, kfold_shuffle=None):
    n_chunks = n // chunk  # folding.py:82 -- the n % chunk tail rows belong to no fold
    order = list(range(n_chunks))
    if shuffle:
        random.shuffle(order)  # folding.py:85-86: Python's global RNG
    per = n_chunks // k
    if per == 0:  # folding.py:90-96 / 157-165: fall back to KFold
        do_shuffle = shuffle if kfold_shuffle is None else kfold_shuffle
        return _kfold_shuffled(n, k) if do_shuffle else _kfold_contiguous(n, k)
    splits = []
    for i in range(k):
        lo = i * per
        hi = (i + 1) * per if i < k - 1 else n_chunks  # last fold takes the remainder (folding.py:102-104)
        test_chunks = order[lo:hi]
        tset = set(test_chunks)
        train_chunks = [c for c in order if c not in tset]
        test_idx: List[int] = []
        for c in test_chunks:
            s, e = c * chunk, min(c * chunk + chunk, n)
            if trim is not None:
                s, e = s + trim, e - trim  # folding.py:184-190
                if s >= e:
                    continue
            test_idx.extend(range(s, e))
        train_idx: List[int] = []
        for c in train_chunks:
            train_idx.extend(range(c * chunk, min(c * chunk + chunk, n)))
        splits.append((train_idx, test_idx))
    return splits


def create_folds(n_samples, fold_type, n_folds, chunk_length=None, trim_size=None, groups=None):
    """folding.py:8-64 (same positional signature)."""
    if fold_type == "chunked":
        return _chunked(n_samples, n_folds, chunk_length, shuffle=True)
    if fold_type == "chunked_trimmed":
        t = 5 if trim_size is None else trim_size
        return _chunked(n_samples, n_folds, chunk_length, shuffle=True, trim=t, kfold_shuffle=False)
    if fold_type == "chunked_contiguous":
        return _chunked(n_samples, n_folds, chunk_length, shuffle=False)
    if fold_type == "kfold":
        return _kfold_contiguous(n_samples, n_folds)
    if fold_type == "kfold_trimmed":
        t = 5 if trim_size is None else trim_size
        out = []
        for tr, te in _kfold_contiguous(n_samples, n_folds):  # folding.py:226-253
            tr, te = list(tr), list(te)
            out.append((tr, te[t:-t] if len(te) > 2 * t else te))
        return out
    if fold_type == "timeseries":
        # sklearn TimeSeriesSplit(n_splits=k): test_size = n // (k+1), expanding train window
        k = n_folds
        ts = n_samples // (k + 1)
        idx = np.arange(n_samples)
        starts = range(n_samples - k * ts, n_samples, ts)
        return [(idx[:s], idx[s:s + ts]) for s in starts]
    if fold_type == "group":
        if groups is None:
            raise ValueError("Groups must be provided for group folding")
        return _group_kfold(np.asarray(groups), n_folds)
    raise ValueError(f"Unknown folding type: {fold_type}")


def _group_kfold(groups: np.ndarray, k: int):
    """sklearn GroupKFold (no shuffle): biggest groups first onto the lightest fold."""
    uniq, inv = np.unique(groups, return_inverse=True)
    counts = np.bincount(inv)
    order = np.argsort(counts, kind="stable")[::-1]  # sklearn >= 1.4: stable sort, then reversed
    counts_sorted = counts[order]
    load = np.zeros(k)
    g2f = np.zeros(len(uniq), dtype=int)
    for gi, w in enumerate(counts_sorted):
        f = int(np.argmin(load))
        load[f] += w
        g2f[order[gi]] = f
    fold_of = g2f[inv]
    idx = np.arange(len(groups))
    return [(idx[fold_of != f], idx[fold_of == f]) for f in range(k)]


# ----------------------------------------------------------------------------------------------
# Ridge kernels  (encoding/models/ridge_utils.py, ridge_regression.py)
# ----------------------------------------------------------------------------------------------
def z_score_f32(x: np.ndarray, eps: float = 1e-8) -> np.ndarray:
    """torch branch of ridge_utils.z_score (:11-15): UNBIASED std, eps added to the std."""
    x = x.astype(F32, copy=False)
    m = x.mean(axis=0, keepdims=True, dtype=F32)
    s = x.std(axis=0, keepdims=True, ddof=1, dtype=F32)
    return ((x - m) / (s + F32(eps))).astype(F32)


def svd_truncated(X: np.ndarray, singcutoff: float):
    """ridge_utils.svd_wrapper (:49-67): thin SVD, keep the S > singcutoff prefix."""
    U, S, Vh = np.linalg.svd(X.astype(F32, copy=False), full_matrices=False)
    k = int(np.sum(S > singcutoff))
    return U[:, :k], S[:k], Vh[:k]


def ridge_corr(Rstim, Pstim, Rresp, Presp, alphas, singcutoff=1e-30, use_corr=True, normalpha=False) -> np.ndarray:
    """ridge_regression.ridge_corr_torch (:66-141) -> (n_alphas, n_voxels) float32."""
    U, S, Vh = svd_truncated(Rstim, singcutoff)
    norm = float(S[0])
    nalphas = [a * norm for a in alphas] if normalpha else list(alphas)
    UR = U.T @ Rresp.astype(F32, copy=False)
    PVh = Pstim.astype(F32, copy=False) @ Vh.T
    Presp = Presp.astype(F32, copy=False)
    zP = z_score_f32(Presp)
    Pvar = Presp.var(axis=0, ddof=1, dtype=F32)
    out = []
    for na in nalphas:
        D = (S / (S ** 2 + F32(na ** 2))).astype(F32)
        pred = (PVh * D[None, :]) @ UR
        if use_corr:
            c = (zP * z_score_f32(pred)).mean(axis=0, dtype=F32)
        else:
            resvar = (Presp - pred).var(axis=0, ddof=1, dtype=F32)
            with np.errstate(divide="ignore", invalid="ignore"):
                rsq = 1 - resvar / Pvar
            c = np.sqrt(np.abs(rsq)) * np.sign(rsq)
        out.append(np.nan_to_num(c).astype(F32))
    return np.stack(out)


def ridge_weights(Rstim, Rresp, valphas, singcutoff=1e-30, normalpha=False) -> np.ndarray:
    """ridge_regression.ridge_torch (:9-63) -> (n_features, n_voxels) float32."""
    U, S, Vh = svd_truncated(Rstim, singcutoff)
    Rresp = Rresp.astype(F32, copy=False)
    UR = U.T @ Rresp
    if np.isscalar(valphas):
        valphas = np.full(Rresp.shape[1], valphas, dtype=F32)
    valphas = np.asarray(valphas, dtype=F32)
    norm = float(S[0])
    nal = (valphas * F32(norm)).astype(F32) if normalpha else valphas
    wt = np.zeros((Rstim.shape[1], Rresp.shape[1]), dtype=F32)
    for ua in np.unique(nal):
        sel = np.nonzero(nal == ua)[0]
        D = (S / (S ** 2 + ua ** 2)).astype(F32)
        wt[:, sel] = (Vh.T * D[None, :]) @ UR[:, sel]
    return wt


def ridge_corr_pred(Rstim, Pstim, Rresp, Presp, valphas, singcutoff=1e-30, use_corr=True, normalpha=True) -> np.ndarray:
    """ridge_regression.ridge_corr_pred_torch (:144-216) -> (n_voxels,) float32; NaNs are NOT scrubbed."""
    U, S, Vh = svd_truncated(Rstim, singcutoff)
    valphas = np.asarray(valphas, dtype=F32)
    nal = (valphas * F32(S[0])).astype(F32) if normalpha else valphas
    UR = U.T @ Rresp.astype(F32, copy=False)
    PVh = Pstim.astype(F32, copy=False) @ Vh.T
    Presp = Presp.astype(F32, copy=False)
    zP = z_score_f32(Presp)
    Pvar = Presp.var(axis=0, ddof=1, dtype=F32)
    corr = np.zeros(Rresp.shape[1], dtype=F32)
    for ua in np.unique(nal):
        sel = np.nonzero(nal == ua)[0]
        D = (S / (S ** 2 + ua ** 2)).astype(F32)
        pred = (PVh * D[None, :]) @ UR[:, sel]
        with np.errstate(divide="ignore", invalid="ignore"):
            if use_corr:
                corr[sel] = (zP[:, sel] * z_score_f32(pred)).mean(axis=0, dtype=F32)
            else:
                rsq = 1 - (Presp[:, sel] - pred).var(axis=0, ddof=1, dtype=F32) / Pvar[sel]
                corr[sel] = np.sqrt(np.abs(rsq)) * np.sign(rsq)
    return corr


def zs(v: np.ndarray) -> np.ndarray:
    """encoding/utils.py:23-29 `zscore`: population std; zero-std columns are centred only."""
    s = v.std(0)
    m = v - v.mean(0)
    nz = s != 0.0
    m[:, nz] /= s[nz]
    return m


# ----------------------------------------------------------------------------------------------
# the trainers' data structuring between FIR and fit_predict  (encoding/trainer.py:203-282)
# ----------------------------------------------------------------------------------------------
def apply_fir_delays(features: dict, delays: Sequence[int]) -> dict:
    """trainer.py:203-209: FIR.make_delayed per story (insertion order kept)."""
    return {story: fir_make_delayed(feat, delays) for story, feat in features.items()}


def create_train_test_split(features: dict, brain_data: dict, trimming_config: dict) -> dict:
    """trainer.py:223-262 (LeBel style): the last story is the test set; per story: trim, zs; the stimulus
    side additionally goes through nan_to_num after the vstack."""
    stories = list(features.keys())
    train, test = stories[:-1], stories[-1:]
    g = trimming_config.get

    def side(names, prefix):
        fs, fe = g(f"{prefix}_features_start", 0), g(f"{prefix}_features_end", None)
        ts, te = g(f"{prefix}_targets_start", 0), g(f"{prefix}_targets_end", None)
        X = np.nan_to_num(np.vstack([zs(features[s][fs:fe]) for s in names]))
        Y = np.vstack([zs(brain_data[s][ts:te]) for s in names])
        return X, Y

    X_train, Y_train = side(train, "train")
    X_test, Y_test = side(test, "test")
    return {"Rstim": X_train, "Rresp": Y_train, "Pstim": X_test, "Presp": Y_test}


def create_concatenated_data(features: dict, brain_data: dict, story_order: Sequence[str], trimming_config: dict) -> dict:
    """trainer.py:264-282 (LPP / Narratives style): concatenate the stories, then trim ONCE; no z-scoring."""
    g = trimming_config.get
    X = np.concatenate([features[s] for s in story_order], axis=0)
    Y = np.concatenate([brain_data[s] for s in story_order], axis=0)
    return {"X": X[g("features_start", 0):g("features_end", None)], "Y": Y[g("targets_start", 0):g("targets_end", None)]}


def find_best_alphas(X, Y, splits, alphas, single_alpha=False, normalpha=False, use_corr=True,
                     singcutoff=1e-10, return_corrs=False):
    """nested_cv._find_best_alphas (:334-415)."""
    corrs = []
    for tr, va in splits:
        tr, va = np.asarray(tr, dtype=np.int64), np.asarray(va, dtype=np.int64)
        corrs.append(ridge_corr(X[tr], X[va], Y[tr], Y[va], alphas, singcutoff=singcutoff, use_corr=use_corr,
                                normalpha=normalpha))
    acc = corrs[0].copy()
    for c in corrs[1:]:
        acc = acc + c
    mean_corr = (acc / F32(len(corrs))).astype(F32)  # torch.stack(...).mean(0)
    if single_alpha:
        j = int(np.argmax(mean_corr.mean(axis=1, dtype=F32)))
        # torch.tensor([alphas[j]] * V) (:399-401) takes its dtype from the element: float64 for an element of a
        # float64 ndarray, torch's default float32 for a Python float
        a = alphas[j]
        best = np.full(Y.shape[1], a, dtype=np.float64 if isinstance(a, np.float64) else F32)
    else:
        best = np.asarray(alphas, dtype=np.float64)[np.argmax(mean_corr, axis=0)].astype(F32)
    return (best, mean_corr) if return_corrs else best


# ----------------------------------------------------------------------------------------------
# Statistics  (nested_cv.py:418-477; statsmodels fdrcorrection)
# ----------------------------------------------------------------------------------------------
def correlations_pvalues(y_true: np.ndarray, y_pred: np.ndarray):
    """nested_cv._calculate_correlations_pvalues (:418-438): SciPy pearsonr per voxel; NaN -> (0, 1)."""
    from scipy.stats import pearsonr
    import warnings

    r_out, p_out = [], []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(y_true.shape[1]):
            r, p = pearsonr(y_true[:, i], y_pred[:, i])
            r_out.append(0.0 if np.isnan(r) else r)
            p_out.append(1.0 if np.isnan(p) else p)
    return r_out, p_out


def correlations_pvalues_vectorised(y_true: np.ndarray, y_pred: np.ndarray, p_dtype=np.float32):
    """Same statistic as correlations_pvalues for all voxels at once (used for large V in tests)."""
    from scipy.special import betainc

    n = y_true.shape[0]
    a = y_true.astype(np.float64) - y_true.astype(np.float64).mean(0)
    b = y_pred.astype(np.float64) - y_pred.astype(np.float64).mean(0)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = (a * b).sum(0) / np.sqrt((a * a).sum(0) * (b * b).sum(0))
    bad = np.isnan(r)
    r = np.clip(np.where(bad, 0.0, r), -1.0, 1.0)
    ab = n / 2.0 - 1.0
    p = 2.0 * betainc(ab, ab, 0.5 * (1.0 - np.abs(r)))
    p = np.minimum(p, 1.0)
    p[bad] = 1.0
    return r.astype(np.float32), p.astype(p_dtype)


def fdr_bh(pvals, alpha=0.05):
    """statsmodels.stats.multitest.fdrcorrection(method='indep') v0.14.4, restated."""
    p = np.asarray(pvals)
    n = len(p)
    order = np.argsort(p)
    ps = p[order]
    ecdf = np.arange(1, n + 1) / float(n)
    reject = ps <= ecdf * alpha
    if reject.any():
        reject[: np.max(np.nonzero(reject)[0]) + 1] = True
    adj = np.minimum.accumulate((ps / ecdf)[::-1])[::-1]
    adj[adj > 1] = 1
    r_out = np.empty_like(reject)
    a_out = np.empty_like(adj)
    r_out[order] = reject
    a_out[order] = adj
    return r_out, a_out


def fisher_combine(fold_pvalues) -> np.ndarray:
    """nested_cv._combine_pvalues_across_folds (:441-477)."""
    from scipy.stats import combine_pvalues
    import warnings

    out = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(len(fold_pvalues[0])):
            pv = [f[i] for f in fold_pvalues]
            if all(x == 1.0 for x in pv):
                out.append(1.0)
            else:
                out.append(combine_pvalues(pv, method="fisher")[1])
    return np.array(out)


def fisher_combine_vectorised(fold_pvalues: np.ndarray) -> np.ndarray:
    """Closed form of Fisher's method for 2K degrees of freedom (all voxels at once)."""
    P = np.asarray(fold_pvalues, dtype=np.float64)
    K = P.shape[0]
    with np.errstate(divide="ignore"):
        x = -np.log(P).sum(0)
    term = np.ones_like(x)
    s = np.ones_like(x)
    with np.errstate(invalid="ignore", over="ignore"):
        for j in range(1, K):
            term = term * x / j
            s = s + term
        out = np.exp(-x) * s
    out[np.isinf(x)] = 0.0
    out[(P == 1.0).all(0)] = 1.0
    return np.minimum(out, 1.0)


# ----------------------------------------------------------------------------------------------
# Metrics dictionaries  (nested_cv.py:480-616)
# ----------------------------------------------------------------------------------------------
def _summary(prefix: str, values: np.ndarray) -> dict:
    return {
        f"median_{prefix}score": float(np.median(values)),
        f"mean_{prefix}score": float(np.mean(values)),
        f"min_{prefix}score": float(np.min(values)),
        f"max_{prefix}score": float(np.max(values)),
    }


def metrics_train_test(corr, pvals, padj, sig, best_alphas, n_sig) -> dict:
    m = {
        "median_score": float(np.median(corr)), "mean_score": float(np.mean(corr)),
        "std_score": float(np.std(corr)), "min_score": float(np.min(corr)), "max_score": float(np.max(corr)),
        "best_alphas": np.asarray(best_alphas).tolist(), "correlations": corr, "p_values": pvals,
        "corrected_p_values": np.asarray(padj).tolist(), "significant_mask": np.asarray(sig).tolist(),
        "n_significant": int(n_sig), "percent_significant": float(n_sig / len(corr) * 100),
    }
    if n_sig > 0:
        m.update(_summary("significant_", np.array(corr)[np.asarray(sig)]))
    return m


def metrics_full_cv(corr, pvals, padj, sig, maj, mean_alphas, n_sig, n_maj) -> dict:
    m = {
        "median_score": float(np.median(corr)), "mean_score": float(np.mean(corr)),
        "std_score": float(np.std(corr)), "min_score": float(np.min(corr)), "max_score": float(np.max(corr)),
        "best_alphas": mean_alphas.tolist(), "correlations": corr.tolist(), "p_values": pvals.tolist(),
        "corrected_p_values": padj.tolist(), "significant_mask": sig.tolist(),
        "majority_significant_mask": maj.tolist(), "n_significant": int(n_sig),
        "n_majority_significant": int(n_maj), "percent_significant": float(n_sig / len(corr) * 100),
        "percent_majority_significant": float(n_maj / len(corr) * 100),
    }
    if n_sig > 0:
        m.update(_summary("significant_", corr[sig]))
    if n_maj > 0:
        m.update(_summary("majority_significant_", corr[maj]))
    return m


# ----------------------------------------------------------------------------------------------
# Driver  (nested_cv.NestedCVModel.fit_predict, :18-331)
# ----------------------------------------------------------------------------------------------
def _normalise(Xtr, Ytr, Xte, Yte, nf, nt, eps=1e-8):
    """ridge_utils.DataNormalizer (:70-180): train statistics, unbiased std."""
    if nf:
        m, s = Xtr.mean(0, keepdims=True, dtype=F32), Xtr.std(0, keepdims=True, ddof=1, dtype=F32)
        Xtr, Xte = (Xtr - m) / (s + F32(eps)), (Xte - m) / (s + F32(eps))
    if nt:
        m, s = Ytr.mean(0, keepdims=True, dtype=F32), Ytr.std(0, keepdims=True, ddof=1, dtype=F32)
        Ytr, Yte = (Ytr - m) / (s + F32(eps)), (Yte - m) / (s + F32(eps))
    return Xtr.astype(F32), Ytr.astype(F32), Xte.astype(F32), Yte.astype(F32)


def fit_predict(features, targets, X_test=None, y_test=None, groups=None, folding_type="chunked", n_outer_folds=5,
                n_inner_folds=5, chunk_length=20, alphas=None, alpha_fdr=0.05, single_alpha=False, normalpha=True,
                use_corr=True, normalize_features=False, normalize_targets=False, singcutoff=1e-10,
                vectorised_stats=False, details=None):
    """Restatement of NestedCVModel.fit_predict; returns (metrics, weights, best_alphas).

    vectorised_stats=True swaps the per-voxel SciPy loops for their closed forms (identical
    statistics, needed to keep large-V test cases within seconds).
    details: optional list; receives one dict per outer fold (one in train/test mode) with what the parity proofs
    of tests/parity.py need: the fold's (normalised) arrays, the fold-mean inner score curves `mean_corr`
    (n_alphas x V), the selected alphas, the test r / p and the weights."""
    if alphas is None:
        alphas = np.logspace(-1, 8, 10)
    X = np.asarray(features).astype(F32)
    Y = np.asarray(targets).astype(F32)

    def stats(y_true, y_pred):
        if vectorised_stats:
            r, p = correlations_pvalues_vectorised(y_true, y_pred)
            return list(r), list(p)
        return correlations_pvalues(y_true, y_pred)

    if X_test is not None and y_test is not None:  # nested_cv.py:105-171
        Xt, Yt = np.asarray(X_test).astype(F32), np.asarray(y_test).astype(F32)
        if normalize_features or normalize_targets:
            X, Y, Xt, Yt = _normalise(X, Y, Xt, Yt, normalize_features, normalize_targets)
        splits = create_folds(len(X), folding_type, n_inner_folds, chunk_length, groups)
        best, mean_corr = find_best_alphas(X, Y, splits, alphas, single_alpha, normalpha, use_corr, singcutoff,
                                           return_corrs=True)
        wt = ridge_weights(X, Y, best, singcutoff=singcutoff, normalpha=normalpha)
        r, p = stats(Yt, Xt @ wt)
        if details is not None:
            details.append(dict(Xtr=X, Ytr=Y, Xte=Xt, Yte=Yt, mean_corr=mean_corr, best=best, r=np.asarray(r),
                                p=np.asarray(p), wt=wt))
        sig, padj = fdr_bh(p, alpha_fdr)
        return metrics_train_test(r, p, padj, sig, best, int(np.sum(sig))), wt, best

    if groups is not None and folding_type == "group":  # nested_cv.py:176-186
        outer = create_folds(len(X), "group", n_outer_folds, groups=groups)
    else:
        outer = create_folds(len(X), folding_type, n_outer_folds, chunk_length, groups)
    scores, pvals, valphas, masks, weights = [], [], [], [], []
    for tr, te in outer:
        tr, te = np.asarray(tr, dtype=np.int64), np.asarray(te, dtype=np.int64)
        Xtr, Xte, Ytr, Yte = X[tr], X[te], Y[tr], Y[te]
        if normalize_features or normalize_targets:
            Xtr, Ytr, Xte, Yte = _normalise(Xtr, Ytr, Xte, Yte, normalize_features, normalize_targets)
        if groups is not None and folding_type == "group":
            inner = create_folds(len(tr), "group", n_inner_folds, groups=[groups[i] for i in tr])
        else:
            inner = create_folds(len(tr), folding_type, n_inner_folds, chunk_length)
        best, mean_corr = find_best_alphas(Xtr, Ytr, inner, alphas, single_alpha, normalpha, use_corr, singcutoff,
                                           return_corrs=True)
        valphas.append(best)
        wt = ridge_weights(Xtr, Ytr, best, singcutoff=singcutoff, normalpha=normalpha)
        weights.append(wt)
        r, p = stats(Yte, Xte @ wt)
        if details is not None:
            details.append(dict(Xtr=Xtr, Ytr=Ytr, Xte=Xte, Yte=Yte, mean_corr=mean_corr, best=best, r=np.asarray(r),
                                p=np.asarray(p), wt=wt))
        scores.append(r)
        pvals.append(p)
        masks.append(fdr_bh(p, alpha_fdr)[0])
    corr = np.mean(scores, axis=0)
    if vectorised_stats:
        comb = fisher_combine_vectorised(np.asarray(pvals))
    else:
        comb = fisher_combine(pvals)
    sig, padj = fdr_bh(comb, alpha_fdr)
    maj = np.sum(masks, axis=0) >= (n_outer_folds // 2 + 1)
    mean_alphas = np.mean(valphas, axis=0)
    mean_w = np.mean(weights, axis=0)
    m = metrics_full_cv(corr, comb, padj, sig, maj, mean_alphas, int(np.sum(sig)), int(np.sum(maj)))
    return m, mean_w, mean_alphas
