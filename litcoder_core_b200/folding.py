"""Cross-validation fold construction (host-side integer work).

Same call signature and index semantics as the reference's `create_folds`
(encoding/models/folding.py:8-64), so that fold membership -- and therefore every number
downstream -- is identical for identical RNG state:

* ``chunked`` / ``chunked_trimmed`` shuffle the chunk order with Python's global ``random``
  module exactly once per call (folding.py:85-86, 153-154); seed it with ``random.seed``.
* the ``n_samples % chunk_length`` tail rows belong to no fold (folding.py:82-83);
* the last fold takes the remainder chunks (folding.py:102-104);
* with fewer chunks than folds the reference falls back to scikit-learn's ``KFold``
  (shuffled with NumPy's global RNG for ``chunked``; unshuffled for the other two), which is
  restated here without the scikit-learn dependency (folding.py:90-96, 157-165);
* ``kfold``, ``kfold_trimmed``, ``timeseries`` and ``group`` follow scikit-learn's
  ``KFold(shuffle=False)``, ``TimeSeriesSplit`` and ``GroupKFold`` (folding.py:45-62, 205-255).

Folds are returned as int64 NumPy index arrays instead of Python lists; they are uploaded to the
device as int32 row-index vectors by the engine.
"""
from __future__ import annotations

import logging
import random
from typing import List, Optional, Sequence, Tuple

import numpy as np

Fold = Tuple[np.ndarray, np.ndarray]

FOLD_TYPES = ("chunked", "chunked_trimmed", "chunked_contiguous", "kfold", "kfold_trimmed", "timeseries", "group")


def _fold_sizes(n: int, k: int) -> np.ndarray:
    sizes = np.full(k, n // k, dtype=np.int64)
    sizes[: n % k] += 1
    return sizes


def _check_kfold(n: int, k: int) -> None:
    if k < 2:
        raise ValueError(f"k-fold cross-validation requires at least one train/test split (n_splits={k})")
    if k > n:
        raise ValueError(f"Cannot have number of splits n_splits={k} greater than the number of samples: n_samples={n}.")


def kfold_indices(n: int, k: int, shuffle: bool = False) -> List[Fold]:
    """scikit-learn KFold: contiguous test blocks, the first n % k folds one sample longer.
    shuffle=True permutes with NumPy's global RNG (random_state=None) and returns SORTED indices."""
    _check_kfold(n, k)
    perm = np.arange(n, dtype=np.int64)
    if shuffle:
        np.random.shuffle(perm)
    bounds = np.concatenate([[0], np.cumsum(_fold_sizes(n, k))])
    folds = []
    for f in range(k):
        mask = np.zeros(n, dtype=bool)
        mask[perm[bounds[f]:bounds[f + 1]]] = True
        folds.append((np.flatnonzero(~mask), np.flatnonzero(mask)))
    return folds


def timeseries_indices(n: int, k: int) -> List[Fold]:
    """scikit-learn TimeSeriesSplit(n_splits=k): expanding training prefix, test blocks of n // (k + 1)."""
    if k + 1 > n:
        raise ValueError(f"Cannot have number of folds={k + 1} greater than the number of samples={n}.")
    size = n // (k + 1)
    idx = np.arange(n, dtype=np.int64)
    return [(idx[:s], idx[s:s + size]) for s in range(n - k * size, n, size)]


def group_kfold_indices(groups: Sequence, k: int) -> List[Fold]:
    """scikit-learn GroupKFold: groups sorted by size (largest first) go to the lightest fold."""
    groups = np.asarray(groups)
    uniq, inv = np.unique(groups, return_inverse=True)
    if k > len(uniq):
        raise ValueError(f"Cannot have number of splits n_splits={k} greater than the number of groups: {len(uniq)}.")
    _check_kfold(len(groups), k)
    counts = np.bincount(inv.ravel())
    # kind="stable" as scikit-learn >= 1.7 (model_selection/_split.py:640 in this image's 1.9, which is what the
    # golden fold sets were generated with).  The reference pins 1.6.0, whose default-kind argsort may break ties
    # between MANY equal-sized groups differently (introsort is unstable beyond 16 elements).
    by_size = np.argsort(counts, kind="stable")[::-1]
    load = np.zeros(k)
    fold_of_group = np.empty(len(uniq), dtype=np.int64)
    for g in by_size:
        f = int(np.argmin(load))
        load[f] += counts[g]
        fold_of_group[g] = f
    fold_of = fold_of_group[inv.ravel()]
    idx = np.arange(len(groups), dtype=np.int64)
    return [(idx[fold_of != f], idx[fold_of == f]) for f in range(k)]


def _chunk_rows(chunks: np.ndarray, chunk_length: int, n_samples: int, trim: int = 0) -> np.ndarray:
    """Row indices of the given chunks, in the given chunk order, each optionally trimmed at both ends."""
    width = chunk_length - 2 * trim
    if len(chunks) == 0 or width <= 0:
        return np.empty(0, dtype=np.int64)
    rows = chunks[:, None] * chunk_length + trim + np.arange(width, dtype=np.int64)[None, :]
    return rows.reshape(-1)  # complete chunks never run past n_samples


def chunked_folds(n_samples: int, n_folds: int, chunk_length: int, shuffle: bool = True,
                  trim_size: Optional[int] = None) -> List[Fold]:
    """folding.py:67-124 (trim_size=None) and :127-202 (trimmed test chunks)."""
    n_chunks = n_samples // chunk_length
    order = list(range(n_chunks))
    if shuffle:
        random.shuffle(order)
    per_fold = n_chunks // n_folds
    if per_fold == 0:
        logging.warning("Not enough chunks for the requested folds, falling back to regular KFold")
        return kfold_indices(n_samples, n_folds, shuffle=shuffle if trim_size is None else False)
    order = np.asarray(order, dtype=np.int64)
    folds = []
    for f in range(n_folds):
        lo = f * per_fold
        hi = (f + 1) * per_fold if f < n_folds - 1 else n_chunks
        test_chunks = order[lo:hi]
        train_chunks = np.concatenate([order[:lo], order[hi:]])
        test_rows = _chunk_rows(test_chunks, chunk_length, n_samples, 0 if trim_size is None else int(trim_size))
        folds.append((_chunk_rows(train_chunks, chunk_length, n_samples), test_rows))
    return folds


def kfold_trimmed(n_samples: int, n_folds: int, trim_size: int = 5) -> List[Fold]:
    """folding.py:205-255: KFold(shuffle=False) with trim_size rows cut from both ends of each test block
    (kept whole when it has <= 2 * trim_size rows)."""
    out = []
    for train, test in kfold_indices(n_samples, n_folds, shuffle=False):
        if len(test) > 2 * trim_size:
            test = test[trim_size:len(test) - trim_size] if trim_size else test[0:0]
        else:
            logging.warning("Test fold too small (%d samples) to trim %d from each end, keeping original test set",
                            len(test), trim_size)
        out.append((train, test))
    return out


def create_folds(n_samples: int, fold_type: str, n_folds: int, chunk_length: Optional[int] = None,
                 trim_size: Optional[int] = None, groups: Optional[Sequence] = None) -> List[Fold]:
    """Train/test row indices for every fold (encoding/models/folding.py:8-64, same positional order)."""
    if fold_type == "chunked":
        return chunked_folds(n_samples, n_folds, chunk_length, shuffle=True)
    if fold_type == "chunked_trimmed":
        return chunked_folds(n_samples, n_folds, chunk_length, shuffle=True, trim_size=5 if trim_size is None else trim_size)
    if fold_type == "chunked_contiguous":
        return chunked_folds(n_samples, n_folds, chunk_length, shuffle=False)
    if fold_type == "kfold":
        return kfold_indices(n_samples, n_folds, shuffle=False)
    if fold_type == "kfold_trimmed":
        return kfold_trimmed(n_samples, n_folds, 5 if trim_size is None else trim_size)
    if fold_type == "timeseries":
        return timeseries_indices(n_samples, n_folds)
    if fold_type == "group":
        if groups is None:
            raise ValueError("Groups must be provided for group folding")
        return group_kfold_indices(groups, n_folds)
    raise ValueError(f"Unknown folding type: {fold_type}")
