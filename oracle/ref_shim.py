"""Loader for the UNMODIFIED reference package (test / bench infrastructure; never imported by the product).

`python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref /root/reference` puts the
reference's pure-Python `encoding` package under baseline/_ref (git-ignored, travels to the GPU box with the
snapshot; dependency resolution fails offline, hence --no-deps).  Its package __init__ files import optional
dependencies that are absent from this image: transformer_lens, gensim, h5py are stubbed in
sys.modules (imported, never executed on this path); statsmodels' `fdrcorrection` (nested_cv.py:158,263,282) is the
only stubbed function that runs -- it is restated here exactly as in scripts/make_golden.py and pinned by the
known-answer tests of tests/test_oracle_golden.py.

    ref = load_reference()          # None when neither baseline/_ref nor /root/reference exists
    ref.NestedCVModel, ref.ridge_corr_torch, ref.ridge_torch, ref.FIR, ref.Downsampler, ref.path
"""
from __future__ import annotations

import os
import sys
import types
from types import SimpleNamespace
from typing import Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = (os.path.join(ROOT, "baseline", "_ref"), "/root/reference")


def _fdrcorrection(pvals, alpha=0.05, method="indep", is_sorted=False):
    """statsmodels 0.14.4 `fdrcorrection(method='indep')` (Benjamini-Hochberg), restated."""
    p = np.asarray(pvals)
    m = len(p)
    o = np.argsort(p)
    ps = p[o]
    c = np.arange(1, m + 1) / float(m)
    rej = ps <= c * alpha
    if rej.any():
        rej[: np.max(np.nonzero(rej)[0])] = True
    adj = np.minimum.accumulate((ps / c)[::-1])[::-1]
    adj[adj > 1] = 1
    r = np.empty_like(rej)
    a = np.empty_like(adj)
    r[o] = rej
    a[o] = adj
    return r, a


def _stub(name, **attrs):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m


_loaded: Optional[SimpleNamespace] = None


def load_reference() -> Optional[SimpleNamespace]:
    global _loaded
    if _loaded is not None:
        return _loaded
    path = next((p for p in CANDIDATES if os.path.isdir(os.path.join(p, "encoding", "models"))), None)
    if path is None:
        return None
    _stub("transformer_lens", HookedTransformer=object)
    _stub("gensim")
    _stub("gensim.models", KeyedVectors=object)
    _stub("h5py")
    try:
        import statsmodels.stats.multitest  # noqa: F401  (the real thing, when present)
    except Exception:  # noqa: BLE001
        _stub("statsmodels")
        _stub("statsmodels.stats")
        _stub("statsmodels.stats.multitest", fdrcorrection=_fdrcorrection)
    sys.path.insert(0, path)
    try:
        from encoding.downsample.downsampling import Downsampler
        from encoding.features.FIR_expander import FIR
        from encoding.models.nested_cv import NestedCVModel
        from encoding.models.ridge_regression import ridge_corr_torch, ridge_torch
        try:
            from encoding.utils import ModelSaver  # the consumer of fit_predict's triple (utils.py:288-354)
        except Exception:  # noqa: BLE001
            ModelSaver = None
    finally:
        sys.path.remove(path)
    _loaded = SimpleNamespace(NestedCVModel=NestedCVModel, ridge_corr_torch=ridge_corr_torch, ridge_torch=ridge_torch,
                              FIR=FIR, Downsampler=Downsampler, ModelSaver=ModelSaver, path=path)
    return _loaded
