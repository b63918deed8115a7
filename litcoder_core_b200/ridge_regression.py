"""Drop-ins for the stand-alone ridge kernels of the reference (encoding/models/ridge_regression.py):

    ridge_torch(Rstim, Rresp, alphas, singcutoff=1e-30, normalpha=False)                      :9-63
    ridge_corr_torch(Rstim, Pstim, Rresp, Presp, alphas, singcutoff, use_corr, normalpha)      :66-141
    ridge_corr_pred_torch(Rstim, Pstim, Rresp, Presp, valphas, singcutoff, use_corr, normalpha) :144-216

and of the column z-scoring the reference's trainers apply before the fit (`zs`, encoding/utils.py:23-34).
Arguments may be NumPy arrays or torch tensors; results come back as NumPy arrays for NumPy input and as torch
tensors on the input's device for torch input.  The arithmetic is the Gram-side formulation of engine.py on the
B200 (no CPU path).
"""
from __future__ import annotations

from typing import Optional, Sequence, Union

import numpy as np

from .engine import EPS, RidgeConfig, RidgeCVEngine


def _ops(ops):
    if ops is None:
        from .device import default_ops

        ops = default_ops()
    return ops


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def _host(x) -> np.ndarray:
    return x.detach().cpu().numpy() if _is_torch(x) else np.asarray(x)


def _like(result: np.ndarray, template):
    if _is_torch(template):
        import torch

        return torch.from_numpy(np.ascontiguousarray(result)).to(template.device)
    return result


def _stack(ops, a, b):
    """Device matrix holding the rows of a followed by the rows of b (columns must agree)."""
    a, b = _host(a), _host(b)
    if a.shape[1] != b.shape[1]:
        raise ValueError("training and prediction matrices must have the same number of columns")
    return ops.upload_matrix(np.vstack([a.astype(np.float32, copy=False), b.astype(np.float32, copy=False)]))


def ridge_corr(Rstim, Pstim, Rresp, Presp, alphas: Sequence[float], singcutoff: float = 1e-30, use_corr: bool = True,
               normalpha: bool = False, logger=None, ops=None):
    """Validation score of every voxel for every alpha -> (n_alphas, n_voxels) float32."""
    ops = _ops(ops)
    cfg = RidgeConfig(alphas=[float(a) for a in alphas], normalpha=normalpha, use_corr=use_corr, singcutoff=singcutoff)
    eng = RidgeCVEngine(ops)
    corr = eng.ridge_corr(_stack(ops, Rstim, Pstim), _stack(ops, Rresp, Presp), _host(Rstim).shape[0], cfg)
    out = ops.download_matrix(corr)
    ops.check_eig()
    if logger is not None:
        for a, row in zip(alphas, out):
            logger.info("Alpha=%.3f, mean corr=%.5f, max corr=%.5f", a, float(row.mean()), float(row.max()))
    return _like(out, Rstim)


def ridge(Rstim, Rresp, alphas: Union[float, Sequence[float]], singcutoff: float = 1e-30, normalpha: bool = False,
          ops=None):
    """Ridge weights for a scalar alpha or one alpha per voxel -> (n_features, n_voxels) float32."""
    ops = _ops(ops)
    Y = ops.upload_matrix(_host(Rresp))
    av = _host(alphas) if not isinstance(alphas, (int, float)) else np.full(Y.cols, float(alphas))
    if av.ndim == 0:
        av = np.full(Y.cols, float(av))
    cfg = RidgeConfig(alphas=[0.0], normalpha=normalpha, singcutoff=singcutoff)
    eng = RidgeCVEngine(ops)
    Wt = eng.ridge_weights(ops.upload_matrix(_host(Rstim)), Y, ops.upload_vector(av.astype(np.float32), "f32"), cfg)
    full = ops.zeros(Wt.rows, Wt.cols)
    ops.axpy(1.0, Wt, full)  # recombine the split pair
    out = ops.download_matrix(ops.transpose(full))
    ops.check_eig()
    return _like(out, Rstim)


def ridge_corr_pred(Rstim, Pstim, Rresp, Presp, valphas, singcutoff: float = 1e-30, use_corr: bool = True,
                    normalpha: bool = True, ops=None):
    """Score of the per-voxel-alpha predictions without forming the weights -> (n_voxels,) float32."""
    ops = _ops(ops)
    YY = _stack(ops, Rresp, Presp)
    av = _host(valphas).astype(np.float32).reshape(-1)
    if av.size == 1:
        av = np.full(YY.cols, av[0], dtype=np.float32)
    cfg = RidgeConfig(alphas=[0.0], normalpha=normalpha, use_corr=use_corr, singcutoff=singcutoff)
    eng = RidgeCVEngine(ops)
    corr = eng.ridge_corr_pred(_stack(ops, Rstim, Pstim), YY, _host(Rstim).shape[0], ops.upload_vector(av, "f32"), cfg)
    out = ops.download_matrix(corr)[0]
    ops.check_eig()
    return _like(out, Rstim)


def zs(v, ops=None):
    """Column z-score with the population std; zero-variance columns are only centred
    (`zscore` / `zs`, encoding/utils.py:23-34).  Keeps the input's floating dtype."""
    ops = _ops(ops)
    a = _host(v)
    M = ops.upload_matrix(a)
    mean, std = ops.col_stats(M, None, M.rows, ddof=0)
    out = ops.download_matrix(ops.gather_normalize(M, None, M.rows, mean, std, 3, EPS))
    return _like(out.astype(a.dtype if a.dtype.kind == "f" else np.float64), v)


# the reference's names
ridge_torch = ridge
ridge_corr_torch = ridge_corr
ridge_corr_pred_torch = ridge_corr_pred
zscore = zs
