"""Parity PROOFS for whole fits (test infrastructure; imports the oracle, never imported by the product).

North star: "per-voxel test r within 1e-4 absolute; selected alpha identical except on documented near-ties;
significant-voxel count exact".  Round 1's tests bounded the FRACTION of disagreeing alphas and then dropped those
voxels from every other comparison; here every disagreement has to be justified and no voxel is dropped:

  alphas   for every outer fold and every voxel whose selected alpha differs from the reference's, the
           reference's own fold-mean inner score curve (oracle `mean_corr`, nested_cv.py:391-393) must rate the
           product's alpha within TIE_TOL = 1e-6 of its own maximum (|delta mean inner corr| < 1e-6);
           with single_alpha the same on the voxel-mean curve (nested_cv.py:396-400).  use_corr=False: the score is
           sqrt(|R^2|) sign(R^2), whose slope is unbounded at 0, so the tie is judged on R^2 itself.
  r, W     compared on ALL voxels against the reference evaluated at the PRODUCT's alphas: where the alphas agree
           that is the reference's result itself, elsewhere (proven near-ties) the oracle's ridge_torch + Pearson r
           for that alpha.  |dr| < 1e-4 (north star), and < R_FP32 = 3e-5 (what fp32 delivers).
  masks    Benjamini-Hochberg is a hard threshold on p = f(r).  A voxel is AMBIGUOUS iff the oracle's own p-value,
           re-evaluated at r -+ R_FP32, straddles the oracle's BH cut (k +- #ambiguous) alpha / V.  Masks must be
           identical and n_significant EXACT on all other voxels; the number of ambiguous voxels is returned (and
           bounded by the callers).
"""
from __future__ import annotations

import random

import numpy as np

from oracle import ridge_oracle as O

TIE_TOL = 1e-6
R_TOL = 1e-4
R_FP32 = 3e-5


def p_of_r(r, n: int, f32: bool = True):
    """Two-sided p of Pearson's r with n samples (scipy.stats.pearsonr: Beta(n/2-1, n/2-1) on (-1, 1))."""
    from scipy.special import betainc

    ab = n / 2.0 - 1.0
    p = np.minimum(2.0 * betainc(ab, ab, 0.5 * (1.0 - np.minimum(np.abs(np.asarray(r, dtype=np.float64)), 1.0))), 1.0)
    return p.astype(np.float32).astype(np.float64) if f32 else p


def _grid_index(values, grid):
    grid = np.asarray(grid, dtype=np.float64)
    return np.argmin(np.abs(np.log(grid)[None, :] - np.log(np.asarray(values, dtype=np.float64))[:, None]), axis=1)


def _tie_scale(curves, use_corr: bool):
    c = np.asarray(curves, dtype=np.float64)
    return c if use_corr else np.sign(c) * c * c


def prove_alpha_ties(idx_prod, idx_ref, curves, use_corr=True, single_alpha=False, tol=TIE_TOL, what="",
                     exempt=None):
    """Every voxel whose alpha index differs must be a near-tie on the reference's score curves (A x V).
    exempt: exactly constant response columns (DESIGN.md divergence (i): the reference's z-scored constant column is
    rounding noise, its R^2 score 0/0 or -inf, and its "selected" alpha arbitrary; r = 0 and p = 1 on both sides)."""
    S = _tie_scale(curves, use_corr)
    if single_alpha:
        Sm = _tie_scale(np.asarray(curves, dtype=np.float32).mean(axis=1, dtype=np.float32), use_corr)
        jp, jr = int(idx_prod[0]), int(idx_ref[0])
        assert Sm[jp] >= Sm[jr] - tol, f"{what}: single alpha {jp} vs reference {jr}: gap {Sm[jr] - Sm[jp]:.3e}"
        return int(jp != jr) * len(idx_prod)
    differ = idx_prod != idx_ref
    if exempt is not None:
        differ &= ~exempt
    dis = np.nonzero(differ)[0]
    if len(dis) == 0:
        return 0
    gap = S[idx_ref[dis], dis] - S[idx_prod[dis], dis]
    worst = int(np.argmax(gap))
    assert gap.max() <= tol, (f"{what}: {int((gap > tol).sum())} of {len(dis)} disagreeing alphas are NOT near-ties; worst "
                              f"voxel {dis[worst]}: product alpha #{idx_prod[dis][worst]} scores {gap.max():.3e} below the "
                              f"reference's #{idx_ref[dis][worst]}")
    return len(dis)


def _bh_ambiguous(p_lo, p_hi, p_star, alpha_fdr):
    """Voxels whose p-interval [p_lo, p_hi] meets the band of BH cuts reachable when the ambiguous voxels flip."""
    V = len(p_star)
    k = int(O.fdr_bh(p_star, alpha_fdr)[0].sum())
    amb = np.zeros(V, dtype=bool)
    for _ in range(8):
        n_amb = int(amb.sum())
        t_lo, t_hi = max(k - n_amb, 0) * alpha_fdr / V, (k + n_amb + 1) * alpha_fdr / V
        new = (p_lo <= t_hi * (1 + 1e-6)) & (p_hi >= t_lo * (1 - 1e-6))
        if (new == amb).all():
            break
        amb = new | amb
    return amb


def golden_folds(g, name):
    """Per-outer-fold observations of the unmodified reference (tests/golden/fit_predict_folds.npz)."""
    return [{k: g[f"{name}__f{f}__{k}"] for k in ("mean_corr", "best", "r", "p")} for f in range(int(g[f"{name}__n_folds"]))]


def fold_results_of_oracle(details, alpha_fdr=0.05):
    """The oracle's own per-fold results in the shape of NestedCVModel.last_fold_results (to prove the ORACLE
    against the reference's recorded folds with the same tool)."""
    return {"alphas": np.stack([np.asarray(d["best"], dtype=np.float32) for d in details]),
            "correlations": np.stack([np.asarray(d["r"], dtype=np.float32) for d in details]),
            "p_values": np.stack([np.asarray(d["p"], dtype=np.float64) for d in details]),
            "masks": np.stack([O.fdr_bh(np.asarray(d["p"]), alpha_fdr)[0] for d in details])}


def _bh_check(p_star, lo, hi, mask, alpha_fdr, cut, what):
    """Compare a product BH mask with the oracle's decisions on p_star; returns the ambiguous-voxel mask.
    cut = None: the voxels at hand are ALL voxels of the fit (BH is recomputed on p_star).  cut = (k, V): they are a
    STRIPE of a larger fit that rejected k of V hypotheses, so the cut k alpha / V comes from the full fit (the band
    allows k to move by 0.1 %)."""
    if cut is None:
        amb = _bh_ambiguous(lo, hi, p_star, alpha_fdr)
        star = O.fdr_bh(p_star, alpha_fdr)[0]
    else:
        tau = cut[0] * alpha_fdr / cut[1]
        amb = (lo <= tau * 1.001) & (hi >= tau * 0.999)
        star = p_star <= tau
    bad = (np.asarray(mask, dtype=bool) != star) & ~amb
    assert not bad.any(), f"{what}: {int(bad.sum())} BH decisions differ away from the threshold"
    return amb, star


def prove_fit_parity(fold_results, metrics, weights, features, targets, seed, X_test=None, y_test=None, ref_folds=None,
                     w_tol=1e-4, r_tol=R_TOL, r_fp32=R_FP32, max_ambiguous=None, bh_cuts=None, **kw):
    """Run the oracle on the same inputs / seed and prove parity of a fit (see the module docstring).
    fold_results: NestedCVModel.last_fold_results of the fit that returned (metrics, weights); kw: the fit_predict
    keyword arguments both sides received.  ref_folds: per-fold observations of the UNMODIFIED reference
    (golden_folds); when given, alphas, score curves, r and p of the reference itself are what the product is held
    to, and the oracle only supplies the refits at the product's alphas.  bh_cuts: when `targets` is a voxel STRIPE
    of a larger fit, {"V": total voxels, "folds": [k per outer fold], "final": k}: the numbers of hypotheses the full
    fit rejected (Benjamini-Hochberg is global over voxels).  Returns a dict of what was observed."""
    alphas = kw.get("alphas")
    alphas = np.logspace(-1, 8, 10) if alphas is None else alphas
    use_corr, single = kw.get("use_corr", True), kw.get("single_alpha", False)
    alpha_fdr = kw.get("alpha_fdr", 0.05)
    details = []
    random.seed(seed)
    np.random.seed(seed)
    okw = {k: v for k, v in kw.items() if k not in ("use_gpu",)}
    mo, wo, ao = O.fit_predict(features, targets, X_test=X_test, y_test=y_test, vectorised_stats=True, details=details,
                               **okw)
    fr = fold_results
    n_folds = len(details)
    assert fr["alphas"].shape[0] == n_folds and (ref_folds is None or len(ref_folds) == n_folds)
    V = np.asarray(targets).shape[1]
    r_star, p_star, w_star, n_dis, max_dr = [], [], [], 0, 0.0
    for f, d in enumerate(details):
        ref = ref_folds[f] if ref_folds is not None else d
        idx_p, idx_r, idx_o = (_grid_index(v, alphas) for v in (fr["alphas"][f], ref["best"], d["best"]))
        const = d["Ytr"].max(axis=0) == d["Ytr"].min(axis=0)
        n_dis += prove_alpha_ties(idx_p, idx_r, ref["mean_corr"], use_corr, single, what=f"outer fold {f}", exempt=const)
        rs, ps, ws = np.array(ref["r"], dtype=np.float64), np.array(ref["p"], dtype=np.float64), d["wt"].copy()
        # the reference evaluated at the PRODUCT's alphas: refit (oracle ridge_torch + Pearson) where they differ
        dis = np.nonzero((idx_p != idx_r) | (idx_p != idx_o))[0]
        if len(dis):
            a_dis = np.asarray(alphas, dtype=np.float64)[idx_p[dis]].astype(np.float32)
            wt = O.ridge_weights(d["Xtr"], d["Ytr"][:, dis], a_dis, singcutoff=kw.get("singcutoff", 1e-10),
                                 normalpha=kw.get("normalpha", True))
            r_d, p_d = O.correlations_pvalues_vectorised(d["Yte"][:, dis], d["Xte"] @ wt)
            keep = idx_p[dis] == idx_r[dis]  # the reference's own r / p stand wherever ITS alpha is the product's
            rs[dis], ps[dis] = np.where(keep, rs[dis], r_d), np.where(keep, ps[dis], p_d)
            ws[:, dis] = wt
        dr = np.abs(fr["correlations"][f].astype(np.float64) - rs)
        assert dr.max() < r_tol, f"outer fold {f}: |dr| = {dr.max():.3e} at voxel {int(dr.argmax())} (north star 1e-4)"
        assert dr.max() < r_fp32, f"outer fold {f}: |dr| = {dr.max():.3e} at voxel {int(dr.argmax())} (fp32 level)"
        max_dr = max(max_dr, float(dr.max()))
        n_te = d["Yte"].shape[0]
        lo, hi = p_of_r(np.abs(rs) + r_fp32, n_te), p_of_r(np.maximum(np.abs(rs) - r_fp32, 0.0), n_te)
        _bh_check(ps, lo, hi, fr["masks"][f], alpha_fdr, None if bh_cuts is None else (bh_cuts["folds"][f], bh_cuts["V"]),
                  f"outer fold {f}")
        r_star.append(rs), p_star.append(ps), w_star.append(ws)
    # aggregation (nested_cv.py:276-296)
    r_star, p_star = np.asarray(r_star), np.asarray(p_star)
    nts = [d["Yte"].shape[0] for d in details]
    if n_folds == 1:
        corr, comb, w_ref = r_star[0], p_star[0], w_star[0]
        lo, hi = p_of_r(np.abs(corr) + r_fp32, nts[0]), p_of_r(np.maximum(np.abs(corr) - r_fp32, 0.0), nts[0])
    else:
        corr = np.mean(r_star.astype(np.float32), axis=0).astype(np.float64)
        comb = O.fisher_combine_vectorised(p_star)
        w_ref = np.mean(w_star, axis=0)
        lo = O.fisher_combine_vectorised(np.asarray([p_of_r(np.abs(r_star[f]) + r_fp32, nts[f]) for f in range(n_folds)]))
        hi = O.fisher_combine_vectorised(np.asarray([p_of_r(np.maximum(np.abs(r_star[f]) - r_fp32, 0.0), nts[f])
                                                     for f in range(n_folds)]))
    r_out = np.asarray(metrics["correlations"], dtype=np.float64)
    assert np.abs(r_out - corr).max() < r_fp32
    sig = np.asarray(metrics["significant_mask"], dtype=bool)
    amb, sig_star = _bh_check(comb, lo, hi, sig, alpha_fdr, None if bh_cuts is None else (bh_cuts["final"], bh_cuts["V"]),
                              "combined p-values")
    assert int(sig[~amb].sum()) == int(sig_star[~amb].sum())  # n_significant exact off the threshold band
    assert abs(int(sig.sum()) - int(sig_star.sum())) <= int(amb.sum())
    if max_ambiguous is not None:
        assert int(amb.sum()) <= max_ambiguous, f"{int(amb.sum())} voxels sit on the BH threshold"
    werr = float(np.abs(np.asarray(weights, dtype=np.float64) - w_ref).max() / max(np.abs(w_ref).max(), 1e-30))
    assert werr < w_tol, f"weights differ by {werr:.3e} of max|W| on some voxel"
    return {"disagreeing_alphas": n_dis, "voxel_folds": V * n_folds, "max_dr": max_dr, "ambiguous_bh": int(amb.sum()),
            "n_significant": int(sig.sum()), "n_significant_oracle": int(mo["n_significant"]),
            "n_significant_at_product_alphas": int(sig_star.sum()), "weights_rel_err": werr,
            "oracle": (mo, wo, ao), "details": details}
