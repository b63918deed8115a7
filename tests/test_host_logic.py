"""CPU tests of the product's host logic: fold construction against the reference's golden index
sets, and the nested-CV engine / API driven through a NumPy stand-in of the C ABI (tests/fake_ops.py)
against the oracle and the reference's golden fit_predict outputs."""
import inspect
import random

import numpy as np
import pytest

from conftest import load_golden, unpack_folds
from fake_ops import FakeOps
from parity import prove_fit_parity
from oracle import ridge_oracle as O

import litcoder_core_b200 as L
from litcoder_core_b200 import folding
from litcoder_core_b200.engine import removed_rows
from litcoder_core_b200.nested_cv import NestedCVModel, shard_bounds


def _cases(npz, suffix):
    return sorted({k.split("__")[0] for k in npz.files if k.endswith(suffix)})


# ------------------------------------------------------------------------------------------ folds
def test_product_folds_match_reference_golden():
    g = load_golden("folds.npz")
    for name in _cases(g, "__flat"):
        n, k, chunk, trim, seed = (int(x) for x in g[f"{name}__args"])
        ftype = str(g[f"{name}__type"])
        groups = g[f"{name}__groups"] if f"{name}__groups" in g.files else None
        random.seed(seed)
        np.random.seed(seed)
        folds = folding.create_folds(n, ftype, k, None if chunk < 0 else chunk, None if trim < 0 else trim, groups)
        ref = unpack_folds(g[f"{name}__flat"], g[f"{name}__offs"])
        assert len(folds) == len(ref), name
        for (tr, te), (rtr, rte) in zip(folds, ref):
            np.testing.assert_array_equal(np.asarray(tr), rtr, err_msg=name)
            np.testing.assert_array_equal(np.asarray(te), rte, err_msg=name)


@pytest.mark.parametrize("ftype", ["chunked", "chunked_trimmed", "chunked_contiguous", "kfold", "kfold_trimmed",
                                   "timeseries", "group"])
@pytest.mark.parametrize("n,k,chunk", [(9407, 5, 20), (233, 4, 10), (50, 5, 20), (61, 3, 7)])
def test_product_folds_match_oracle(ftype, n, k, chunk):
    groups = np.random.default_rng(n).integers(0, 9, n) if ftype == "group" else None
    for trim in (None, 2):
        random.seed(n + k)
        np.random.seed(n + k)
        a = folding.create_folds(n, ftype, k, chunk, trim, groups)
        random.seed(n + k)
        np.random.seed(n + k)
        b = O.create_folds(n, ftype, k, chunk, trim, groups)
        assert len(a) == len(b)
        for (tr, te), (rtr, rte) in zip(a, b):
            np.testing.assert_array_equal(tr, np.asarray(rtr, dtype=np.int64))
            np.testing.assert_array_equal(te, np.asarray(rte, dtype=np.int64))


def test_fold_errors():
    with pytest.raises(ValueError, match="Unknown folding type"):
        folding.create_folds(100, "nope", 5, 20)
    with pytest.raises(ValueError, match="Groups must be provided"):
        folding.create_folds(100, "group", 5, 20)
    # kfold_trimmed with trim 0 reproduces the reference's `test[0:-0]` quirk: an empty test set
    assert all(len(te) == 0 for _, te in folding.create_folds(100, "kfold_trimmed", 5, None, 0))


def test_removed_rows():
    outer = np.array([5, 6, 7, 0, 1, 2, 10, 11])
    np.testing.assert_array_equal(removed_rows(outer, np.array([0, 1, 2, 5, 6, 7])), [10, 11])
    assert removed_rows(outer, np.array([0, 1, 99])) is None  # not a subset
    assert removed_rows(outer, np.array([0, 0, 1])) is None  # duplicates
    assert len(removed_rows(outer, outer)) == 0


def test_shard_bounds():
    assert shard_bounds(95000, 1) == [0, 95000]
    b = shard_bounds(95000, 8)
    assert b[0] == 0 and b[-1] == 95000 and all(x % 128 == 0 for x in b[:-1])
    assert all(b[i + 1] >= b[i] for i in range(8))
    assert shard_bounds(100, 4) == [0, 100, 100, 100, 100]  # fewer than one tile per rank: trailing ranks are empty


# ------------------------------------------------------------------------------------------ API surface
def test_signatures_match_reference_docs():
    sig = inspect.signature(NestedCVModel.fit_predict)
    names = list(sig.parameters)
    ref = ["self", "features", "targets", "X_test", "y_test", "groups", "folding_type", "n_outer_folds",
           "n_inner_folds", "chunk_length", "alphas", "alpha_fdr", "use_gpu", "single_alpha", "normalpha", "use_corr",
           "normalize_features", "normalize_targets", "singcutoff"]
    assert names[: len(ref)] == ref  # nested_cv.py:18-37 (extensions only after the reference's arguments)
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d["folding_type"], d["n_outer_folds"], d["n_inner_folds"], d["chunk_length"]) == ("chunked", 5, 5, 20)
    assert (d["alpha_fdr"], d["single_alpha"], d["normalpha"], d["use_corr"], d["singcutoff"]) == \
        (0.05, False, True, True, 1e-10)
    assert list(inspect.signature(L.FIR.make_delayed).parameters)[:3] == ["stim", "delays", "circpad"]
    assert list(inspect.signature(L.Downsampler.downsample).parameters)[:5] == \
        ["self", "data", "data_times", "tr_times", "method"]
    ds = L.Downsampler(ops=FakeOps())
    assert ds.available_methods == ["rect", "average", "sinc", "lanczos", "last", "gabor", "legacy_average",
                                    "legacy_last", "sum", "legacy_sum"]  # downsampling.py:348-359
    assert ds.get_method_params("lanczos") == {"required": ["window", "cutoff_mult"], "optional": ["rectify"]}
    with pytest.raises(ValueError, match="Unsupported downsampling method"):
        ds.downsample(np.zeros((3, 2)), np.arange(3.0), np.arange(2.0), method="cubic")
    with pytest.raises(ValueError, match="Required parameter 'window' missing"):
        ds.downsample(np.zeros((3, 2)), np.arange(3.0), np.arange(2.0), method="lanczos", cutoff_mult=1.0)
    with pytest.raises(ValueError, match="delays must be provided"):
        L.FIR().expand(np.zeros((3, 2)))
    fir = L.FIR(delays=[1, 2, 3, 4])
    assert fir.n_delays() == 4 and fir.output_dim(768) == 3072 and fir.valid_length(100) == 96
    assert "Output dim: 8" in fir.summary(input_dim=2)


# ------------------------------------------------------------------------------------------ FIR / Lanczos host glue
def test_fir_facade_on_fake_ops_matches_golden():
    g = load_golden("fir.npz")
    ops = FakeOps()
    for name in _cases(g, "__out"):
        out = L.FIR.make_delayed(g[f"{name}__stim"], g[f"{name}__delays"].tolist(), bool(g[f"{name}__circpad"]), ops=ops)
        ref = g[f"{name}__out"]
        assert out.dtype == ref.dtype and out.shape == ref.shape, name
        np.testing.assert_array_equal(out, ref, err_msg=name)


def test_lanczos_facade_on_fake_ops_matches_golden():
    g = load_golden("lanczos.npz")
    ds = L.Downsampler(ops=FakeOps())
    for name in _cases(g, "__out"):
        w, cm, rect = g[f"{name}__params"]
        out = ds.downsample(g[f"{name}__data"], g[f"{name}__data_times"], g[f"{name}__tr_times"], method="lanczos",
                            window=int(w), cutoff_mult=float(cm), rectify=bool(rect), split_indices=None)
        ref = g[f"{name}__out"]
        assert out.dtype == np.float64 and out.shape == ref.shape, name
        np.testing.assert_allclose(out, ref, rtol=1e-10, atol=1e-12, err_msg=name)


# ------------------------------------------------------------------------------------------ engine end to end
RUNS = {
    "tt_default": dict(train_test=True),
    "tt_single": dict(train_test=True, single_alpha=True),
    "tt_norm": dict(train_test=True, normalize_features=True, normalize_targets=True),
    "tt_nonormalpha": dict(train_test=True, normalpha=False),
    "tt_rsq": dict(train_test=True, use_corr=False),
    "cv_default": dict(train_test=False),
    "cv_single": dict(train_test=False, single_alpha=True),
    "cv_kfold": dict(train_test=False, folding_type="kfold"),
    "cv_norm": dict(train_test=False, normalize_targets=True),
}


def _run_product(name, **extra):
    g = load_golden("fit_predict.npz")
    X, Y, alphas = g["X"], g["Y"], list(g["alphas"])  # np.float64 elements, as the generator passed them
    kw = dict(RUNS[name])
    tt = kw.pop("train_test")
    common = dict(folding_type="chunked", n_outer_folds=4, n_inner_folds=3, chunk_length=10, alphas=alphas)
    common.update(kw)
    common.update(extra)
    random.seed(7)
    np.random.seed(7)
    model = NestedCVModel("ridge_regression", ops=FakeOps())
    if tt:
        return g, model.fit_predict(X[:400], Y[:400], X_test=X[400:], y_test=Y[400:], **common)
    return g, model.fit_predict(X[:400], Y[:400], **common)


GRID20 = {"tt_grid20": dict(train_test=True), "cv_grid20": dict(train_test=False),
          "cv_grid20_single": dict(train_test=False, single_alpha=True)}


def _golden_args(name):
    """(golden file of the run, X, Y, test-set kwargs, fit_predict kwargs) of one of the 12 golden runs."""
    g0 = load_golden("fit_predict.npz")
    X, Y = g0["X"], g0["Y"]
    if name in GRID20:
        g, kw = load_golden("fit_predict_grid20.npz"), dict(GRID20[name])
        alphas = g["alphas"].tolist()
    else:
        g, kw = g0, dict(RUNS[name])
        alphas = list(g0["alphas"])  # np.float64 elements, as the generator passed them
    tt = kw.pop("train_test")
    common = dict(folding_type="chunked", n_outer_folds=4, n_inner_folds=3, chunk_length=10, alphas=alphas)
    common.update(kw)
    test = dict(X_test=X[400:], y_test=Y[400:]) if tt else {}
    return g, X[:400], Y[:400], test, common


def check_against_reference_golden(name, fold_results, m, w, va, max_ambiguous=3):
    """Shared by the oracle, CPU (fake ops) and GPU suites: parity PROOF (tests/parity.py) against the per-fold
    observations of the unmodified reference (fit_predict_folds.npz), plus the return contract."""
    from parity import golden_folds, prove_fit_parity

    g, X, Y, test, common = _golden_args(name)
    gf = load_golden("fit_predict_folds.npz")
    ref_va, ref_r = g[f"{name}__best_alphas"], g[f"{name}__m__correlations"]
    np.testing.assert_array_equal(gf[f"{name}__best_alphas"], ref_va)  # the recorded folds belong to this very run
    assert va.dtype == ref_va.dtype and va.shape == ref_va.shape
    info = prove_fit_parity(fold_results, m, w, X, Y, 7, ref_folds=golden_folds(gf, name), max_ambiguous=max_ambiguous,
                            **test, **common)
    same = np.isclose(va, ref_va, rtol=1e-6)
    r = np.asarray(m["correlations"], dtype=np.float64)
    np.testing.assert_allclose(r[same], ref_r[same], atol=3e-5)
    assert set(m.keys()) == {k.split("__m__")[1] for k in g.files if k.startswith(f"{name}__m__")}
    assert list(m.keys())[:5] == ["median_score", "mean_score", "std_score", "min_score", "max_score"]
    wref = g[f"{name}__weights"]
    assert w.shape == wref.shape and w.dtype == wref.dtype
    if info["disagreeing_alphas"] == 0:  # then everything must match the reference's returned values directly
        assert abs(m["n_significant"] - int(g[f"{name}__m__n_significant"])) <= info["ambiguous_bh"]
        assert np.abs(w - wref).max() < 1e-4 * np.abs(wref).max()
    return info


@pytest.mark.parametrize("name", sorted(RUNS))
def test_engine_on_fake_ops_matches_reference_golden(name):
    g, X, Y, test, common = _golden_args(name)
    random.seed(7)
    np.random.seed(7)
    model = NestedCVModel("ridge_regression", ops=FakeOps())
    m, w, va = model.fit_predict(X, Y, **test, **common)
    info = check_against_reference_golden(name, model.last_fold_results, m, w, va)
    assert info["disagreeing_alphas"] <= (0.3 if name == "tt_rsq" else 0.1) * info["voxel_folds"]


@pytest.mark.parametrize("name", ["tt_default", "cv_default"])
def test_results_hand_off_to_the_reference_trainer(name, tmp_path):
    """What AbstractTrainer.train does with the triple after fit_predict: log_metrics reads scalars
    and arrays out of the metrics dict (trainer.py:322-336), ModelSaver pickles the dict and np.saves the weights
    (encoding/utils.py:324-354)."""
    import pickle

    _, (m, w, a) = _run_product(name)
    for k in ("median_score", "mean_score", "std_score"):
        assert isinstance(float(m[k]), float)
    corr = np.array(m["correlations"])
    mask = np.array(m["significant_mask"], dtype=bool)
    assert corr.shape == mask.shape == (w.shape[1],) and float(m["n_significant"]) == mask.sum()
    back = pickle.loads(pickle.dumps(m))
    assert back.keys() == m.keys() and back["correlations"] == m["correlations"]
    np.save(tmp_path / "weights.npy", w)
    np.testing.assert_array_equal(np.load(tmp_path / "weights.npy"), w)
    assert isinstance(m["correlations"], list) and isinstance(m["significant_mask"], list)
    assert a.shape == (w.shape[1],)


@pytest.mark.parametrize("name", ["tt_default", "cv_default", "cv_kfold"])
def test_downdate_and_overlap_do_not_change_results(name):
    """The Gram / cross-product downdates and the asynchronous eigendecompositions are pure
    re-orderings: results must agree with the direct, synchronous schedule to fp32 noise."""
    from litcoder_core_b200 import engine as E

    _, (m1, w1, a1) = _run_product(name)
    orig = E.RidgeConfig.__init__

    def patched(self, *a, **k):
        orig(self, *a, **k)
        self.downdate = False
        self.overlap_eig = False

    E.RidgeConfig.__init__ = patched
    try:
        _, (m2, w2, a2) = _run_product(name)
    finally:
        E.RidgeConfig.__init__ = orig
    same = np.isclose(a1, a2)
    assert same.mean() > 0.95
    np.testing.assert_allclose(np.asarray(m1["correlations"])[same], np.asarray(m2["correlations"])[same], atol=1e-5)
    assert np.abs(w1[:, same] - w2[:, same]).max() <= 2e-5 * np.abs(w2).max()


def test_bound_scales_cannot_overflow_fp16():
    """The a-priori bounds behind the producer-written fp16 pairs (DeviceOps.f16_bound_scales, restated in
    tests/fake_ops.py): gathered response rows never exceed the column maximum over all rows, and a downdated cross
    product C_o^T - Y_R^T X_R never exceeds max_j |C_o^T[v][j]| + |y_(v,R)|_2 max_j |x_(j,R)|_2 -- also when the outer
    cross product cancels to rounding noise, when a voxel is zero outside the removed rows, and with outliers."""
    from fake_ops import FakeOps, FMat

    ops = FakeOps()
    rng = np.random.default_rng(12)
    N, p, V, nR = 600, 40, 64, 130
    X = rng.standard_normal((N, p)).astype(np.float32)
    R = np.sort(rng.choice(N, nR, replace=False))
    Y = rng.standard_normal((N, V)).astype(np.float32)
    # voxel 0: orthogonal to every feature over all rows (C_o row = rounding noise), large on the removed rows
    y0 = rng.standard_normal(N)
    y0[R] *= 50.0
    y0 -= X.astype(np.float64) @ np.linalg.lstsq(X.astype(np.float64), y0, rcond=None)[0]
    Y[:, 0] = y0
    Y[:, 1] = 0.0
    Y[R, 1] = 1e4          # zero outside the removed rows
    Y[R[3], 2] = 3e7       # one outlier inside the removed rows
    Y[5, 3] = -2e6         # one outlier outside them
    Y[:, 4] = 0.0          # all-zero voxel: scale 1
    Y[:, 5] *= 1e-20       # tiny voxel
    Ct_o = (Y.T.astype(np.float64) @ X.astype(np.float64)).astype(np.float32)
    exact = Ct_o.astype(np.float64) - Y[R].T.astype(np.float64) @ X[R].astype(np.float64)
    ys = ops.f16_bound_scales(V, absmax=ops.col_reduce(FMat(Y), None, N, sumsq=False, absmax=True)[1])
    assert ys[0][4] == 1.0 and np.all(np.abs(Y).max(0) * ys[0] < 2.0 ** 15)
    T = ops.gather_rows_T_f16(FMat(Y), R, nR, ys)  # asserts |scaled value| < 65504 inside
    np.testing.assert_allclose(T.a[[0, 6, 7]], Y[R].T[[0, 6, 7]], rtol=2.0 ** -20)
    sc = ops.f16_bound_scales(V, absmax=ops.row_absmax(FMat(Ct_o)), row_sumsq=ops.col_reduce(FMat(Y), R, nR)[0],
                              col_sumsq=ops.col_reduce(FMat(X), R, nR)[0])
    bound = 2.0 ** 15 / sc[0].astype(np.float64)
    assert np.all(np.abs(exact).max(1) <= bound)  # the bound holds for every voxel, the adversarial ones included ...
    ordinary = np.arange(V) >= 6
    assert np.all(bound[ordinary] <= 2.0 ** 4 * np.abs(exact).max(1)[ordinary])  # ... and is tight for ordinary ones
    H = ops.gemm(T, FMat(X[R].T.copy(), split=True), alpha=-1.0, Cin=FMat(Ct_o), beta=1.0, precision="f16x3",
                 pair_out=sc)  # asserts no overflow inside
    # what the pair keeps: 2^-21 of the element or 2^-39 of the bound, whichever is larger (lit_split_f16's contract
    # with the bound in place of the row maximum) -- measured against the fp32 result the epilogue converts
    d32 = (Ct_o.astype(np.float64) - T.a.astype(np.float64) @ ops._f16_pair_value(X[R].T, 1).T).astype(np.float32)
    tol = np.maximum(np.abs(d32) * 2.0 ** -21, bound[:, None] * 2.0 ** -39)
    assert np.all(np.abs(H.a.astype(np.float64) - d32) <= tol)
    err = np.abs(H.a.astype(np.float64) - exact).max(1)
    assert np.all(err[ordinary] <= 2e-6 * np.abs(exact).max(1)[ordinary])


def test_engine_edge_cases_on_fake_ops():
    """Constant voxels (zero variance), duplicated voxels and a rank-deficient design."""
    rng = np.random.default_rng(3)
    N, p, V = 240, 12, 40
    X = rng.standard_normal((N, p)).astype(np.float32)
    X[:, 5] = X[:, 4]  # duplicate column -> rank deficient
    Y = (X @ rng.standard_normal((p, V)) * 0.3 + rng.standard_normal((N, V))).astype(np.float32)
    Y[:, 7] = 0.0
    Y[:, 8] = 2.5
    Y[:, 9] = Y[:, 10]
    random.seed(1)
    model = NestedCVModel("ridge_regression", ops=FakeOps())
    kw = dict(n_outer_folds=3, n_inner_folds=3, chunk_length=10, alphas=np.logspace(-1, 3, 6))
    m, w, a = model.fit_predict(X, Y, **kw)
    r = np.asarray(m["correlations"])
    assert np.isfinite(r).all() and np.isfinite(w).all()
    assert r[7] == 0.0 and r[8] == 0.0  # pearsonr -> NaN -> 0.0 (nested_cv.py:435)
    assert m["p_values"][7] == 1.0 and m["p_values"][8] == 1.0  # all folds 1.0 -> Fisher shortcut 1.0
    assert r[9] == r[10] and a[9] == a[10]
    assert not m["significant_mask"][7]
    # every alpha that differs from the oracle's is a proven near-tie; r, weights, masks on ALL voxels (parity.py)
    info = prove_fit_parity(model.last_fold_results, m, w, X, Y, 1, w_tol=2e-4, **kw)
    assert info["disagreeing_alphas"] <= 0.2 * info["voxel_folds"]


# ------------------------------------------------------------------------------------------ stand-alone ridge kernels
@pytest.mark.parametrize("name", ["tall", "dupcol", "wide"])
def test_ridge_functions_on_fake_ops_match_reference_golden(name):
    g, e = load_golden("ridge_kernels.npz"), load_golden("ridge_extra.npz")
    alphas = g["alphas"].tolist()
    X, Y, n = g[f"{name}__X"], g[f"{name}__Y"], int(g[f"{name}__n_train"])
    ops = FakeOps()
    nonconst = Y[n:].std(0) > 0
    for normalpha in (True, False):
        tol = 2e-4 if (name != "tall" and not normalpha) else 3e-5
        for use_corr in (True, False):
            out = L.ridge_corr(X[:n], X[n:], Y[:n], Y[n:], alphas, singcutoff=1e-10, use_corr=use_corr,
                               normalpha=normalpha, ops=ops)
            ref = g[f"{name}_n{int(normalpha)}_c{int(use_corr)}__corr"]
            assert out.shape == ref.shape and out.dtype == np.float32
            if not use_corr:
                out, ref = np.sign(out) * out ** 2, np.sign(ref) * ref ** 2
            np.testing.assert_allclose(out[:, nonconst], ref[:, nonconst], rtol=0, atol=tol)
            va = g[f"{name}_n{int(normalpha)}__valphas"]
            cp = L.ridge_corr_pred(X[:n], X[n:], Y[:n], Y[n:], va, singcutoff=1e-10, use_corr=use_corr,
                                   normalpha=normalpha, ops=ops)
            refp = e[f"{name}_n{int(normalpha)}_c{int(use_corr)}__corrpred"]
            ok = np.isfinite(refp) & nonconst
            if not use_corr:
                cp, refp = np.sign(cp) * cp ** 2, np.sign(refp) * refp ** 2
            np.testing.assert_allclose(cp[ok], refp[ok], rtol=0, atol=tol)
        w = L.ridge(X[:n], Y[:n], g[f"{name}_n{int(normalpha)}__valphas"], singcutoff=1e-10, normalpha=normalpha, ops=ops)
        ref = g[f"{name}_n{int(normalpha)}__wt"]
        assert w.shape == ref.shape and w.dtype == np.float32
        assert np.abs(w - ref).max() <= (5e-3 if (name != "tall" and not normalpha) else 1e-4) * np.abs(ref).max()
    z = L.zs(e["zs__in64"], ops=ops)
    assert z.dtype == np.float64
    np.testing.assert_allclose(z, e["zs__out64"], atol=2e-6)
    np.testing.assert_allclose(L.zs(e["zs__in32"], ops=ops), e["zs__out32"], atol=2e-6)
    # scalar alpha and torch tensors in -> torch tensor out
    import torch

    w1 = L.ridge_torch(torch.from_numpy(X[:n]), torch.from_numpy(Y[:n]), 10.0, ops=ops)
    assert torch.is_tensor(w1) and tuple(w1.shape) == (X.shape[1], Y.shape[1])
    np.testing.assert_allclose(w1.numpy(), O.ridge_weights(X[:n], Y[:n], 10.0), atol=1e-4 * np.abs(w1.numpy()).max())


# ------------------------------------------------------------------------------------------ the other downsamplers
def test_other_downsamplers_on_fake_ops_match_reference_golden():
    from test_oracle_golden import _run_extra_downsampler

    ds = L.Downsampler(ops=FakeOps())
    g = load_golden("downsample_extra.npz")
    names = _cases(g, "__out")
    assert len(names) == 12
    for name in names:
        out, ref = _run_extra_downsampler(lambda m, d, t, tr, kw: ds.downsample(d, t, tr, method=m, **kw), g, name)
        assert out.shape == ref.shape and out.dtype == np.float64, name
        np.testing.assert_allclose(out, ref, rtol=1e-9, atol=1e-11, err_msg=name)
        if not name.startswith(("sinc", "gabor")):
            np.testing.assert_array_equal(out, ref, err_msg=name)  # membership reductions are bit-exact
    for method, msg in [("average", "average"), ("sum", "sum"), ("last", "last point"), ("legacy_sum", "Legacy")]:
        with pytest.raises(ValueError, match=f"split_indices must be provided for {msg} downsampling"):
            ds.downsample(np.zeros((4, 2)), np.arange(4.0), np.arange(2.0), method=method, split_indices=None)
    with pytest.raises(ValueError, match="Required parameter 'freqs' missing"):
        ds.downsample(np.zeros((4, 2)), np.arange(4.0), np.arange(2.0), method="gabor", sigma=1.0)


@pytest.mark.parametrize("N,p,label", [(150, 260, "dual everywhere"), (130, 100, "outer primal, inner dual"),
                                       (300, 40, "primal everywhere"), (232, 260, "dual everywhere, ragged outer folds")])
def test_dual_form_matches_oracle_on_fake_ops(N, p, label):
    """Folds with fewer training rows than features are solved through the n x n kernel matrix."""
    from litcoder_core_b200 import engine as E

    rng = np.random.default_rng(N + p)
    V = 30
    X = rng.standard_normal((N, p)).astype(np.float32)
    Y = (X[:, :20] @ rng.standard_normal((20, V)) * 0.5 + rng.standard_normal((N, V))).astype(np.float32)
    kw = dict(n_outer_folds=5, n_inner_folds=4, chunk_length=5, alphas=np.logspace(-1, 3, 6))
    random.seed(2)
    ops = FakeOps()
    model = NestedCVModel("ridge_regression", ops=ops)
    m, w, a = model.fit_predict(X, Y, **kw)
    n_o = (N // 5 // 5) * 5 * 4
    # default route, primal or dual (kernel-matrix) form: batched direct solves inside, grouped direct fit outside, no
    # decomposition anywhere; the systems have min(n, p) unknowns, as the reference's thin SVD has components
    assert n_o < p or not label.startswith("dual everywhere")
    assert not ops.eig_sizes and ops.outer_direct == 5 and ops.direct_solved > 0
    info = prove_fit_parity(model.last_fold_results, m, w, X, Y, 2, w_tol=2e-4, **kw)
    assert info["disagreeing_alphas"] <= 0.15 * info["voxel_folds"], (label, info["disagreeing_alphas"])
    # and the dual path agrees with the primal path on the same problem
    orig = E.RidgeConfig.__init__

    def patched(self, *a_, **k_):
        orig(self, *a_, **k_)
        self.allow_dual = False

    E.RidgeConfig.__init__ = patched
    try:
        random.seed(2)
        m2, w2, a2 = NestedCVModel("ridge_regression", ops=FakeOps()).fit_predict(X, Y, **kw)
    finally:
        E.RidgeConfig.__init__ = orig
    same2 = np.isclose(a, a2)
    assert same2.mean() > 0.85
    np.testing.assert_allclose(np.asarray(m["correlations"])[same2], np.asarray(m2["correlations"])[same2], atol=5e-5)


@pytest.mark.parametrize("name", ["tt_default", "cv_default"])
def test_inner_solvers_agree_on_fake_ops(name):
    """The GEMM-only inner solver and the eigendecomposition route compute the same validation scores."""
    ops_c, ops_e = FakeOps(), FakeOps()
    g = load_golden("fit_predict.npz")
    X, Y, alphas = g["X"], g["Y"], g["alphas"].tolist()
    kw = dict(folding_type="chunked", n_outer_folds=4, n_inner_folds=3, chunk_length=10, alphas=alphas)
    if name.startswith("tt"):
        kw.update(X_test=X[400:], y_test=Y[400:])
    random.seed(7)
    mc, wc, ac = NestedCVModel("ridge_regression", ops=ops_c).fit_predict(X[:400], Y[:400], inner_solver="chebyshev", **kw)
    random.seed(7)
    me, we, ae = NestedCVModel("ridge_regression", ops=ops_e).fit_predict(X[:400], Y[:400], inner_solver="eig", **kw)
    n_outer = 1 if name.startswith("tt") else 4
    # GEMM-only route: no decomposition at all (batched direct solves inside, grouped direct fit per outer fold)
    assert getattr(ops_c, "solver_calls", 0) == 3 * n_outer and ops_c.eig_calls == 0 and ops_c.outer_direct == n_outer
    assert getattr(ops_e, "solver_calls", 0) == 0 and ops_e.eig_calls == 4 * n_outer
    same = np.isclose(ac, ae)
    assert same.mean() > 0.95
    np.testing.assert_allclose(np.asarray(mc["correlations"])[same], np.asarray(me["correlations"])[same], atol=1e-5)
    with pytest.raises(ValueError, match="Unknown inner_solver"):
        NestedCVModel("ridge_regression", ops=FakeOps()).fit_predict(X[:400], Y[:400], inner_solver="qr", **kw)


def test_series_moments_match_full_stack_on_fake_ops(monkeypatch):
    """GEMM-only folds: the compact stack (four series terms + 14 per-voxel sums, combined per alpha at finalize)
    scores the alphas like the one-block-per-alpha stack; both follow the golden reference run."""
    g = load_golden("fit_predict.npz")
    X, Y = g["X"], g["Y"]
    kw = dict(folding_type="chunked", n_outer_folds=4, n_inner_folds=3, chunk_length=10, alphas=np.logspace(-1, 8, 20))
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("LIT_SERIES_MOMENTS", flag)
        for prec in ("tf32x3", "f16x3"):
            ops = FakeOps()
            seen = []
            orig = ops.gemm_corr
            ops.gemm_corr = lambda *a, _o=orig, _s=seen, **k: (_s.append(type(a[1]).__name__), _o(*a, **k))[1]
            random.seed(7)
            m, w, a = NestedCVModel("ridge_regression", ops=ops).fit_predict(X[:400], Y[:400], inner_solver="chebyshev",
                                                                             corr_precision=prec, **kw)
            out[flag, prec] = (np.asarray(m["correlations"]), np.asarray(a))
            assert ("FakeSeriesStack" in seen) == (flag == "1")
    for prec in ("tf32x3", "f16x3"):
        same = out["0", prec][1] == out["1", prec][1]
        assert same.mean() > 0.97
        np.testing.assert_allclose(out["0", prec][0][same], out["1", prec][0][same], atol=1e-5)


def test_leave_block_out_matches_chebyshev_route_on_fake_ops(monkeypatch):
    """GEMM-only folds whose validation rows are the rows removed from the outer training set: the small alphas
    solved through the leave-block-out identity on the outer eigendecomposition give the scores of the direct
    p x p solves; folds that are not leave-block-out folds (timeseries layout) keep the direct route."""
    g = load_golden("fit_predict.npz")
    X, Y, alphas = g["X"], g["Y"], g["alphas"].tolist()
    kw = dict(folding_type="chunked", n_outer_folds=4, n_inner_folds=3, chunk_length=10, alphas=alphas)
    monkeypatch.setenv("LIT_DIRECT_SOLVER", "0")  # the round-1 iterative routes (kept as options)
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("LIT_LEAVE_BLOCK_OUT", flag)
        ops = FakeOps()
        random.seed(7)
        m, w, a = NestedCVModel("ridge_regression", ops=ops).fit_predict(X[:400], Y[:400], inner_solver="chebyshev", **kw)
        out[flag] = (np.asarray(m["correlations"]), w, np.asarray(a))
        n_small = 3  # alphas below sqrt(60): solved, not served by the Neumann series
        assert getattr(ops, "lbo_solved", 0) == (12 * n_small if flag == "1" else 0)
        assert getattr(ops, "lbo_prepared", 0) == (12 if flag == "1" else 0)
        assert ops.eig_calls == 4 and ops.solver_calls == 12
    same = out["0"][2] == out["1"][2]
    assert same.mean() > 0.97
    np.testing.assert_allclose(out["0"][0][same], out["1"][0][same], atol=1e-5)
    ref_va, ref_r = g["cv_default__best_alphas"], g["cv_default__m__correlations"]
    same = np.isclose(out["1"][2], ref_va, rtol=1e-6)
    assert same.mean() >= 0.9
    assert np.abs(out["1"][0][same] - ref_r[same]).max() < 3e-5
    # invalid spectral bounds (a Lanczos value outside [0, 1)) send the fold back to the direct route
    ops = FakeOps()
    ops.lambda_max_batched = lambda mats, steps=96: (
        np.full(len(mats), np.nan) if steps == 48 else FakeOps.lambda_max_batched(ops, mats, steps))
    random.seed(7)
    m, w, a = NestedCVModel("ridge_regression", ops=ops).fit_predict(X[:400], Y[:400], inner_solver="chebyshev", **kw)
    assert getattr(ops, "lbo_solved", 0) == 0 and ops.solver_calls == 12
    np.testing.assert_array_equal(np.asarray(a), out["0"][2])
    # a fold whose validation rows are only PART of the rows removed from the outer training set is downdated but
    # keeps the direct solve; validating on all of them makes it a leave-block-out fold
    from fake_ops import FMat
    from litcoder_core_b200 import engine as E
    from litcoder_core_b200.engine import FoldPlan, RidgeCVEngine

    tr_o, te = np.arange(300), np.arange(300, 400)
    for val, n_lbo in ((np.arange(250, 300), 0), (np.arange(200, 300), 3)):
        ops = FakeOps()
        cfg = E.RidgeConfig(alphas=alphas, inner_solver="chebyshev", n_outer_folds=1, direct_solver=False)
        RidgeCVEngine(ops).fit_shard(FMat(X[:400]), FMat(Y[:400]), [FoldPlan(tr_o, te, [(np.arange(200), val)])], cfg)
        assert getattr(ops, "lbo_solved", 0) == n_lbo and ops.solver_calls == 1


GRID20 = {"tt_grid20": dict(train_test=True), "cv_grid20": dict(train_test=False),
          "cv_grid20_single": dict(train_test=False, single_alpha=True)}


@pytest.mark.parametrize("name", sorted(GRID20))
@pytest.mark.parametrize("solver", ["auto", "eig"])
def test_engine_on_the_baseline_alpha_grid_matches_reference_golden(name, solver):
    """The reference's own output on np.logspace(-1, 8, 20) (fit_predict_grid20.npz): with the default solver the
    inner folds run the compact stack (16 series alphas) and the batched direct solves (4 small alphas)."""
    g, X, Y, test, common = _golden_args(name)
    tt = bool(test)
    ops = FakeOps()
    random.seed(7)
    np.random.seed(7)
    model = NestedCVModel("ridge_regression", ops=ops)
    m, w, va = model.fit_predict(X, Y, inner_solver=solver, **test, **common)
    # the 4 small alphas of every inner fold go through the batched direct (Cholesky) solver, all folds at once
    assert getattr(ops, "direct_solved", 0) == (4 * (3 if tt else 12) if solver == "auto" else 0)
    assert getattr(ops, "lbo_solved", 0) == 0
    # ... and every downdated cross product leaves its GEMM as the fp16 pair the prediction GEMM reads
    # (a fold whose removed rows outnumber its training rows forms the product directly instead)
    assert (getattr(ops, "pair_out_gemms", 0) >= (2 if tt else 8)) == (solver == "auto")
    info = check_against_reference_golden(name, model.last_fold_results, m, w, va)
    assert info["disagreeing_alphas"] <= 0.05 * info["voxel_folds"]


def _structure_golden():
    g = load_golden("structure.npz")
    stories = [str(x) for x in g["stories"]]
    feats = {s: g[f"feat__{s}"] for s in stories}
    brain = {s: g[f"brain__{s}"] for s in stories}

    def cfg(prefix):
        out = {}
        for k in g.files:
            if k.startswith(prefix):
                v = float(g[k])
                out[k[len(prefix):]] = None if np.isnan(v) else int(v)
        return out

    return g, stories, feats, brain, [int(d) for d in g["delays"]], cfg("cfg_tt__"), cfg("cfg_cc__")


def test_structure_train_test_split_matches_reference_trainer():
    """create_train_test_split (FIR fused) on the NumPy stand-in against the unmodified trainer's outputs."""
    from litcoder_core_b200.structure import create_train_test_split

    g, stories, feats, brain, delays, cfg_tt, _ = _structure_golden()
    out = create_train_test_split(feats, brain, cfg_tt, fir_delays=delays, ops=FakeOps())
    for k in ("Rstim", "Rresp", "Pstim", "Presp"):
        want = g[f"tt__{k}"]
        assert out[k].dtype == np.float32 and out[k].shape == want.shape
        np.testing.assert_allclose(out[k], want.astype(np.float32), rtol=2e-6, atol=2e-6)
    # already-delayed input (fir_delays=None) gives the same matrices
    delayed = {s: O.fir_make_delayed(feats[s], delays) for s in stories}
    out2 = create_train_test_split(delayed, brain, cfg_tt, ops=FakeOps())
    np.testing.assert_array_equal(out2["Rstim"], out["Rstim"])
    np.testing.assert_array_equal(out2["Pstim"], out["Pstim"])
    # the NaN feature of story s2 is scrubbed: its delayed columns are all zero in that story's rows
    assert not np.isnan(out["Rstim"]).any()
    with pytest.raises(ValueError, match="at least one training story"):
        create_train_test_split({"a": feats["s0"]}, {"a": brain["s0"]}, cfg_tt, ops=FakeOps())


def test_structure_concatenated_matches_reference_trainer():
    from litcoder_core_b200.structure import create_concatenated_data

    g, stories, feats, brain, delays, _, cfg_cc = _structure_golden()
    brain_cc = {s: np.vstack([brain[s], brain[s][:15]]) for s in stories}
    out = create_concatenated_data(feats, brain_cc, stories, cfg_cc, fir_delays=delays, ops=FakeOps())
    with np.errstate(invalid="ignore"):
        np.testing.assert_array_equal(out["X"], g["cc__X"].astype(np.float32))  # copies: exact (NaN kept)
        np.testing.assert_array_equal(out["Y"], g["cc__Y"].astype(np.float32))
    # a window that starts inside the second story and ends inside the third
    cfg = {"features_start": 80, "features_end": 200, "targets_start": 80, "targets_end": 200}
    out = create_concatenated_data(feats, brain_cc, stories, cfg, fir_delays=delays, ops=FakeOps())
    want = np.concatenate([O.fir_make_delayed(feats[s], delays) for s in stories])[80:200]
    np.testing.assert_array_equal(out["X"], want.astype(np.float32))
    np.testing.assert_array_equal(out["Y"], np.concatenate([brain_cc[s] for s in stories])[80:200].astype(np.float32))
