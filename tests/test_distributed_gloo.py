"""World-size-2 test of the multi-rank path on CPU (gloo): voxel sharding, eigenproblem ownership +
broadcast, the single-alpha all-reduce and the final all-gathers must reproduce the single-process
result.  The arithmetic runs on the NumPy stand-in of the C ABI (tests/fake_ops.py)."""
import os
import random
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem(wide=False):
    rng = np.random.default_rng(5)
    N, p, V = 300, 10, 300  # 300 voxels -> blocks of 256 + 44 (shards start on 128-voxel tile boundaries)
    if wide:
        p = 330  # more features than training rows: dual form (n x n kernel matrix, contraction over features)
    X = rng.standard_normal((N, p)).astype(np.float32)
    Y = (X @ rng.standard_normal((p, V)) * 0.4 + rng.standard_normal((N, V))).astype(np.float32)
    return X, Y


def _worker(rank, world, port, single_alpha, out_dir, row_shard_gram=False, wide=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch.distributed as dist

    from fake_ops import FakeOps
    from litcoder_core_b200.nested_cv import NestedCVModel, TorchDistComm

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        X, Y = _problem(wide)
        random.seed(11)
        ops = FakeOps()
        model = NestedCVModel("ridge_regression", ops=ops, comm=TorchDistComm())
        m, w, a = model.fit_predict(X, Y, n_outer_folds=3, n_inner_folds=3, chunk_length=10,
                                    alphas=np.logspace(-1, 3, 6), single_alpha=single_alpha,
                                    row_shard_gram=row_shard_gram)
        assert model.last_stats["world"] == world and model.last_stats["voxels_this_rank"] in (256, 44)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), r=np.asarray(m["correlations"]), w=w, a=a,
                 p=np.asarray(m["p_values"]), sig=np.asarray(m["significant_mask"]), n_sig=m["n_significant"],
                 lbo_solved=getattr(ops, "direct_solved", 0), solver_calls=getattr(ops, "solver_calls", 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("single_alpha", [False, True])
def test_two_ranks_match_single_process(tmp_path, single_alpha):
    import torch.multiprocessing as mp

    sys.path.insert(0, HERE)
    from fake_ops import FakeOps
    from litcoder_core_b200.nested_cv import NestedCVModel

    X, Y = _problem()
    random.seed(11)
    m, w, a = NestedCVModel("ridge_regression", ops=FakeOps()).fit_predict(
        X, Y, n_outer_folds=3, n_inner_folds=3, chunk_length=10, alphas=np.logspace(-1, 3, 6), single_alpha=single_alpha)
    mp.spawn(_worker, args=(2, _free_port(), single_alpha, str(tmp_path)), nprocs=2, join=True)
    # the 9 inner folds are solved once each, by their owners, with the batched direct solver (3 small alphas each)
    per_rank = [np.load(tmp_path / f"rank{rank}.npz") for rank in range(2)]
    assert sum(int(g["solver_calls"]) for g in per_rank) == 9 and all(int(g["solver_calls"]) >= 3 for g in per_rank)
    assert sum(int(g["lbo_solved"]) for g in per_rank) == 27
    for rank in range(2):
        g = per_rank[rank]
        np.testing.assert_array_equal(g["a"], a)
        np.testing.assert_allclose(g["r"], np.asarray(m["correlations"]), atol=1e-6)
        np.testing.assert_allclose(g["w"], w, atol=1e-6 * np.abs(w).max())
        np.testing.assert_allclose(g["p"], np.asarray(m["p_values"]), rtol=1e-4, atol=1e-12)
        assert int(g["n_sig"]) == m["n_significant"]
        np.testing.assert_array_equal(g["sig"], np.asarray(m["significant_mask"]))


@pytest.mark.parametrize("wide", [False, True])
def test_row_sharded_gram_matches_single_process(tmp_path, wide):
    """row_shard_gram=True: each rank forms the outer Gram over its half of the TRs, one all-reduce; the fit then
    agrees with the single-process one to fp32 rounding of that sum (not bit for bit)."""
    import torch.multiprocessing as mp

    sys.path.insert(0, HERE)
    from fake_ops import FakeOps
    from litcoder_core_b200.nested_cv import NestedCVModel

    X, Y = _problem(wide)
    random.seed(11)
    m, w, a = NestedCVModel("ridge_regression", ops=FakeOps()).fit_predict(
        X, Y, n_outer_folds=3, n_inner_folds=3, chunk_length=10, alphas=np.logspace(-1, 3, 6))
    mp.spawn(_worker, args=(2, _free_port(), False, str(tmp_path), True, wide), nprocs=2, join=True)
    for rank in range(2):
        g = np.load(tmp_path / f"rank{rank}.npz")
        same = g["a"] == a
        assert same.mean() > 0.98
        np.testing.assert_allclose(g["r"][same], np.asarray(m["correlations"])[same], atol=1e-5)
        np.testing.assert_allclose(g["w"][:, same], w[:, same], atol=1e-5 * np.abs(w).max())
    np.testing.assert_array_equal(np.load(tmp_path / "rank0.npz")["a"], np.load(tmp_path / "rank1.npz")["a"])


def _problem_5x5():
    rng = np.random.default_rng(5)
    N, p, V = 500, 12, 1100  # 1,100 voxels over 4 ranks: blocks of 384, 384, 332 and 0 (tile-aligned starts)
    X = rng.standard_normal((N, p)).astype(np.float32)
    Y = (X @ rng.standard_normal((p, V)) * 0.4 + rng.standard_normal((N, V))).astype(np.float32)
    return X, Y


_KW_5X5 = dict(n_outer_folds=5, n_inner_folds=5, chunk_length=10, alphas=np.logspace(-1, 8, 20))


def _worker_5x5(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch.distributed as dist

    from fake_ops import FakeOps
    from litcoder_core_b200.nested_cv import NestedCVModel, TorchDistComm

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        X, Y = _problem_5x5()
        random.seed(11)
        ops = FakeOps()
        m, w, a = NestedCVModel("ridge_regression", ops=ops, comm=TorchDistComm()).fit_predict(X, Y, **_KW_5X5)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), r=np.asarray(m["correlations"]), w=w, a=a,
                 n_sig=m["n_significant"], lbo=getattr(ops, "direct_solved", 0), eig=ops.eig_calls,
                 solves=getattr(ops, "solver_calls", 0))
    finally:
        dist.destroy_process_group()


def test_four_ranks_bench_layout(tmp_path):
    """The BASELINE layout (5 x 5 chunked folds, 20 alphas: 4 solved + 16 series alphas per inner fold) on 4 ranks:
    25 batched direct solves dealt out evenly, 5 grouped direct outer fits on every rank, compact stacks broadcast, same result."""
    import torch.multiprocessing as mp

    sys.path.insert(0, HERE)
    from fake_ops import FakeOps
    from litcoder_core_b200.nested_cv import NestedCVModel

    X, Y = _problem_5x5()
    random.seed(11)
    m, w, a = NestedCVModel("ridge_regression", ops=FakeOps()).fit_predict(X, Y, **_KW_5X5)
    mp.spawn(_worker_5x5, args=(4, _free_port(), str(tmp_path)), nprocs=4, join=True)
    per = [np.load(tmp_path / f"rank{r}.npz") for r in range(4)]
    assert sorted(int(g["eig"]) for g in per) == [0, 0, 0, 0]  # no decomposition is left on the default route
    assert sorted(int(g["solves"]) for g in per) == [6, 6, 6, 7]
    assert sum(int(g["lbo"]) for g in per) == 25 * 4
    for g in per:
        np.testing.assert_array_equal(g["a"], a)
        np.testing.assert_allclose(g["r"], np.asarray(m["correlations"]), atol=1e-6)
        np.testing.assert_allclose(g["w"], w, atol=1e-6 * np.abs(w).max())
        assert int(g["n_sig"]) == m["n_significant"]
