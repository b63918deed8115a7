import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(),'tests'))
import numpy as np
from conftest import load_golden
from litcoder_core_b200.device import default_ops
from litcoder_core_b200.engine import FoldPlan, RidgeConfig, RidgeCVEngine
ops = default_ops()
g = load_golden("ridge_kernels.npz")
alphas = g["alphas"].tolist()
name="tall"
X, Y, n = g[f"{name}__X"], g[f"{name}__Y"], int(g[f"{name}__n_train"])
eng = RidgeCVEngine(ops)
Xd, Yd = ops.upload_matrix(X), ops.upload_matrix(Y)
tr, va = np.arange(n), np.arange(n, X.shape[0])
plan = FoldPlan(tr, va, [(tr, va)])
for use_corr in (True, False):
    cfg = RidgeConfig(alphas=alphas, normalpha=True, use_corr=use_corr, singcutoff=1e-10)
    outer, inners = eng._design_side(Xd, plan, cfg)
    corr, _ = eng._inner_scores(Xd, Yd, plan, outer, inners, ops.upload_vector(np.asarray(alphas), "f64"), len(alphas), cfg)
    eng._eig_ready(outer)
    out = ops.download_matrix(corr)
    ref = g[f"{name}_n1_c{int(use_corr)}__corr"]
    np.set_printoptions(linewidth=200, precision=5)
    print("use_corr", use_corr)
    print("ours last 3 voxels:\n", out[:, -3:])
    print("ref  last 3 voxels:\n", ref[:, -3:])
    fin = np.isfinite(out) & (np.abs(ref) < 1e30) & (np.abs(out) < 1e30)
    print("max abs diff finite:", np.abs(out-ref)[fin].max())
ops.check_eig()
