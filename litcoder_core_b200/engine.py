"""Nested-CV ridge engine: the reference's fit_predict algebra re-planned for B200.

Reference formulation (encoding/models/ridge_regression.py, nested_cv.py): a thin SVD of every
training design, predictions materialised per alpha (n_val x V fp32 each), several elementwise
passes for the z-scored correlation, and per-voxel SciPy loops for the test statistics.

This engine computes the same quantities from the Gram side, with every matrix kept in the layout
the tensor cores want (K-major operands, voxels on the accumulator rows):

    G   = X_tr^T X_tr                    (p x p)     tcgen05 3xTF32 GEMM
    G   = V diag(lam) V^T                            cuSOLVER syevd  (lam = S^2, rows of Vt = Vh)
    C^T = Y_tr^T X_tr                    (V x p)     GEMM, K = training TRs
    Z^T = C^T V                          (V x k)     GEMM           (= (S U^T Y)^T)
    L   = P_val V, column-centred        (n_v x k)   GEMM + streaming kernel
    pred_a = L diag(1/(lam + a'^2)) Z    never materialised: ONE GEMM over the alpha-stacked L with the
                                         per-voxel reduction (sum pred*zY, sum pred^2) in its epilogue
    W^T = (Z^T * 1/(lam + a_v'^2)) V^T   (V x p)     per-voxel shrinkage + GEMM
    r   = corr(P_test W, Y_test)                     same fused GEMM, then p-values / BH / Fisher kernels

With X = U S V^T:  U^T Y = S^-1 V^T X^T Y, so (PVh * D) @ UR with D = S/(S^2+a^2) equals
P V diag(1/(lam+a^2)) V^T X^T Y -- identical algebra, no SVD of the tall matrix, no n_v x V buffer.

Two further savings inside one outer fold (both exact up to fp32 rounding of a sum):

  * the Gram and the cross product of an inner fold are DOWNDATES of the outer fold's:
        G_i = G_o - X_R^T X_R,   C_i^T = C_o^T - Y_R^T X_R,   R = outer-train rows not in inner-train
    (|R| ~ n/5), so Y is streamed once per outer fold plus once per removed block instead of once
    per inner fold, and the K extent of those GEMMs drops by 4x;
  * all eigendecompositions of an outer fold depend only on X, so they are queued up front on a
    side stream and overlap the response-side GEMMs on the main stream.

Inner folds are by default solved WITHOUT an eigendecomposition of their own: the large
alphas through a 4-term Neumann series whose terms share four stacked row blocks (compact alpha stack), the small
ones through the leave-block-out identity on the outer fold's decomposition, or Chebyshev iteration on the p x p Gram.

The engine is written against the small `ops` interface of device.DeviceOps so that its control
flow can be unit-tested on CPU with a NumPy stand-in (tests/fake_ops.py).
"""
from __future__ import annotations

import logging
from dataclasses import dataclass, field, replace
from typing import List, Optional, Sequence, Tuple

import numpy as np

logger = logging.getLogger(__name__)

EPS = 1e-8  # ridge_utils.z_score / DataNormalizer eps


class SolverAccuracyError(RuntimeError):
    """A GEMM-only inner solve failed its a-posteriori check (DeviceOps.check_solver)."""


@dataclass
class RidgeConfig:
    alphas: Sequence[float]
    alpha_fdr: float = 0.05
    single_alpha: bool = False
    normalpha: bool = True
    use_corr: bool = True
    normalize_features: bool = False
    normalize_targets: bool = False
    singcutoff: float = 1e-10
    n_outer_folds: int = 5
    p_round_f32: bool = True  # SciPy >= 1.14 keeps float32 inputs' dtype for p-values
    allow_dual: bool = True  # folds with fewer training rows than features use the n x n kernel matrix
    downdate: bool = True  # inner-fold Gram / cross product by subtraction from the outer fold's
    overlap_eig: bool = True  # eigendecompositions on a side stream
    # inner-fold solver: "eig" (syevd of every inner Gram), "chebyshev" (GEMM-only, no inner eigendecomposition),
    # "auto" = chebyshev for primal folds when alphas are normalised and the smallest is >= 0.05 (kappa <= 401)
    inner_solver: str = "auto"
    # operand format of the fused inner-CV prediction + correlation GEMM: "tf32x3" or "f16x3" (scaled fp16 split
    # pairs, lit_split_f16: same product accuracy, twice the tensor-core rate)
    corr_precision: str = "f16x3"
    # operand format of the other voxel-side GEMMs (cross products Y^T X and their downdates, rotations of the
    # coefficients, weights): same choice, through lit_gemm_f16x3_nt
    voxel_gemm_precision: str = "f16x3"
    # GEMM-only folds: the alphas served by the Neumann series (a^2 >= 60 lambda_max; 16 of the 20 BASELINE alphas)
    # share FOUR stacked row blocks P_c G^q instead of one block each; their scores are 4-term combinations of 14
    # per-voxel sums taken in the GEMM epilogue (DeviceOps.assemble_series_stack).  2.5x fewer prediction flops.
    series_moments: bool = True
    # GEMM-only folds whose validation rows are exactly the rows removed from the outer training set (the nested
    # chunked / k-fold layouts of the reference): solve the small alphas through the leave-block-out identity on
    # the outer fold's eigendecomposition (DeviceOps.lbo_prepare) instead of Chebyshev iteration on the p x p Gram
    leave_block_out: bool = True
    # several ranks: each forms the outer Gram / kernel matrix over its slice of the contraction axis, one all-reduce
    row_shard_gram: bool = False
    # GEMM-only folds: solve the small alphas with the batched blocked-Cholesky solver (DeviceOps.solve_blocks_many:
    # all systems of this rank's folds together, a few batched launches per 128-column panel) instead of Chebyshev
    # iteration; supersedes leave_block_out (no dependence on the outer fold's eigendecomposition)
    direct_solver: bool = True
    # primal outer folds under the same conditions: weights without an eigendecomposition -- (G_o + a^2 I)^-1 per grid
    # alpha (batched Cholesky / Neumann polynomials, DeviceOps.outer_inverses) and ONE grouped GEMM over the voxels
    # sorted by their selected alpha (DeviceOps.gemm_grouped); no syevd is left on the default path
    direct_outer: bool = True
    # fp16-pair GEMMs: operands that a kernel of ours produces (gathered response rows, downdated cross products) are
    # written as scaled fp16 pairs by that kernel, with scales from a-priori bounds (DeviceOps.f16_bound_scales),
    # instead of an fp32 plane that lit_split_f16 then reads twice
    producer_pairs: bool = True
    # keep the fold-mean inner score curves (n_alphas x V_r per outer fold) for the caller (tests: near-tie proofs)
    record_scores: bool = False


@dataclass
class FoldPlan:
    """Host-side description of one outer fold; all indices are absolute rows of the source arrays."""
    train_rows: np.ndarray
    test_rows: np.ndarray
    inner: List[Tuple[np.ndarray, np.ndarray]]  # (train rows, validation rows)


@dataclass
class ShardResult:
    """Per-voxel-shard outputs of the fit (device handles)."""
    r: list = field(default_factory=list)  # per outer fold: (V_r,) f32
    p: list = field(default_factory=list)  # per outer fold: (V_r,) f64
    alpha: list = field(default_factory=list)  # per outer fold: (V_r,) f32
    Wt_mean: object = None  # Mat (V_r x p): fold-mean of the weights, voxel-major
    scores: list = field(default_factory=list)  # per outer fold: (n_alphas x V_r) host array (cfg.record_scores)
    n_test: list = field(default_factory=list)


class SingleProcess:
    """Degenerate communicator (one voxel shard)."""
    rank, world = 0, 1

    def all_reduce_sum(self, arr: np.ndarray) -> np.ndarray:
        return arr

    def all_gather_concat(self, arr: np.ndarray, counts=None) -> np.ndarray:
        return arr

    def broadcast_inplace(self, buffers, src: int) -> None:
        pass

    def all_reduce_sum_inplace(self, buffers) -> None:
        pass

    def broadcast_async(self, buffers, src: int) -> list:
        return []

    @staticmethod
    def wait_all(works) -> None:
        pass


def removed_rows(outer_rows: np.ndarray, inner_rows: np.ndarray) -> Optional[np.ndarray]:
    """Rows of the outer training set that are absent from the inner one, or None when the inner set
    is not a duplicate-free subset of the outer set (then the downdate identity does not hold)."""
    outer_rows = np.asarray(outer_rows, dtype=np.int64)
    inner_rows = np.asarray(inner_rows, dtype=np.int64)
    if len(outer_rows) == 0 or len(inner_rows) == 0 or min(outer_rows.min(), inner_rows.min()) < 0:
        return None
    # multiplicity tables instead of np.unique / np.isin (those cost ~6 ms per fold: 150 ms of host time per
    # 5 x 5 fit, all of it on the critical path of every rank)
    size = int(max(outer_rows.max(), inner_rows.max())) + 1
    in_outer = np.bincount(outer_rows, minlength=size)
    in_inner = np.bincount(inner_rows, minlength=size)
    if in_outer.max() > 1 or in_inner.max() > 1 or np.any(in_inner > in_outer):
        return None
    return outer_rows[in_inner[outer_rows] == 0]


class RidgeCVEngine:
    def __init__(self, ops, comm=None):
        self.ops = ops
        self.comm = comm if comm is not None else SingleProcess()
        self._eig_jobs = 0  # running count of inner-fold solves / eigenproblems; job j belongs to rank j % world
        self._outer_jobs = 0  # outer eigendecompositions are dealt out separately, from the last rank downwards

    # ------------------------------------------------------------------------------------------
    # row-index vectors of every fold, staged to the device in ONE asynchronous copy
    # ------------------------------------------------------------------------------------------
    def stage_plans(self, plans: List[FoldPlan], cfg: RidgeConfig, n_rows_total: Optional[int] = None):
        """Per plan: device int32 row-index vectors for the outer train / test rows and, per inner fold,
        the train and validation rows plus the removed rows R when the fold can be downdated.
        (A synchronous upload per fold would drain the stream ~150 times per fit.)"""
        arrays, staged = [], []
        for plan in plans:
            tr_o = np.asarray(plan.train_rows, dtype=np.int64)
            entry = {"train_rows": tr_o, "test_rows": np.asarray(plan.test_rows, dtype=np.int64), "inner": []}
            arrays += [entry["train_rows"], entry["test_rows"]]
            # rows of the WHOLE matrix that are not outer-training rows (test rows + dropped tail): with several outer
            # folds the outer cross product is a downdate of the all-rows product, computed once per fit
            Ro = None
            if cfg.downdate and n_rows_total is not None and len(plans) > 1:
                Ro = removed_rows(np.arange(n_rows_total, dtype=np.int64), tr_o)
                if Ro is not None and not (0 < len(Ro) <= len(tr_o) // 2):
                    Ro = None
            entry["Ro_rows"] = Ro
            if Ro is not None:
                arrays.append(Ro)
            for tr_i, va_i in plan.inner:
                tr_i, va_i = np.asarray(tr_i, dtype=np.int64), np.asarray(va_i, dtype=np.int64)
                R = removed_rows(tr_o, tr_i) if cfg.downdate else None
                if R is not None and not (0 < len(R) <= len(tr_i) // 2):
                    R = None  # the downdate only pays when it is much shorter than the direct product
                entry["inner"].append({"train_rows": tr_i, "val_rows": va_i, "R_rows": R})
                arrays += [tr_i, va_i] + ([R] if R is not None else [])
            staged.append(entry)
        dev = iter(self.ops.stage_indices(arrays))
        for entry in staged:
            entry["train"], entry["test"] = next(dev), next(dev)
            entry["Ro"] = next(dev) if entry["Ro_rows"] is not None else None
            for d in entry["inner"]:
                d["train"], d["val"] = next(dev), next(dev)
                d["R"] = next(dev) if d["R_rows"] is not None else None
        return staged

    # ------------------------------------------------------------------------------------------
    # design side: Grams and their eigendecompositions (depend on X only)
    # ------------------------------------------------------------------------------------------
    def _gram(self, A, cfg: RidgeConfig):
        """A A^T for a split pair A (rows x K).  With `row_shard_gram` and several ranks: this rank's slice of the
        contraction axis only (slice boundaries on multiples of 32 values), then ONE all-reduce of the partial sums
        (the row-sharded Gram of the wide designs)."""
        ops, comm = self.ops, self.comm
        if not (cfg.row_shard_gram and comm.world > 1):
            return ops.gemm(A, A)
        per = -(-(-(-A.cols // comm.world)) // 32) * 32
        k0, k1 = min(comm.rank * per, A.cols), min((comm.rank + 1) * per, A.cols)
        if k1 > k0:
            part = ops.col_view(A, k0, k1)
            G = ops.gemm(part, part)
        else:
            G = ops.zeros(A.rows, A.rows)
        comm.all_reduce_sum_inplace([ops.raw(G)])
        return G

    def _design_side(self, X, sp, cfg: RidgeConfig):
        """Gram + syevd for the outer training set and for every inner fold of one staged plan `sp`.

        A fold with at least as many training rows as features is solved in the PRIMAL form (p x p Gram
        X^T X, eigenvectors = right singular vectors); a fold with fewer rows than features in the DUAL form
        (n x n kernel matrix X X^T, eigenvectors = left singular vectors), which keeps the eigenproblem at
        min(n, p) -- the reference's thin SVD has exactly that many components (ridge_utils.py:52).

        Returns (outer, inners): dicts with `dual`, XtT / XRt (primal operands), G (holds the eigenvectors as
        rows after the eig), lam and the eig ticket."""
        ops, comm = self.ops, self.comm
        p = X.cols
        n_o = len(sp["train_rows"])
        outer = {"dual": bool(cfg.allow_dual and n_o < p), "owner": None}
        if outer["dual"]:
            XoR = ops.gather_rows(X, sp["train"], n_o, split=True)  # (n_o x p)
            G_o = None
            outer["G"] = self._gram(XoR, cfg)  # kernel matrix, n_o x n_o (needed on every rank only after the eig)
            del XoR
        else:
            XoT = ops.gather_rows_T_split(X, sp["train"], n_o)  # (p x n_o)
            G_o = self._gram(XoT, cfg)  # outer Gram, p x p
            outer["XtT"] = XoT
        inners = []
        for d in sp["inner"]:
            d = dict(d)
            n_i = len(d["train_rows"])
            d["owner"] = self._next_eig_owner()
            d["dual"] = bool(cfg.allow_dual and n_i < p)
            d["cheb"] = False
            mine = d["owner"] == comm.rank
            if d["dual"]:
                # kernel-matrix form: K = X_tr X_tr^T (n x n).  GEMM-only as well when the alphas allow it: the
                # solves and the series then run on K instead of the Gram (see _inner_scores)
                d["R"] = None
                d["cheb"] = bool(cfg.direct_solver and self._use_chebyshev(cfg))
                d["lbo"] = False
                if mine:
                    XiR = ops.gather_rows(X, d["train"], n_i, split=True)  # (n_i x p)
                    d["G"] = ops.gemm(XiR, XiR)
                    del XiR
                else:
                    d["G"] = ops.empty(n_i, n_i)
            else:
                d["cheb"] = self._use_chebyshev(cfg)
                need_G = mine  # the Gram is only needed by the rank that solves this fold
                d["lbo"] = bool(d["cheb"] and cfg.leave_block_out and not cfg.direct_solver and d["R"] is not None
                                and G_o is not None
                                and len(d["R_rows"]) == len(d["val_rows"])
                                and np.array_equal(np.sort(d["R_rows"]), np.sort(d["val_rows"])))
                if d["R"] is not None and G_o is not None:
                    XRt = ops.gather_rows_T_split(X, d["R"], len(d["R_rows"]))  # (p x |R|)
                    d.update(XRt=XRt, XtT=None,
                             G=ops.gemm(XRt, XRt, alpha=-1.0, Cin=G_o, beta=1.0) if need_G else ops.empty(p, p))
                else:
                    d["R"] = None
                    XtT = ops.gather_rows_T_split(X, d["train"], n_i)
                    d.update(XtT=XtT, XRt=None, G=ops.gemm(XtT, XtT) if need_G else ops.empty(p, p))
            inners.append(d)
        outer["owner"] = self._next_outer_owner()
        outer["direct"] = bool(cfg.direct_outer and cfg.direct_solver and self._use_chebyshev(cfg))
        if outer["direct"]:
            # no decomposition: the Gram (primal) or kernel matrix (dual) itself is what the outer fit uses
            outer["G_keep"] = outer["G"] if outer["dual"] else G_o
            outer["G"] = outer["G_keep"]
            outer["cheb"] = True  # lambda_max by Lanczos with the inner folds' (see _finish_design)
        elif not outer["dual"]:
            outer["G"] = ops.copy(G_o) if any(d["R"] is not None for d in inners) else G_o
            outer["G_keep"] = G_o
        # Queue the eigendecompositions this rank owns (inner folds first: they are needed first).  With
        # several ranks every eigenproblem is solved once, by rank (job index mod world), and broadcast when
        # it is consumed: X is replicated, so the 30 decompositions of a fit would otherwise be redundant.
        for d in inners + [outer]:
            if d.get("cheb"):  # (a direct outer fold counts as one: no decomposition of its own)
                # GEMM-only fold: lambda_max by Lanczos now (read back once for all folds by fit_shard)
                d["lam"], d["ticket"] = None, None
                d["lmax_dev"] = None  # all folds' Lanczos runs advance together: see _finish_design
            elif d["owner"] != comm.rank:
                d["lam"], d["ticket"] = ops.vec(d["G"].rows), None
            elif cfg.overlap_eig:
                d["lam"], d["ticket"] = ops.syevd_async(d["G"])
            else:
                d["lam"], d["ticket"] = ops.syevd(d["G"]), None
        return outer, inners

    # The GEMM-only inner solvers invert (G + a^2 I) over the FULL spectrum: they cannot drop the directions with
    # S <= singcutoff the way svd_wrapper does (ridge_utils.py:62-65).  That is invisible while the cutoff is far
    # below the smallest shift (a >= 0.05 S[0] in the 'auto' case), so 'auto' keeps them only for cutoffs that small.
    SINGCUTOFF_NEGLIGIBLE = 1e-6

    @classmethod
    def _use_chebyshev(cls, cfg: RidgeConfig) -> bool:
        if cfg.inner_solver == "chebyshev":
            if cfg.singcutoff > cls.SINGCUTOFF_NEGLIGIBLE:
                raise ValueError("inner_solver='chebyshev' solves over the full spectrum and cannot honour "
                                 f"singcutoff={cfg.singcutoff!r}; use inner_solver='eig'")
            return True
        if cfg.inner_solver == "auto":
            ok = bool(cfg.normalpha and len(cfg.alphas) and min(cfg.alphas) >= 0.05
                      and cfg.singcutoff <= cls.SINGCUTOFF_NEGLIGIBLE)
            if not ok and not getattr(cfg, "_warned_eig", False):
                cfg._warned_eig = True
                logger.warning("inner_solver='auto': alphas are not normalised / below 0.05 or singcutoff is not "
                               "negligible -> every inner fold is decomposed with cuSOLVER syevd (about 2.7x slower "
                               "than the GEMM-only inner solver on the BASELINE config)")
            return ok
        return False

    def _scaled_alphas_sq(self, lam_max: float, alphas, cfg: RidgeConfig):
        s0 = np.sqrt(np.float32(lam_max)) if cfg.normalpha else 1.0  # S[0] as the reference's fp32 scalar
        return [(float(a) * float(s0)) ** 2 for a in alphas]

    def _centred_val_design(self, X, d):
        """J P (n_v x p, primal) or J P X_tr^T (n_v x n, dual): the validation side of the fold's predictor, centred
        over the validation rows so that predictions come out mean-free."""
        ops = self.ops
        n_va = len(d["val_rows"])
        if d.get("dual"):
            Pv = ops.gather_rows(X, d["val"], n_va, split=True)
            XiR = ops.gather_rows(X, d["train"], len(d["train_rows"]), split=True)
            Kvt = ops.gemm(Pv, XiR)  # (n_v x n), K = p
            km, _ = ops.col_stats(Kvt, None, n_va, ddof=0)
            return ops.gather_normalize(Kvt, None, n_va, km, None, 2, EPS)
        pm, _ = ops.col_stats(X, d["val"], n_va, ddof=0)
        return ops.gather_normalize(X, d["val"], n_va, pm, None, 2, EPS)  # (n_v x p) fp32

    def _check_lmax(self, lam_max: float, cfg: RidgeConfig) -> None:
        if not (lam_max > 0.0) or not np.isfinite(lam_max):
            raise FloatingPointError("inner-fold Gram has no positive eigenvalue (degenerate design)")
        if min(self._scaled_alphas_sq(lam_max, cfg.alphas, cfg)) * 1e4 < lam_max:
            raise ValueError("inner_solver='chebyshev' needs alpha^2 >= 1e-4 * lambda_max; use inner_solver='eig'")

    def _solve_blocks(self, X, d, alphas, cfg: RidgeConfig):
        """Owner rank: the compact solution block of a GEMM-only fold (see DeviceOps.solve_blocks)."""
        ops = self.ops
        lam_max = float(d["lmax"])
        a2 = self._scaled_alphas_sq(lam_max, alphas, cfg)
        block = ops.solve_blocks(ops.split(d["G"]), self._centred_val_design(X, d), len(d["val_rows"]), lam_max, a2,
                                 lbo=d.pop("lbo_args", None))
        d["G"] = None
        return block

    def _prepare_lbo(self, X, outer, inners, cfg: RidgeConfig) -> None:
        """Leave-block-out folds of one outer fold (see DeviceOps.lbo_prepare): make the outer eigendecomposition
        available (every rank takes part in its broadcast), queue B / H / Lanczos for the folds this rank owns,
        read the Lanczos values and the top outer eigenvalue back in ONE synchronisation, and -- with several
        ranks -- solve the owned folds right away so that no rank waits for another one's solve."""
        ops, comm = self.ops, self.comm
        if not any(d.get("lbo") for d in inners):
            return
        Vt_s, Vt, lam = self._eig_ready(outer)
        todo = []
        for d in inners:
            if not (d.get("lbo") and d["owner"] == comm.rank and "block" not in d):
                continue
            a2 = self._scaled_alphas_sq(float(d["lmax"]), cfg.alphas, cfg)
            cheb, _ = ops.solver_partition(float(d["lmax"]), a2)
            if not cheb:
                continue
            if "V_split" not in outer:
                outer["V_split"] = ops.transpose(Vt, split=True)  # V (p x k): rows = features
            Pv = ops.gather_rows(X, d["val"], len(d["val_rows"]), split=True)  # X_R (|R| x p)
            d["lbo_args"] = {"prep": ops.lbo_prepare(Pv, Vt_s, lam, a2[cheb[0]], lanczos=False), "V": outer["V_split"],
                             "lam": lam}
            todo.append(d)
        if not todo:
            return
        runs = [(idx, ops.lambda_max_batched([todo[i]["lbo_args"]["prep"]["H"] for i in idx], 48))
                for idx in self._same_shape_groups([(i, d["lbo_args"]["prep"]["H"]) for i, d in enumerate(todo)])]
        lam_top = float(np.asarray(ops.download(lam)).reshape(-1)[Vt.rows - 1])
        for idx, h_dev in runs:
            h = np.asarray(ops.download(h_dev)).reshape(-1)
            for k, i in enumerate(idx):
                todo[i]["lbo_args"]["h0"] = float(h[k])
        for d in todo:
            d["lbo_args"]["lam_top"] = lam_top
            if not (0.0 <= d["lbo_args"]["h0"] < 1.0 + 1e-3) or not (lam_top > 0.0):
                # lambda_max(H) must lie in [0, 1): a degenerate outer decomposition -- this fold keeps the direct
                # Chebyshev route on its own Gram (an owner-local decision: the other ranks only receive the block)
                logger.warning("leave-block-out bounds invalid (lambda_max(H) = %r); solving the p x p system",
                               d["lbo_args"]["h0"])
                del d["lbo_args"]
        if comm.world > 1:
            for d in todo:
                d["block"] = self._solve_blocks(X, d, cfg.alphas, cfg)

    def _stack_from_blocks(self, X, d, n_alphas: int, rows_pad: int, alphas, cfg: RidgeConfig):
        """Every rank: alpha stack of a GEMM-only fold from the owner's solution block (broadcast if needed)."""
        ops, comm = self.ops, self.comm
        n_va = len(d["val_rows"])
        block = d.pop("block", None)
        if d["owner"] == comm.rank and block is None:
            with ops.timed("phase_inner_solve"):
                block = self._solve_blocks(X, d, alphas, cfg)
        lam_max = float(d["lmax"])
        a2 = self._scaled_alphas_sq(lam_max, alphas, cfg)
        if comm.world > 1 and "block_work" in d:
            comm.wait_all(d.pop("block_work"))  # the broadcast was started in _finish_design
        elif comm.world > 1:
            if block is None:
                width = len(d["train_rows"]) if d.get("dual") else X.cols
                block = ops.zeros(ops.solver_block_rows(n_va, lam_max, a2), width)
            comm.broadcast_inplace(ops.planes(block), src=d["owner"])
        return ops.assemble_stack(block, self._centred_val_design(X, d), n_va, rows_pad, lam_max, a2,
                                  series_moments=cfg.series_moments)

    def _finish_design(self, groups, cfg: RidgeConfig, outers=None) -> None:
        """After the design side of every plan is queued: read the Lanczos lambda_max of all GEMM-only folds back
        in ONE synchronisation (and, with several ranks, exchange them in ONE all-reduce), then -- with several
        ranks -- solve this rank's folds right away so that no rank waits for another one's solve when the fold
        is consumed.  groups: list of (X, inners)."""
        ops, comm = self.ops, self.comm
        jobs = [(X, d) for X, inners in groups for d in inners if d.get("cheb")]
        n_inner_jobs = len(jobs)
        jobs += [(None, o) for o in (outers or []) if o.get("direct")]  # every rank needs every outer lambda_max
        if not jobs:
            return
        vals = np.zeros(len(jobs), dtype=np.float64)
        mine = lambda i, d: d["owner"] == comm.rank  # noqa: E731  (outer folds: their owner, then all-reduced)
        for idx in self._same_shape_groups([(i, d["G"]) for i, (_, d) in enumerate(jobs) if mine(i, d)]):
            got = np.asarray(ops.download(ops.lambda_max_batched([jobs[i][1]["G"] for i in idx]))).reshape(-1)
            vals[idx] = got[: len(idx)]
        if comm.world > 1:
            vals = comm.all_reduce_sum(vals)
        for (X, d), v in zip(jobs, vals):
            d["lmax"] = float(v)
            # every rank holds every lambda_max after the all-reduce: all of them raise together (an owner-only
            # check would leave the other ranks waiting in the broadcast of the fold's solution block)
            self._check_lmax(float(v), cfg)
        direct_outers = [o for _, o in jobs[n_inner_jobs:]]
        jobs = jobs[:n_inner_jobs]
        if direct_outers:
            # (G_o + a^2 I)^-1 for every outer fold and grid alpha, all Cholesky systems in one batch (every rank:
            # 4 small systems per outer fold are cheaper than shipping 20 p x p inverses)
            # on a second stream: this chain of short launches and the inner solves' below run side by side
            with ops.side_stream() as inv_ready, ops.timed("phase_outer_inverses"):
                lms = [float(o["lmax"]) for o in direct_outers]
                a2s = [self._scaled_alphas_sq(lm, cfg.alphas, cfg) for lm in lms]
                owned = None
                if comm.world > 1:  # the Cholesky systems are dealt out; their inverses travel by broadcast
                    systems = ops.outer_inverse_systems(lms, a2s)
                    owned = {k for k in range(len(systems)) if k % comm.world == comm.rank}
                # (outer folds of different sizes -- ragged chunk counts in the dual form -- go in separate batches)
                invs = [None] * len(direct_outers)
                for idx in self._same_shape_groups([(i, o["G_keep"]) for i, o in enumerate(direct_outers)]):
                    sub_owned = None
                    if owned is not None:
                        sub_sys = ops.outer_inverse_systems([lms[i] for i in idx], [a2s[i] for i in idx])
                        glob = {s: k for k, s in enumerate(systems)}
                        sub_owned = {k for k, (ii, j) in enumerate(sub_sys) if glob[(idx[ii], j)] in owned}
                    got = ops.outer_inverses_many([direct_outers[i]["G_keep"] for i in idx], [lms[i] for i in idx],
                                                  [a2s[i] for i in idx], owned=sub_owned)
                    for i, inv in zip(idx, got):
                        invs[i] = inv
                if comm.world > 1:
                    for k, (i, j) in enumerate(systems):
                        comm.broadcast_inplace(ops.inverse_slot(invs[i], j), src=k % comm.world)
            for o, inv in zip(direct_outers, invs):
                o["inv"], o["inv_ready"] = inv, inv_ready
                ops.adopt([x for x in inv if not isinstance(x, int)] if isinstance(inv, tuple) else [])
        if cfg.direct_solver:
            # every fold this rank owns, all outer folds at once: batched Cholesky solves (128 systems per launch)
            with ops.timed("phase_inner_solve"):
                self._solve_direct([(X, d) for X, d in jobs if d["owner"] == comm.rank], cfg)
            if comm.world > 1 and hasattr(comm, "broadcast_async"):
                # all solution blocks start travelling now (NCCL's stream), in fold order on every rank; each is
                # awaited where its fold is consumed (_stack_from_blocks)
                for X, d in jobs:
                    n_va, lam_max = len(d["val_rows"]), float(d["lmax"])
                    if d["owner"] != comm.rank:
                        width = len(d["train_rows"]) if d.get("dual") else X.cols
                        a2 = self._scaled_alphas_sq(lam_max, cfg.alphas, cfg)
                        d["block"] = ops.empty(ops.solver_block_rows(n_va, lam_max, a2), width)
                    d["block_work"] = comm.broadcast_async(ops.planes(d["block"]), src=d["owner"])
        elif comm.world > 1:
            for X, d in jobs:
                if d["owner"] == comm.rank and not d.get("lbo"):  # leave-block-out folds: see _prepare_lbo
                    d["block"] = self._solve_blocks(X, d, cfg.alphas, cfg)

    def _solve_direct(self, pairs, cfg: RidgeConfig) -> None:
        ops = self.ops
        jobs = []
        for X, d in pairs:
            lam_max = float(d["lmax"])
            jobs.append(dict(G=d["G"], Pc=self._centred_val_design(X, d), n_rows=len(d["val_rows"]), lam_max=lam_max,
                             a2=self._scaled_alphas_sq(lam_max, cfg.alphas, cfg)))
        for (X, d), block in zip(pairs, ops.solve_blocks_many(jobs)):
            d["block"] = block
            d["G"] = None

    @staticmethod
    def _same_shape_groups(items):
        """Index lists of the (index, Mat) pairs that share one size and pitch (batched Lanczos runs)."""
        groups = {}
        for i, m in items:
            groups.setdefault((m.rows, m.ld), []).append(i)
        return list(groups.values())

    def _next_outer_owner(self) -> int:
        """Owner of an outer fold's eigendecomposition: outer k -> rank (world - 1 - k) mod world, so that the long
        cuSOLVER calls land on distinct ranks (and on the ranks with the fewest inner-fold solves) instead of
        wherever a single running job count happens to put them (with 2 ranks and 5 + 1 jobs per outer fold: all
        on rank 1)."""
        owner = (self.comm.world - 1 - self._outer_jobs) % self.comm.world
        self._outer_jobs += 1
        return owner

    def _next_eig_owner(self) -> int:
        owner = self._eig_jobs % self.comm.world
        self._eig_jobs += 1
        return owner

    def _eig_ready(self, d):
        """Wait (on the main stream) for d's eigendecomposition; returns (Vt split, Vt fp32, lam)."""
        ops = self.ops
        if not d.get("eig_ready"):
            if d["ticket"] is not None:
                ops.wait(d["ticket"])
                d["ticket"] = None
            if self.comm.world > 1:
                self.comm.broadcast_inplace([ops.raw(d["G"]), ops.raw(d["lam"])], src=d["owner"])
            d["eig_ready"] = True
            d["Vt_split"] = ops.split(d["G"])
        return d["Vt_split"], d["G"], d["lam"]

    # ------------------------------------------------------------------------------------------
    # inner CV
    # ------------------------------------------------------------------------------------------
    def _response_scales(self, Y):
        """fp16-pair scales of the response COLUMNS from their |max| over all rows: a bound for every gathered subset
        of rows, so that gather_rows_T_f16 needs no pass over its own output.  One reduction per response matrix."""
        c = getattr(self, "_y_scales", None)
        if c is None or c[0] is not Y:
            _, am = self.ops.col_reduce(Y, None, Y.rows, sumsq=False, absmax=True)
            c = self._y_scales = (Y, self.ops.f16_bound_scales(Y.cols, absmax=am))
        return c[1]

    def _response_rows_T(self, Y, idx, n: int, cfg: RidgeConfig, consumer: Optional[str] = None):
        """(V_r x n) transposed gather of response rows in the form its consumer multiplies: an fp16 pair, one fp32
        plane (re-split by the GEMM), or a TF32 pair.  consumer: precision of the GEMM (default: voxel GEMMs)."""
        prec = cfg.voxel_gemm_precision if consumer is None else consumer
        if prec == "f16x3" and cfg.producer_pairs:
            return self.ops.gather_rows_T_f16(Y, idx, n, self._response_scales(Y))
        return self.ops.gather_rows_T_split(Y, idx, n, split=prec != "f16x3")

    def _inner_scores(self, X, Y, sp, outer, inners, alphas_dev, n_alphas: int, cfg: RidgeConfig, all_rows=None):
        """Sum over inner folds of the (n_alphas x V_r) validation scores (ridge_corr_torch per fold).
        Also returns C_o^T (V_r x p fp32) for a primal outer fit (None when the outer fold is dual)."""
        ops = self.ops
        vp = cfg.voxel_gemm_precision
        # operands of the fp16-pair GEMMs: written as fp16 pairs by their producers where a bound on their magnitude is
        # known beforehand (_response_rows_T, the downdate below), else as ONE fp32 plane that the GEMM re-splits (never
        # as a TF32 pair: half the bytes written, half read -- twice -- by lit_split_f16)
        pair_ct = cfg.corr_precision != "f16x3"
        fuse_ct = cfg.producer_pairs and vp == "f16x3" and cfg.corr_precision == "f16x3"
        ct_absmax = None  # max_j |C_o^T[v][j]|, formed on the first downdated inner fold
        Ct_o = None
        if not outer["dual"] and all_rows is not None and sp.get("Ro") is not None:
            # C_o^T = C_all^T - Y_Ro^T X_Ro: the all-rows cross product is formed once per fit (fit_shard)
            n_ro = len(sp["Ro_rows"])
            YRoT = self._response_rows_T(Y, sp["Ro"], n_ro, cfg)  # (V_r x |Ro|)
            XRoT = ops.gather_rows_T_split(X, sp["Ro"], n_ro)  # (p x |Ro|)
            Ct_o = ops.gemm(YRoT, XRoT, alpha=-1.0, Cin=all_rows["Ct"], beta=1.0, precision=vp)
            del YRoT, XRoT
        elif not outer["dual"]:
            YoT = self._response_rows_T(Y, sp["train"], len(sp["train_rows"]), cfg)  # (V_r x n_o)
            Ct_o = ops.gemm(YoT, outer["XtT"], precision=vp)  # (V_r x p), K = n_o
            del YoT
        with ops.timed("phase_lbo_prepare"):
            self._prepare_lbo(X, outer, inners, cfg)
        corr_sum = ops.empty(n_alphas, Y.cols)
        metric = 0 if cfg.use_corr else 1
        for i, d in enumerate(inners):
            n_tr, n_va = len(d["train_rows"]), len(d["val_rows"])
            if n_va < 2 or n_tr < 1:
                raise ValueError("inner fold needs >= 1 training and >= 2 validation samples")
            rows_pad = -(-n_va // ops.TILE_N) * ops.TILE_N
            # (n_v x p) validation design: only the decomposition routes rotate it into an eigenbasis
            Pv = None if d["cheb"] else ops.gather_rows(X, d["val"], n_va, split=True)
            if d["dual"] and d["cheb"]:
                # dual GEMM-only fold: pred_a = [J K_vt (K_tt + a^2 I)^-1] Y_tr -- the stack lives in R^n and the
                # "coefficients" are the training responses themselves (no cross product, no decomposition)
                Zt = self._response_rows_T(Y, d["train"], n_tr, cfg, consumer=cfg.corr_precision)  # (V_r x n)
                Lst = self._stack_from_blocks(X, d, n_alphas, rows_pad, cfg.alphas, cfg)
                L = None
            elif d["dual"]:
                # dual form: Z^T = Y_tr^T U (V_r x n), L = (P X_tr^T) U (n_v x n); lam = eig(X_tr X_tr^T) = S^2
                YtT = ops.gather_rows_T_split(Y, d["train"], n_tr)  # (V_r x n)
                Ut, _, lam = self._eig_ready(d)
                Zt = ops.gemm(YtT, Ut, split_out=True, precision=vp)
                del YtT
                XiR = ops.gather_rows(X, d["train"], n_tr, split=True)  # (n x p)
                Kpt = ops.gemm(Pv, XiR, split_out=True)  # (n_v x n), K = p
                del XiR
                L = ops.gemm(Kpt, Ut)
                del Kpt, Ut
            else:
                # cross product of the inner training rows (downdated from the outer fold's when possible)
                if d["R"] is not None:
                    n_r = len(d["R_rows"])
                    YRt = self._response_rows_T(Y, d["R"], n_r, cfg)  # (V_r x |R|)
                    if fuse_ct and d["cheb"]:
                        # the downdate GEMM writes the fp16 pair the prediction GEMM reads; its row scales come
                        # from |C_i^T[v][j]| <= max_j |C_o^T[v][j]| + |y_(v,R)|_2 max_j |x_(j,R)|_2
                        if ct_absmax is None:
                            ct_absmax = ops.row_absmax(Ct_o)
                        scales = ops.f16_bound_scales(Y.cols, absmax=ct_absmax,
                                                      row_sumsq=ops.col_reduce(Y, d["R"], n_r)[0],
                                                      col_sumsq=ops.col_reduce(X, d["R"], n_r)[0])
                        Ct = ops.gemm(YRt, d["XRt"], alpha=-1.0, Cin=Ct_o, beta=1.0, precision=vp, pair_out=scales)
                    else:
                        Ct = ops.gemm(YRt, d["XRt"], alpha=-1.0, Cin=Ct_o, beta=1.0,
                                      split_out=pair_ct or not d["cheb"], precision=vp)
                    del YRt
                else:
                    YtT = self._response_rows_T(Y, d["train"], n_tr, cfg)
                    Ct = ops.gemm(YtT, d["XtT"], split_out=pair_ct or not d["cheb"], precision=vp)
                    del YtT
                if d["cheb"]:
                    # GEMM-only fold: pred_a^T = C^T [P_c (G + a^2 I)^-1]^T, no rotation into an eigenbasis
                    Zt = Ct
                    Lst = self._stack_from_blocks(X, d, n_alphas, rows_pad, cfg.alphas, cfg)
                    L = None
                else:
                    Vt, _, lam = self._eig_ready(d)
                    Zt = ops.gemm(Ct, Vt, split_out=True, precision=vp)  # (V_r x k), K = p
                    L = ops.gemm(Pv, Vt)  # P V  (n_v x k)
                    del Vt
                del Ct
            # validation design in the eigenbasis, centred and stacked over alphas
            if L is not None:
                Lst = ops.build_alpha_stack(L, n_va, rows_pad, lam, alphas_dev, n_alphas, cfg.normalpha, cfg.singcutoff)
            del Pv, L
            mean, std = ops.col_stats(Y, d["val"], n_va, ddof=1)
            Yz = ops.gather_normalize(Y, d["val"], n_va, mean, std, 0 if cfg.use_corr else 2, EPS, rows_out=rows_pad)
            parts = ops.gemm_corr(Zt, Lst, n_alphas, rows_pad, Yz, precision=cfg.corr_precision)
            ops.corr_finalize(parts, rows_pad // ops.PART_N, n_alphas, Y.cols, n_va, EPS, corr_sum,
                              accumulate=(i > 0), metric=metric, resp_std=std)
            del Zt, Lst, Yz, parts
            d["G"] = d["XRt"] = d["XtT"] = d["Vt_split"] = None
        return corr_sum, Ct_o

    def _select_alphas(self, corr_sum, n_folds: int, alphas_dev, cfg: RidgeConfig, n_vox_total: int):
        """nested_cv._find_best_alphas :391-413 -> device vector of per-voxel alpha values."""
        ops = self.ops
        best, alpha_v, sums = ops.argmax_alpha(corr_sum, n_folds, alphas_dev, want_sums=cfg.single_alpha)
        if cfg.single_alpha:
            tot = self.comm.all_reduce_sum(np.asarray(ops.download(sums), dtype=np.float64)[: len(cfg.alphas)])
            j = int(np.argmax((tot / float(n_vox_total)).astype(np.float32)))
            alpha_v = ops.upload_vector(np.full(corr_sum.cols, np.float32(cfg.alphas[j]), dtype=np.float32), "f32")
            best = ops.upload_vector(np.full(corr_sum.cols, j, dtype=np.int32), "i32")
        self._best_idx = best  # alpha INDEX per voxel (the grouped outer fit sorts the voxels by it)
        return alpha_v

    # ------------------------------------------------------------------------------------------
    # outer fit + test scoring
    # ------------------------------------------------------------------------------------------
    def _shrunk_coefficients(self, X, Y, sp, outer, Ct_o, alpha_v, cfg: RidgeConfig):
        """(Z_o^T * keep / (lam_o + (alpha_v s0)^2))  (V_r x k split) and the eigenvector matrix (rows)."""
        ops = self.ops
        Vt, G, lam = self._eig_ready(outer)
        if outer["dual"]:
            YoT = ops.gather_rows_T_split(Y, sp["train"], len(sp["train_rows"]))  # (V_r x n_o)
            Zt = ops.gemm(YoT, Vt, split_out=True, precision=cfg.voxel_gemm_precision)  # Y_tr^T U
            del YoT
        else:
            Zt = ops.gemm(ops.split(Ct_o), Vt, split_out=True, precision=cfg.voxel_gemm_precision)  # (V_r x k)
        del Vt
        return ops.scale_rows_by_alpha(Zt, lam, alpha_v, cfg.normalpha, cfg.singcutoff), G

    def _outer_weights(self, X, Y, sp, outer, Ct_o, alpha_v, cfg: RidgeConfig):
        """ridge_torch (ridge_regression.py:29-63) for every voxel at once -> W^T (V_r x p split).
        Primal: W^T = ZS V^T.  Dual: W^T = ZS (X_tr^T U)^T with U the eigenvectors of X_tr X_tr^T."""
        ops = self.ops
        if outer.get("direct") and outer["dual"] and getattr(self, "_best_idx", None) is not None:
            # dual form: D^T[v] = y_v^T (K_o + a_v^2 I)^-1 (grouped GEMM over the sorted voxels), W^T = D^T X_tr
            n_o = len(sp["train_rows"])
            inv = outer.pop("inv", None)
            ops.wait_copy(outer.pop("inv_ready", None))
            if inv is None:
                lam_max = float(outer["lmax"])
                inv = ops.outer_inverses(outer["G_keep"], lam_max, self._scaled_alphas_sq(lam_max, cfg.alphas, cfg))
            YoT = ops.gather_rows_T_split(Y, sp["train"], n_o)  # (V_r x n_o) split pair = fp32 exactly (hi + lo)
            Yf = ops.zeros(YoT.rows, YoT.cols)
            ops.axpy(1.0, YoT, Yf)
            del YoT
            pos, perm, tile_group, cap = ops.group_plan(self._best_idx, Yf.rows, len(cfg.alphas))
            Dt = ops.gather_rows(ops.gemm_grouped(ops.gather_rows(Yf, perm, cap, split=True), inv, tile_group,
                                                  split_out=False), pos, Yf.rows, split=True)
            del Yf, inv
            XoT = ops.gather_rows_T_split(X, sp["train"], n_o)  # (p x n_o)
            return ops.gemm(Dt, XoT, split_out=True, precision=cfg.voxel_gemm_precision)
        if outer.get("direct") and getattr(self, "_best_idx", None) is not None:
            # ridge_torch without a decomposition: W^T[v] = C^T[v] (G_o + a_v^2 I)^-1, one grouped GEMM over the voxels
            # sorted by alpha index (a_v = alpha_v * S[0] of the OUTER training set, S[0]^2 = lambda_max by Lanczos)
            inv = outer.pop("inv", None)
            ops.wait_copy(outer.pop("inv_ready", None))
            if inv is None:
                lam_max = float(outer["lmax"])
                inv = ops.outer_inverses(outer["G_keep"], lam_max, self._scaled_alphas_sq(lam_max, cfg.alphas, cfg))
            pos, perm, tile_group, cap = ops.group_plan(self._best_idx, Ct_o.rows, len(cfg.alphas))
            sorted_ct = ops.gather_rows(Ct_o, perm, cap, split=True)
            return ops.gather_rows(ops.gemm_grouped(sorted_ct, inv, tile_group, split_out=False), pos, Ct_o.rows,
                                   split=True)
        ZS, G = self._shrunk_coefficients(X, Y, sp, outer, Ct_o, alpha_v, cfg)
        if outer["dual"]:
            XoT = ops.gather_rows_T_split(X, sp["train"], len(sp["train_rows"]))  # (p x n_o)
            basis = ops.gemm(XoT, ops.split(G), split_out=True)  # X_tr^T U  (p x n_o)
        else:
            basis = outer.get("V_split")
            if basis is None:
                basis = ops.transpose(G, split=True)  # V (p x k): rows = features
        return ops.gemm(ZS, basis, split_out=True, precision=cfg.voxel_gemm_precision)

    def _outer_fit_and_score(self, X, Y, Xte_src, Yte_src, sp, outer, Ct_o, alpha_v, cfg: RidgeConfig):
        """ridge_torch on the outer training set with the selected alphas, then test r / p."""
        ops = self.ops
        n_te = len(sp["test_rows"])
        te_dev = sp["test"]
        Wt = self._outer_weights(X, Y, sp, outer, Ct_o, alpha_v, cfg)  # (V_r x p) voxel-major weights
        # test predictions, centred so that the fused epilogue yields Pearson's r directly
        rows_pad = -(-n_te // ops.TILE_N) * ops.TILE_N
        pm, _ = ops.col_stats(Xte_src, te_dev, n_te, ddof=0)
        Pt = ops.gather_normalize(Xte_src, te_dev, n_te, pm, None, 2, EPS, rows_out=rows_pad, split=True)
        ym, ys = ops.col_stats(Yte_src, te_dev, n_te, ddof=1)
        Ytz = ops.gather_normalize(Yte_src, te_dev, n_te, ym, ys, 1, EPS, rows_out=rows_pad)
        parts = ops.gemm_corr(Wt, Pt, 1, rows_pad, Ytz)
        r, p = ops.pearson_finalize(parts, Wt.rows, n_te, cfg.p_round_f32)
        return Wt, r, p

    # ------------------------------------------------------------------------------------------
    # stand-alone ridge kernels of the reference (encoding/models/ridge_regression.py)
    # ------------------------------------------------------------------------------------------
    def _single_split(self, n_train: int, n_val: int, cfg: RidgeConfig, inner: bool = True):
        """One staged plan whose outer training rows are [0, n_train) and whose test rows -- and, when
        `inner`, the validation rows of a single inner fold on the same training rows -- follow them."""
        tr = np.arange(n_train, dtype=np.int64)
        va = np.arange(n_train, n_train + n_val, dtype=np.int64)
        self._eig_jobs = self._outer_jobs = 0
        return self.stage_plans([FoldPlan(tr, va, [(tr, va)] if (inner and n_val) else [])], cfg)[0]

    def ridge_corr(self, XX, YY, n_train: int, cfg: RidgeConfig):
        """ridge_corr_torch (ridge_regression.py:66-141): XX / YY hold the training rows followed by the
        prediction rows; returns the (n_alphas x V) score matrix (device)."""
        ops = self.ops
        cfg = replace(cfg, direct_outer=False)  # alphas arrive as values here: the decomposition route
        sp = self._single_split(n_train, XX.rows - n_train, cfg)
        outer, inners = self._design_side(XX, sp, cfg)
        self._finish_design([(XX, inners)], cfg)
        alphas = ops.upload_vector(np.asarray(cfg.alphas, dtype=np.float64), "f64")
        corr, _ = self._inner_scores(XX, YY, sp, outer, inners, alphas, len(cfg.alphas), cfg)
        self._eig_ready(outer)  # consume the (unused) outer ticket
        self._y_scales = None
        return corr

    def ridge_weights(self, X, Y, alpha_v, cfg: RidgeConfig):
        """ridge_torch (ridge_regression.py:9-63): returns W^T (V x p, split pair) on the device."""
        ops = self.ops
        cfg = replace(cfg, direct_outer=False)  # alphas arrive as values here: the decomposition route
        sp = self._single_split(X.rows, 0, cfg)
        outer, _ = self._design_side(X, sp, cfg)
        Ct_o = None
        if not outer["dual"]:
            YoT = ops.gather_rows_T_split(Y, sp["train"], X.rows)
            Ct_o = ops.gemm(YoT, outer["XtT"])
            del YoT
        return self._outer_weights(X, Y, sp, outer, Ct_o, alpha_v, cfg)

    def ridge_corr_pred(self, XX, YY, n_train: int, alpha_v, cfg: RidgeConfig):
        """ridge_corr_pred_torch (ridge_regression.py:144-216): per-voxel alphas, no weights formed:
        pred^T = (Z^T * d_v) (P V)^T reduced against the z-scored responses in the GEMM epilogue.
        Returns the (1 x V) score (device); NaNs are kept, as in the reference."""
        ops = self.ops
        n_va = XX.rows - n_train
        cfg = replace(cfg, direct_outer=False)  # alphas arrive as values here: the decomposition route
        sp = self._single_split(n_train, n_va, cfg, inner=False)
        outer, _ = self._design_side(XX, sp, cfg)
        val = sp["test"]
        Ct_o = None
        if not outer["dual"]:
            YoT = ops.gather_rows_T_split(YY, sp["train"], n_train)
            Ct_o = ops.gemm(YoT, outer["XtT"])
            del YoT
        ZS, G = self._shrunk_coefficients(XX, YY, sp, outer, Ct_o, alpha_v, cfg)
        rows_pad = -(-n_va // ops.TILE_N) * ops.TILE_N
        Pv = ops.gather_rows(XX, val, n_va, split=True)
        if outer["dual"]:
            XiR = ops.gather_rows(XX, sp["train"], n_train, split=True)
            Pv = ops.gemm(Pv, XiR, split_out=True)  # P X_tr^T (n_v x n)
            del XiR
        L = ops.gemm(Pv, ops.split(G))  # (n_v x k)
        lm, _ = ops.col_stats(L, None, n_va, ddof=0)
        Lc = ops.gather_normalize(L, None, n_va, lm, None, 2, EPS, rows_out=rows_pad, split=True)
        mean, std = ops.col_stats(YY, val, n_va, ddof=1)
        Yz = ops.gather_normalize(YY, val, n_va, mean, std, 0 if cfg.use_corr else 2, EPS, rows_out=rows_pad)
        parts = ops.gemm_corr(ZS, Lc, 1, rows_pad, Yz)
        corr = ops.empty(1, YY.cols)
        ops.corr_finalize(parts, rows_pad // ops.PART_N, 1, YY.cols, n_va, EPS, corr, accumulate=False,
                          metric=(0 if cfg.use_corr else 1) | 2, resp_std=std)
        return corr

    def _normalised(self, M, Mte_src, train_rows, train_dev, enabled: bool, same_source: bool):
        """DataNormalizer (ridge_utils.py:70-180): z-score a matrix (and its test source) with the
        training rows' statistics (unbiased std, eps added to the std)."""
        if not enabled:
            return M, Mte_src
        ops = self.ops
        m, s = ops.col_stats(M, train_dev, len(train_rows), ddof=1)
        Mn = ops.gather_normalize(M, None, M.rows, m, s, 0, EPS)
        Mtn = Mn if same_source else ops.gather_normalize(Mte_src, None, Mte_src.rows, m, s, 0, EPS)
        return Mn, Mtn

    # ------------------------------------------------------------------------------------------
    # drivers
    # ------------------------------------------------------------------------------------------
    def fit_shard(self, X, Y, plans: List[FoldPlan], cfg: RidgeConfig, X_test=None, Y_test=None,
                  n_vox_total: Optional[int] = None, y_ready=None) -> ShardResult:
        """Run every outer fold on this rank's voxel shard.

        X (N x p) and Y (N x V_r) are device matrices.  In train/test mode there is one plan whose
        test_rows index X_test / Y_test; in nested mode test rows index X / Y themselves."""
        ops = self.ops
        self._best_idx = self._y_scales = None
        alphas = np.asarray(cfg.alphas, dtype=np.float64)
        alphas_f32 = ops.upload_vector(alphas.astype(np.float32), "f32")
        alphas_f64 = ops.upload_vector(alphas, "f64")
        n_vox_total = Y.cols if n_vox_total is None else n_vox_total
        res = ShardResult()
        same_source = X_test is None
        inv = 1.0 / len(plans)
        # Design side of EVERY outer fold first: Grams are cheap and depend on X only, and queueing all the
        # eigendecompositions now lets the side stream run ahead of the response-side GEMMs.
        Xte_src, Yte_src = (X, Y) if same_source else (X_test, Y_test)
        downdate_outer = same_source and not (cfg.normalize_features or cfg.normalize_targets)
        staged = self.stage_plans(plans, cfg, n_rows_total=X.rows if downdate_outer else None)
        self._eig_jobs = self._outer_jobs = 0
        prepared = []
        with ops.timed("phase_design"):  # phase_* categories: coarse CUDA-event brackets for the bench's report
            for sp in staged:
                Xs, Xts = self._normalised(X, Xte_src, sp["train_rows"], sp["train"], cfg.normalize_features,
                                           same_source)
                prepared.append((Xs, Xts) + self._design_side(Xs, sp, cfg))
            self._finish_design([(pr[0], pr[3]) for pr in prepared], cfg, outers=[pr[2] for pr in prepared])
        ops.wait_copy(y_ready)  # the responses may still be in flight on the copy stream: first use is below
        all_rows = None
        if any(sp.get("Ro") is not None for sp in staged) and not all(pr[2]["dual"] for pr in prepared):
            # Y^T X over ALL rows, once: every outer fold's cross product is this minus its removed rows' share
            # (Y is then streamed once + 1/5 per outer fold instead of 4/5 per outer fold)
            with ops.timed("phase_cross_all"):
                idx_all = ops.stage_indices([np.arange(X.rows, dtype=np.int64)])[0]
                YaT = self._response_rows_T(Y, idx_all, X.rows, cfg)
                XaT = ops.gather_rows_T_split(X, idx_all, X.rows)
                all_rows = {"Ct": ops.gemm(YaT, XaT, precision=cfg.voxel_gemm_precision)}
                del YaT, XaT
        for plan, sp in zip(plans, staged):
            Xs, Xts, outer, inners = prepared.pop(0)
            Ys, Yts = self._normalised(Y, Yte_src, sp["train_rows"], sp["train"], cfg.normalize_targets, same_source)
            with ops.timed("phase_inner_cv"):
                corr_sum, Ct_o = self._inner_scores(Xs, Ys, sp, outer, inners, alphas_f64, len(alphas), cfg,
                                                    all_rows=all_rows)
                alpha_v = self._select_alphas(corr_sum, len(plan.inner), alphas_f32, cfg, n_vox_total)
                if cfg.record_scores:  # torch.stack(all_corrs).mean(dim=0), nested_cv.py:391-393
                    res.scores.append(ops.download_matrix(corr_sum) / np.float32(len(plan.inner)))
            del corr_sum, inners
            with ops.timed("phase_outer_fit"):
                Wt, r, p = self._outer_fit_and_score(Xs, Ys, Xts, Yts, sp, outer, Ct_o, alpha_v, cfg)
            del outer, Ct_o, Xs, Xts, Ys, Yts
            if len(plans) == 1:
                res.Wt_mean = Wt
            else:
                if res.Wt_mean is None:
                    res.Wt_mean = ops.zeros(Wt.rows, Wt.cols)
                ops.axpy(inv, Wt, res.Wt_mean)  # np.mean(fold_weights, axis=0), nested_cv.py:296
            del Wt
            res.r.append(r)
            res.p.append(p)
            res.alpha.append(alpha_v)
            res.n_test.append(len(plan.test_rows))
        self._y_scales = None  # holds a reference to the response matrix
        return res

    def weights_matrix(self, res: ShardResult):
        """(p x V_r) fp32 device matrix of this shard's (fold-mean) weights."""
        ops = self.ops
        W = res.Wt_mean
        if W.is_split:  # single fold: the GEMM wrote a split pair -> recombine hi + lo
            full = ops.zeros(W.rows, W.cols)
            ops.axpy(1.0, W, full)
            W = full
        return ops.transpose(W)

    def significance(self, p_folds: np.ndarray, cfg: RidgeConfig):
        """BH per fold, Fisher across folds, BH on the combined p (nested_cv.py:263-290) on device.

        p_folds: (n_folds x V) float64 host array holding ALL voxels (gathered across shards).
        Returns per-fold masks, combined p, its BH mask and adjusted p (host arrays)."""
        ops = self.ops
        K, V = p_folds.shape
        masks = []
        p_dev = [ops.upload_vector(p_folds[f], "f64") for f in range(K)]
        padj0 = None
        for f in range(K):
            rej, padj, _ = ops.bh_fdr(p_dev[f], V, cfg.alpha_fdr)
            masks.append(ops.download(rej)[:V].astype(bool))
            if K == 1:
                padj0 = ops.download(padj)[:V]
        if K == 1:
            return masks, None, masks[0], padj0
        stack = ops.stack_vectors(p_dev, V, "f64")
        comb = ops.fisher(stack, K, V, cfg.p_round_f32)
        rej, padj, _ = ops.bh_fdr(comb, V, cfg.alpha_fdr)
        return masks, ops.download(comb)[:V], ops.download(rej)[:V].astype(bool), ops.download(padj)[:V]
