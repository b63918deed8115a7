"""Synthetic inputs of SURVEY.md section 8(d) for the bench, the full-scale parity run and the wide GPU tests.

    stories     25 "stories" with n_TR ~ U{300..450} summing to N (TR = 2 s); word times = cumulative Exp(mean
                0.33 s); embeddings e_t = g_t + 0.7 e_{t-1}, g ~ N(0, I_D) fp32 (AR(1): delayed features are
                realistically correlated).  Speech streams (config 3) are 10 Hz frames of the same AR(1) kind.
    design      Lanczos(window=3, cutoff_mult=1.0) -> FIR delays (1, 2, 3, 4) -> per-story column z-score (ddof=0)
                -> vstack, either through the PRODUCT's kernels (design_device: lit_lanczos_downsample,
                lit_fir_zscore_rows) or through the CPU oracle (design_host) -- the same matrix up to fp32 rounding.
    responses   Y = X W + eps, W ~ N(0, 1/p) on a random 30 % of the voxels (the rest pure noise), eps ~ N(0, s^2)
                with s chosen so that the median true r of the signal voxels is 0.1, plus 0.1 % exactly constant and
                0.1 % duplicated voxels (tie / NaN paths).

Data generation is not the product: responses are drawn with torch on the device (bench) or NumPy on the host
(CPU arms); only the design goes through the kernels under test.  Not imported by the package.
"""
from __future__ import annotations

import numpy as np

DELAYS = (1, 2, 3, 4)
# name: (TRs, [(stream kind, width)], voxels, TR seconds, stories)
CONFIGS = {
    "config1_wordrate_9400x4x95000": (9400, [("wordrate", 1)], 95000, 2.0, 25),
    "config2_gpt2_9400x3072x95000": (9400, [("words", 768)], 95000, 2.0, 25),
    "config3_whisper_gpt2_9400x5120x95000": (9400, [("frames", 512), ("words", 768)], 95000, 2.0, 25),
    "config3b_whisper_gpt2_9400x10240x95000": (9400, [("frames", 1792), ("words", 768)], 95000, 2.0, 25),
    "config4_narratives_2226x3072x81924": (2226, [("words", 768)], 81924, 1.5, 1),
    "config5_llama_9400x16384x95000": (9400, [("words", 4096)], 95000, 2.0, 25),
    "dev_small_2000x256x4096": (2000, [("words", 64)], 4096, 2.0, 6),
}


def story_lengths(rng, n_rows: int, n_stories: int):
    """n_TR per story ~ U{300..450} (scaled when n_rows / n_stories lies outside that range), summing to n_rows."""
    if n_stories == 1:
        return [n_rows]
    lens = rng.integers(300, 451, n_stories).astype(np.float64)
    lens = np.maximum(8, np.floor(lens * n_rows / lens.sum())).astype(np.int64)
    lens[-1] += n_rows - lens.sum()
    assert lens.min() > 4 and lens.sum() == n_rows
    return [int(x) for x in lens]


def _ar1(rng, n: int, D: int, phi: float = 0.7):
    from scipy.signal import lfilter

    g = rng.standard_normal((n, D), dtype=np.float32)
    return lfilter([1.0], [1.0, -phi], g, axis=0).astype(np.float32)


def make_stories(workload: str, seed: int = 0):
    """Per story: {"tr_times", "streams": [(kind, times or None, values (n_samples x D) float32)]}."""
    N, streams, _, tr, n_stories = CONFIGS[workload]
    rng = np.random.default_rng(seed)
    out = []
    for n_tr in story_lengths(rng, N, n_stories):
        tr_times = np.arange(n_tr, dtype=np.float64) * tr + tr / 2
        st = []
        for kind, D in streams:
            if kind == "wordrate":  # per-TR word counts, no resampling (trainer.py:168-172)
                st.append((kind, None, rng.poisson(4.0, (n_tr, D)).astype(np.float32)))
                continue
            if kind == "words":
                t = np.cumsum(rng.exponential(0.33, int(n_tr * tr / 0.33 * 1.2) + 16))
                t = t[t < n_tr * tr]
            else:  # 10 Hz frames
                t = np.arange(0.05, n_tr * tr, 0.1)
            st.append((kind, t, _ar1(rng, len(t), D)))
        out.append({"tr_times": tr_times, "streams": st})
    return out


def _downsampled(stories, lanczos):
    """Per story: the (n_TR x sum D) float64 features after resampling every stream onto the TR grid."""
    feats = []
    for s in stories:
        cols = [v.astype(np.float64) if t is None else lanczos(v, t, s["tr_times"]) for _, t, v in s["streams"]]
        feats.append(cols[0] if len(cols) == 1 else np.concatenate(cols, axis=1))
    return feats


def design_host(stories, delays=DELAYS):
    """The design through the CPU oracle (reference semantics): float32 (N x p)."""
    from oracle import ridge_oracle as O

    feats = _downsampled(stories, lambda v, t, tr: O.lanczos_interp2d(v, t, tr, 3, 1.0))
    return np.nan_to_num(np.vstack([O.zs(O.fir_make_delayed(f, list(delays))) for f in feats])).astype(np.float32)


def design_device(stories, ops, delays=DELAYS):
    """The design through the product's kernels; returns a torch CUDA float32 tensor (N x p), resident."""
    import litcoder_core_b200 as L
    from litcoder_core_b200 import structure

    ds = L.Downsampler(ops=ops)
    feats = _downsampled(stories, lambda v, t, tr: ds.downsample(v, t, tr, method="lanczos", window=3, cutoff_mult=1.0))
    names = [str(i) for i in range(len(feats))]
    X = structure._stim_side(ops, dict(zip(names, feats)), names, None, None, [int(d) for d in delays], False)
    return ops.as_tensor(X)


def voxel_roles(rng, V: int, frac_signal: float = 0.3):
    """(signal mask, constant voxel ids, (duplicate ids, their sources))."""
    signal = rng.random(V) < frac_signal
    n_special = max(1, V // 1000)
    special = rng.choice(V, size=min(V, 2 * n_special), replace=False)
    const, dup = special[:n_special], special[n_special:]
    src = (dup + 1 + rng.integers(0, max(V - 1, 1), len(dup))) % V
    return signal, const, (dup, src)


def responses_host(X: np.ndarray, V: int, seed: int = 0, true_r: float = 0.1, v0: int = 0, v1: int = None):
    """Columns [v0, v1) of Y on the host (NumPy): the same voxel roles as responses_device for the same seed; the
    noise stream is NumPy's (the device arm draws its own)."""
    rng = np.random.default_rng(seed + 1)
    signal, const, (dup, src) = voxel_roles(rng, V)
    v1 = V if v1 is None else v1
    N, p = X.shape
    rs = np.random.default_rng([seed, 7, v0])
    W = (rs.standard_normal((p, v1 - v0), dtype=np.float32) / np.float32(np.sqrt(p))) * signal[v0:v1][None, :]
    S = X @ W
    sd = np.median(S.std(0)[signal[v0:v1]]) if signal[v0:v1].any() else 1.0
    sigma = np.float32(sd * np.sqrt(1.0 / true_r ** 2 - 1.0))
    Y = S + sigma * rs.standard_normal((N, v1 - v0), dtype=np.float32)
    for c in const:
        if v0 <= c < v1:
            Y[:, c - v0] = np.float32(1.5)
    for d, s in zip(dup, src):
        if v0 <= d < v1 and v0 <= s < v1:
            Y[:, d - v0] = Y[:, s - v0]
    return Y.astype(np.float32)


def responses_device(torch, X, V: int, seed: int = 0, true_r: float = 0.1):
    """Y (N x V) float32 on X's device, drawn with torch (data generation is not the product)."""
    rng = np.random.default_rng(seed + 1)
    signal, const, (dup, src) = voxel_roles(rng, V)
    N, p = X.shape
    g = torch.Generator(device=X.device).manual_seed(seed)
    Y = torch.empty((N, V), device=X.device, dtype=torch.float32)
    sig_t = torch.from_numpy(signal).to(X.device)
    sds = []
    for c0 in range(0, V, 16384):  # column blocks: bounded scratch for W
        c1 = min(V, c0 + 16384)
        W = torch.randn((p, c1 - c0), device=X.device, generator=g) / p ** 0.5
        W *= sig_t[c0:c1][None, :]
        Y[:, c0:c1] = X @ W
        sds.append(Y[:, c0:c1].std(0)[sig_t[c0:c1]])
    sd = torch.cat(sds).median().item() if signal.any() else 1.0
    sigma = sd * (1.0 / true_r ** 2 - 1.0) ** 0.5
    for r0 in range(0, N, 2048):
        Y[r0:r0 + 2048] += sigma * torch.randn((min(2048, N - r0), V), device=X.device, generator=g)
    Y[:, torch.from_numpy(const).to(X.device)] = 1.5
    Y[:, torch.from_numpy(dup).to(X.device)] = Y[:, torch.from_numpy(src).to(X.device)]
    return Y
