"""The C-ABI library loads without a GPU and exports every symbol include/litridge.h declares; the
ctypes prototype table covers exactly the header's entry points; the package fails loudly (no CPU
fallback) when asked to compute without a CUDA device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "litridge.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lit_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from litcoder_core_b200 import _lib

    names = _declared()
    assert len(names) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in litridge.h but not exported"
    assert sorted(_lib.PROTOTYPES) == names, "ctypes prototype table and header disagree"
    loaded = _lib.load()
    assert loaded.lit_abi_version() == _lib.ABI_VERSION == 6
    assert loaded.lit_last_error() is not None


def test_argument_checks_without_gpu():
    """Entry points validate their arguments before touching the device."""
    from litcoder_core_b200 import _lib

    lib = _lib.load()
    rc = lib.lit_gemm_tf32x3_nt(None, None, 0, None, None, 0, -1, 4, 4, 1.0, None, 0, 0.0, None, None, 4, 0, None)
    assert rc == -22 and b"negative" in lib.lit_last_error()
    rc = lib.lit_gemm_tf32x3_nt_corr(None, None, 0, None, None, 0, 8, 1, 100, 8, None, 0, None, None, 0, 0, None)
    assert rc == -22 and b"multiple of 256" in lib.lit_last_error()
    rc = lib.lit_bh_fdr(None, 0, 0.05, None, None, None, None, 0, None)
    assert rc == -22
    need = ctypes.c_size_t(0)
    assert lib.lit_bh_workspace(95000, ctypes.byref(need)) == 0 and need.value == 131072 * 16
    with pytest.raises(_lib.LitRidgeError):
        _lib.check(lib.lit_fir_make_delayed(None, 7, 4, 4, 4, None, 1, 0, None, 4, None), "fir")


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np

    import litcoder_core_b200 as L

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        L.FIR.make_delayed(np.zeros((4, 2)), [1, 2])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        L.fit_nested_cv(features=np.zeros((100, 4), np.float32), targets=np.zeros((100, 8), np.float32))
