"""litcoder_core_b200 -- the B200 (sm_100a) hot path of LITcoder's encoding models.

Drop-in replacements for the three reference entry points on that path:

    NestedCVModel / fit_nested_cv     encoding/models/nested_cv.py   (nested-CV ridge regression)
    Downsampler                       encoding/downsample/downsampling.py   (Lanczos TR resampling)
    FIR                               encoding/features/FIR_expander.py     (delay stacking)
    create_folds                      encoding/models/folding.py
    apply_fir_delays, create_train_test_split, create_concatenated_data
                                      encoding/trainer.py:203-282   (per-story trim / z-score / stacking)

All arithmetic runs in hand-written CUDA kernels behind the C ABI of `liblitridge.so`
(include/litridge.h); importing this package does not need a GPU, calling it does.
"""
from .downsample import Downsampler
from .fir import FIR
from .folding import create_folds
from .nested_cv import NestedCVModel, fit_nested_cv
from .structure import apply_fir_delays, create_concatenated_data, create_train_test_split
from .ridge_regression import ridge, ridge_corr, ridge_corr_pred, ridge_corr_pred_torch, ridge_corr_torch, ridge_torch, zs

__all__ = ["NestedCVModel", "fit_nested_cv", "Downsampler", "FIR", "create_folds", "ridge", "ridge_corr",
           "ridge_corr_pred", "ridge_torch", "ridge_corr_torch", "ridge_corr_pred_torch", "zs", "apply_fir_delays",
           "create_train_test_split", "create_concatenated_data"]
__version__ = "0.1.0"
