"""Drop-in for the reference's FIR delay expander (encoding/features/FIR_expander.py:6-73).

`FIR.make_delayed(stim, delays, circpad=False)` stacks shifted copies of the stimulus matrix
side by side (delay-major column order).  The copy runs on the B200 (lit_fir_make_delayed: one
coalesced read of the stimulus, `ndelays` coalesced float64 writes); the host code only mirrors
the reference's dtype rule: the output is float64 unless every delay is 0, in which case NumPy's
hstack keeps the input dtype (FIR_expander.py:31,41-43).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Optional

import numpy as np


@dataclass
class FIR:
    """Finite impulse response expander.

    Static use: ``FIR.make_delayed(stim, delays, circpad=False)``; instance use:
    ``FIR(delays, circpad).expand(stim)``.
    """

    delays: Optional[Iterable[int]] = None
    circpad: bool = False

    def expand(self, stim: np.ndarray) -> np.ndarray:
        if self.delays is None:
            raise ValueError("delays must be provided for instance usage of FIR")
        return FIR.make_delayed(stim, self.delays, self.circpad)

    @staticmethod
    def make_delayed(stim: np.ndarray, delays: Iterable[int], circpad: bool = False, ops=None) -> np.ndarray:
        stim = np.asarray(stim)
        nt, ndim = stim.shape  # a 1-D stimulus raises here, as in the reference
        delays = [int(d) for d in delays]
        if not delays:
            raise ValueError("need at least one array to concatenate")  # np.hstack([]) in the reference
        if ops is None:
            from .device import default_ops

            ops = default_ops()
        out = ops.fir_make_delayed(stim, delays, bool(circpad))  # float64 (nt, ndim * ndelays)
        if all(d == 0 for d in delays) and stim.dtype != np.float64:
            # every block is stim.copy(): hstack keeps the input dtype
            out = out.astype(stim.dtype)
        return out

    def n_delays(self) -> int:
        """Number of delays used."""
        return len(self.delays) if self.delays is not None else 0

    def output_dim(self, input_dim: int) -> int:
        """Output dimensionality after expansion."""
        return input_dim * self.n_delays()

    def valid_length(self, nt: int) -> int:
        """Number of time points that contain no padding (nt with circular padding)."""
        if self.delays is None:
            raise ValueError("delays must be provided")
        if self.circpad:
            return nt
        return max(0, nt - max(abs(d) for d in self.delays))

    def summary(self, input_dim: Optional[int] = None, nt: Optional[int] = None) -> str:
        """Readable summary of the configuration."""
        msg = f"FIR(delays={list(self.delays)}, circpad={self.circpad})"
        if input_dim is not None:
            msg += f"\n- Output dim: {self.output_dim(input_dim)}"
        if nt is not None:
            msg += f"\n- Valid length: {self.valid_length(nt)}"
        return msg
