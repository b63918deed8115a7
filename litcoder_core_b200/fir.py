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

    # ---- bookkeeping helpers of the reference class (FIR_expander.py:45-73): same names, same answers ----
    def _shifts(self):
        if self.delays is None:
            raise ValueError("delays must be provided")
        return [int(d) for d in self.delays]

    def n_delays(self) -> int:
        """How many shifted copies `expand` stacks (0 for an instance built without delays)."""
        return 0 if self.delays is None else len(self.delays)

    def output_dim(self, input_dim: int) -> int:
        """Width of the expanded matrix for an `input_dim`-wide stimulus."""
        return self.n_delays() * input_dim

    def valid_length(self, nt: int) -> int:
        """Rows of an nt-row expansion that hold no zero padding: all of them with circular padding, otherwise nt
        minus the largest shift in either direction (never negative)."""
        shifts = self._shifts()
        return nt if self.circpad else max(nt - max(abs(d) for d in shifts), 0)

    def summary(self, input_dim: Optional[int] = None, nt: Optional[int] = None) -> str:
        """One line naming the configuration, plus the output width / unpadded length when asked for."""
        lines = [f"FIR(delays={list(self.delays)}, circpad={self.circpad})"]
        if input_dim is not None:
            lines.append(f"- Output dim: {self.output_dim(input_dim)}")
        if nt is not None:
            lines.append(f"- Valid length: {self.valid_length(nt)}")
        return "\n".join(lines)
