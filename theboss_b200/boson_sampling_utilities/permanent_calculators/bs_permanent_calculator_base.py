"""Shared plumbing of the GPU-backed single-permanent calculators.

API parity with the reference base class
(theboss/boson_sampling_utilities/permanent_calculators/bs_permanent_calculator_base.py:22-72):

* ctor ``(matrix, input_state=None, output_state=None)``; the three properties hand back the very
  object they were given, and the matrix is re-read at every ``compute_permanent`` call because
  reference callers mutate it in place (``calculator.matrix *= sqrt(eta)``).
* shape check of :61-72, ``AttributeError`` on mismatch (:179-180).
* instances are deep-copied by the strategy factory and pickled into worker processes by the BOBS
  strategies, so the native handle is never part of the object state: it is looked up lazily per
  (process, device).

Extra, GPU-only keyword: ``device`` (CUDA ordinal, default 0).
"""
from typing import Optional, Sequence

import numpy as np

from ... import _native
from .bs_permanent_calculator_interface import BSPermanentCalculatorInterface


class BSPermanentCalculatorBase(BSPermanentCalculatorInterface):
    #: selector forwarded to the C ABI (include/bossperm.h BP_FORMULA_*)
    _formula = _native.FORMULA_GLYNN

    def __init__(self, matrix, input_state: Optional[Sequence[int]] = None,
                 output_state: Optional[Sequence[int]] = None, device: int = 0) -> None:
        self._matrix = matrix
        self._input_state = [] if input_state is None else input_state
        self._output_state = [] if output_state is None else output_state
        self._device = int(device)

    # -- the reference's three read/write properties ---------------------------------------------
    @property
    def matrix(self):
        return self._matrix

    @matrix.setter
    def matrix(self, matrix) -> None:
        self._matrix = matrix

    @property
    def input_state(self):
        return self._input_state

    @input_state.setter
    def input_state(self, input_state) -> None:
        self._input_state = input_state

    @property
    def output_state(self):
        return self._output_state

    @output_state.setter
    def output_state(self, output_state) -> None:
        self._output_state = output_state

    @property
    def device(self) -> int:
        return self._device

    # -- helpers -----------------------------------------------------------------------------------
    def _handle(self) -> "_native.Handle":
        return _native.default_handle(self._device)

    def _can_calculation_be_performed(self) -> bool:
        m = self._matrix
        return (len(m) == len(m[0]) and len(self._output_state) == len(self._input_state)
                and len(self._output_state) <= len(m[0]))

    def _device_operands(self):
        """(U complex128 C-contiguous, s int32[m], t int32[m]) for the C ABI."""
        U = _native.as_matrix(self._matrix)
        m = U.shape[0]
        return U, _native.as_state(self._input_state, m), _native.as_state(self._output_state, m)

    def _multiplicity_permanent(self) -> np.complex128:
        """One item through the batched multiplicity kernel (K2)."""
        U, s, t = self._device_operands()
        if int(s.sum()) != int(t.sum()):
            raise AttributeError("input and output states hold different particle numbers")
        out = self._handle().perm_batched(U, s[None, :], t[None, :], self._formula)
        return np.complex128(out[0])


# The reference derives Ryser / Chin-Huh from a second base that adds the Guan-code driver
# (bs_permanent_calculator_base.py:75-209); the walk lives in kernel K2 here, so both names denote the same class.
BSGuanCodeBasedPermanentCalculatorBase = BSPermanentCalculatorBase
