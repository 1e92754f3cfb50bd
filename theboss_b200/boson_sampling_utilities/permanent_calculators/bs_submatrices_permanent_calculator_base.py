"""GPU-backed base of the submatrices permanent calculators.

API parity with ``BSSubmatricesPermanentCalculatorBase`` /
``BSGuanBasedSubmatricesPermanentCalculatorBase`` of the reference
(theboss/boson_sampling_utilities/permanent_calculators/bs_submatrices_permanent_calculator_base.py:21-189):
ctor ``(matrix, input_state=None, output_state=None)``, the three read/write properties,
``compute_permanents() -> list of m numpy.complex128`` with sum(input) = k, sum(output) = k - 1 and
the k == 1 shortcut returning the occupations as complex numbers (:157-158).  One launch of kernel K3
(theboss_b200/csrc/minors_kernel.cu) replaces the reference's Guan sweep.
"""
from typing import List, Optional, Sequence

import numpy as np

from ... import _native
from .bs_submatrices_permanent_calculator_interface import BSSubmatricesPermanentCalculatorInterface


class BSSubmatricesPermanentCalculatorBase(BSSubmatricesPermanentCalculatorInterface):
    _formula = _native.FORMULA_CHIN_HUH

    def __init__(self, matrix, input_state: Optional[Sequence[int]] = None,
                 output_state: Optional[Sequence[int]] = None, device: int = 0) -> None:
        self._matrix = matrix
        self._input_state = [] if input_state is None else input_state
        self._output_state = [] if output_state is None else output_state
        self._device = int(device)

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

    def compute_permanents(self) -> List[np.complex128]:
        if sum(self.input_state) == 1:
            return [np.complex128(v) for v in self.input_state]
        U = _native.as_matrix(self._matrix)
        m = U.shape[0]
        s, t = _native.as_state(self._input_state, m), _native.as_state(self._output_state, m)
        out = _native.default_handle(self._device).minors(U, s, t, self._formula)
        return [np.complex128(v) for v in out[: len(self._input_state)]]


# The reference splits the boilerplate into two bases; both names are kept importable.
BSGuanBasedSubmatricesPermanentCalculatorBase = BSSubmatricesPermanentCalculatorBase
