"""Gray-code Glynn permanent calculator on the B200.

Drop-in for ``GlynnGrayPermanentCalculator``
(theboss/boson_sampling_utilities/permanent_calculators/glynn_gray_permanent_calculator.py:27-84).
The effective scattering matrix (rows repeated by the output occupation, columns by the input
occupation, boson_sampling_utilities.py:595-626) is expanded on the device and its 2^(N-1) Gray
steps are evaluated by kernel K1 (theboss_b200/csrc/glynn_kernel.cu).
"""
import numpy as np

from ... import _native
from .bs_permanent_calculator_base import BSPermanentCalculatorBase


class GlynnGrayPermanentCalculator(BSPermanentCalculatorBase):
    _formula = _native.FORMULA_GLYNN

    def compute_permanent(self) -> np.complex128:
        U, s, t = self._device_operands()
        # Like the reference (:48-53) no shape validation happens here: an empty effective matrix
        # (no particles on either side) gives 1.
        return np.complex128(self._handle().glynn_single(U, s, t))

    def compute_permanent_of_matrix(self, A) -> np.complex128:
        """perm(A) of an explicit square matrix (what the reference computes once A is built)."""
        return np.complex128(self._handle().glynn_matrix(A))
