"""``ClassicPermanentCalculator`` of the reference factory, served by the GPU engine.

The reference class (theboss/boson_sampling_utilities/permanent_calculators/
classic_permanent_calculator.py:19-67) is an O(n!) Laplace expansion used as ground truth for n <= 4 in
its tests.  Here the name maps onto kernel K2 (same quantity, O(n 2^n)); the O(n!) recursion itself lives
in oracle/bossperm_oracle.c (orc_classic) as a test oracle only.  The reference's edge cases are kept:
no input particles gives 1 if the output is empty too, else 0 (:28-33).
"""
import numpy as np

from ... import _native
from .bs_permanent_calculator_base import BSPermanentCalculatorBase


class ClassicPermanentCalculator(BSPermanentCalculatorBase):
    _formula = _native.FORMULA_GLYNN

    def compute_permanent(self) -> np.complex128:
        if sum(self.input_state) == 0:
            return np.complex128(1) if sum(self.output_state) == 0 else np.complex128(0)
        return self._multiplicity_permanent()
