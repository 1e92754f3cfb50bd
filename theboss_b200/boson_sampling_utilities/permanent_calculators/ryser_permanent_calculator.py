"""Ryser-named permanent calculator on the B200.

Drop-in for ``RyserPermanentCalculator``
(theboss/boson_sampling_utilities/permanent_calculators/ryser_permanent_calculator.py:19-64), the default
of the reference's factory and test-suite.  It returns the same mathematical quantity, but the device
engine evaluates it in Glynn/Chin-Huh form (kernel K2): Ryser-form float64 arithmetic loses ~1 bit per
photon to cancellation (the reference's own Ryser is 2e-10 off at n=20) and could
not meet the 1e-10 parity bar at the headline sizes.
"""
import numpy as np

from ... import _native
from .bs_permanent_calculator_base import BSPermanentCalculatorBase


class RyserPermanentCalculator(BSPermanentCalculatorBase):
    _formula = _native.FORMULA_RYSER

    def compute_permanent(self) -> np.complex128:
        if not self._can_calculation_be_performed():
            raise AttributeError   # bs_permanent_calculator_base.py:179-180
        return self._multiplicity_permanent()
