"""Chin-Huh permanent calculator (generalised Glynn with mode multiplicities) on the B200.

Drop-in for ``ChinHuhPermanentCalculator``
(theboss/boson_sampling_utilities/permanent_calculators/chin_huh_permanent_calculator.py:19-59 with the
Guan-code driver of bs_permanent_calculator_base.py:166-209).  One item through kernel K2
(theboss_b200/csrc/guan_kernel.cu), which walks the Guan code of the cheaper side and exploits the
r <-> s - r symmetry the reference leaves unused.
"""
import numpy as np

from ... import _native
from .bs_permanent_calculator_base import BSPermanentCalculatorBase


class ChinHuhPermanentCalculator(BSPermanentCalculatorBase):
    _formula = _native.FORMULA_CHIN_HUH

    def compute_permanent(self) -> np.complex128:
        if not self._can_calculation_be_performed():
            raise AttributeError   # bs_permanent_calculator_base.py:179-180
        return self._multiplicity_permanent()
