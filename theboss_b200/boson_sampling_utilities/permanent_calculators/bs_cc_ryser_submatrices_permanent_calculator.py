"""Clifford & Clifford submatrices calculator under its Ryser name, on the B200.

Drop-in for ``BSCCRyserSubmatricesPermanentCalculator``
(theboss/boson_sampling_utilities/permanent_calculators/bs_cc_ryser_submatrices_permanent_calculator.py:16-119),
the variant the GCC-B sampler is hard-wired to (generalized_cliffords_b_simulation_strategy.py:73-77).
Same outputs; the device engine (kernel K3) evaluates them in Glynn/Chin-Huh form over the output
particles, which needs 4x fewer terms than the reference's sweep over the input particles and does not
suffer Ryser-form cancellation in float64.
"""
from ... import _native
from .bs_submatrices_permanent_calculator_base import BSSubmatricesPermanentCalculatorBase


class BSCCRyserSubmatricesPermanentCalculator(BSSubmatricesPermanentCalculatorBase):
    _formula = _native.FORMULA_RYSER
