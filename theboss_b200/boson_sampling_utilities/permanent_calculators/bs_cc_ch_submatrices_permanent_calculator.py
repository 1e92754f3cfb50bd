"""Clifford & Clifford submatrices calculator, Chin-Huh form, on the B200.

Drop-in for ``BSCCCHSubmatricesPermanentCalculator``
(theboss/boson_sampling_utilities/permanent_calculators/bs_cc_ch_submatrices_permanent_calculator.py:27-105).
The reference spends O(m * #outputs) per Guan term; kernel K3 obtains all column-deleted products of
a term from prefix x suffix products in O(k).
"""
from ... import _native
from .bs_submatrices_permanent_calculator_base import BSSubmatricesPermanentCalculatorBase


class BSCCCHSubmatricesPermanentCalculator(BSSubmatricesPermanentCalculatorBase):
    _formula = _native.FORMULA_CHIN_HUH
