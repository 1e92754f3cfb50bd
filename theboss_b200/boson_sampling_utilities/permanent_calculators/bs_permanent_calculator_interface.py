"""Interface of the single-permanent calculators.

Mirrors ``BSPermanentCalculatorInterface`` of the reference
(theboss/boson_sampling_utilities/permanent_calculators/bs_permanent_calculator_interface.py:18-52):
``compute_permanent()`` plus read/write ``matrix``, ``input_state``, ``output_state``.
"""
import abc
from typing import Sequence


class BSPermanentCalculatorInterface(abc.ABC):
    """Permanent of the effective scattering matrix of (matrix, input_state, output_state)."""

    @abc.abstractmethod
    def compute_permanent(self) -> complex:
        ...

    matrix: Sequence[Sequence[complex]] = abc.abstractproperty()
    input_state: Sequence[int] = abc.abstractproperty()
    output_state: Sequence[int] = abc.abstractproperty()
