"""Interface of the submatrices ("all minors") permanent calculators; mirrors
``BSSubmatricesPermanentCalculatorInterface`` of the reference
(theboss/boson_sampling_utilities/permanent_calculators/bs_submatrices_permanent_calculator_interface.py:12-50)."""
import abc
from typing import List, Sequence


class BSSubmatricesPermanentCalculatorInterface(abc.ABC):
    """``compute_permanents()[i]`` = permanent of the effective scattering matrix with one particle
    removed from input mode i (0 where that mode is empty)."""

    @abc.abstractmethod
    def compute_permanents(self) -> List[complex]:
        ...

    matrix: Sequence[Sequence[complex]] = abc.abstractproperty()
    input_state: Sequence[int] = abc.abstractproperty()
    output_state: Sequence[int] = abc.abstractproperty()
