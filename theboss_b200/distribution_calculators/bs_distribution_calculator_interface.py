"""Configuration dataclass + interface of the distribution calculators; mirrors
``theboss/distribution_calculators/bs_distribution_calculator_interface.py:11-52``."""
import abc
from dataclasses import dataclass
from typing import Any, List, Sequence, Tuple


@dataclass
class BosonSamplingExperimentConfiguration:
    interferometer_matrix: Sequence[Sequence[complex]]
    initial_state: Sequence[int]
    initial_number_of_particles: int
    number_of_modes: int
    number_of_particles_lost: int
    number_of_particles_left: int
    uniform_transmissivity: float = 1
    network_simulation_strategy: Any = None
    hierarchy_level: int = 0


class BSDistributionCalculatorInterface(abc.ABC):
    @abc.abstractmethod
    def calculate_distribution(self) -> List[float]:
        ...

    @abc.abstractmethod
    def calculate_probabilities_of_outcomes(self, outcomes) -> List[float]:
        ...

    @abc.abstractmethod
    def get_outcomes_in_proper_order(self) -> List[Tuple[int, ...]]:
        ...
