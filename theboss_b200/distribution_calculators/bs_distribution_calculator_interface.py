"""Experiment description and the calculator interface used by the exact-distribution classes.

API surface mirrored from the reference (theboss/distribution_calculators/bs_distribution_calculator_interface.py:11-52):
``BosonSamplingExperimentConfiguration`` is a plain record with the same field names, order and defaults, so code
that builds it by keyword or position keeps working; ``BSDistributionCalculatorInterface`` names the three methods
every distribution calculator offers.
"""
import abc
import dataclasses
from typing import Any, List, Sequence, Tuple

# (name, type, default) -- a field without default is required
_FIELDS = (
    ("interferometer_matrix", Sequence[Sequence[complex]]),          # m x m, possibly lossy
    ("initial_state", Sequence[int]),                                 # input occupation
    ("initial_number_of_particles", int),
    ("number_of_modes", int),
    ("number_of_particles_lost", int),
    ("number_of_particles_left", int),
    ("uniform_transmissivity", float, dataclasses.field(default=1)),
    ("network_simulation_strategy", Any, dataclasses.field(default=None)),
    ("hierarchy_level", int, dataclasses.field(default=0)),          # k of the Brod-Oszmaniec papers
)

BosonSamplingExperimentConfiguration = dataclasses.make_dataclass("BosonSamplingExperimentConfiguration", _FIELDS)
BosonSamplingExperimentConfiguration.__module__ = __name__          # picklable / deep-copyable like a class statement
BosonSamplingExperimentConfiguration.__doc__ = "Settings of one (lossy) boson-sampling experiment."


class BSDistributionCalculatorInterface(abc.ABC):
    """What every exact / sample-based distribution calculator provides."""

    @abc.abstractmethod
    def get_outcomes_in_proper_order(self) -> List[Tuple[int, ...]]:
        """Output Fock states, in the order the distribution vector uses."""

    @abc.abstractmethod
    def calculate_probabilities_of_outcomes(self, outcomes) -> List[float]:
        """Probabilities of the given output states."""

    @abc.abstractmethod
    def calculate_distribution(self) -> List[float]:
        """Probabilities of all outcomes of ``get_outcomes_in_proper_order``."""
