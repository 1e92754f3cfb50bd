"""Exact output distribution under uniform losses, on the B200.

Drop-in for ``BSDistributionCalculatorWithUniformLosses``
(theboss/distribution_calculators/bs_exact_distribution_with_uniform_losses.py:23-104): outcomes with l surviving
particles are weighted by C(n,l) eta^l (1-eta)^(n-l) (:43-60).  The reference fans the outcomes out over a
``multiprocessing.Pool`` (:62-71); here outcomes are grouped by particle number and every group is one batched
K2 launch.
"""
from typing import Iterable, List, Tuple

from scipy import special

from ..boson_sampling_utilities.boson_sampling_utilities import generate_possible_states
from .bs_distribution_calculator_interface import BosonSamplingExperimentConfiguration
from .bs_distribution_calculator_with_fixed_losses import BSDistributionCalculatorWithFixedLosses

__all__ = ["BSDistributionCalculatorWithUniformLosses", "BSDistributionCalculatorWithFixedLosses",
           "BosonSamplingExperimentConfiguration"]


class BSDistributionCalculatorWithUniformLosses(BSDistributionCalculatorWithFixedLosses):
    def __init__(self, configuration: BosonSamplingExperimentConfiguration, permanent_calculator) -> None:
        super().__init__(configuration, permanent_calculator)
        self.weights = self._initialize_weights()
        self.weightless = False

    def set_weightless(self, weightless: bool) -> None:
        self.weights = [1 for _ in self.weights] if weightless else self._initialize_weights()
        self.weightless = weightless

    def _initialize_weights(self) -> List[float]:
        n, eta = self.configuration.initial_number_of_particles, self.configuration.uniform_transmissivity
        return [pow(eta, l) * special.binom(n, l) * pow(1.0 - eta, n - l) for l in range(n + 1)]

    def get_outcomes_in_proper_order(self) -> List[Tuple[int, ...]]:
        return generate_possible_states(self.configuration.initial_number_of_particles,
                                        self.configuration.number_of_modes, losses=True)

    def calculate_probabilities_of_outcomes(self, outcomes: Iterable[Iterable[int]]) -> List[float]:
        outcomes = [tuple(int(v) for v in o) for o in outcomes]
        result = [0.0] * len(outcomes)
        by_l = {}
        for idx, o in enumerate(outcomes):
            by_l.setdefault(sum(o), []).append(idx)
        for l, idxs in by_l.items():
            if l == 0:
                for i in idxs:
                    result[i] = float(self.weights[0])
                continue
            probs = self._batched_probabilities([outcomes[i] for i in idxs], l)
            for i, p in zip(idxs, probs):
                result[i] = p * self.weights[l]
        return result
