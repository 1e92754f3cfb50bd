"""Generalized Clifford & Clifford sampler (version A) with uniform losses, on the B200.

Drop-in for ``GeneralizedCliffordsUniformLossesSimulationStrategy``
(theboss/simulation_strategies/generalized_cliffords_uniform_losses_simulation_strategy.py:19-173): every input
particle survives with probability ``transmissivity`` (one stdlib ``random.random()`` per particle, :133-135 --
the reference's third RNG entry point), surviving particles are placed by the version-A
chain rule, and the exact (un)weighted probabilities of the visited outcomes are recorded on the way (:140-173).
Each new layer of pmfs is one batched launch of kernel K2, like the lossless strategy.
"""
import random as _stdlib_random
from math import factorial
from typing import List

import numpy as np
from scipy import special

from ..boson_sampling_utilities.boson_sampling_utilities import generate_possible_states
from .generalized_cliffords_simulation_strategy import GeneralizedCliffordsSimulationStrategy


class GeneralizedCliffordsUniformLossesSimulationStrategy(GeneralizedCliffordsSimulationStrategy):
    def __init__(self, bs_permanent_calculator, transmissivity: float = 0) -> None:
        super().__init__(bs_permanent_calculator)
        self._transmissivity = transmissivity
        self.distribution: List[float] = []
        self.unweighted_distribution: List[float] = []
        self._possible_outputs = []
        self._outcome_index = {}
        self._binomial_weights: List[float] = []
        self.missing_values_in_distribution = False

    def _initialize_simulation(self, input_state) -> None:
        self.input_state = input_state
        self.number_of_input_photons = int(sum(input_state))
        self._prepare_substates()
        self.pmfs = {}
        init = -1 if self.missing_values_in_distribution else 0
        n, eta = self.number_of_input_photons, self._transmissivity
        self._possible_outputs = generate_possible_states(n, len(input_state), losses=True)
        self._outcome_index = {o: i for i, o in enumerate(self._possible_outputs)}
        self.distribution = [init for _ in self._possible_outputs]
        self.unweighted_distribution = [init for _ in self._possible_outputs]
        self._binomial_weights = [pow(eta, left) * special.binom(n, left) * pow(1 - eta, n - left) for left in range(n + 1)]
        self.distribution[0] = self._binomial_weights[0]

    def simulate(self, input_state, samples_number: int = 1) -> List[np.ndarray]:
        self._initialize_simulation(input_state)
        samples = []
        while len(samples) < samples_number:
            self._fill_r_sample()
            samples.append(np.array(self.r_sample, dtype=np.int64))
        return samples

    def compute_distribution_up_to_accuracy(self, input_state, accuracy: float = 1.0) -> List[float]:
        self._initialize_simulation(input_state)
        while not np.isclose(max(accuracy - sum(self.distribution), 0), 0):
            self._fill_r_sample()
        return self.distribution

    def compute_unweighted_distribution_up_to_accuracy(self, input_state, accuracy: float = 1.0) -> List[float]:
        self._initialize_simulation(input_state)
        while not np.isclose(max(accuracy - sum(self.unweighted_distribution) / sum(input_state), 0), 0):
            self._fill_r_sample()
        return self.unweighted_distribution

    def _record_layer(self, r_sample, pmf) -> None:
        """Exact probabilities of the outcomes r_sample + e_j that this layer touches (:157-171)."""
        for j, p in enumerate(pmf):
            output = list(r_sample)
            output[j] += 1
            i = self._outcome_index.get(tuple(output))
            if i is None:
                continue
            value = p * factorial(sum(output))
            for occ in output:
                value /= factorial(occ)
            self.unweighted_distribution[i] = value
            self.distribution[i] = value * self._binomial_weights[sum(output)]

    def _fill_r_sample(self) -> None:
        self.r_sample = [0 for _ in self.input_state]
        for _ in range(self.number_of_input_photons):
            if _stdlib_random.random() >= self._transmissivity:
                continue
            pmf, new = self._pmf_for(tuple(self.r_sample))
            if new:
                self._record_layer(self.r_sample, pmf)
            threshold = np.random.random() * sum(pmf)
            running, index = 0, 0
            for p in pmf:
                running += p
                if running > threshold:
                    break
                index += 1
            self.r_sample[index] += 1
