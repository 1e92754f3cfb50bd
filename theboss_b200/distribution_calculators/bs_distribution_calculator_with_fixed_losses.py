"""Exact boson-sampling output distribution with a fixed number of lost particles, on the B200.

Drop-in for ``BSDistributionCalculatorWithFixedLosses``
(theboss/distribution_calculators/bs_distribution_calculator_with_fixed_losses.py:30-169).  The reference
sets (input, output) on its permanent calculator and calls ``compute_permanent()`` once per
(outcome, lossy input) pair (:119-137); here ALL pairs of a request go through ONE batched launch of kernel
K2 (``bp_perm_batched``) using the matrix held by the injected calculator, and the normalisations of
:99-106 / :139-151 are applied on the host.
"""
import math
from copy import deepcopy
from typing import Iterable, List, Tuple

import numpy as np
from scipy.special import binom

from .. import _native
from ..boson_sampling_utilities.boson_sampling_utilities import (
    generate_lossy_n_particle_input_states,
    generate_possible_states,
)
from .bs_distribution_calculator_interface import (
    BosonSamplingExperimentConfiguration,
    BSDistributionCalculatorInterface,
)


class BSDistributionCalculatorWithFixedLosses(BSDistributionCalculatorInterface):
    def __init__(self, configuration: BosonSamplingExperimentConfiguration, permanent_calculator) -> None:
        # a snapshot, like the reference (:41): its tests go on mutating the configuration they passed in
        # (tests/gcc_based_strategies_tests_base.py:84-91) and the calculator must keep describing the original experiment
        self.configuration = deepcopy(configuration)
        self._permanent_calculator = permanent_calculator

    @property
    def permanent_calculator(self):
        return self._permanent_calculator

    @permanent_calculator.setter
    def permanent_calculator(self, permanent_calculator) -> None:
        self._permanent_calculator = permanent_calculator

    def get_outcomes_in_proper_order(self) -> List[Tuple[int, ...]]:
        return generate_possible_states(self.configuration.number_of_particles_left, self.configuration.number_of_modes)

    def calculate_distribution(self) -> List[float]:
        return self.calculate_probabilities_of_outcomes(self.get_outcomes_in_proper_order())

    def _batched_probabilities(self, outcomes: List[Tuple[int, ...]], particles_left: int) -> List[float]:
        """Probabilities of ``outcomes`` (all holding ``particles_left`` particles) given that exactly that many
        input particles survived."""
        cfg = self.configuration
        n, m = cfg.initial_number_of_particles, cfg.number_of_modes
        lossy_inputs = generate_lossy_n_particle_input_states(cfg.initial_state, particles_left)
        multiplicity = [int(np.prod([binom(cfg.initial_state[i], cfg.initial_state[i] - li[i]) for i in range(m)]))
                        for li in lossy_inputs]
        self._permanent_calculator.matrix = cfg.interferometer_matrix      # like the reference (:133)
        U = _native.as_matrix(cfg.interferometer_matrix)
        S = np.zeros((len(outcomes) * len(lossy_inputs), U.shape[0]), dtype=np.uint8)
        T = np.zeros_like(S)
        S[:, :m] = np.tile(np.array(lossy_inputs, dtype=np.uint8).reshape(len(lossy_inputs), m), (len(outcomes), 1))
        T[:, :m] = np.repeat(np.array(outcomes, dtype=np.uint8).reshape(len(outcomes), m), len(lossy_inputs), axis=0)
        device = getattr(self._permanent_calculator, "device", 0)
        perms = _native.default_handle(device).perm_batched(U, S, T)
        out = []
        for k, outcome in enumerate(outcomes):
            p = 0
            for j, li in enumerate(lossy_inputs):
                sub = abs(perms[k * len(lossy_inputs) + j]) ** 2
                for occ in li:
                    sub /= math.factorial(occ)
                p += sub * multiplicity[j]
            p /= math.factorial(particles_left)
            p /= binom(n, particles_left)
            p *= math.factorial(particles_left)           # particle-basis multiplicity of the outcome (:99-104)
            for occ in outcome:
                p /= math.factorial(int(occ))
            out.append(float(p))
        return out

    def calculate_probabilities_of_outcomes(self, outcomes: Iterable[Iterable[int]]) -> List[float]:
        outcomes = [tuple(int(v) for v in o) for o in outcomes]
        return self._batched_probabilities(outcomes, self.configuration.number_of_particles_left)
