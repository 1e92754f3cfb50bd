"""GCC-B with uniform losses on the B200.

Drop-in for ``GeneralizedCliffordsBUniformLossesSimulationStrategy``
(theboss/simulation_strategies/generalized_cliffords_b_uniform_losses_simulation_strategy.py:23-121):
the number of surviving particles l is drawn from Binomial(n, transmissivity) by inverse CDF (:50-85),
then l steps of GCC-B run on the LOSSLESS matrix.  Returns int64 ndarrays like the reference (:108).
"""
from typing import List, Optional

import numpy as np
from scipy.special import binom

from .generalized_cliffords_b_simulation_strategy import GeneralizedCliffordsBSimulationStrategy


class GeneralizedCliffordsBUniformLossesSimulationStrategy(GeneralizedCliffordsBSimulationStrategy):
    def __init__(self, bs_permanent_calculator, transmissivity: float = 1.0, rng_mode: str = "philox",
                 device: Optional[int] = None) -> None:
        super().__init__(bs_permanent_calculator, rng_mode=rng_mode, device=device)
        self._transmissivity = float(transmissivity)

    def _eta(self) -> float:
        return self._transmissivity

    def _uniform_losses_weights(self):
        n, eta = self.number_of_input_photons, self._transmissivity
        return [binom(n, l) * pow(eta, l) * pow(1 - eta, n - l) for l in range(n + 1)]

    def simulate(self, input_state, samples_number: int = 1,
                 decision_tape: Optional[np.ndarray] = None) -> List[np.ndarray]:
        out = self._run(input_state, samples_number, decision_tape)
        return list(out.astype(np.int64))   # one int64 row per sample, like the reference (:108)
