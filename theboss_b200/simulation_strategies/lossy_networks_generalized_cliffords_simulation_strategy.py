"""GCC-B for networks with non-uniform (mode-dependent) losses on the B200.

Drop-in for ``LossyNetworksGeneralizedCliffordsSimulationStrategy``
(theboss/simulation_strategies/lossy_networks_generalized_cliffords_simulation_strategy.py:22-88): the
lossy m x m matrix held by the calculator is replaced by its 2m x 2m dilation (:41-44), GCC-B runs
on the input padded with m empty modes (:69-74) and the outputs are trimmed back to m modes (:78-81).
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np

from ..boson_sampling_utilities.boson_sampling_utilities import prepare_interferometer_matrix_in_expanded_space
from .generalized_cliffords_b_simulation_strategy import GeneralizedCliffordsBSimulationStrategy
from .simulation_strategy_interface import SimulationStrategyInterface


class LossyNetworksGeneralizedCliffordsSimulationStrategy(SimulationStrategyInterface):
    def __init__(self, bs_permanent_calculator, rng_mode: str = "philox", device: Optional[int] = None) -> None:
        bs_permanent_calculator.matrix = prepare_interferometer_matrix_in_expanded_space(bs_permanent_calculator.matrix)
        self._helper_strategy = GeneralizedCliffordsBSimulationStrategy(bs_permanent_calculator, rng_mode=rng_mode,
                                                                        device=device)

    def simulate(self, input_state: Sequence[int], samples_number: int = 1,
                 decision_tape: Optional[np.ndarray] = None) -> List[Tuple[int, ...]]:
        m = len(input_state)
        expanded = np.concatenate([np.asarray(input_state).astype(np.int64), np.zeros(m, dtype=np.int64)])
        samples = self._helper_strategy.simulate(expanded, samples_number, decision_tape=decision_tape)
        return [tuple(sample[:m]) for sample in samples]

    def set_new_matrix(self, matrix) -> None:
        self._helper_strategy.set_new_matrix(prepare_interferometer_matrix_in_expanded_space(matrix))
