"""Entry point of a sampling experiment, kept for callers of the reference's ``BosonSamplingSimulator``
(theboss/boson_sampling_simulator.py:19-26 there): wraps any object with a ``simulate(input_state, samples_number)``
method -- here one of the GPU-backed strategies of ``theboss_b200.simulation_strategies`` -- and forwards requests to it.
``get_outcome_frequencies`` adds the tally the reference's TODO asks for."""
from collections import Counter
from typing import Dict, List, Sequence, Tuple

from .simulation_strategies.simulation_strategy_interface import SimulationStrategyInterface

Outcome = Tuple[int, ...]


class BosonSamplingSimulator:
    __slots__ = ("_simulation_strategy",)

    def __init__(self, simulation_strategy: SimulationStrategyInterface) -> None:
        if not callable(getattr(simulation_strategy, "simulate", None)):
            raise TypeError("a simulation strategy must provide simulate(input_state, samples_number)")
        self._simulation_strategy = simulation_strategy

    @property
    def simulation_strategy(self) -> SimulationStrategyInterface:
        return self._simulation_strategy

    def get_classical_simulation_results(self, input_state: Sequence[int], samples_number: int = 1) -> List[Outcome]:
        """``samples_number`` output occupations for the given input occupation, in the strategy's own container types."""
        return self._simulation_strategy.simulate(input_state, samples_number)

    def get_outcome_frequencies(self, input_state: Sequence[int], samples_number: int) -> Dict[Outcome, float]:
        """Relative frequency of every outcome that occurred in ``samples_number`` fresh samples."""
        samples = self.get_classical_simulation_results(input_state, samples_number)
        tally = Counter(tuple(int(v) for v in sample) for sample in samples)
        return {outcome: count / len(samples) for outcome, count in tally.items()} if samples else {}
