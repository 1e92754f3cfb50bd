"""``SimulationStrategyInterface`` of the reference
(theboss/simulation_strategies/simulation_strategy_interface.py:7-23): anything with a callable
``simulate(input_state, samples_number=1)`` returning a list of output occupations."""
import abc
from typing import List, Sequence, Tuple


class SimulationStrategyInterface(abc.ABC):
    @classmethod
    def __subclasshook__(cls, subclass):
        return callable(getattr(subclass, "simulate", None))

    @abc.abstractmethod
    def simulate(self, input_state: Sequence[int], samples_number: int = 1) -> List[Tuple[int, ...]]:
        ...
