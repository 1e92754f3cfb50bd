"""The one-method protocol of every sampler.

Same contract as the reference's ``SimulationStrategyInterface``
(theboss/simulation_strategies/simulation_strategy_interface.py:7-23): ``simulate(input_state, samples_number=1)``
returns ``samples_number`` output occupations; any object with a callable ``simulate`` counts as a strategy
(structural ``issubclass`` check, like the reference's ``__subclasshook__``).
"""
import abc
from typing import List, Sequence, Tuple


class SimulationStrategyInterface(abc.ABC):
    @abc.abstractmethod
    def simulate(self, input_state: Sequence[int], samples_number: int = 1) -> List[Tuple[int, ...]]:
        """Draw ``samples_number`` output states for the given input occupation."""

    @classmethod
    def __subclasshook__(cls, candidate):
        method = getattr(candidate, "simulate", None)
        return True if callable(method) else NotImplemented
