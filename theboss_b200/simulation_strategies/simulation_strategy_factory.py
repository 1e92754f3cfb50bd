"""Factory of the GPU-backed simulation strategies.

Mirrors ``StrategyType`` / ``SimulationStrategyFactory`` of the reference
(theboss/simulation_strategies/simulation_strategy_factory.py:36-226) for the strategies on the permanent
hot path.  The enum keeps every reference member (so values match), but only the GCC family is built
here (GCC, lossy-network GCC-B, uniform-loss GCC and both BOBS variants); the mean-field and R-backed strategies
never evaluate a permanent and are outside this package's scope (SURVEY.md section 8 marks them out of scope): for them the
factory raises ``NotImplementedError`` naming the reference class to use.
Like the reference the factory deep-copies the calculator it is given (:58, :107, :186).
"""
import enum
from copy import deepcopy

from .generalized_cliffords_simulation_strategy import GeneralizedCliffordsSimulationStrategy
from .generalized_cliffords_uniform_losses_simulation_strategy import (
    GeneralizedCliffordsUniformLossesSimulationStrategy,
)
from .lossy_networks_generalized_cliffords_simulation_strategy import (
    LossyNetworksGeneralizedCliffordsSimulationStrategy,
)
from .lossy_state_approximated_simulation_strategy import LossyStateApproximationSimulationStrategy
from .nonuniform_losses_approximation_strategy import NonuniformLossesApproximationStrategy
from .simulation_strategy_interface import SimulationStrategyInterface


class StrategyType(enum.IntEnum):
    FIXED_LOSS = enum.auto()
    UNIFORM_LOSS = enum.auto()
    CLIFFORD_R = enum.auto()
    GCC = enum.auto()
    LOSSY_NET_GCC = enum.auto()
    LOSSLESS_MODES_STRATEGY = enum.auto()
    UNIFORM_LOSSES_GCC = enum.auto()
    BOBS = enum.auto()
    UNIFORM_LOSSES_BOBS = enum.auto()


# Members that never evaluate a permanent (mean-field approximations, the R-backed sampler) are not rebuilt here and there is
# no CPU path in this package: the factory raises NotImplementedError naming the reference class to use instead.
_OUT_OF_SCOPE = {
    StrategyType.FIXED_LOSS: "theboss.simulation_strategies.fixed_loss_simulation_strategy.FixedLossSimulationStrategy",
    StrategyType.UNIFORM_LOSS: "theboss.simulation_strategies.uniform_loss_simulation_strategy.UniformLossSimulationStrategy",
    StrategyType.CLIFFORD_R: "theboss.simulation_strategies.cliffords_r_simulation_strategy.CliffordsRSimulationStrategy",
    # the reference maps no builder to this member and falls back to the fixed-loss strategy (:116-119)
    StrategyType.LOSSLESS_MODES_STRATEGY: "theboss.simulation_strategies.fixed_loss_simulation_strategy.FixedLossSimulationStrategy",
}


class SimulationStrategyFactory:
    def __init__(self, experiment_configuration, bs_permanent_calculator,
                 strategy_type: StrategyType = StrategyType.FIXED_LOSS) -> None:   # same default as the reference (:52)
        self.experiment_configuration = experiment_configuration
        self.strategy_type = strategy_type
        self._bs_permanent_calculator = deepcopy(bs_permanent_calculator)
        # :74, :79-85 of the reference: size of the BOBS process pool.  Kept so that callers can set it; the samples of a
        # request run as one GPU batch here, so the value is not used.
        self.available_threads_number = -1

    @property
    def bs_permanent_calculator(self):
        return self._bs_permanent_calculator

    @bs_permanent_calculator.setter
    def bs_permanent_calculator(self, bs_permanent_calculator) -> None:
        self._bs_permanent_calculator = deepcopy(bs_permanent_calculator)

    def generate_strategy(self) -> SimulationStrategyInterface:
        kind = self.strategy_type
        calc = deepcopy(self._bs_permanent_calculator)
        if kind == StrategyType.GCC:
            return GeneralizedCliffordsSimulationStrategy(calc)
        if kind == StrategyType.LOSSY_NET_GCC:
            return LossyNetworksGeneralizedCliffordsSimulationStrategy(calc)
        if kind == StrategyType.UNIFORM_LOSSES_GCC:
            # simulation_strategy_factory.py:175-181 of the reference: version A with per-particle Bernoulli losses
            return GeneralizedCliffordsUniformLossesSimulationStrategy(
                calc, getattr(self.experiment_configuration, "uniform_transmissivity", 1.0))
        cfg = self.experiment_configuration
        if kind == StrategyType.BOBS:            # simulation_strategy_factory.py:183-192 of the reference
            return NonuniformLossesApproximationStrategy(calc, cfg.number_of_modes - cfg.hierarchy_level)
        if kind == StrategyType.UNIFORM_LOSSES_BOBS:   # :194-205
            return LossyStateApproximationSimulationStrategy(calc, cfg.uniform_transmissivity, cfg.hierarchy_level)
        raise NotImplementedError(
            f"{kind.name} is outside the permanent hot path built here; use the reference class "
            f"{_OUT_OF_SCOPE.get(kind, _OUT_OF_SCOPE[StrategyType.FIXED_LOSS])}")
