"""'Next' rows of SURVEY.md section 8(f) on the GPU path: exact distribution calculators on top of the batched
permanent kernel (f2), the version-A uniform-loss sampler (f3), and the reference's acceptance criterion for
samplers (empirical frequencies within the statistical TVD bound of the exact distribution,
quantum_computations_utilities.py:95-127 in the reference)."""
import json
import os

import numpy as np
import pytest

from tests import workloads

pytestmark = pytest.mark.gpu


def _classes():
    from theboss_b200.boson_sampling_utilities.permanent_calculators.bs_permanent_calculator_factory import (
        BSPermanentCalculatorFactory, PermanentCalculatorType)
    from theboss_b200.distribution_calculators.bs_exact_distribution_with_uniform_losses import (
        BosonSamplingExperimentConfiguration, BSDistributionCalculatorWithFixedLosses, BSDistributionCalculatorWithUniformLosses)
    return (BSPermanentCalculatorFactory, PermanentCalculatorType, BosonSamplingExperimentConfiguration,
            BSDistributionCalculatorWithFixedLosses, BSDistributionCalculatorWithUniformLosses)


def _config(U, s, lost=0, eta=1.0):
    _, _, Config, _, _ = _classes()
    return Config(interferometer_matrix=U, initial_state=list(s), initial_number_of_particles=int(sum(s)),
                  number_of_modes=len(s), number_of_particles_lost=lost, number_of_particles_left=int(sum(s)) - lost,
                  uniform_transmissivity=eta)


def _tvd_bound(n_outcomes, samples, delta=1e-3):
    return np.sqrt((-np.log(delta) + n_outcomes * np.log(2)) / (2 * samples))


def test_reference_known_answer_vector(golden_dir):
    """The reference's literal 56-entry distribution (tests/test_exact_distribution_calculator.py:75-142)."""
    Factory, Type, _, _, Uniform = _classes()
    with open(os.path.join(golden_dir, "exact_distribution.json")) as f:
        g = json.load(f)
    P = np.array(g["matrix_real"], dtype=np.complex128)
    calc = Factory(None, None, None, Type.CHIN_HUH).generate_calculator()
    dist = Uniform(_config(P, g["initial_state"], lost=2, eta=g["eta"]), calc).calculate_distribution()
    assert len(dist) == 56
    assert np.allclose(dist, g["reference_literal"])
    assert np.allclose(dist, g["reference_computed"], rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("state,lost", [([1, 1, 1, 1, 0], 0), ([3, 1, 2, 0], 0), ([3, 1, 2, 0], 2), ([1, 1, 1, 1, 0], 3)])
def test_distributions_sum_to_one(state, lost):
    """tests/test_distribution_calculators.py:22-99 and tests/test_exact_distribution_calculator.py:61-73."""
    Factory, Type, _, Fixed, Uniform = _classes()
    U = workloads.haar(len(state), 31)
    for typ in (Type.RYSER, Type.CHIN_HUH, Type.GLYNN):
        calc = Factory(None, None, None, typ).generate_calculator()
        assert abs(sum(Fixed(_config(U, state, lost), calc).calculate_distribution()) - 1) < 1e-10
        assert abs(sum(Uniform(_config(U, state, lost, eta=0.7), calc).calculate_distribution()) - 1) < 1e-10


def test_version_a_uniform_losses_sampler_against_exact_distribution():
    from theboss_b200.simulation_strategies.generalized_cliffords_uniform_losses_simulation_strategy import (
        GeneralizedCliffordsUniformLossesSimulationStrategy)
    Factory, Type, _, _, Uniform = _classes()
    import random
    U, s, eta, N = workloads.haar(4, 17), [1, 2, 0, 1], 0.6, 20000
    calc = Factory(U, None, None, Type.CHIN_HUH).generate_calculator()
    exact_calc = Uniform(_config(U, s, eta=eta), calc)
    outcomes, exact = exact_calc.get_outcomes_in_proper_order(), np.array(exact_calc.calculate_distribution())
    strat = GeneralizedCliffordsUniformLossesSimulationStrategy(calc, eta)
    random.seed(5); np.random.seed(5)
    samples = strat.simulate(s, N)
    assert isinstance(samples[0], np.ndarray) and samples[0].dtype == np.int64
    counts = {o: 0 for o in outcomes}
    for x in samples:
        counts[tuple(int(v) for v in x)] += 1
    freq = np.array([counts[o] / N for o in outcomes])
    assert 0.5 * np.abs(freq - exact).sum() <= _tvd_bound(len(outcomes), N)
    # the exact probabilities recorded on the way agree with the exact calculator
    dist = np.array(strat.compute_distribution_up_to_accuracy(s, 1.0))
    assert np.abs(dist - exact).max() <= 1e-12


def test_lossy_network_sampler_matches_uniform_loss_distribution():
    """A uniformly lossy matrix sqrt(eta) U through the dilated-network sampler reproduces the uniform-loss exact
    distribution (the reference checks the same in tests/gcc_based_strategies_tests_base.py:75-110)."""
    from theboss_b200.simulation_strategies.lossy_networks_generalized_cliffords_simulation_strategy import (
        LossyNetworksGeneralizedCliffordsSimulationStrategy)
    from theboss_b200.simulation_strategies.generalized_cliffords_b_uniform_losses_simulation_strategy import (
        GeneralizedCliffordsBUniformLossesSimulationStrategy)
    Factory, Type, _, _, Uniform = _classes()
    U, s, eta, N = workloads.haar(4, 23), [1, 1, 1, 0], 0.5, 20000
    calc = Factory(U.copy(), None, None).generate_calculator()
    exact_calc = Uniform(_config(U, s, eta=eta), calc)
    outcomes, exact = exact_calc.get_outcomes_in_proper_order(), np.array(exact_calc.calculate_distribution())
    bound = _tvd_bound(len(outcomes), N)
    np.random.seed(9)
    lossy_calc = Factory(U.copy(), None, None).generate_calculator()
    lossy_calc.matrix *= np.sqrt(eta)                      # in-place scaling, like the reference's tests
    for strat in (LossyNetworksGeneralizedCliffordsSimulationStrategy(lossy_calc),
                  GeneralizedCliffordsBUniformLossesSimulationStrategy(Factory(U.copy(), None, None).generate_calculator(), eta)):
        samples = strat.simulate(np.array(s), N)
        counts = {o: 0 for o in outcomes}
        for x in samples:
            counts[tuple(int(v) for v in x)] += 1
        freq = np.array([counts[o] / N for o in outcomes])
        assert 0.5 * np.abs(freq - exact).sum() <= bound, type(strat).__name__


def test_gcc_strategies_against_exact_distribution_bunched_input():
    from theboss_b200.simulation_strategies.generalized_cliffords_b_simulation_strategy import GeneralizedCliffordsBSimulationStrategy
    from theboss_b200.simulation_strategies.generalized_cliffords_simulation_strategy import GeneralizedCliffordsSimulationStrategy
    Factory, Type, _, Fixed, _ = _classes()
    U, s, N = workloads.haar(4, 41), [2, 1, 0, 1], 20000
    calc = Factory(U, None, None).generate_calculator()
    exact_calc = Fixed(_config(U, s), calc)
    outcomes, exact = exact_calc.get_outcomes_in_proper_order(), np.array(exact_calc.calculate_distribution())
    bound = _tvd_bound(len(outcomes), N)
    np.random.seed(2)
    for strat in (GeneralizedCliffordsSimulationStrategy(calc), GeneralizedCliffordsBSimulationStrategy(calc)):
        samples = strat.simulate(s, N)
        counts = {o: 0 for o in outcomes}
        for x in samples:
            counts[tuple(int(v) for v in x)] += 1
        freq = np.array([counts[o] / N for o in outcomes])
        assert 0.5 * np.abs(freq - exact).sum() <= bound, type(strat).__name__
