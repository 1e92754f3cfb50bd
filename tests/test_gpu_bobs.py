"""Row f1 of SURVEY.md section 8(f): the batched one-matrix-per-sample entry point and the BOBS strategies on
top of it.  Acceptance follows the reference's tests/test_bobs_strategy.py: exact when nothing is approximated
(lossless and uniformly lossy matrices), and within the Brod-Oszmaniec bound (formula (22) of their paper)
otherwise."""
import numpy as np
import pytest

from tests import workloads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle():
    from theboss_b200 import _native
    return _native.default_handle(0)


def _tvd_bound(n_outcomes, samples, delta=1e-3):
    return np.sqrt((-np.log(delta) + n_outcomes * np.log(2)) / (2 * samples))


def _frequencies(samples, outcomes):
    counts = {o: 0 for o in outcomes}
    for x in samples:
        counts[tuple(int(v) for v in x)] += 1
    return np.array([counts[o] / len(samples) for o in outcomes])


def _exact(U, s, eta):
    from theboss_b200.boson_sampling_utilities.permanent_calculators.bs_permanent_calculator_factory import BSPermanentCalculatorFactory
    from theboss_b200.distribution_calculators.bs_exact_distribution_with_uniform_losses import (
        BosonSamplingExperimentConfiguration, BSDistributionCalculatorWithUniformLosses)
    cfg = BosonSamplingExperimentConfiguration(interferometer_matrix=U, initial_state=list(s), initial_number_of_particles=int(sum(s)),
                                               number_of_modes=len(s), number_of_particles_lost=0,
                                               number_of_particles_left=int(sum(s)), uniform_transmissivity=eta)
    calc = BSDistributionCalculatorWithUniformLosses(cfg, BSPermanentCalculatorFactory(U, None, None).generate_calculator())
    return calc.get_outcomes_in_proper_order(), np.array(calc.calculate_distribution())


def test_batch_entry_point_equals_single_matrix_calls(handle):
    """Same tape rows through bp_gccb_simulate_batch (one matrix + state per sample) and bp_gccb_simulate."""
    rng = np.random.RandomState(3)
    m, S = 7, 24
    Us = np.array([workloads.haar(m, 100 + i) for i in range(S)])
    states = np.zeros((S, m), dtype=np.int32)
    for i in range(S):
        for j in rng.randint(0, m, rng.randint(0, 6)):     # 0..5 particles, bunching allowed
            states[i, j] += 1
    n_max = int(states.sum(axis=1).max())
    tape = rng.random_sample((S, 1 + 2 * n_max))
    got = handle.gccb_simulate_batch(Us, states, tape=tape)
    for i in range(S):
        n_i = int(states[i].sum())
        if n_i == 0:
            assert not got[i].any()
            continue
        want = handle.gccb_simulate(np.ascontiguousarray(Us[i]), states[i], 1, tape=tape[i:i + 1, : 1 + 2 * n_i])
        assert np.array_equal(got[i], want[0]), i
    # shared matrix + state through the batch entry point == the plain call
    U, s = workloads.haar(m, 5), np.array([1, 1, 0, 2, 0, 1, 0], dtype=np.int32)
    tape = rng.random_sample((16, 11))
    a = handle.gccb_simulate(U, s, 16, tape=tape)
    b = handle.gccb_simulate_batch(np.repeat(U[None], 16, axis=0), np.repeat(s[None], 16, axis=0), tape=tape)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("eta", [1.0, 0.6])
def test_bobs_is_exact_when_nothing_is_approximated(eta):
    """tests/test_bobs_strategy.py:20-48 of the reference (hierarchy_level = number_of_modes)."""
    from theboss_b200.boson_sampling_utilities.permanent_calculators.bs_permanent_calculator_factory import BSPermanentCalculatorFactory
    from theboss_b200.distribution_calculators.bs_distribution_calculator_interface import BosonSamplingExperimentConfiguration
    from theboss_b200.simulation_strategies.simulation_strategy_factory import SimulationStrategyFactory, StrategyType
    U, s, N = workloads.haar(4, 61), [1, 1, 2, 0], 20000
    outcomes, exact = _exact(U, s, eta)
    calc = BSPermanentCalculatorFactory(U * np.sqrt(eta), None, None).generate_calculator()
    cfg = BosonSamplingExperimentConfiguration(interferometer_matrix=U * np.sqrt(eta), initial_state=s, initial_number_of_particles=4,
                                               number_of_modes=4, number_of_particles_lost=0, number_of_particles_left=4,
                                               uniform_transmissivity=eta, hierarchy_level=4)
    strat = SimulationStrategyFactory(cfg, calc, StrategyType.BOBS).generate_strategy()
    np.random.seed(12)
    samples = strat.simulate(s, N)
    assert len(samples) == N and len(samples[0]) == 4
    assert 0.5 * np.abs(_frequencies(samples, outcomes) - exact).sum() <= _tvd_bound(len(outcomes), N)


def test_bobs_approximation_stays_within_the_brod_oszmaniec_bound():
    """tests/test_bobs_strategy.py:50-90: TVD <= eta^2 (n - k) / 2 + eta (1 - eta) / 2 (+ statistical term)."""
    from theboss_b200.boson_sampling_utilities.permanent_calculators.bs_permanent_calculator_factory import BSPermanentCalculatorFactory
    from theboss_b200.simulation_strategies.nonuniform_losses_approximation_strategy import NonuniformLossesApproximationStrategy
    U, eta, N = workloads.haar(5, 71), 0.5, 20000
    s = [1, 1, 1, 0, 0]
    approximated = 2
    outcomes, exact = _exact(U, s, eta)
    calc = BSPermanentCalculatorFactory(U * np.sqrt(eta), None, None).generate_calculator()
    strat = NonuniformLossesApproximationStrategy(calc, approximated)
    np.random.seed(21)
    samples = strat.simulate(s, N)
    tvd = 0.5 * np.abs(_frequencies(samples, outcomes) - exact).sum()
    n, k = 3, 5 - approximated
    bound = eta ** 2 / 2 * abs(n - k) + eta * (1 - eta) / 2
    assert tvd <= bound + _tvd_bound(len(outcomes), N)
    assert tvd > 0   # it IS an approximation


def test_lossy_state_approximation_runs_and_conserves_particles():
    from theboss_b200.boson_sampling_utilities.permanent_calculators.bs_permanent_calculator_factory import BSPermanentCalculatorFactory
    from theboss_b200.simulation_strategies.lossy_state_approximated_simulation_strategy import LossyStateApproximationSimulationStrategy
    U, s = workloads.haar(6, 81), [1, 1, 1, 1, 0, 0]
    calc = BSPermanentCalculatorFactory(U, list(s), list(s)).generate_calculator()
    np.random.seed(4)
    out = np.array(LossyStateApproximationSimulationStrategy(calc, 0.7, 2).simulate(s, 4000))
    assert out.shape == (4000, 6) and out.sum(axis=1).max() <= 4
    assert abs(out.sum(axis=1).mean() - 0.7 * 4) < 0.1       # Binomial(4, 0.7) particles survive on average
    full = np.array(LossyStateApproximationSimulationStrategy(calc, 1.0, 6).simulate(s, 50))
    assert np.all(full.sum(axis=1) == 4)


@pytest.mark.parametrize("m,a,with_perms", [(6, 0, False), (6, 3, False), (12, 12, True), (9, 4, True), (40, 17, False), (120, 60, False)])
def test_device_side_matrix_build_matches_the_reference_construction(handle, m, a, with_perms):
    """Row f4: bp_bobs_build / bp_gccb_simulate_bobs build (B[:, perm]) @ diag(phases, 1 ...) @ QFT_a per sample on the device --
    M0 @ random_phases @ QFT of nonuniform_losses_approximation_strategy.py:331-347 and U[:, perm] @ phases @ QFT of
    lossy_state_approximated_simulation_strategy.py:329-362 -- checked against the NumPy restatement of those lines."""
    from oracle.handle_standin import bobs_matrices
    from theboss_b200.boson_sampling_utilities.boson_sampling_utilities import generate_qft_matrix_for_first_m_modes
    rng = np.random.RandomState(m * 31 + a)
    S = 9
    B = workloads.haar(m, m + a)
    qft = generate_qft_matrix_for_first_m_modes(a, m)[:a, :a]
    phases = np.exp(2j * np.pi * rng.rand(S, a))
    perms = np.argsort(rng.rand(S, m), axis=1).astype(np.int32) if with_perms else None
    got = handle.bobs_build(B, qft, phases, perms)
    want = bobs_matrices(B, qft, phases, perms)
    assert got.shape == want.shape == (S, m, m)
    assert np.abs(got - want).max() <= 1e-13
    assert np.array_equal(got[:, :, a:], want[:, :, a:])                        # untouched columns are copies
    # the sampler on device-built matrices == the sampler on the same matrices shipped from the host (same decision tape)
    if m <= 12:
        states = np.zeros((S, m), dtype=np.int32)
        for i in range(S):
            for j in rng.randint(0, m, rng.randint(0, 5)):
                states[i, j] += 1
        tape = rng.random_sample((S, 1 + 2 * max(1, int(states.sum(axis=1).max()))))
        a_dev = handle.gccb_simulate_bobs(B, qft, phases, perms, states, tape=tape)
        a_host = handle.gccb_simulate_batch(got, states, tape=tape)
        assert np.array_equal(a_dev, a_host)
