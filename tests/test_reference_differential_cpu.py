"""Differential run of the reference's calculators and the drop-in classes on the same inputs, edge cases included
(build container only: imports the unmodified reference from /root/reference; the arithmetic under the drop-in comes from
the CPU stand-in oracle/handle_standin.py, so this pins the Python layer: accepted argument types, padding of short states,
shortcuts, exceptions and return types).  The deliberate deviations are listed in DESIGN.md section 5."""
import os
import sys

import numpy as np
import pytest

from tests import workloads

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("THEBOSS_REFERENCE", "/root/reference")
_PC = "boson_sampling_utilities.permanent_calculators."


@pytest.fixture(scope="module")
def reference():
    if not os.path.isdir(os.path.join(REF, "theboss")):
        pytest.skip("reference checkout not available on this machine")
    import importlib
    sys.path[:0] = [REF, os.path.join(REPO, "oracle", "refshim")]      # taken off again by _restore_path
    return lambda name: importlib.import_module("theboss." + name)


@pytest.fixture(autouse=True)
def _restore_path():
    before = list(sys.path)
    yield
    sys.path[:] = before


def _outcome(f):
    try:
        return "ok", f()
    except Exception as error:          # noqa: BLE001 -- the exception type is what is compared
        return "raised", type(error).__name__


U4 = workloads.haar(4, 3)
SINGLE_CASES = {
    "plain": (U4, [1, 1, 0, 1], [0, 2, 1, 0]),
    "collision-free": (U4, [1, 1, 1, 0], [0, 1, 1, 1]),
    "all in one mode": (U4, [0, 3, 0, 0], [0, 0, 3, 0]),
    "no particles": (U4, [0, 0, 0, 0], [0, 0, 0, 0]),
    "short states of equal length": (U4, [1, 1, 0], [0, 1, 1]),
    "states of different length": (U4, [1, 1, 0], [0, 1, 1, 0]),
    "float ndarray states": (U4, np.array([1., 0., 2., 0.]), np.array([0., 1., 1., 1.])),
    "tuple states": (U4, (1, 0, 2, 0), (0, 1, 1, 1)),
    "list-of-lists matrix": (U4.tolist(), [1, 0, 2, 0], [0, 1, 1, 1]),
    "one particle": (U4, [0, 1, 0, 0], [0, 0, 1, 0]),
    "1x1 matrix, three particles": (np.array([[0.5 + 0.5j]]), [3], [3]),
}
SINGLE_CLASSES = {
    "ryser_permanent_calculator": "RyserPermanentCalculator",
    "chin_huh_permanent_calculator": "ChinHuhPermanentCalculator",
    "glynn_gray_permanent_calculator": "GlynnGrayPermanentCalculator",
    "classic_permanent_calculator": "ClassicPermanentCalculator",
}


@pytest.mark.parametrize("module", sorted(SINGLE_CLASSES))
def test_single_permanent_calculators_behave_like_the_reference(reference, monkeypatch, module):
    import importlib
    from oracle import handle_standin
    handle_standin.install(monkeypatch)
    ref_cls = getattr(reference(_PC + module), SINGLE_CLASSES[module])
    our_cls = getattr(importlib.import_module("theboss_b200." + _PC + module), SINGLE_CLASSES[module])
    for name, (M, s, t) in SINGLE_CASES.items():
        want = _outcome(lambda: ref_cls(M, s, t).compute_permanent())
        got = _outcome(lambda: our_cls(M, s, t).compute_permanent())
        assert got[0] == want[0], (module, name, want, got)
        if want[0] == "raised":
            assert got[1] == want[1], (module, name, want, got)
        else:
            assert type(got[1]) is type(want[1]) is np.complex128, (module, name)
            assert abs(got[1] - want[1]) <= 1e-10 * max(1.0, abs(want[1])), (module, name, want, got)
    # the three properties hand back the objects they were given (reference tests mutate `calculator.matrix` in place)
    M, s, t = U4.copy(), [1, 1, 0, 1], [0, 2, 1, 0]
    calc = our_cls(M, s, t)
    assert calc.matrix is M and calc.input_state is s and calc.output_state is t
    before = calc.compute_permanent()
    calc.matrix *= 2.0
    assert abs(calc.compute_permanent() - before * 2 ** 3) <= 1e-10 * abs(before) * 8


SUB_CASES = {
    "plain": (U4, [1, 1, 0, 1], [0, 1, 1, 0]),
    "bunched": (U4, [2, 1, 0, 1], [0, 2, 1, 0]),
    "k = 1": (U4, [0, 1, 0, 0], [0, 0, 0, 0]),
    "k = 1, no output state": (U4, [0, 1, 0, 0], None),
    "float ndarray states": (U4, np.array([1., 1., 0., 1.]), np.array([0., 1., 1., 0.])),
    "short states": (U4, [1, 1, 1], [0, 1, 1]),
}
SUB_CLASSES = {
    "bs_cc_ryser_submatrices_permanent_calculator": "BSCCRyserSubmatricesPermanentCalculator",
    "bs_cc_ch_submatrices_permanent_calculator": "BSCCCHSubmatricesPermanentCalculator",
}


@pytest.mark.parametrize("module", sorted(SUB_CLASSES))
def test_submatrices_calculators_behave_like_the_reference(reference, monkeypatch, module):
    import importlib
    from oracle import handle_standin
    handle_standin.install(monkeypatch)
    ref_cls = getattr(reference(_PC + module), SUB_CLASSES[module])
    our_cls = getattr(importlib.import_module("theboss_b200." + _PC + module), SUB_CLASSES[module])
    for name, (M, s, t) in SUB_CASES.items():
        want = _outcome(lambda: ref_cls(M, s, t).compute_permanents())
        got = _outcome(lambda: our_cls(M, s, t).compute_permanents())
        assert got[0] == want[0] == "ok", (module, name, want, got)
        assert type(got[1]) is type(want[1]) is list and len(got[1]) == len(want[1]), (module, name)
        assert all(type(v) is np.complex128 for v in got[1]), (module, name)
        assert np.allclose(np.array(got[1]), np.array(want[1], dtype=complex), rtol=1e-9, atol=1e-12), (module, name, want, got)


def test_gccb_family_replays_the_reference_under_the_same_numpy_seed(reference, monkeypatch):
    """`rng_mode="numpy"`: the drop-in strategies consume NumPy's global generator in the reference's call order, so the same
    seed gives the reference's samples (generalized_cliffords_b_simulation_strategy.py:94-110, the uniform-loss subclass
    :67-121, the lossy-network wrapper :41-88), in the reference's container types."""
    from oracle import handle_standin
    from theboss_b200.boson_sampling_utilities.permanent_calculators.ryser_permanent_calculator import RyserPermanentCalculator
    from theboss_b200.simulation_strategies.generalized_cliffords_b_simulation_strategy import GeneralizedCliffordsBSimulationStrategy
    from theboss_b200.simulation_strategies.generalized_cliffords_b_uniform_losses_simulation_strategy import (
        GeneralizedCliffordsBUniformLossesSimulationStrategy)
    from theboss_b200.simulation_strategies.lossy_networks_generalized_cliffords_simulation_strategy import (
        LossyNetworksGeneralizedCliffordsSimulationStrategy)
    handle_standin.install(monkeypatch)
    ref_calc = reference(_PC + "ryser_permanent_calculator").RyserPermanentCalculator
    ref_b = reference("simulation_strategies.generalized_cliffords_b_simulation_strategy").GeneralizedCliffordsBSimulationStrategy
    ref_bu = reference("simulation_strategies.generalized_cliffords_b_uniform_losses_simulation_strategy").GeneralizedCliffordsBUniformLossesSimulationStrategy
    ref_ln = reference("simulation_strategies.lossy_networks_generalized_cliffords_simulation_strategy").LossyNetworksGeneralizedCliffordsSimulationStrategy
    U, s = workloads.haar(6, 21), [1, 2, 0, 1, 1, 0]
    lossy = U @ np.diag(np.sqrt(np.linspace(0.4, 0.9, 6)))

    def both(ref_strategy, our_strategy, state, seed, n=40):
        np.random.seed(seed)
        want = ref_strategy.simulate(state, n)
        np.random.seed(seed)
        got = our_strategy.simulate(state, n)
        assert len(got) == len(want) == n
        assert [tuple(int(v) for v in x) for x in got] == [tuple(int(v) for v in x) for x in want]
        return want, got

    want, got = both(ref_b(ref_calc(U.copy(), None, None)), GeneralizedCliffordsBSimulationStrategy(RyserPermanentCalculator(U.copy()), rng_mode="numpy"), s, 3)
    assert type(got) is type(want) is list and type(got[0]) is type(want[0]) is tuple
    want, got = both(ref_bu(ref_calc(U.copy(), None, None), 0.6),
                     GeneralizedCliffordsBUniformLossesSimulationStrategy(RyserPermanentCalculator(U.copy()), 0.6, rng_mode="numpy"), np.array(s), 4)
    assert isinstance(got[0], np.ndarray) and isinstance(want[0], np.ndarray) and got[0].dtype == want[0].dtype
    want, got = both(ref_ln(ref_calc(lossy.copy(), None, None)),
                     LossyNetworksGeneralizedCliffordsSimulationStrategy(RyserPermanentCalculator(lossy.copy()), rng_mode="numpy"), s, 5)
    assert type(got[0]) is type(want[0]) is tuple and len(got[0]) == len(want[0]) == 6
    # the generator is left in the same state: the NEXT draw agrees too
    np.random.seed(9)
    ref_b(ref_calc(U.copy(), None, None)).simulate(s, 3)
    after_ref = np.random.random()
    np.random.seed(9)
    GeneralizedCliffordsBSimulationStrategy(RyserPermanentCalculator(U.copy()), rng_mode="numpy").simulate(s, 3)
    assert np.random.random() == after_ref


def test_utilities_equal_the_reference_over_a_grid(reference):
    """boson_sampling_utilities.py of the reference vs the drop-in module: state enumeration (order included), lossy input
    states, occupation <-> assignment conversions, effective scattering matrices, the 2m-mode dilation, QFT / random-phase
    matrices (same NumPy draws) and the state-space counting helpers."""
    from theboss_b200.boson_sampling_utilities import boson_sampling_utilities as ours
    ref = reference("boson_sampling_utilities.boson_sampling_utilities")
    for n in range(0, 6):
        for m in range(1, 6):
            for losses in (False, True):
                assert [tuple(x) for x in ref.generate_possible_states(n, m, losses)] == ours.generate_possible_states(n, m, losses), (n, m, losses)
                assert ref.bosonic_space_dimension(n, m, losses) == ours.bosonic_space_dimension(n, m, losses)
                assert ref.generate_state_types(m, n, losses) == ours.generate_state_types(m, n, losses)
                assert ref.compute_number_of_state_types(m, n, losses) == ours.compute_number_of_state_types(m, n, losses)
            for k in range(0, 7):
                assert (ref.compute_number_of_k_element_integer_partitions_of_n(k, n)
                        == ours.compute_number_of_k_element_integer_partitions_of_n(k, n))
    rng = np.random.RandomState(5)
    for _ in range(40):
        m = int(rng.randint(1, 6))
        state = [int(x) for x in rng.randint(0, 3, m)]
        assert tuple(ref.mode_occupation_to_mode_assignment(state)) == ours.mode_occupation_to_mode_assignment(state)
        assignment = ours.mode_occupation_to_mode_assignment(state)
        assert tuple(ref.mode_assignment_to_mode_occupation(assignment, m)) == ours.mode_assignment_to_mode_occupation(assignment, m)
        assert ref.compute_number_of_states_of_given_type(state) == ours.compute_number_of_states_of_given_type(state)
        for left in range(sum(state) + 1):
            want = [tuple(int(v) for v in x) for x in ref.generate_lossy_n_particle_input_states(state, left)]
            assert want == ours.generate_lossy_n_particle_input_states(state, left), (state, left)
        U = workloads.haar(m, 100 + m)
        out_state = [int(x) for x in rng.randint(0, 3, m)]
        want = ref.EffectiveScatteringMatrixCalculator(U, state, out_state).calculate()
        got = ours.EffectiveScatteringMatrixCalculator(U, state, out_state).calculate()
        assert len(want) == len(got) and all(np.array_equal(np.asarray(a), np.asarray(b)) for a, b in zip(want, got))
        lossy = U @ np.diag(np.sqrt(rng.uniform(0.2, 1.0, m)))
        assert np.allclose(ref.prepare_interferometer_matrix_in_expanded_space(lossy),
                           ours.prepare_interferometer_matrix_in_expanded_space(lossy), atol=1e-13)
        assert np.allclose(ref.get_modes_transmissivity_values_from_matrix(lossy), ours.get_modes_transmissivity_values_from_matrix(lossy))
        k = int(rng.randint(0, m + 1))
        assert np.allclose(ref.generate_qft_matrix_for_first_m_modes(k, m), ours.generate_qft_matrix_for_first_m_modes(k, m), atol=1e-14)
        np.random.seed(17)
        want = ref.generate_random_phases_matrix_for_first_m_modes(k, m)
        after = np.random.random()
        np.random.seed(17)
        assert np.allclose(want, ours.generate_random_phases_matrix_for_first_m_modes(k, m), atol=1e-15)
        assert np.random.random() == after


def test_distribution_calculators_equal_the_reference(reference, monkeypatch):
    """Exact distributions with fixed and with uniform losses (row f2): outcome order and every probability against the
    reference's calculators on bunched inputs (the reference loops over single compute_permanent calls and a process pool;
    the drop-in sends all (outcome, lossy input) pairs through one batched call)."""
    from oracle import handle_standin
    from theboss_b200.boson_sampling_utilities.permanent_calculators.chin_huh_permanent_calculator import ChinHuhPermanentCalculator
    from theboss_b200.distribution_calculators import bs_exact_distribution_with_uniform_losses as ours
    handle_standin.install(monkeypatch)
    ref = reference("distribution_calculators.bs_exact_distribution_with_uniform_losses")
    ref_calc = reference(_PC + "chin_huh_permanent_calculator").ChinHuhPermanentCalculator

    def config(module, U, s, lost, eta):
        return module.BosonSamplingExperimentConfiguration(
            interferometer_matrix=U, initial_state=list(s), initial_number_of_particles=int(sum(s)), number_of_modes=len(s),
            number_of_particles_lost=lost, number_of_particles_left=int(sum(s)) - lost, uniform_transmissivity=eta)

    for seed, s, lost, eta in ((31, [2, 1, 0, 1], 0, 1.0), (32, [2, 1, 0, 1], 2, 0.7), (33, [1, 1, 1], 1, 0.4), (34, [0, 3, 0, 0, 1], 3, 0.9)):
        U = workloads.haar(len(s), seed)
        for name in ("BSDistributionCalculatorWithFixedLosses", "BSDistributionCalculatorWithUniformLosses"):
            want_calc = getattr(ref, name)(config(ref, U, s, lost, eta), ref_calc(U.copy(), None, None))
            got_calc = getattr(ours, name)(config(ours, U, s, lost, eta), ChinHuhPermanentCalculator(U.copy()))
            assert [tuple(o) for o in want_calc.get_outcomes_in_proper_order()] == [tuple(o) for o in got_calc.get_outcomes_in_proper_order()]
            want, got = want_calc.calculate_distribution(), got_calc.calculate_distribution()
            assert type(got) is list and len(got) == len(want)
            assert np.allclose(got, want, rtol=1e-10, atol=1e-14), (name, s, lost, eta)
            some = [tuple(o) for o in got_calc.get_outcomes_in_proper_order()][::3]
            assert np.allclose(got_calc.calculate_probabilities_of_outcomes(some), want_calc.calculate_probabilities_of_outcomes(some),
                               rtol=1e-10, atol=1e-14)


def test_gcc_version_a_replays_the_reference_under_the_same_seeds(reference, monkeypatch):
    """GeneralizedCliffordsSimulationStrategy (row a11) and its uniform-loss subclass (row f3) on further inputs than the
    committed fixtures: same NumPy / stdlib seeds -> same samples, same memoised pmf layers."""
    import random
    from oracle import handle_standin
    from theboss_b200.boson_sampling_utilities.permanent_calculators.chin_huh_permanent_calculator import ChinHuhPermanentCalculator
    from theboss_b200.simulation_strategies.generalized_cliffords_simulation_strategy import GeneralizedCliffordsSimulationStrategy
    from theboss_b200.simulation_strategies.generalized_cliffords_uniform_losses_simulation_strategy import (
        GeneralizedCliffordsUniformLossesSimulationStrategy)
    handle_standin.install(monkeypatch)
    ref_calc = reference(_PC + "chin_huh_permanent_calculator").ChinHuhPermanentCalculator
    ref_a = reference("simulation_strategies.generalized_cliffords_simulation_strategy").GeneralizedCliffordsSimulationStrategy
    ref_au = reference("simulation_strategies.generalized_cliffords_uniform_losses_simulation_strategy").GeneralizedCliffordsUniformLossesSimulationStrategy
    for seed, s in ((1, [1, 1, 1, 0, 0]), (2, [0, 4, 0, 0]), (3, [1, 0, 0]), (4, [2, 2, 0, 1, 0, 0])):
        U = workloads.haar(len(s), 40 + seed)
        want_strategy, got_strategy = ref_a(ref_calc(U.copy(), None, None)), GeneralizedCliffordsSimulationStrategy(ChinHuhPermanentCalculator(U.copy()))
        np.random.seed(seed)
        want = want_strategy.simulate(list(s), 60)
        np.random.seed(seed)
        got = got_strategy.simulate(list(s), 60)
        assert [tuple(int(v) for v in x) for x in got] == [tuple(int(v) for v in x) for x in want], s
        assert type(got) is type(want) and type(got[0]) is type(want[0])
        for key, pmf in want_strategy.pmfs.items():
            assert np.allclose(got_strategy.pmfs[tuple(key)], pmf, rtol=1e-10, atol=1e-14), (s, key)
        want_strategy, got_strategy = ref_au(ref_calc(U.copy(), None, None), 0.7), GeneralizedCliffordsUniformLossesSimulationStrategy(ChinHuhPermanentCalculator(U.copy()), 0.7)
        random.seed(seed), np.random.seed(seed)
        want = want_strategy.simulate(list(s), 60)
        random.seed(seed), np.random.seed(seed)
        got = got_strategy.simulate(list(s), 60)
        assert np.array_equal(np.array(got), np.array(want)), s
        assert np.allclose(got_strategy.distribution, want_strategy.distribution, rtol=1e-10, atol=1e-14)
