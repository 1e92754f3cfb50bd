"""K3+K4 parity (GPU): the device-resident GCC-B sampling loops and the GPU-backed strategy classes.

Bit-exactness is tested the way SURVEY.md Appendix B prescribes: the reference was run with its random
decisions injected from a decision tape (tests/golden/make_golden.py); the same tape is fed to the CUDA
path and the samples must be identical."""
import os

import numpy as np
import pytest

from tests import workloads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle():
    from theboss_b200 import _native
    return _native.default_handle(0)


@pytest.fixture(scope="module")
def orc():
    from oracle import pyoracle
    return pyoracle


@pytest.fixture(scope="module")
def gccb_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "gccb_samples.npz"))


def _strategy_classes():
    from theboss_b200.boson_sampling_utilities.permanent_calculators.ryser_permanent_calculator import RyserPermanentCalculator
    from theboss_b200.simulation_strategies.generalized_cliffords_b_simulation_strategy import GeneralizedCliffordsBSimulationStrategy
    from theboss_b200.simulation_strategies.generalized_cliffords_b_uniform_losses_simulation_strategy import (
        GeneralizedCliffordsBUniformLossesSimulationStrategy)
    from theboss_b200.simulation_strategies.lossy_networks_generalized_cliffords_simulation_strategy import (
        LossyNetworksGeneralizedCliffordsSimulationStrategy)
    return (RyserPermanentCalculator, GeneralizedCliffordsBSimulationStrategy,
            GeneralizedCliffordsBUniformLossesSimulationStrategy, LossyNetworksGeneralizedCliffordsSimulationStrategy)


def test_gccb_samples_bit_exact_against_reference_tapes(gccb_golden):
    Calc, GCCB, _, _ = _strategy_classes()
    z = gccb_golden
    for name in z["plain_names"]:
        U, s, tape = z[f"{name}_U"], z[f"{name}_s"], z[f"{name}_tape"]
        got = GCCB(Calc(U.copy(), None, None)).simulate(list(s), tape.shape[0], decision_tape=tape)
        assert isinstance(got, list) and isinstance(got[0], tuple) and len(got[0]) == len(s)
        assert np.array_equal(np.array(got), z[f"{name}_samples"]), name


def test_step_pmfs_match_reference(gccb_golden, handle, orc):
    """Layer (i) of Appendix B: the pmfs themselves, replayed along the reference's trajectories."""
    z = gccb_golden
    name = "plain_m6_n4"
    U, s, tape, pmfs = z[f"{name}_U"], z[f"{name}_s"], z[f"{name}_tape"], z[f"{name}_pmfs"]
    n, m = int(s.sum()), len(s)
    idx = 0
    for i in range(8):
        cur, r, remaining = np.zeros(m, dtype=np.int32), np.zeros(m, dtype=np.int32), orc.mode_assignment(s)
        for k in range(n):
            cur[remaining.pop(int(tape[i, 1 + 2 * k] * len(remaining)))] += 1
            got = handle.gccb_pmf(np.ascontiguousarray(U), cur, r)
            assert np.abs(got - pmfs[idx]).max() <= 1e-12
            # identical probabilities => identical index (the draw itself is bit-exact)
            assert orc.numpy_choice(got, tape[i, 2 + 2 * k]) == orc.numpy_choice(pmfs[idx], tape[i, 2 + 2 * k])
            r[orc.numpy_choice(pmfs[idx], tape[i, 2 + 2 * k])] += 1
            idx += 1


def test_uniform_losses_bit_exact(gccb_golden):
    Calc, _, GCCBU, _ = _strategy_classes()
    z = gccb_golden
    U, s, tape, eta = z["uniform_U"], z["uniform_s"], z["uniform_tape"], float(z["uniform_eta"])
    got = GCCBU(Calc(U.copy(), None, None), eta).simulate(np.array(s), tape.shape[0], decision_tape=tape)
    assert isinstance(got[0], np.ndarray) and got[0].dtype == np.int64   # ...b_uniform_losses...:108
    assert np.array_equal(np.array(got), z["uniform_samples"])
    assert len({int(g.sum()) for g in got}) > 2   # particle numbers really vary


def test_lossy_network_bit_exact(gccb_golden):
    Calc, _, _, LossyNet = _strategy_classes()
    z = gccb_golden
    U, s, tape = z["lossynet_U"], z["lossynet_s"], z["lossynet_tape"]
    calc = Calc(U.copy(), None, None)
    strat = LossyNet(calc)
    assert np.abs(calc.matrix - z["lossynet_expanded"]).max() <= 1e-14   # dilated in place (:41-44)
    got = strat.simulate(list(s), tape.shape[0], decision_tape=tape)
    assert len(got[0]) == len(s)
    assert np.array_equal(np.array(got), z["lossynet_samples"])


def test_oracle_and_device_agree_on_fresh_tapes(handle, orc):
    """Seeded tapes the reference never saw, n = 9 with bunched input, vs the oracle's sampling loop."""
    rng = np.random.RandomState(99)
    U = workloads.haar(12, 12)
    s = np.array([2, 1, 0, 1, 1, 0, 1, 0, 2, 0, 1, 0], dtype=np.int32)
    tape = rng.random_sample((48, 1 + 2 * int(s.sum())))
    want = np.array(orc.gccb_simulate(U, s, tape))
    got = handle.gccb_simulate(U, s, 48, tape=tape)
    assert np.array_equal(got, want)
    want_u = np.array(orc.gccb_uniform_losses_simulate(U, s, 0.7, tape))
    got_u = handle.gccb_simulate(U, s, 48, eta=0.7, tape=tape)
    assert np.array_equal(got_u, want_u)


def test_philox_mode_is_deterministic_and_split_invariant(handle):
    U = workloads.haar(8, 5)
    s = np.array([1, 1, 1, 1, 1, 0, 0, 0], dtype=np.int32)
    a = handle.gccb_simulate(U, s, 300, seed=1234)
    b = handle.gccb_simulate(U, s, 300, seed=1234)
    assert np.array_equal(a, b)
    parts = [handle.gccb_simulate(U, s, 100, seed=1234, first_sample=off) for off in (0, 100, 200)]
    assert np.array_equal(np.concatenate(parts), a)          # sharding over GPUs does not change samples
    c = handle.gccb_simulate(U, s, 300, seed=1235)
    assert not np.array_equal(a, c)
    assert np.all(a.sum(axis=1) == 5)


@pytest.mark.parametrize("eta", [-1.0, 0.6])
def test_large_batch_equals_the_same_job_cut_into_pieces(handle, eta):
    """The block sizing of the minors kernel depends on the batch size (k3_plan: 192 .. 2048 terms per lane group);
    the samples must not: a 4097-sample job equals the same job cut into three smaller calls."""
    U = workloads.haar(12, 9)
    s = np.array([1] * 7 + [0] * 5, dtype=np.int32)
    total = 4097                                                # odd: the halves differ in size
    whole = handle.gccb_simulate(U, s, total, eta=eta, seed=99)
    cuts = [0, 1500, 3000, total]
    parts = [handle.gccb_simulate(U, s, b - a, eta=eta, seed=99, first_sample=a) for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.array_equal(np.concatenate(parts), whole)
    assert np.all(whole.sum(axis=1) <= 7)
    if eta < 0:
        assert np.all(whole.sum(axis=1) == 7)


def test_sampler_statistics_against_exact_distribution(handle):
    """Same acceptance criterion as the reference's strategy tests
    (quantum_computations_utilities.py:95-127): TVD within sqrt((-ln delta + K ln 2) / (2N))."""
    from itertools import product
    from math import factorial
    m, n, N = 4, 3, 20000
    U = workloads.haar(m, 44)
    s = np.array([1, 1, 1, 0], dtype=np.int32)
    outcomes = [o for o in product(range(n + 1), repeat=m) if sum(o) == n]
    S = np.repeat(s[None].astype(np.uint8), len(outcomes), axis=0)
    T = np.array(outcomes, dtype=np.uint8)
    perms = handle.perm_batched(U, S, T)
    exact = np.array([abs(p) ** 2 / np.prod([factorial(x) for x in o]) for p, o in zip(perms, outcomes)])
    assert abs(exact.sum() - 1) < 1e-12
    samples = handle.gccb_simulate(U, s, N, seed=7)
    counts = {o: 0 for o in outcomes}
    for row in samples:
        counts[tuple(int(x) for x in row)] += 1
    freq = np.array([counts[o] / N for o in outcomes])
    tvd = 0.5 * np.abs(freq - exact).sum()
    bound = np.sqrt((-np.log(1e-3) + len(outcomes) * np.log(2)) / (2 * N))
    assert tvd <= bound, (tvd, bound)


def test_numpy_rng_mode_consumes_the_global_generator_like_the_reference():
    Calc, GCCB, _, _ = _strategy_classes()
    U = workloads.haar(6, 6)
    s = [1, 1, 1, 1, 0, 0]
    np.random.seed(11)
    a = GCCB(Calc(U, None, None), rng_mode="numpy").simulate(s, 10)
    state_after = np.random.random()
    # replay the same stream by hand: per sample, per particle one randint and one random_sample
    np.random.seed(11)
    for _ in range(10):
        for k in range(4):
            np.random.randint(0, 4 - k)
            np.random.random_sample()
    assert np.random.random() == state_after
    np.random.seed(11)
    b = GCCB(Calc(U, None, None), rng_mode="numpy").simulate(s, 10)
    assert a == b


def test_gcc_strategy_config1_against_reference_samples(golden_dir):
    """BASELINE config 1 (GCC, n=5, m=10, Haar seed 2024, 1000 samples, Glynn calculator): with numpy's
    generator seeded like the fixture the GPU-backed strategy must return the reference's samples."""
    from theboss_b200.boson_sampling_utilities.permanent_calculators.glynn_gray_permanent_calculator import GlynnGrayPermanentCalculator
    from theboss_b200.boson_sampling_utilities.permanent_calculators.ryser_permanent_calculator import RyserPermanentCalculator
    from theboss_b200.simulation_strategies.generalized_cliffords_simulation_strategy import GeneralizedCliffordsSimulationStrategy
    z = np.load(os.path.join(golden_dir, "gcc_samples.npz"))
    for name, cls in (("c1_glynn", GlynnGrayPermanentCalculator), ("bunched_ryser", RyserPermanentCalculator)):
        U, s = z[f"{name}_U"], [int(x) for x in z[f"{name}_s"]]
        ref = z[f"{name}_samples"].astype(np.int64)
        strat = GeneralizedCliffordsSimulationStrategy(cls(U.copy(), None, None))
        np.random.seed(7)   # the fixture's uniforms are RandomState(7).random_sample in call order
        got = np.array(strat.simulate(s, ref.shape[0]))
        assert np.array_equal(got, ref), name
        keys, vals = z[f"{name}_pmf_keys"], z[f"{name}_pmf_vals"]
        for kk, vv in zip(keys, vals):
            assert np.abs(strat.pmfs[tuple(int(x) for x in kk)] - vv).max() <= 1e-12 * vv.max()


def test_strategy_factory_and_deepcopy():
    import copy
    import pickle
    from theboss_b200.boson_sampling_utilities.permanent_calculators.bs_permanent_calculator_factory import BSPermanentCalculatorFactory
    from theboss_b200.simulation_strategies.simulation_strategy_factory import SimulationStrategyFactory, StrategyType
    U = workloads.haar(5, 8)
    calc = BSPermanentCalculatorFactory(U, None, None).generate_calculator()
    calc2 = pickle.loads(pickle.dumps(copy.deepcopy(calc)))        # simulation_strategy_factory.py:58, BOBS pools
    strat = SimulationStrategyFactory(None, calc2, StrategyType.GCC).generate_strategy()
    out = strat.simulate([1, 1, 0, 1, 0], 5)
    assert len(out) == 5 and all(sum(o) == 3 for o in out)
    with pytest.raises(NotImplementedError):
        SimulationStrategyFactory(None, calc2, StrategyType.FIXED_LOSS).generate_strategy()
