"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares, sharding helpers, decision tapes, host utilities, fail-loud behaviour without a GPU."""
import os
import re

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(REPO, "include", "bossperm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from theboss_b200 import _native
    lib = _native.load_library()
    declared = _header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/bossperm.h but not exported"
        assert name in _native.SIGNATURES, f"{name} has no ctypes prototype"
    assert lib.bp_abi_version() == 1


def test_library_is_built_for_sm_100a_only():
    import subprocess
    from theboss_b200 import _native
    out = subprocess.run(["cuobjdump", "-lelf", _native.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from theboss_b200 import _native
    from theboss_b200.boson_sampling_utilities.permanent_calculators.ryser_permanent_calculator import RyserPermanentCalculator
    with pytest.raises(_native.BossPermError):
        RyserPermanentCalculator(np.eye(2), [1, 1], [1, 1]).compute_permanent()


def test_product_package_never_imports_the_oracle():
    for root, _, files in os.walk(os.path.join(REPO, "theboss_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                hits = re.findall(r"^\s*(?:from|import)\s+[^\n]*oracle[^\n]*$", src, flags=re.M)
                assert not hits, (os.path.join(root, f), hits)
                assert "libbossperm_oracle" not in src and "pyoracle" not in src


def test_shard_bounds_cover_without_overlap():
    from theboss_b200.distributed import gray_shard, shard_bounds
    for total in (0, 1, 7, 64, 1000, 2 ** 29):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_bounds(total, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(edges[:-1], edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    assert gray_shard(30, 8, 3) == (3 * 2 ** 26, 4 * 2 ** 26)


def test_combine_partials_is_exact_double_double():
    from theboss_b200.distributed import combine_partials
    parts = np.array([[1.0, 1e-20, -2.0, 0.0], [1e-17, 0.0, 2.0, 3e-18], [-1.0, 0.0, 1e-30, 0.0]])
    got = combine_partials(parts, 1)
    assert got.real == 1e-17 + 1e-20
    assert abs(got.imag - (3e-18 + 1e-30)) <= 1e-33


def test_expanded_space_matrix_is_an_isometry_on_the_physical_inputs():
    from theboss_b200.boson_sampling_utilities.boson_sampling_utilities import prepare_interferometer_matrix_in_expanded_space
    from tests import workloads
    m = 5
    A = workloads.haar(m, 3) @ np.diag(np.sqrt(np.linspace(0.2, 1.0, m)))
    E = prepare_interferometer_matrix_in_expanded_space(A)
    assert E.shape == (2 * m, 2 * m)
    # like the reference's [[S, L], [L, S]] block (boson_sampling_utilities.py:331-341) the result is an
    # isometry on the m physical input modes (all the sampler uses), not a full unitary
    assert np.abs(E[:, :m].conj().T @ E[:, :m] - np.eye(m)).max() < 1e-13
    assert np.abs(E[:m, :m] - A).max() < 1e-13
    z = np.load(os.path.join(REPO, "tests", "golden", "gccb_samples.npz"))
    assert np.abs(prepare_interferometer_matrix_in_expanded_space(z["lossynet_U"]) - z["lossynet_expanded"]).max() < 1e-14


def test_effective_scattering_matrix_convention():
    from theboss_b200.boson_sampling_utilities.boson_sampling_utilities import (
        EffectiveScatteringMatrixCalculator, mode_occupation_to_mode_assignment)
    from oracle import pyoracle as orc
    U = np.arange(16).reshape(4, 4) + 1j
    A = np.array(EffectiveScatteringMatrixCalculator(U, [2, 0, 1, 0], [0, 1, 0, 2]).calculate())
    assert np.array_equal(A, U[np.ix_([1, 3, 3], [0, 0, 2])])   # rows = outputs, columns = inputs
    assert np.array_equal(A, orc.effective_matrix(U, [2, 0, 1, 0], [0, 1, 0, 2]))
    assert EffectiveScatteringMatrixCalculator(U, [0, 0, 0, 0], [0, 0, 0, 0]).calculate() == []
    assert mode_occupation_to_mode_assignment([2, 0, 1, 3]) == (0, 0, 2, 3, 3, 3)


def test_numpy_compatible_tape_replays_reference_decisions():
    """With /root/reference present (build container) run the reference GCC-B under a numpy seed and check
    that the tape generated under the same seed reproduces its samples through the oracle's loop."""
    ref = os.environ.get("THEBOSS_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "theboss")):
        pytest.skip("reference checkout not available on this machine")
    import sys
    sys.path[:0] = [ref, os.path.join(REPO, "oracle", "refshim")]
    try:
        from theboss.boson_sampling_utilities.permanent_calculators.ryser_permanent_calculator import RyserPermanentCalculator
        from theboss.simulation_strategies.generalized_cliffords_b_simulation_strategy import GeneralizedCliffordsBSimulationStrategy
        from theboss.simulation_strategies.generalized_cliffords_b_uniform_losses_simulation_strategy import (
            GeneralizedCliffordsBUniformLossesSimulationStrategy)
    finally:
        del sys.path[:2]
    from oracle import pyoracle as orc
    from tests import workloads
    from theboss_b200.simulation_strategies.decision_tape import numpy_compatible_tape
    U, s = workloads.haar(6, 21), [1, 2, 0, 1, 1, 0]
    np.random.seed(3)
    want = GeneralizedCliffordsBSimulationStrategy(RyserPermanentCalculator(U, None, None)).simulate(s, 12)
    np.random.seed(3)
    tape = numpy_compatible_tape(sum(s), 12)
    assert orc.gccb_simulate(U, s, tape) == [tuple(int(x) for x in w) for w in want]
    np.random.seed(4)
    want = GeneralizedCliffordsBUniformLossesSimulationStrategy(RyserPermanentCalculator(U, None, None), 0.6).simulate(np.array(s), 12)
    np.random.seed(4)
    tape = numpy_compatible_tape(sum(s), 12, orc.binomial_weights(sum(s), 0.6))
    assert np.array_equal(np.array(orc.gccb_uniform_losses_simulate(U, s, 0.6, tape)), np.array(want))


def _install_oracle_handle(monkeypatch):
    """CPU stand-in for the native handle (oracle/handle_standin.py): the arithmetic below the C-ABI boundary comes from
    the oracle, so that the host-side logic above it can be tested without a GPU."""
    from oracle import handle_standin
    return handle_standin.install(monkeypatch)


def test_gcc_host_loop_reproduces_reference_samples_with_oracle_permanents(golden_dir, monkeypatch):
    """Host logic of the GCC (version A) strategy on the CPU: layer memo, speculative prefetch, the draw loop on cached
    Python floats and the consumption of numpy's generator must reproduce the reference's BASELINE config 1 run
    (tests/golden/gcc_samples.npz) when the layer permanents are exact.  (The GPU test of the same name checks the
    same fixture with kernel K2 underneath.)"""
    from theboss_b200 import _native
    from theboss_b200.simulation_strategies.generalized_cliffords_simulation_strategy import GeneralizedCliffordsSimulationStrategy
    _install_oracle_handle(monkeypatch)

    class Calc:   # minimal calculator interface: the strategy only reads .matrix (and .device)
        def __init__(self, U):
            self.matrix = U

    z = np.load(os.path.join(golden_dir, "gcc_samples.npz"))
    for name in ("c1_glynn", "bunched_ryser"):
        U, s = z[f"{name}_U"], [int(x) for x in z[f"{name}_s"]]
        ref = z[f"{name}_samples"].astype(np.int64)
        strat = GeneralizedCliffordsSimulationStrategy(Calc(U.copy()))
        np.random.seed(7)
        got = np.array(strat.simulate(s, ref.shape[0]))
        assert np.array_equal(got, ref), name
        for kk, vv in zip(z[f"{name}_pmf_keys"], z[f"{name}_pmf_vals"]):
            assert np.abs(strat.pmfs[tuple(int(x) for x in kk)] - vv).max() <= 1e-12 * vv.max()
        # a second run on the same object starts from an empty memo and draws the same samples again
        np.random.seed(7)
        assert np.array_equal(np.array(strat.simulate(s, 50)), ref[:50]), name


def test_exact_distribution_host_logic_with_oracle_permanents(golden_dir, monkeypatch):
    """Host logic of the exact distribution calculators (outcome enumeration, lossy-input weights, normalisation,
    binomial loss weights) on the CPU against the reference's literal 56-entry vector
    (tests/test_exact_distribution_calculator.py:75-142 in the reference) and against sum = 1."""
    import json
    from tests import workloads
    from theboss_b200 import _native
    from theboss_b200.boson_sampling_utilities.permanent_calculators.bs_permanent_calculator_factory import (
        BSPermanentCalculatorFactory, PermanentCalculatorType)
    from theboss_b200.distribution_calculators.bs_exact_distribution_with_uniform_losses import (
        BosonSamplingExperimentConfiguration, BSDistributionCalculatorWithFixedLosses, BSDistributionCalculatorWithUniformLosses)
    _install_oracle_handle(monkeypatch)

    def config(U, s, lost=0, eta=1.0):
        return BosonSamplingExperimentConfiguration(
            interferometer_matrix=U, initial_state=list(s), initial_number_of_particles=int(sum(s)), number_of_modes=len(s),
            number_of_particles_lost=lost, number_of_particles_left=int(sum(s)) - lost, uniform_transmissivity=eta)

    with open(os.path.join(golden_dir, "exact_distribution.json")) as f:
        g = json.load(f)
    P = np.array(g["matrix_real"], dtype=np.complex128)
    calc = BSPermanentCalculatorFactory(None, None, None, PermanentCalculatorType.CHIN_HUH).generate_calculator()
    dist = BSDistributionCalculatorWithUniformLosses(config(P, g["initial_state"], lost=2, eta=g["eta"]), calc).calculate_distribution()
    assert len(dist) == 56
    assert np.allclose(dist, g["reference_literal"])
    assert np.allclose(dist, g["reference_computed"], rtol=1e-12, atol=1e-15)
    U = workloads.haar(4, 31)
    for state, lost in (([3, 1, 2, 0], 0), ([3, 1, 2, 0], 2)):
        assert abs(sum(BSDistributionCalculatorWithFixedLosses(config(U, state, lost), calc).calculate_distribution()) - 1) < 1e-10
        assert abs(sum(BSDistributionCalculatorWithUniformLosses(config(U, state, lost, eta=0.7), calc).calculate_distribution()) - 1) < 1e-10


def test_bobs_host_logic_matches_reference_frequencies_with_oracle_sampling(monkeypatch):
    """Row f1 on the CPU: the vectorised per-sample input states and matrices of the two BOBS strategies (random approximated
    mode, binomial thinning, random phases x QFT, column permutations, SVD-free dilation) must give the outcome statistics of
    the UNMODIFIED reference (tests/golden/bobs_frequencies.json, 50 000 samples per case from
    tests/golden/make_bobs_golden.py) when the per-sample GCC-B draw underneath is exact (oracle loop).  The GPU test
    tests/test_gpu_zz_reference_runs.py checks the same fixture with bp_gccb_simulate_batch underneath."""
    from tests import bobs_cases
    from theboss_b200 import _native
    _install_oracle_handle(monkeypatch)
    ref_samples, cases = bobs_cases.load_cases()
    N = 12000
    for i, (name, case) in enumerate(cases.items()):
        K = bobs_cases.outcomes_count(case)
        bound = bobs_cases.tvd_bound(K, N) + bobs_cases.tvd_bound(K, ref_samples)
        np.random.seed(50 + i)
        samples = bobs_cases.build_strategy(case).simulate(case["state"], N)
        assert len(samples) == N and all(len(x) == len(case["state"]) for x in samples[:10])
        tvd = bobs_cases.tvd_to_reference(samples, case)
        assert tvd <= bound, (name, tvd, bound)                                  # the reference's own acceptance criterion
        p = bobs_cases.chi2_pvalue(samples, case, ref_samples)
        assert p > 1e-6, (name, p)                                               # the sharper two-sample test
    # power of the check: a strategy with one parameter off is rejected (p-values measured: 1e-18 ... 1e-90)
    wrong = [("nla_m5_k2", dict(approximated_modes=0)), ("nla_m5_k2", dict(etas=[0.6, 0.6, 0.7, 0.8, 0.9])),
             ("lsa_m5_hl2", dict(eta=0.65)), ("lsa_m4_hl3_bunched", dict(hierarchy_level=2))]
    for name, override in wrong:
        case = cases[name]
        np.random.seed(7)
        samples = bobs_cases.build_strategy(case, **override).simulate(case["state"], N)
        assert bobs_cases.chi2_pvalue(samples, case, ref_samples) < 1e-9, (name, override)


def _check_uniform_losses_a_fixture(golden_dir):
    """Shared by the CPU (oracle permanents) and GPU (kernel K2) tests of row f3: same seeds of the stdlib and NumPy
    generators as tests/golden/make_uniform_losses_a_golden.py -> the reference's samples bit for bit, and the recorded
    distributions to 1e-12."""
    import random
    from theboss_b200.simulation_strategies.generalized_cliffords_uniform_losses_simulation_strategy import (
        GeneralizedCliffordsUniformLossesSimulationStrategy)

    class Calc:
        def __init__(self, U):
            self.matrix = U

    z = np.load(os.path.join(golden_dir, "gcc_uniform_losses_samples.npz"))
    assert len(z["names"]) == 3
    for name in z["names"]:
        U, s, eta = z[f"{name}_U"], [int(x) for x in z[f"{name}_s"]], float(z[f"{name}_eta"])
        ref = z[f"{name}_samples"]
        strat = GeneralizedCliffordsUniformLossesSimulationStrategy(Calc(U.copy()), eta)
        random.seed(int(z[f"{name}_seeds"][0]))
        np.random.seed(int(z[f"{name}_seeds"][1]))
        got = np.array(strat.simulate(s, ref.shape[0]))
        assert np.array_equal(got, ref), name
        assert np.allclose(strat.distribution, z[f"{name}_distribution"], rtol=1e-12, atol=1e-15), name
        assert np.allclose(strat.unweighted_distribution, z[f"{name}_unweighted"], rtol=1e-12, atol=1e-15), name


def test_version_a_uniform_losses_host_loop_reproduces_reference_samples(golden_dir, monkeypatch):
    """Row f3 on the CPU: per-particle stdlib `random.random()` loss draws, NumPy draws for the kept particles, layer memo
    and the distribution bookkeeping of generalized_cliffords_uniform_losses_simulation_strategy.py:126-173 in the
    reference, with exact (oracle) layer permanents underneath."""
    from theboss_b200 import _native
    _install_oracle_handle(monkeypatch)
    _check_uniform_losses_a_fixture(golden_dir)


def test_state_space_helpers():
    """Known values of the bookkeeping helpers mirrored from the reference's boson_sampling_utilities.py (:206-261,
    :345-502); identity with the reference's functions over a grid of sizes is part of tests/test_reference_suite_cpu.py
    (its tests/test_boson_sampling_utilities.py runs against these)."""
    from theboss_b200.boson_sampling_utilities import boson_sampling_utilities as bsu
    assert bsu.bosonic_space_dimension(3, 5) == 35 and bsu.bosonic_space_dimension(3, 5, losses=True) == 1 + 5 + 15 + 35
    assert bsu.bosonic_space_dimension(0, 4) == 1
    assert bsu.generate_state_types(3, 4) == [(4, 0, 0), (3, 1, 0), (2, 1, 1), (2, 2, 0)]
    assert bsu.generate_state_types(2, 2, losses=True) == [(2, 0), (1, 1), (0, 0), (1, 0)]
    assert [bsu.compute_number_of_k_element_integer_partitions_of_n(k, 7) for k in range(1, 8)] == [1, 3, 4, 3, 2, 1, 1]
    for m, n, losses in ((3, 4, False), (5, 6, True), (2, 7, True)):
        types = bsu.generate_state_types(m, n, losses)
        assert bsu.compute_number_of_state_types(m, n, losses) == len(types)
        # every Fock state belongs to exactly one type
        assert sum(bsu.compute_number_of_states_of_given_type(t) for t in types) == bsu.bosonic_space_dimension(n, m, losses)
    assert bsu.compute_number_of_states_of_given_type((2, 1, 1, 0)) == 12
    lossy = __import__("tests.workloads", fromlist=["haar"]).haar(4, 9) @ np.diag(np.sqrt([0.2, 0.5, 0.9, 1.0]))
    assert np.allclose(bsu.get_modes_transmissivity_values_from_matrix(lossy), [0.2, 0.5, 0.9, 1.0])


def test_distribution_calculator_snapshots_its_configuration(monkeypatch):
    """bs_distribution_calculator_with_fixed_losses.py:41 of the reference deep-copies the configuration; its tests rely on
    that (they re-point `initial_state` of the object they passed in before asking for the distribution)."""
    from tests import workloads
    from theboss_b200.boson_sampling_utilities.permanent_calculators.chin_huh_permanent_calculator import ChinHuhPermanentCalculator
    from theboss_b200.distribution_calculators.bs_exact_distribution_with_uniform_losses import (
        BosonSamplingExperimentConfiguration, BSDistributionCalculatorWithUniformLosses)
    _install_oracle_handle(monkeypatch)
    U = workloads.haar(3, 5)
    cfg = BosonSamplingExperimentConfiguration(interferometer_matrix=U, initial_state=[1, 1, 0], initial_number_of_particles=2,
                                               number_of_modes=3, number_of_particles_lost=0, number_of_particles_left=2,
                                               uniform_transmissivity=0.7)
    calc = BSDistributionCalculatorWithUniformLosses(cfg, ChinHuhPermanentCalculator(U))
    before = calc.calculate_distribution()
    cfg.initial_state, cfg.uniform_transmissivity = [0, 0, 2], 0.1
    assert calc.configuration is not cfg and list(calc.configuration.initial_state) == [1, 1, 0]
    assert np.allclose(calc.calculate_distribution(), before) and abs(sum(before) - 1) < 1e-12


def test_strategy_factory_defaults_and_out_of_scope_members():
    """Same default member as the reference (simulation_strategy_factory.py:52); the members that never touch a permanent are
    not part of the drop-in and raise NotImplementedError naming the reference class (there is no CPU path in the package)."""
    from theboss_b200.simulation_strategies import simulation_strategy_factory as ssf
    factory = ssf.SimulationStrategyFactory(None, None)
    assert factory.strategy_type == ssf.StrategyType.FIXED_LOSS and factory.available_threads_number == -1
    assert [t.value for t in ssf.StrategyType] == list(range(1, 10))
    for member, cls in ((ssf.StrategyType.FIXED_LOSS, "FixedLossSimulationStrategy"), (ssf.StrategyType.UNIFORM_LOSS, "UniformLossSimulationStrategy"),
                        (ssf.StrategyType.CLIFFORD_R, "CliffordsRSimulationStrategy"), (ssf.StrategyType.LOSSLESS_MODES_STRATEGY, "FixedLossSimulationStrategy")):
        factory.strategy_type = member
        with pytest.raises(NotImplementedError, match=cls):
            factory.generate_strategy()



def test_bench_flop_accounting_of_a_sampling_run_matches_the_true_draw_order():
    """bench.py -> gcc_sampling.roofline counts the algorithmic flops of a GCC-B run from its output samples alone, using
    the exchangeability of the chain-rule sequence (a uniformly random draw order per sample).  Check against the TRUE
    draw order, recorded from the oracle's sampling loop: same total within the statistical spread, and the estimator
    is exact on collision-free outputs."""
    import bench
    from oracle import pyoracle as orc
    from tests import workloads
    n, m, S = 6, 9, 1500
    U = workloads.haar(m, 77)
    s = np.array([1] * n + [0] * (m - n), dtype=np.int32)
    rng = np.random.RandomState(8)
    true_flops, samples = 0.0, np.zeros((S, m), dtype=np.int64)
    for i in range(S):
        cur, r, remaining = np.zeros(m, dtype=np.int32), np.zeros(m, dtype=np.int32), orc.mode_assignment(s)
        for k in range(1, n + 1):                                   # the loop of orc.gccb_simulate, keeping the order
            if k >= 2:
                true_flops += float((np.prod(r + 1) + 1) // 2) * (22 * k - 36)
            cur[remaining.pop(int(rng.random_sample() * len(remaining)))] += 1
            r[orc.numpy_choice(orc.gccb_pmf(U, cur, r), rng.random_sample())] += 1
        samples[i] = r
    estimate = bench.sampling_algorithmic_flops(samples, seed=1)
    assert abs(estimate - true_flops) <= 0.02 * true_flops, (estimate, true_flops)
    assert abs(bench.sampling_algorithmic_flops(samples, seed=2) - estimate) <= 0.02 * estimate      # spread between orders
    free = np.zeros((4, m), dtype=np.int64)
    free[:, :n] = 1
    assert bench.sampling_algorithmic_flops(free) == 4 * sum(2.0 ** (k - 2) * (22 * k - 36) for k in range(2, n + 1))
    assert bench.sampling_algorithmic_flops(np.zeros((0, m))) == 0.0


def test_bobs_per_sample_matrices_and_slicing(monkeypatch):
    """What the two BOBS strategies hand to bp_gccb_simulate_bobs: every per-sample matrix (rebuilt here from the operands with the
    NumPy restatement of the device build) must be what the reference
    builds -- M0 @ diag(phases) @ QFT in the top-left block of an isometric 2m-mode dilation
    (nonuniform_losses_approximation_strategy.py:331-347), resp. a column-permuted unitary times phases times QFT
    (lossy_state_approximated_simulation_strategy.py:329-362) -- although only the columns that change are recomputed per
    sample; requests are cut into slices that continue one Philox sample counter."""
    from tests import workloads
    from theboss_b200 import _native
    from theboss_b200.boson_sampling_utilities.boson_sampling_utilities import generate_qft_matrix_for_first_m_modes
    from theboss_b200.simulation_strategies.lossy_state_approximated_simulation_strategy import (
        LossyStateApproximationSimulationStrategy)
    from theboss_b200.simulation_strategies.nonuniform_losses_approximation_strategy import (
        NonuniformLossesApproximationStrategy)
    calls = []

    from oracle.handle_standin import bobs_matrices

    class Capture:
        def gccb_simulate_bobs(self, B, qft, phases, perms, states, seed=0, first_sample=0, tape=None):
            calls.append((bobs_matrices(B, qft, phases, perms), np.array(states), seed, first_sample))
            return np.array(states, dtype=np.int32)          # echo: lets the test see which rows came from which slice

    monkeypatch.setattr(_native, "default_handle", lambda device=0: Capture())

    class Calc:
        def __init__(self, U, s):
            self.matrix, self.input_state = U, list(s)

    m, s = 6, [2, 1, 0, 1, 1, 0]
    lossy = workloads.haar(m, 91) @ np.diag(np.sqrt([0.5, 0.9, 0.7, 0.6, 0.8, 0.95]))
    for k in (0, 2, 6):
        calls.clear()
        strat = NonuniformLossesApproximationStrategy(Calc(lossy.copy(), s), k)
        monkeypatch.setattr(strat, "_SLICE_SAMPLES", 8)                              # 8 samples per slice
        np.random.seed(k)
        out = strat.simulate(s, 20)
        assert [c[0].shape[0] for c in calls] == [8, 8, 4] and [c[3] for c in calls] == [0, 8, 16]
        assert len({c[2] for c in calls}) == 1                                       # one seed for the whole request
        Us, states = np.concatenate([c[0] for c in calls]), np.concatenate([c[1] for c in calls])
        assert np.array_equal(np.array(out), states[:, :m])
        qft = generate_qft_matrix_for_first_m_modes(k, m)
        for U2, st in zip(Us, states):
            # isometry on the physical inputs, and the physical block is M0 @ (unit-modulus diagonal) @ QFT
            assert np.abs(U2[:, :m].conj().T @ U2[:, :m] - np.eye(m)).max() < 1e-12
            D = np.linalg.solve(strat._initial_matrix, U2[:m, :m] @ qft.conj().T)
            assert np.abs(D - np.diag(np.diag(D))).max() < 1e-10 and np.allclose(np.abs(np.diag(D)), 1.0)
            assert np.allclose(np.diag(D)[k:], 1.0)
            assert st[m:].sum() == 0 and np.all(st[k:m] <= np.array(s[k:])) and st[:k].sum() <= sum(s[:k])
            assert (st[:k] > 0).sum() <= 1                                           # approximated particles sit in one mode
    U = workloads.haar(m, 92)
    for hl in (0, 2, 6):
        calls.clear()
        strat = LossyStateApproximationSimulationStrategy(Calc(U.copy(), s), 0.6, hl)
        monkeypatch.setattr(strat, "_SLICE_SAMPLES", 8)
        np.random.seed(hl)
        strat.simulate(s, 20)
        assert [c[0].shape[0] for c in calls] == [8, 8, 4] and [c[3] for c in calls] == [0, 8, 16]
        qft = generate_qft_matrix_for_first_m_modes(m - hl, m)
        for U2, st in zip(np.concatenate([c[0] for c in calls]), np.concatenate([c[1] for c in calls])):
            assert np.abs(U2 @ U2.conj().T - np.eye(m)).max() < 1e-12
            P = U.conj().T @ U2 @ qft.conj().T                                       # = permutation x unit-modulus diagonal
            assert np.allclose(np.sort(np.abs(P), axis=0)[-1], 1.0) and np.allclose(np.abs(P).sum(axis=0), 1.0, atol=1e-10)
            assert np.all(st[:hl] <= np.array(s[:hl])) and st[hl + 1:].sum() == 0


def _check_seeded_gccb_fixture(golden_dir):
    """Shared by the CPU (oracle loop) and GPU (kernels K3 / K4) tests: `rng_mode="numpy"` strategies under the seeds of
    tests/golden/make_gccb_seeded_golden.py must return the unmodified reference's samples (n = 12 .. 16) bit for bit."""
    from theboss_b200.boson_sampling_utilities.permanent_calculators.ryser_permanent_calculator import RyserPermanentCalculator
    from theboss_b200.simulation_strategies.generalized_cliffords_b_simulation_strategy import GeneralizedCliffordsBSimulationStrategy
    from theboss_b200.simulation_strategies.generalized_cliffords_b_uniform_losses_simulation_strategy import (
        GeneralizedCliffordsBUniformLossesSimulationStrategy)
    from theboss_b200.simulation_strategies.lossy_networks_generalized_cliffords_simulation_strategy import (
        LossyNetworksGeneralizedCliffordsSimulationStrategy)
    z = np.load(os.path.join(golden_dir, "gccb_seeded_samples.npz"))
    assert len(z["names"]) == 6
    for name in z["names"]:
        kind, U, s, want = str(z[f"{name}_kind"]), z[f"{name}_U"], [int(x) for x in z[f"{name}_s"]], z[f"{name}_samples"]
        calc = RyserPermanentCalculator(U.copy())
        if kind == "plain":
            strategy = GeneralizedCliffordsBSimulationStrategy(calc, rng_mode="numpy")
        elif kind == "uniform":
            strategy = GeneralizedCliffordsBUniformLossesSimulationStrategy(calc, float(z[f"{name}_eta"]), rng_mode="numpy")
        else:
            strategy = LossyNetworksGeneralizedCliffordsSimulationStrategy(calc, rng_mode="numpy")
        np.random.seed(int(z[f"{name}_seed"]))
        got = np.array([[int(v) for v in x] for x in strategy.simulate(s, want.shape[0])], dtype=np.int64)
        assert np.array_equal(got, want), name


def test_seeded_gccb_runs_reproduce_reference_samples_with_the_oracle_loop(golden_dir, monkeypatch):
    _install_oracle_handle(monkeypatch)
    _check_seeded_gccb_fixture(golden_dir)


def test_every_entry_point_rejects_a_null_handle_without_touching_the_device():
    """include/bossperm.h: "every function returns 0 on success or a negative bp_status; nothing throws".  With a NULL
    handle each compute / timing entry point must come back with BP_ERR_INVALID and a message (no crash, no CUDA call),
    which is also what a caller sees after a failed bp_create on a machine without a GPU."""
    import ctypes as C
    from theboss_b200 import _native
    lib = _native.load_library()
    out = (C.c_double * 4)()
    A = np.eye(2, dtype=np.complex128)
    s = np.array([1, 1], dtype=np.int32)
    S = np.array([[1, 1]], dtype=np.uint8)
    o, oi, pmf = np.zeros(4, dtype=np.complex128), np.zeros((1, 2), dtype=np.int32), np.zeros(2)
    calls = {
        "bp_synchronize": lambda: lib.bp_synchronize(None),
        "bp_device_info": lambda: lib.bp_device_info(None, None, None, None, None),
        "bp_timer_start": lambda: lib.bp_timer_start(None),
        "bp_timer_stop": lambda: lib.bp_timer_stop(None, C.byref(C.c_float())),
        "bp_fp64_peak": lambda: lib.bp_fp64_peak(None, 1.0, out),
        "bp_glynn_matrix": lambda: lib.bp_glynn_matrix(None, A.ctypes.data, 2, out),
        "bp_glynn_matrix_range": lambda: lib.bp_glynn_matrix_range(None, A.ctypes.data, 2, 0, 2, out),
        "bp_glynn_matrix_range_dev": lambda: lib.bp_glynn_matrix_range_dev(None, A.ctypes.data, 2, 0, 2, None),
        "bp_glynn_set_resident": lambda: lib.bp_glynn_set_resident(None, None),
        "bp_exchange_create": lambda: lib.bp_exchange_create(None, 2, 0, (C.c_ubyte * 64)()),
        "bp_exchange_connect": lambda: lib.bp_exchange_connect(None, (C.c_ubyte * 128)()),
        "bp_exchange_destroy": lambda: lib.bp_exchange_destroy(None),
        "bp_glynn_matrix_range_exchange": lambda: lib.bp_glynn_matrix_range_exchange(None, A.ctypes.data, 2, 0, 2, None),
        "bp_glynn_matrix_range_exchange_host": lambda: lib.bp_glynn_matrix_range_exchange_host(None, A.ctypes.data, 2, 0, 2, None),
        "bp_glynn_single": lambda: lib.bp_glynn_single(None, A.ctypes.data, 2, s.ctypes.data, s.ctypes.data, out),
        "bp_perm_batched": lambda: lib.bp_perm_batched(None, A.ctypes.data, 2, S.ctypes.data, S.ctypes.data, 1, 1, o.ctypes.data),
        "bp_perm_batched_dev": lambda: lib.bp_perm_batched_dev(None, A.ctypes.data, 2, S.ctypes.data, S.ctypes.data, 1, 1, o.ctypes.data),
        "bp_minors": lambda: lib.bp_minors(None, A.ctypes.data, 2, s.ctypes.data, s.ctypes.data, 1, o.ctypes.data),
        "bp_gccb_pmf": lambda: lib.bp_gccb_pmf(None, A.ctypes.data, 2, s.ctypes.data, s.ctypes.data, pmf.ctypes.data, None),
        "bp_gccb_simulate": lambda: lib.bp_gccb_simulate(None, A.ctypes.data, 2, s.ctypes.data, 1, -1.0, 0, 0, None, oi.ctypes.data),
        "bp_gccb_simulate_batch": lambda: lib.bp_gccb_simulate_batch(None, A.ctypes.data, 2, s.ctypes.data, 1, 0, 0, None, 0, oi.ctypes.data),
        "bp_gccb_simulate_bobs": lambda: lib.bp_gccb_simulate_bobs(None, A.ctypes.data, 2, A.ctypes.data, 2, A.ctypes.data, None, s.ctypes.data, 1, 0, 0,
                                                                   None, 0, oi.ctypes.data),
        "bp_bobs_build": lambda: lib.bp_bobs_build(None, A.ctypes.data, 2, A.ctypes.data, 2, A.ctypes.data, None, 1, o.ctypes.data),
    }
    declared = set(_header_symbols()) - {"bp_abi_version", "bp_create", "bp_create_on_stream", "bp_destroy", "bp_last_error", "bp_launch_count"}
    assert declared == set(calls), declared ^ set(calls)
    for name, call in calls.items():
        assert call() == _native.BP_ERR_INVALID, name
        assert b"NULL" in lib.bp_last_error(None), name
    assert lib.bp_launch_count(None) == 0 and lib.bp_destroy(None) == _native.BP_OK
