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
