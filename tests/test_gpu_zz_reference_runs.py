"""Rows f1 and f3 of SURVEY.md section 8(f) against runs of the reference itself (this module sorts last on purpose: it
was added after the last GPU visit of round 1).  f3: see the last test.  f1: outcome statistics of the two BOBS strategies, with
bp_gccb_simulate_batch (one matrix + one input state per sample, Philox decisions) underneath, against the frequencies the
UNMODIFIED reference produced on the same seeded networks (tests/golden/bobs_frequencies.json, 50 000 samples per case,
tests/golden/make_bobs_golden.py).  The strategies draw fresh random phases, permutations and lossy inputs per sample, so
parity is statistical: two-sample total-variation distance below the bound the reference's own sampling tests use
(tests/simulation_strategies_tests_common.py:192-208 there), at delta = 1e-3 per side, and a two-sample chi-square test.  tests/test_host_logic.py runs the
same comparison on the CPU with the oracle's sampling loop underneath, including a check that a wrong parameter is rejected."""
import numpy as np
import pytest

from tests import bobs_cases

pytestmark = pytest.mark.gpu

_REF_SAMPLES, _CASES = bobs_cases.load_cases()


@pytest.mark.parametrize("name", sorted(_CASES))
def test_bobs_strategy_matches_reference_frequencies(name):
    case = _CASES[name]
    N = 30000   # one device batch (the C ABI cuts requests at 32768 samples)
    K = bobs_cases.outcomes_count(case)
    np.random.seed(300 + sorted(_CASES).index(name))
    samples = bobs_cases.build_strategy(case).simulate(case["state"], N)
    assert len(samples) == N
    out = np.array(samples)
    assert out.shape == (N, len(case["state"])) and out.min() >= 0 and out.sum(axis=1).max() <= sum(case["state"])
    tvd = bobs_cases.tvd_to_reference(samples, case)
    assert tvd <= bobs_cases.tvd_bound(K, N) + bobs_cases.tvd_bound(K, _REF_SAMPLES), (name, tvd)
    # sharper: two-sample chi-square; wrong parameters give p < 1e-18 already at 12 000 samples (tests/test_host_logic.py)
    p = bobs_cases.chi2_pvalue(samples, case, _REF_SAMPLES)
    assert p > 1e-6, (name, p)


def test_version_a_uniform_losses_sampler_reproduces_reference_samples(golden_dir):
    """Row f3 against the reference itself (tests/golden/gcc_uniform_losses_samples.npz): identical seeds of the stdlib
    and NumPy generators give the reference's samples bit for bit, with the layer permanents from kernel K2."""
    from tests.test_host_logic import _check_uniform_losses_a_fixture
    _check_uniform_losses_a_fixture(golden_dir)


def test_kernels_against_reference_outputs_on_config_2_and_3_workloads(golden_dir):
    """tests/golden/reference_large.json (the unmodified reference on BASELINE config 2 items at n = 20 and config 3 steps at
    k = 12 .. 16, tests/golden/make_reference_large_golden.py): the drop-in classes, i.e. kernels K1 / K2 / K3, within
    BASELINE's relative tolerance of 1e-10 of the reference's own outputs (whose float64 error is up to 6.5e-11 here)."""
    import json
    import os
    from tests import workloads
    from theboss_b200.boson_sampling_utilities.permanent_calculators.bs_cc_ch_submatrices_permanent_calculator import (
        BSCCCHSubmatricesPermanentCalculator)
    from theboss_b200.boson_sampling_utilities.permanent_calculators.bs_cc_ryser_submatrices_permanent_calculator import (
        BSCCRyserSubmatricesPermanentCalculator)
    from theboss_b200.boson_sampling_utilities.permanent_calculators.chin_huh_permanent_calculator import ChinHuhPermanentCalculator
    from theboss_b200.boson_sampling_utilities.permanent_calculators.glynn_gray_permanent_calculator import GlynnGrayPermanentCalculator
    with open(os.path.join(golden_dir, "reference_large.json")) as f:
        g = json.load(f)
    U, S, T = workloads.c2_batch(g["c2"]["n"], g["c2"]["m"], g["c2"]["items"])
    for which, cls in (("chin_huh", ChinHuhPermanentCalculator), ("glynn", GlynnGrayPermanentCalculator)):
        for i, (re, im) in g["c2"][which].items():
            want = complex(re, im)
            got = cls(U, [int(x) for x in S[int(i)]], [int(x) for x in T[int(i)]]).compute_permanent()
            assert abs(got - want) <= 1e-10 * abs(want), (which, i, got, want)
    for key, values in g["c3"].items():
        k, free = int(key[1:key.index("_")]), key.endswith("free")
        U3, s, t = workloads.c3_step(k, 2 * k, collision_free=free)
        for which, v in values.items():
            cls = BSCCRyserSubmatricesPermanentCalculator if which == "ryser" else BSCCCHSubmatricesPermanentCalculator
            want = np.array([complex(*x) for x in v])
            got = np.array(cls(U3, [int(x) for x in s], [int(x) for x in t]).compute_permanents())
            assert got.shape == want.shape and np.abs(got - want).max() <= 1e-10 * np.abs(want).max(), (key, which)


def test_seeded_gccb_runs_reproduce_reference_samples(golden_dir):
    """tests/golden/gccb_seeded_samples.npz: GCC-B, its uniform-loss variant and the lossy-network wrapper at n = 12 .. 16
    (m up to 32), seeded through `numpy.random.seed` exactly like the reference run that produced the fixture; the device
    loop (K3 minors + finish kernel per step) must return the reference's samples bit for bit."""
    from tests.test_host_logic import _check_seeded_gccb_fixture
    _check_seeded_gccb_fixture(golden_dir)
