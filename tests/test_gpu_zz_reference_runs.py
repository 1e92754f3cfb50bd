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
