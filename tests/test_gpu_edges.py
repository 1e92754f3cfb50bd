"""Edge cases and size limits on the GPU path (empty and ragged inputs, maximum sizes, heavy bunching)."""
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


@pytest.mark.parametrize("N", [23, 34, 35, 36, 37])
def test_scaled_permutation_matrices_at_the_size_limits(handle, N):
    """perm(D P) = prod(d) exactly; N = 23 / 34 bracket the block-4 bulk kernel, 35+ run the warp-pair kernel (odd N: halves of 18 + 17 columns).
    (2^(N-1) equal-magnitude terms cancelling down to 2^(N-1) prod(d): a harsh accumulation test.)"""
    rng = np.random.RandomState(N)
    d = np.exp(1j * rng.uniform(0, 2 * np.pi, N)) * rng.uniform(0.8, 1.2, N)
    A = np.zeros((N, N), dtype=np.complex128)
    A[rng.permutation(N), np.arange(N)] = d
    got = handle.glynn_matrix(A)
    want = np.prod(d)
    assert abs(got - want) <= 1e-10 * abs(want)


def test_all_ones_matrix_is_n_factorial(handle):
    from math import factorial
    for N in (1, 2, 5, 12, 20, 24):
        got = handle.glynn_matrix(np.ones((N, N), dtype=np.complex128))
        assert abs(got - factorial(N)) <= 1e-10 * factorial(N)


def test_heavy_bunching_up_to_forty_particles(handle, orc):
    """n = 40 (BP_MAX_N) concentrated in a few modes: tiny walks, large binomial weights.  With 20-fold
    bunching the Chin-Huh sum cancels ~10 digits in ANY float64 evaluation (the reference's included), so
    the bar is 1e-10 or 12x the error of the double-precision restatement of the reference, whichever is
    looser (measured, tests/bunching_accuracy.py: kernel 8e-10 / 1e-9 / 1.5e-11 / 9e-11 against 9e-11 / 9e-10 / 1.3e-12 / 6e-11
    of the reference's arithmetic; round 1, before the product tree: up to 2e-9)."""
    m = 8
    U = workloads.haar(m, 40)
    S = np.array([[20, 20, 0, 0, 0, 0, 0, 0], [40, 0, 0, 0, 0, 0, 0, 0], [10, 10, 10, 10, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 7, 33]], dtype=np.uint8)
    T = np.array([[10, 10, 10, 10, 0, 0, 0, 0], [0, 0, 40, 0, 0, 0, 0, 0], [5, 5, 5, 5, 5, 5, 5, 5], [13, 0, 0, 27, 0, 0, 0, 0]], dtype=np.uint8)
    got = handle.perm_batched(U, S, T)
    for b in range(len(S)):
        want = orc.guan_permanent(U, S[b], T[b], orc.CHIN_HUH, "ld")
        ref_err = abs(orc.guan_permanent(U, S[b], T[b], orc.CHIN_HUH, "d") - want) / abs(want)
        assert abs(got[b] - want) <= max(1e-10, 12 * ref_err) * abs(want), (b, ref_err)


def test_max_mode_count(handle, orc):
    m = 256
    U = workloads.haar(m, 7)
    rng = np.random.RandomState(1)
    s = np.zeros(m, dtype=np.int32); t = np.zeros(m, dtype=np.int32)
    s[rng.choice(m, 6, replace=False)] = 1
    t[rng.choice(m, 5, replace=False)] = 1
    minors = handle.minors(U, s, t)
    want = orc.submatrices(U, s, t, orc.RYSER, "ld")
    assert np.abs(minors - want).max() <= 1e-12 * np.abs(want).max()
    out = handle.gccb_simulate(U, s, 64, seed=3)
    assert out.shape == (64, m) and np.all(out.sum(axis=1) == 6)
    with pytest.raises(Exception):
        handle.perm_batched(np.eye(257, dtype=np.complex128), np.zeros((1, 257), np.uint8), np.zeros((1, 257), np.uint8))


def test_minors_with_41_input_particles(handle, orc):
    """k - 1 = 40 output particles is the walker's limit; keep the walk short by bunching."""
    m = 6
    U = workloads.haar(m, 41)
    s = np.array([10, 9, 8, 7, 7, 0], dtype=np.int32)        # k = 41
    t = np.array([0, 20, 0, 20, 0, 0], dtype=np.int32)       # 40 particles, 21 * 11 terms
    got = handle.minors(U, s, t)
    want = orc.submatrices(U, s, t, orc.CHIN_HUH, "ld")
    ref_err = np.abs(orc.submatrices(U, s, t, orc.CHIN_HUH, "d") - want).max() / np.abs(want).max()
    # 20-fold bunching: same float64 cancellation caveat as test_heavy_bunching_up_to_forty_particles
    assert np.abs(got - want).max() <= max(1e-10, 12 * ref_err) * np.abs(want).max(), ref_err


def test_empty_and_degenerate_sampling_requests(handle):
    U = workloads.haar(5, 5)
    zero = np.zeros(5, dtype=np.int32)
    assert handle.gccb_simulate(U, zero, 7).tolist() == [[0] * 5] * 7          # no particles
    one = np.array([0, 0, 1, 0, 0], dtype=np.int32)
    assert handle.gccb_simulate(U, one, 0).shape == (0, 5)                      # no samples
    out = handle.gccb_simulate(U, one, 2000, seed=1)
    freq = out.mean(axis=0)
    assert np.abs(freq - np.abs(U[:, 2]) ** 2).max() < 0.05                      # single particle: |U[j][i]|^2
    lost = handle.gccb_simulate(U, np.array([1, 1, 1, 0, 0], dtype=np.int32), 50, eta=0.0, seed=2)
    assert not lost.any()                                                        # eta = 0: everything is lost
    kept = handle.gccb_simulate(U, np.array([1, 1, 1, 0, 0], dtype=np.int32), 50, eta=1.0, seed=2)
    assert np.all(kept.sum(axis=1) == 3)


def test_batched_call_with_zero_items(handle):
    U = workloads.haar(4, 4)
    out = handle.perm_batched(U, np.zeros((0, 4), np.uint8), np.zeros((0, 4), np.uint8))
    assert out.shape == (0,)


def test_two_handles_share_the_constant_bank_safely():
    """K1's bulk kernel keeps the matrix in the (per-device) constant bank; two handles on different streams
    must not see each other's matrix."""
    from theboss_b200 import _native
    h1, h2 = _native.Handle(0), _native.Handle(0)
    A1, A2 = workloads.c4_matrix(24), workloads.c4_matrix(25)
    want1, want2 = h1.glynn_matrix(A1), h1.glynn_matrix(A2)
    import threading
    res = {}

    def run(h, A, key):
        res[key] = [h.glynn_matrix(A) for _ in range(6)]
    th = [threading.Thread(target=run, args=(h1, A1, 1)), threading.Thread(target=run, args=(h2, A2, 2))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert all(v == want1 for v in res[1]) and all(v == want2 for v in res[2])
    h1.close(); h2.close()


def test_sharded_sampling_equals_single_call(handle):
    from theboss_b200.distributed import shard_bounds
    U = workloads.haar(9, 9)
    s = np.array([1, 2, 0, 1, 1, 0, 1, 0, 0], dtype=np.int32)
    full = handle.gccb_simulate(U, s, 101, seed=77)
    parts = []
    for r in range(4):
        lo, hi = shard_bounds(101, 4, r)
        parts.append(handle.gccb_simulate(U, s, hi - lo, seed=77, first_sample=lo))
    assert np.array_equal(np.concatenate(parts), full)


def test_a_step_without_a_distribution_raises_like_numpy_random_choice(handle):
    """An all-zero input column of U leaves every output with probability 0: the reference's numpy.random.choice raises
    ValueError ("probabilities do not sum to 1", generalized_cliffords_b_simulation_strategy.py:107-110); the device loop must not
    place the particle in mode 0 instead (BP_ERR_DOMAIN -> ValueError)."""
    U = workloads.haar(6, 6)
    U[:, 1] = 0.0
    s = np.array([1, 1, 1, 0, 0, 0], dtype=np.int32)
    with pytest.raises(ValueError):
        handle.gccb_simulate(U, s, 16, seed=3)
    with pytest.raises(ValueError):
        handle.gccb_pmf(U, np.array([0, 1, 0, 0, 0, 0], dtype=np.int32), np.zeros(6, dtype=np.int32))
    ok = handle.gccb_simulate(workloads.haar(6, 6), s, 16, seed=3)            # the handle stays usable
    assert np.all(ok.sum(axis=1) == 3)


def test_concurrent_calculators_on_the_shared_default_handle(handle):
    """Independent calculator objects in different Python threads share the process-wide handle of a device; its calls are
    serialised by the binding (the C handle holds one staging buffer and one set of scratch slots)."""
    import threading
    from theboss_b200.boson_sampling_utilities.permanent_calculators.chin_huh_permanent_calculator import ChinHuhPermanentCalculator
    from theboss_b200.boson_sampling_utilities.permanent_calculators.glynn_gray_permanent_calculator import GlynnGrayPermanentCalculator
    rng = np.random.RandomState(5)
    jobs = []
    for i in range(4):
        m = 6 + i
        U = workloads.haar(m, 40 + i)
        s = np.zeros(m, dtype=int); t = np.zeros(m, dtype=int)
        for j in rng.randint(0, m, 5): s[j] += 1
        for j in rng.randint(0, m, 5): t[j] += 1
        cls = ChinHuhPermanentCalculator if i % 2 else GlynnGrayPermanentCalculator
        jobs.append(cls(U, list(s), list(t)))
    want = [c.compute_permanent() for c in jobs]
    got = [[] for _ in jobs]

    def run(i):
        for _ in range(25):
            got[i].append(jobs[i].compute_permanent())
    th = [threading.Thread(target=run, args=(i,)) for i in range(len(jobs))]
    [t.start() for t in th]
    [t.join() for t in th]
    for i in range(len(jobs)):
        assert all(v == want[i] for v in got[i]), i
