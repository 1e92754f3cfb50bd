"""Parity at BASELINE.json's full sizes through size-independent properties (the CPU oracle cannot run these
sizes in seconds): invariances of the permanent, transpose symmetry over the whole config-2 batch, and the
Laplace expansion that ties the all-minors kernel (K3) to the single-permanent kernels (K1 / K2)."""
import numpy as np
import pytest

from tests import workloads

pytestmark = pytest.mark.gpu
REL = 1e-10


@pytest.fixture(scope="module")
def handle():
    from theboss_b200 import _native
    return _native.default_handle(0)


def test_n30_row_permutation_and_column_scaling(handle):
    """perm(P A) = perm(A) (a different Gray path through the same sum) and perm(A diag(d)) = prod(d) perm(A)."""
    A = workloads.c4_matrix(30)
    base = handle.glynn_matrix(A)
    rng = np.random.RandomState(30)
    got = handle.glynn_matrix(A[rng.permutation(30)])
    assert abs(got - base) <= REL * abs(base)
    d = np.exp(1j * rng.uniform(0, 2 * np.pi, 30)) * rng.uniform(0.9, 1.1, 30)
    got = handle.glynn_matrix(A * d[None, :])
    assert abs(got - base * np.prod(d)) <= REL * abs(base * np.prod(d))
    got = handle.glynn_matrix(np.ascontiguousarray(A.T))
    assert abs(got - base) <= REL * abs(base)


def test_n30_shards_of_every_world_size_agree(handle):
    """The 1/2/4/8-GPU split of config 4: partial sums over contiguous Gray ranges."""
    from theboss_b200.distributed import combine_partials, gray_shard
    A = workloads.c4_matrix(30)
    base = handle.glynn_matrix(A)
    for world in (2, 8):
        parts = [handle.glynn_matrix_range(A, *gray_shard(30, world, r)) for r in range(world)]
        got = combine_partials(np.array(parts), 30)
        # each evaluation sits ~3e-12 from the long-double truth (incremental column sums over ~1e4 steps per thread,
        # amplified by the 1e7-fold cancellation of the Glynn sum); different splits round differently
        assert abs(got - base) <= 2e-11 * abs(base), world


def test_c2_full_batch_transpose_symmetry(handle):
    """Config 2 at full size (10^4 items, n=20, m=40): perm(U; s, t) = perm(U^T; t, s) item by item -- the two
    calls walk opposite sides of every item."""
    U, S, T = workloads.c2_batch()
    a = handle.perm_batched(U, S, T)
    b = handle.perm_batched(np.ascontiguousarray(U.T), T, S)
    assert np.all(np.isfinite(a.view(np.float64)))
    assert np.max(np.abs(a - b) / np.abs(a)) <= REL


def test_c3_minors_satisfy_the_laplace_expansion(handle):
    """Config 3 (k = 24, m = 48): sum_i s_i P_i U[j][i] = perm(U; s, t + e_j), the identity _compute_pmf relies on
    (generalized_cliffords_b_simulation_strategy.py:82-89); right-hand side from the batched single-permanent kernel."""
    for collision_free in (False, True):
        U, s, t = workloads.c3_step(24, 48, collision_free)
        minors = handle.minors(U, s, t)
        js = [0, 7, 23, 31, 47]
        S = np.repeat(s[None].astype(np.uint8), len(js), axis=0)
        T = np.repeat(t[None].astype(np.uint8), len(js), axis=0)
        T[np.arange(len(js)), js] += 1
        singles = handle.perm_batched(U, S, T)
        for q, j in enumerate(js):
            lhs = np.sum(s * minors * U[j, :])
            assert abs(lhs - singles[q]) <= REL * abs(singles[q]), (collision_free, j)


def test_c5_lossy_runs_conserve_particles(handle):
    from theboss_b200.boson_sampling_utilities.boson_sampling_utilities import prepare_interferometer_matrix_in_expanded_space
    U, U_lossy, s = workloads.c5_lossy(30, 60)
    out = handle.gccb_simulate(U, s, 512, eta=0.5, seed=60)
    kept = out.sum(axis=1)
    assert kept.min() >= 0 and kept.max() <= 30 and abs(kept.mean() - 15) < 1.0     # Binomial(30, 1/2)
    n_small = 12                                                                      # dilated run, fewer photons: seconds
    s12 = np.array([1] * n_small + [0] * (120 - n_small), dtype=np.int32)
    big = np.ascontiguousarray(prepare_interferometer_matrix_in_expanded_space(U_lossy))
    out = handle.gccb_simulate(big, s12, 256, seed=61)
    assert np.all(out.sum(axis=1) == n_small)                                         # every particle lands somewhere in 2m modes
    survived = out[:, :60].sum(axis=1).mean() / n_small
    assert abs(survived - np.linspace(0.3, 0.9, 60)[:n_small].mean()) < 0.08          # mean transmissivity of the used inputs


def test_c5_nonuniform_losses_at_full_size_through_the_strategy(handle):
    """BASELINE config 5(ii) as quoted: n = 30 photons, m = 60, non-uniform losses, through
    LossyNetworksGeneralizedCliffordsSimulationStrategy (lossy_networks_generalized_cliffords_simulation_strategy.py:41-88) -- i.e.
    GCC-B on the 120-mode dilation, steps k = 18 .. 30 in the two-lane kernels.  Size-independent properties: every particle is
    either detected or sits in a loss mode, a loss mode never holds more particles than the single input it couples to sent in,
    the detected fraction follows the transmissivities, and the samples do not depend on how the request is split."""
    from theboss_b200.boson_sampling_utilities.boson_sampling_utilities import prepare_interferometer_matrix_in_expanded_space
    from theboss_b200.boson_sampling_utilities.permanent_calculators.ryser_permanent_calculator import RyserPermanentCalculator
    from theboss_b200.simulation_strategies.lossy_networks_generalized_cliffords_simulation_strategy import (
        LossyNetworksGeneralizedCliffordsSimulationStrategy)
    U, U_lossy, s = workloads.c5_lossy(30, 60)
    big = np.ascontiguousarray(prepare_interferometer_matrix_in_expanded_space(U_lossy))
    s_big = np.concatenate([s, np.zeros(60, dtype=np.int32)])
    full = handle.gccb_simulate(big, s_big, 24, seed=62)
    assert np.all(full.sum(axis=1) == 30)
    loss_mode_of_input = [60 + int(np.argmax(np.abs(big[60:, i]))) for i in range(30)]
    assert len(set(loss_mode_of_input)) == 30
    lost = full[:, 60:]
    assert lost.max() <= 1 and np.all(lost[:, [j - 60 for j in range(60, 120) if j not in loss_mode_of_input]] == 0)
    detected = full[:, :60].sum(axis=1).mean() / 30
    assert abs(detected - np.linspace(0.3, 0.9, 60)[:30].mean()) < 0.12
    parts = [handle.gccb_simulate(big, s_big, 8, seed=62, first_sample=lo) for lo in (0, 8, 16)]
    assert np.array_equal(np.concatenate(parts), full)
    strategy = LossyNetworksGeneralizedCliffordsSimulationStrategy(RyserPermanentCalculator(U_lossy))
    np.random.seed(5)
    samples = strategy.simulate([int(x) for x in s], 6)
    assert len(samples) == 6 and all(len(x) == 60 and 0 <= sum(x) <= 30 for x in samples)
