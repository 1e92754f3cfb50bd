"""K3 parity (GPU): all one-particle-removed permanents + the GCC-B step pmf, vs the reference golden
vectors (tests/test_bs_submatrices_permanent_calculators.py cases) and the oracle."""
import os

import numpy as np
import pytest

from tests import workloads

pytestmark = pytest.mark.gpu
REL_TOL = 1e-10


@pytest.fixture(scope="module")
def handle():
    from theboss_b200 import _native
    return _native.default_handle(0)


@pytest.fixture(scope="module")
def orc():
    from oracle import pyoracle
    return pyoracle


def _occ(rng, m, n):
    out = np.zeros(m, dtype=np.int32)
    for j in rng.randint(0, m, n):
        out[j] += 1
    return out


def test_submatrices_classes_against_reference_golden(golden_dir):
    from theboss_b200.boson_sampling_utilities.permanent_calculators.bs_cc_ch_submatrices_permanent_calculator import (
        BSCCCHSubmatricesPermanentCalculator)
    from theboss_b200.boson_sampling_utilities.permanent_calculators.bs_cc_ryser_submatrices_permanent_calculator import (
        BSCCRyserSubmatricesPermanentCalculator)
    z = np.load(os.path.join(golden_dir, "submatrices_permanents.npz"))
    for i in range(int(z["n_cases"])):
        U, s, t = z[f"U_{i}"], z[f"s_{i}"], z[f"t_{i}"]
        ref = z[f"chin_huh_{i}"]
        scale = max(np.abs(ref).max(), 1e-30)
        for cls in (BSCCCHSubmatricesPermanentCalculator, BSCCRyserSubmatricesPermanentCalculator):
            got = cls(U, s, t).compute_permanents()
            assert isinstance(got, list) and len(got) == len(s) and isinstance(got[0], np.complex128)
            assert np.abs(np.array(got) - ref).max() <= REL_TOL * scale, (i, cls.__name__)
            assert np.abs(np.array(got) - z[f"ryser_{i}"]).max() <= 1e-8 * scale


@pytest.mark.parametrize("m,k", [(4, 2), (5, 3), (8, 6), (12, 7), (12, 10), (16, 13), (20, 14), (24, 16), (28, 17)])
def test_minors_random_occupations_vs_oracle(handle, orc, m, k):
    rng = np.random.RandomState(31 * m + k)
    U = workloads.haar(m, 3 * m + k)
    for rep in range(4):
        s = np.array([1] * k + [0] * (m - k), dtype=np.int32) if rep == 0 else _occ(rng, m, k)
        t = _occ(rng, m, k - 1)
        if rep == 3:
            t[:] = 0
            t[rng.choice(m, k - 1, replace=False)] = 1   # collision-free outputs: the full 2^(k-2) walk
        got = handle.minors(U, s, t)
        want = orc.submatrices(U, s, t, orc.RYSER, "ld")
        scale = np.abs(want).max()
        assert np.abs(got - want).max() <= REL_TOL * scale, (rep, s, t)
        assert np.all(got[s == 0] == 0)


@pytest.mark.parametrize("k,collision_free", [(20, False), (20, True), (24, False)])
def test_minors_headline_sizes(handle, orc, k, collision_free):
    """BASELINE config 3 shape: m = 2k, collision-free input, k-1 outputs placed at random."""
    U, s, t = workloads.c3_step(k, 2 * k, collision_free)
    got = handle.minors(U, s, t)
    want = orc.submatrices(U, s, t, orc.RYSER, "ld")   # 80-bit: Ryser-form error ~1e-13 here
    assert np.abs(got - want).max() <= REL_TOL * np.abs(want).max()


def test_minors_equal_single_permanents_with_one_particle_removed(handle):
    """The property the reference tests (tests/test_bs_submatrices_permanent_calculators.py:75-109), as full
    complex numbers, against this package's own single-permanent kernel."""
    rng = np.random.RandomState(8)
    m, k = 10, 8
    U = workloads.haar(m, 11)
    s, t = _occ(rng, m, k), _occ(rng, m, k - 1)
    minors = handle.minors(U, s, t)
    idx = np.nonzero(s)[0]
    S = np.repeat(s[None, :], len(idx), axis=0).astype(np.uint8)
    S[np.arange(len(idx)), idx] -= 1
    T = np.repeat(t[None, :], len(idx), axis=0).astype(np.uint8)
    singles = handle.perm_batched(U, S, T)
    assert np.abs(minors[idx] - singles).max() <= 1e-13 * np.abs(singles).max()


def test_k_equals_one_edge_case(handle):
    U = workloads.haar(5, 2)
    s = np.array([0, 0, 1, 0, 0], dtype=np.int32)
    got = handle.minors(U, s, np.zeros(5, dtype=np.int32))
    assert np.array_equal(got, s.astype(np.complex128))   # bs_submatrices_permanent_calculator_base.py:157-158


@pytest.mark.parametrize("m,k", [(6, 1), (6, 2), (6, 4), (10, 7), (16, 11), (30, 15)])
def test_step_pmf_vs_oracle(handle, orc, m, k):
    rng = np.random.RandomState(m + k)
    U = workloads.haar(m, m * k)
    s, t = _occ(rng, m, k), _occ(rng, m, k - 1)
    pmf, minors = handle.gccb_pmf(U, s, t, want_minors=True)
    want = orc.gccb_pmf(U, s, t, "ld")
    assert abs(pmf.sum() - 1) <= 1e-14
    assert np.abs(pmf - want).max() <= 1e-12


def test_shape_error(handle):
    U = workloads.haar(4, 1)
    with pytest.raises(AttributeError):
        handle.minors(U, np.array([1, 1, 0, 0], dtype=np.int32), np.array([1, 1, 0, 0], dtype=np.int32))


LARGE_CASES = ["cf_k25", "bunched_k25", "cf_k26", "bunched_k26", "bunchedin_k26", "bunched_k28", "cf_k28", "bunchedin_k31", "dilated_k30"]


@pytest.mark.parametrize("name", LARGE_CASES)
def test_minors_beyond_one_lane_vs_long_double_fixtures(handle, golden_dir, name):
    """k = 25 ... 31 (the steps of BASELINE config 5(ii): two and four lanes share the columns of a term) against the oracle's
    80-bit sub-Ryser sweep (bs_cc_ryser_submatrices_permanent_calculator.py:80-119), computed once by
    tests/golden/make_minors_large_golden.py: m = 2k Haar unitaries with collision-free / bunched outputs, bunched inputs, and
    a k = 30 step in the 120-mode dilation of the config-5 lossy network."""
    z = np.load(os.path.join(golden_dir, "minors_large.npz"))
    if name not in list(z["names"]):
        pytest.fail(f"fixture {name} missing from minors_large.npz")
    k = int(name.split("_k")[1])
    U = z[f"{name}_U"] if f"{name}_U" in z.files else workloads.haar(2 * k, 900 + k)
    s, t, want = z[f"{name}_s"], z[f"{name}_t"], z[f"{name}_minors"]
    assert int(s.sum()) == k and int(t.sum()) == k - 1
    got = handle.minors(U, s, t)
    assert np.abs(got - want).max() <= REL_TOL * np.abs(want).max(), name
    assert np.all(got[s == 0] == 0)
    pmf = handle.gccb_pmf(U, s, t)
    amp = (s * want) @ U.T                                       # sum_i s_i P_i U[j][i]
    ref = np.abs(amp) ** 2
    assert np.abs(pmf - ref / ref.sum()).max() <= 1e-10


def test_k30_minors_in_the_dilated_network_satisfy_the_laplace_expansion(handle, golden_dir):
    """Config 5(ii) at full size (k = 30 in the 120-mode dilation): sum_i s_i P_i U[j][i] = perm(U; s, t + e_j), right-hand side
    from the batched single-permanent kernel (K2 walks the bunched output side of these items)."""
    z = np.load(os.path.join(golden_dir, "minors_large.npz"))
    U, s, t = z["dilated_k30_U"], z["dilated_k30_s"], z["dilated_k30_t"]
    minors = handle.minors(U, s, t)
    # outcomes that CAN happen: any detector mode, and the loss modes of occupied inputs whose particle is not lost yet (a loss
    # mode couples to exactly one input mode; every other placement has permanent 0 and no relative accuracy to speak of)
    loss_modes = [60 + int(np.argmax(np.abs(U[60:, i]))) for i in np.nonzero(s)[0]]
    js = [0, 17, 59] + [j for j in loss_modes if t[j] == 0][:3]
    assert len(js) == 6
    S = np.repeat(s[None].astype(np.uint8), len(js), axis=0)
    T = np.repeat(t[None].astype(np.uint8), len(js), axis=0)
    T[np.arange(len(js)), js] += 1
    singles = handle.perm_batched(U, S, T)
    for q, j in enumerate(js):
        lhs = np.sum(s * minors * U[j, :])
        assert abs(lhs - singles[q]) <= REL_TOL * abs(singles[q]), j
