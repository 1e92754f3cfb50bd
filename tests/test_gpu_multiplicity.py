"""K2 parity (GPU): batched permanents with input/output multiplicities through the C ABI and the
Ryser / Chin-Huh / Classic / Glynn calculator classes, vs the reference golden vectors and the oracle."""
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
    out = np.zeros(m, dtype=np.uint8)
    for j in rng.randint(0, m, n):
        out[j] += 1
    return out


def test_calculator_classes_against_reference_golden(golden_dir):
    """Same cases as the reference's tests/test_bs_permanent_calculators.py:88-122 plus seeded extras; the
    expected values are the reference's own outputs."""
    from theboss_b200.boson_sampling_utilities.permanent_calculators.bs_permanent_calculator_factory import (
        BSPermanentCalculatorFactory, PermanentCalculatorType)
    z = np.load(os.path.join(golden_dir, "single_permanents.npz"))
    for i in range(int(z["n_cases"])):
        U, s, t = z[f"U_{i}"], z[f"s_{i}"], z[f"t_{i}"]
        ref = z[f"chin_huh_{i}"]
        scale = max(abs(ref), 1e-30)
        for typ in PermanentCalculatorType:
            got = BSPermanentCalculatorFactory(U, list(s), list(t), typ).generate_calculator().compute_permanent()
            assert isinstance(got, np.complex128)
            assert abs(got - ref) <= REL_TOL * scale, (i, typ.name, got, ref)
            assert abs(got - z[f"glynn_{i}"]) <= REL_TOL * scale
            assert abs(got - z[f"ryser_{i}"]) <= 1e-8 * scale   # the reference's Ryser is the inaccurate one


def test_shape_errors_follow_the_reference():
    from theboss_b200.boson_sampling_utilities.permanent_calculators.chin_huh_permanent_calculator import ChinHuhPermanentCalculator
    from theboss_b200.boson_sampling_utilities.permanent_calculators.ryser_permanent_calculator import RyserPermanentCalculator
    U = workloads.haar(4, 3)
    with pytest.raises(AttributeError):   # bs_permanent_calculator_base.py:179-180
        RyserPermanentCalculator(U, [1, 1, 0], [1, 1, 0, 0]).compute_permanent()
    with pytest.raises(AttributeError):
        ChinHuhPermanentCalculator(U[:3], [1, 1, 0, 0], [1, 1, 0, 0]).compute_permanent()
    with pytest.raises(AttributeError):   # particle numbers differ: BP_ERR_SHAPE
        ChinHuhPermanentCalculator(U, [1, 1, 0, 0], [1, 0, 0, 0]).compute_permanent()
    # shorter-than-m states of equal length are legal (:61-72)
    a = ChinHuhPermanentCalculator(U, [1, 1, 0], [0, 1, 1]).compute_permanent()
    b = ChinHuhPermanentCalculator(U, [1, 1, 0, 0], [0, 1, 1, 0]).compute_permanent()
    assert a == b


def test_list_of_lists_and_float_states():
    from theboss_b200.boson_sampling_utilities.permanent_calculators.ryser_permanent_calculator import RyserPermanentCalculator
    U = workloads.haar(3, 9)
    a = RyserPermanentCalculator(U.tolist(), (1.0, 2.0, 0.0), np.array([0.0, 1.0, 2.0])).compute_permanent()
    b = RyserPermanentCalculator(U, [1, 2, 0], [0, 1, 2]).compute_permanent()
    assert a == b


@pytest.mark.parametrize("m,n", [(6, 1), (6, 2), (8, 5), (10, 8), (12, 12), (16, 14), (20, 16)])
def test_batched_random_occupations_vs_oracle(handle, orc, m, n):
    rng = np.random.RandomState(100 * m + n)
    U = workloads.haar(m, m + n)
    B = 24
    S = np.array([_occ(rng, m, n) for _ in range(B)])
    T = np.array([_occ(rng, m, n) for _ in range(B)])
    S[0] = T[0] = np.array([1] * n + [0] * (m - n), dtype=np.uint8) if n <= m else S[0]
    S[1, :] = 0; S[1, 2] = n          # everything in one input mode
    T[2, :] = 0; T[2, m - 1] = n      # everything in one output mode
    got = handle.perm_batched(U, S, T)
    for b in range(B):
        want = orc.guan_permanent(U, S[b], T[b], orc.CHIN_HUH, "ld")
        assert abs(got[b] - want) <= REL_TOL * abs(want) + 1e-300, (b, S[b], T[b], got[b], want)
        assert abs(got[b] - want) <= 1e-12 * abs(want) + 1e-300


def test_mixed_particle_numbers_and_empty_items(handle, orc):
    rng = np.random.RandomState(5)
    m = 9
    U = workloads.haar(m, 77)
    ns = [0, 1, 3, 0, 7, 2, 5, 7, 1, 4, 6, 0]
    S = np.array([_occ(rng, m, n) for n in ns])
    T = np.array([_occ(rng, m, n) for n in ns])
    got = handle.perm_batched(U, S, T)
    for b, n in enumerate(ns):
        want = 1.0 if n == 0 else orc.guan_permanent(U, S[b], T[b], orc.CHIN_HUH, "ld")
        assert abs(got[b] - want) <= 1e-12 * abs(want)


def test_c2_subset_n20(handle, orc):
    """BASELINE config 2 (n=20, m=40, repeated rows and columns), first 48 of the 10^4 items."""
    U, S, T = workloads.c2_batch(items=48)
    got = handle.perm_batched(U, S, T)
    for b in range(48):
        want = orc.guan_permanent(U, S[b], T[b], orc.CHIN_HUH, "ld")
        assert abs(got[b] - want) <= REL_TOL * abs(want), (b, got[b], want)


def test_collision_free_n24_equals_k1(handle):
    """K2 on a collision-free item is the same Glynn sum as K1."""
    n = 22
    U = workloads.haar(2 * n, n)
    rows = np.sort(np.random.RandomState(n).choice(2 * n, n, replace=False))
    s = np.zeros(2 * n, dtype=np.uint8); s[:n] = 1
    t = np.zeros(2 * n, dtype=np.uint8); t[rows] = 1
    a = handle.perm_batched(U, s[None], t[None])[0]
    b = handle.glynn_matrix(workloads.c4_matrix(n))
    assert abs(a - b) <= 1e-11 * abs(b)   # both paths sit ~1e-12 from the long-double truth


def test_batched_shape_error(handle):
    U = workloads.haar(4, 1)
    S = np.array([[1, 1, 0, 0], [1, 0, 0, 0]], dtype=np.uint8)
    T = np.array([[0, 1, 1, 0], [1, 1, 0, 0]], dtype=np.uint8)
    with pytest.raises(AttributeError):
        handle.perm_batched(U, S, T)


def test_device_pointer_entry_points_match_host_entry_points(orc):
    """`_dev` variants (inputs resident in HBM, caller-owned stream) against the host-pointer calls."""
    import torch
    from theboss_b200 import _native
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream(0)
    h = _native.Handle(0, stream_ptr=stream.cuda_stream)
    rng = np.random.RandomState(4)
    m, n, B = 10, 7, 40
    U = workloads.haar(m, 3)
    S = np.array([_occ(rng, m, n) for _ in range(B)])
    T = np.array([_occ(rng, m, n) for _ in range(B)])
    dU = torch.from_numpy(U.view(np.float64).copy()).to(dev)
    dS, dT = torch.from_numpy(S).to(dev), torch.from_numpy(T).to(dev)
    d_out = torch.zeros(2 * B, dtype=torch.float64, device=dev)
    h.perm_batched_dev(dU.data_ptr(), m, dS.data_ptr(), dT.data_ptr(), B, _native.FORMULA_GLYNN, d_out.data_ptr())
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().view(np.complex128)
    want = _native.default_handle(0).perm_batched(U, S, T)
    assert np.array_equal(got, want)
    # K1 range on device pointers
    A = workloads.c4_matrix(14)
    dA = torch.from_numpy(A.view(np.float64).copy()).to(dev)
    d_part = torch.zeros(4, dtype=torch.float64, device=dev)
    h.glynn_matrix_range_dev(dA.data_ptr(), 14, 0, 1 << 13, d_part.data_ptr())
    torch.cuda.synchronize()
    p = d_part.cpu().numpy()
    got1 = complex(p[0] + p[1], p[2] + p[3]) / (1 << 13)
    want1 = orc.glynn_matrix(A, "ld")
    assert abs(got1 - want1) <= 1e-12 * abs(want1)
    h.close()
