"""K1 parity (GPU): Gray-code Glynn through the C ABI vs the CPU oracle and the golden fixtures."""
import json
import os

import numpy as np
import pytest

from tests import workloads

pytestmark = pytest.mark.gpu

REL_TOL = 1e-10   # BASELINE.json north_star: relative tolerance on complex permanents


@pytest.fixture(scope="module")
def handle():
    from theboss_b200 import _native
    return _native.default_handle(0)


@pytest.fixture(scope="module")
def orc():
    from oracle import pyoracle
    return pyoracle


def _rel(a, b):
    return abs(a - b) / abs(b)


@pytest.mark.parametrize("N", list(range(1, 19)))
def test_glynn_matrix_matches_oracle(handle, orc, N):
    rng = np.random.RandomState(N)
    for kind in ("gauss", "haar"):
        A = rng.randn(N, N) + 1j * rng.randn(N, N) if kind == "gauss" else workloads.c4_matrix(N) if N > 1 else np.array([[0.3 - 0.4j]])
        got = handle.glynn_matrix(A)
        want = orc.glynn_matrix(A, "ld", nthreads=4 if N > 12 else 1)
        assert _rel(got, want) <= REL_TOL, (N, kind, got, want)
        assert _rel(got, want) <= 1e-12, (N, kind, got, want)   # what the kernel actually achieves


def test_empty_matrix_is_one(handle):
    assert handle.glynn_matrix(np.zeros((0, 0), dtype=np.complex128)) == 1   # glynn_gray_permanent_calculator.py:52-53


@pytest.mark.parametrize("n", [20, 24, 26, 30])
def test_headline_sizes_against_long_double_fixture(handle, golden_dir, n):
    with open(os.path.join(golden_dir, "large_permanents.json")) as f:
        g = json.load(f)[f"glynn_n{n}"]
    got = handle.glynn_matrix(workloads.c4_matrix(n))
    assert _rel(got, complex(g["re"], g["im"])) <= REL_TOL


@pytest.mark.parametrize("N,shards", [(12, 2), (12, 3), (16, 4), (16, 8), (21, 8)])
def test_range_partials_cover_the_term_space(handle, orc, N, shards):
    """bp_glynn_matrix_range shards (the multi-GPU split) sum to the full permanent, and every shard
    equals the oracle's partial over the same Gray steps."""
    A = workloads.c4_matrix(N)
    T = 1 << (N - 1)
    edges = [T * i // shards for i in range(shards + 1)]
    total = 0
    for lo, hi in zip(edges[:-1], edges[1:]):
        p = handle.glynn_matrix_range(A, lo, hi)
        part = complex(p[0] + p[1], p[2] + p[3])
        if N <= 16:
            want = orc.glynn_range(A, lo, hi, "ld")
            assert abs(part - want) <= 1e-11 * max(abs(want), 1e-3)
        total += part
    assert _rel(total / T, handle.glynn_matrix(A)) <= 1e-13


@pytest.mark.parametrize("N", [35, 36, 39, 40])
def test_wide_matrices_on_step_ranges_vs_oracle(handle, orc, N):
    """N = 35 ... 40 (glynn_pair4_kernel: two warps share the columns of a Gray stream and trade half-products under a named barrier):
    partial sums over step ranges against the 80-bit oracle -- aligned bulk, bulk + unaligned head and tail (generic kernel), a range
    deep inside the term space (rows >= 30 flip), and one short enough to leave most warp pairs without work."""
    A = workloads.c4_matrix(N)
    for lo, hi in ((0, 1 << 18), (1 << 20, (1 << 20) + (1 << 17)), (3, (1 << 17) + 77), ((1 << 30) + 4096, (1 << 30) + 4096 + (1 << 16)),
                   ((1 << (N - 1)) - (1 << 16), 1 << (N - 1))):
        p = handle.glynn_matrix_range(A, lo, hi)
        got = complex(p[0] + p[1], p[2] + p[3])
        want = orc.glynn_range(A, lo, hi, "ld")
        # (partial sums cancel less than the permanent: 1e-11 of the largest term magnitude is the meaningful bar)
        assert abs(got - want) <= 1e-11 * max(abs(want), 1e-3 * abs(got) + 1e-300), (N, lo, hi, got, want)


def test_wide_matrix_shards_agree_with_the_unsharded_walk(handle):
    """The multi-GPU split at N = 36: eight range partials over 2^24 steps sum to the partial of the whole range (both through the
    warp-pair kernel, different spans per lane)."""
    A = workloads.c4_matrix(36)
    T = 1 << 24
    whole = handle.glynn_matrix_range(A, 0, T)
    parts = [handle.glynn_matrix_range(A, T * i // 8, T * (i + 1) // 8) for i in range(8)]
    tot = sum(complex(p[0] + p[1], p[2] + p[3]) for p in parts)
    ref = complex(whole[0] + whole[1], whole[2] + whole[3])
    assert abs(tot - ref) <= 1e-12 * abs(ref)


def test_unaligned_ranges(handle, orc):
    A = workloads.c4_matrix(11)
    for lo, hi in [(0, 1), (1, 2), (3, 70), (65, 129), (17, 1024), (1000, 1024), (5, 5)]:
        p = handle.glynn_matrix_range(A, lo, hi)
        got = complex(p[0] + p[1], p[2] + p[3])
        want = orc.glynn_range(A, lo, hi, "ld") if hi > lo else 0
        assert abs(got - want) <= 1e-12 * max(abs(want), 1e-6), (lo, hi)


def test_glynn_calculator_against_reference_golden(golden_dir):
    """The reference's own outputs (all four calculators) on 50 (U, s, t) cases."""
    from theboss_b200.boson_sampling_utilities.permanent_calculators.glynn_gray_permanent_calculator import (
        GlynnGrayPermanentCalculator,
    )
    z = np.load(os.path.join(golden_dir, "single_permanents.npz"))
    for i in range(int(z["n_cases"])):
        U, s, t = z[f"U_{i}"], z[f"s_{i}"], z[f"t_{i}"]
        got = GlynnGrayPermanentCalculator(U, list(s), list(t)).compute_permanent()
        assert isinstance(got, np.complex128)
        ref = z[f"glynn_{i}"]
        scale = max(abs(ref), 1e-30)
        assert abs(got - ref) <= REL_TOL * scale, i
        assert abs(got - z[f"chin_huh_{i}"]) <= REL_TOL * scale, i


def test_matrix_is_reread_after_inplace_mutation():
    """Reference tests scale calculator.matrix in place (tests/gcc_based_strategies_tests_base.py:89-92)."""
    from theboss_b200.boson_sampling_utilities.permanent_calculators.glynn_gray_permanent_calculator import (
        GlynnGrayPermanentCalculator,
    )
    U = workloads.haar(5, 1)
    calc = GlynnGrayPermanentCalculator(U, [1, 1, 1, 0, 0], [0, 1, 1, 1, 0])
    p1 = calc.compute_permanent()
    calc.matrix *= 0.5
    assert calc.matrix is U
    assert abs(calc.compute_permanent() - p1 / 8) <= 1e-14 * abs(p1)
