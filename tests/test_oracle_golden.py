"""Pins the CPU oracle (oracle/) to the Python reference: every routine is compared with fixtures
that tests/golden/make_golden.py produced by running the unmodified reference, and with the
reference's own literal known-answer vector (tests/test_exact_distribution_calculator.py:75-142 in
the reference).  CPU only."""
import json
import os
from math import factorial

import numpy as np
import pytest
from scipy.special import binom

from oracle import pyoracle as orc

TOL_D = 1e-12   # double variant follows the reference's operation order: agreement ~1e-15 expected
TOL_LD = 1e-9   # long double vs the reference: limited by the *reference's* float64 error (Ryser)


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def test_single_permanents_double_variant(golden_dir):
    z = _load(golden_dir, "single_permanents.npz")
    for i in range(int(z["n_cases"])):
        U, s, t = z[f"U_{i}"], z[f"s_{i}"], z[f"t_{i}"]
        scale = max(abs(z[f"glynn_{i}"]), 1e-30)
        assert abs(orc.glynn(U, s, t, "d") - z[f"glynn_{i}"]) <= TOL_D * scale, i
        assert abs(orc.guan_permanent(U, s, t, orc.RYSER, "d") - z[f"ryser_{i}"]) <= 50 * TOL_D * scale, i
        assert abs(orc.guan_permanent(U, s, t, orc.CHIN_HUH, "d") - z[f"chin_huh_{i}"]) <= TOL_D * scale, i
        if f"classic_{i}" in z:
            assert abs(orc.classic(U, s, t) - z[f"classic_{i}"]) <= TOL_D * scale, i


def test_single_permanents_long_double_is_consistent(golden_dir):
    """All four reference calculators and the 80-bit oracle agree on the mathematical quantity."""
    z = _load(golden_dir, "single_permanents.npz")
    for i in range(int(z["n_cases"])):
        U, s, t = z[f"U_{i}"], z[f"s_{i}"], z[f"t_{i}"]
        truth = orc.glynn(U, s, t, "ld")
        scale = max(abs(truth), 1e-30)
        for name in ("glynn", "chin_huh", "ryser", "classic"):
            if f"{name}_{i}" in z:
                assert abs(z[f"{name}_{i}"] - truth) <= TOL_LD * scale, (i, name)
        assert abs(orc.guan_permanent(U, s, t, orc.CHIN_HUH, "ld") - truth) <= 1e-13 * scale
        assert abs(orc.guan_permanent(U, s, t, orc.RYSER, "ld") - truth) <= 1e-11 * scale


def test_zero_particles_and_empty_edge_cases():
    U = np.eye(3, dtype=np.complex128)
    assert orc.glynn(U, [0, 0, 0], [0, 0, 0]) == 1          # glynn_gray_permanent_calculator.py:52-53
    assert orc.classic(U, [0, 0, 0], [0, 0, 0]) == 1        # classic_permanent_calculator.py:29-31
    assert orc.classic(U, [0, 0, 0], [1, 0, 0]) == 0        # classic_permanent_calculator.py:32-33
    assert orc.guan_permanent(U, [0, 0, 0], [0, 0, 0], orc.RYSER) == 1
    assert orc.guan_permanent(U, [0, 0, 0], [0, 0, 0], orc.CHIN_HUH) == 1


def test_submatrices(golden_dir):
    z = _load(golden_dir, "submatrices_permanents.npz")
    for i in range(int(z["n_cases"])):
        U, s, t = z[f"U_{i}"], z[f"s_{i}"], z[f"t_{i}"]
        ry, ch = z[f"ryser_{i}"], z[f"chin_huh_{i}"]
        scale = max(np.abs(ch).max(), 1e-30)
        assert np.abs(orc.submatrices(U, s, t, orc.RYSER, "d") - ry).max() <= 50 * TOL_D * scale, i
        assert np.abs(orc.submatrices(U, s, t, orc.CHIN_HUH, "d") - ch).max() <= TOL_D * scale, i
        truth = orc.submatrices(U, s, t, orc.CHIN_HUH, "ld")
        assert np.abs(truth - ch).max() <= TOL_LD * scale
        assert np.abs(orc.submatrices(U, s, t, orc.RYSER, "ld") - truth).max() <= 1e-11 * scale


def test_submatrices_equal_single_permanents_with_one_particle_removed(golden_dir):
    """The property the reference tests (tests/test_bs_submatrices_permanent_calculators.py:75-109),
    here as full complex numbers instead of abs() to 7 decimals."""
    z = _load(golden_dir, "submatrices_permanents.npz")
    for i in range(int(z["n_cases"])):
        U, s, t = z[f"U_{i}"], z[f"s_{i}"].astype(int), z[f"t_{i}"].astype(int)
        if s.sum() == 1:
            continue
        minors = orc.submatrices(U, s, t, orc.CHIN_HUH, "ld")
        for v in range(len(s)):
            if s[v] == 0:
                assert minors[v] == 0
                continue
            s2 = s.copy()
            s2[v] -= 1
            single = orc.glynn(U, s2, t, "ld")
            assert abs(minors[v] - single) <= 1e-13 * max(abs(single), np.abs(minors).max()), (i, v)


def test_exact_distribution_known_answer(golden_dir):
    """Reference literal: tests/test_exact_distribution_calculator.py:75-142 (through Chin-Huh)."""
    with open(os.path.join(golden_dir, "exact_distribution.json")) as f:
        g = json.load(f)
    P = np.array(g["matrix_real"], dtype=np.complex128)
    s0, eta, n = g["initial_state"], g["eta"], sum(g["initial_state"])
    dist = []
    for outcome in g["outcomes"]:
        l = sum(outcome)
        w = binom(n, l) * eta ** l * (1 - eta) ** (n - l)   # bs_exact_distribution_with_uniform_losses.py:50-60
        if l == 0:
            dist.append(w)
            continue
        p = 0.0
        for li in g["lossy_inputs"][str(l)]:                # bs_distribution_calculator_with_fixed_losses.py:119-151
            mult = int(np.prod([binom(s0[i], s0[i] - li[i]) for i in range(len(s0))]))
            sub = abs(orc.guan_permanent(P, li, outcome, orc.CHIN_HUH, "d")) ** 2
            for occ in li:
                sub /= factorial(occ)
            p += sub * mult
        p /= factorial(l)
        p /= binom(n, l)
        p *= factorial(l)                                    # :99-104
        for occ in outcome:
            p /= factorial(occ)
        dist.append(p * w)
    assert np.allclose(dist, g["reference_literal"])
    assert np.allclose(dist, g["reference_computed"], rtol=1e-13, atol=1e-16)


def test_gccb_tape_samples(golden_dir):
    z = _load(golden_dir, "gccb_samples.npz")
    for name in z["plain_names"]:
        U, s, tape = z[f"{name}_U"], z[f"{name}_s"], z[f"{name}_tape"]
        samples, pmfs = orc.gccb_simulate(U, s, tape, return_pmfs=True)
        assert np.array_equal(np.array(samples), z[f"{name}_samples"]), name
        assert np.abs(np.array(pmfs) - z[f"{name}_pmfs"]).max() <= 1e-12, name
    samples = orc.gccb_uniform_losses_simulate(z["uniform_U"], z["uniform_s"], float(z["uniform_eta"]), z["uniform_tape"])
    assert np.array_equal(np.array(samples), z["uniform_samples"])
    assert np.abs(orc.expanded_matrix(z["lossynet_U"]) - z["lossynet_expanded"]).max() <= 1e-14
    samples = orc.lossy_net_simulate(z["lossynet_U"], list(z["lossynet_s"]), z["lossynet_tape"])
    assert np.array_equal(np.array(samples), z["lossynet_samples"])


def test_gcc_tape_samples(golden_dir):
    z = _load(golden_dir, "gcc_samples.npz")
    for name, calc in (("c1_glynn", "glynn"), ("bunched_ryser", "ryser")):
        U, s, uni = z[f"{name}_U"], [int(x) for x in z[f"{name}_s"]], z[f"{name}_uniforms"]
        samples = orc.gcc_simulate(U, s, uni, calculator=calc)
        assert np.array_equal(np.array(samples), z[f"{name}_samples"].astype(np.int64)), name
        keys, vals = z[f"{name}_pmf_keys"], z[f"{name}_pmf_vals"]
        for k, v in list(zip(keys, vals))[:40]:
            got = orc.gcc_layer_pmf(U, s, [int(x) for x in k], calculator=calc)
            assert np.abs(got - v).max() <= 1e-13 * max(v.max(), 1e-30)


@pytest.mark.parametrize("n", [10, 14])
def test_parallel_glynn_matches_sequential(n):
    rng = np.random.RandomState(n)
    A = rng.randn(n, n) + 1j * rng.randn(n, n)
    seq = orc.glynn_matrix(A, "ld")
    par = orc.glynn_matrix(A, "ld", nthreads=4)
    assert _rel(par, seq) <= 1e-15
    T = 1 << (n - 1)
    parts = [orc.glynn_range(A, lo, lo + T // 4) for lo in range(0, T, T // 4)]
    assert _rel(sum(parts) / T, seq) <= 1e-14


def test_reference_outputs_at_n14_to_n20_pin_the_oracle_and_the_large_ground_truth(golden_dir):
    """The unmodified reference was also RUN on the C4 matrix family at the largest sizes Python finishes
    (scripts/time_reference_python.py, N = 14 .. 20; GlynnGrayPermanentCalculator, glynn_gray_permanent_calculator.py:41-84):
    its outputs must agree with the oracle in both precisions and -- at N = 20, the size where the two overlap --
    with the long-double ground truth that the GPU parity tests of the n = 20 .. 30 permanents are anchored on."""
    from tests import workloads
    path = os.path.join(os.path.dirname(os.path.dirname(golden_dir)), "profiles", "r01_reference_python_cpu.json")
    with open(path) as f:
        ref = json.load(f)["single_core"]["c4_glynn_single_permanent"]
    sizes = sorted(int(k) for k in ref if k.isdigit())
    assert sizes and max(sizes) >= 20
    for n in sizes:
        want = complex(ref[str(n)]["re"], ref[str(n)]["im"])
        A = workloads.c4_matrix(n)
        assert _rel(orc.glynn_matrix(A, "d"), want) <= 1e-10, n     # same algorithm, same precision (summation order differs)
        assert _rel(orc.glynn_matrix(A, "ld"), want) <= 1e-10, n    # the reference's own float64 error at N = 20 is ~1e-11
    with open(os.path.join(golden_dir, "large_permanents.json")) as f:
        big = json.load(f)["glynn_n20"]
    assert _rel(complex(big["re"], big["im"]), complex(ref["20"]["re"], ref["20"]["im"])) <= 1e-10


def test_reference_outputs_on_config_2_and_3_workloads_pin_the_oracle(golden_dir):
    """tests/golden/reference_large.json: the unmodified reference on BASELINE config 2 items (n = 20, m = 40, repeated rows
    and columns; Chin-Huh and Glynn calculators) and config 3 steps (all minors at k = 12 .. 16, sub-Ryser and sub-Chin-Huh).
    The oracle's double variant follows the reference's operation order and reproduces these values to the last bit; the
    80-bit variant differs by the reference's own float64 error (up to 6.5e-11 on these items)."""
    from tests import workloads
    with open(os.path.join(golden_dir, "reference_large.json")) as f:
        g = json.load(f)
    U, S, T = workloads.c2_batch(g["c2"]["n"], g["c2"]["m"], g["c2"]["items"])
    worst = 0.0
    for which in ("chin_huh", "glynn"):
        assert len(g["c2"][which]) == g["c2"]["items"]
        for i, (re, im) in g["c2"][which].items():
            s, t, want = S[int(i)].astype(np.int32), T[int(i)].astype(np.int32), complex(re, im)
            same_order = orc.guan_permanent(U, s, t, orc.CHIN_HUH, "d") if which == "chin_huh" else orc.glynn(U, s, t, "d")
            assert _rel(same_order, want) <= 1e-14, (which, i)
            worst = max(worst, _rel(orc.guan_permanent(U, s, t, orc.CHIN_HUH, "ld"), want))
    assert worst <= 1e-10
    assert len(g["c3"]) == 6
    for key, values in g["c3"].items():
        k, free = int(key[1:key.index("_")]), key.endswith("free")
        U3, s, t = workloads.c3_step(k, 2 * k, collision_free=free)
        truth = orc.submatrices(U3, s, t, orc.RYSER, "ld")
        for which, v in values.items():
            want = np.array([complex(*x) for x in v])
            same_order = orc.submatrices(U3, s, t, orc.RYSER if which == "ryser" else orc.CHIN_HUH, "d")
            assert np.abs(same_order - want).max() <= 1e-14 * np.abs(want).max(), (key, which)
            assert np.abs(truth - want).max() <= 1e-10 * np.abs(want).max(), (key, which)
