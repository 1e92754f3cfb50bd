"""80-bit ground truths of the all-minors calculation at k = 25 ... 30 (the sizes BASELINE config 5(ii) reaches and the
Python reference cannot: 2^k Guan terms in CPython).  The oracle's sub-Ryser restatement
(oracle/oracle_impl.h, following bs_cc_ryser_submatrices_permanent_calculator.py:80-119) in x87 long double is pinned to the
reference at k <= 16 by tests/test_oracle_golden.py; here it runs the sizes whose GPU kernels (lane-split variants of
k3_minors_kernel) no smaller case reaches.  Minutes of CPU per case, so the results are committed:

    python tests/golden/make_minors_large_golden.py [case ...]      -> tests/golden/minors_large.npz

Cases: m = 2k Haar unitaries with collision-free and bunched outputs, a bunched input, and the 120-mode dilation of the
config-5 lossy network at k = 30 (the matrix is stored: it comes out of an SVD).
"""
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [REPO]

from tests import workloads  # noqa: E402


def _occ(rng, m, n):
    out = np.zeros(m, dtype=np.int32)
    for j in rng.randint(0, m, n):
        out[j] += 1
    return out


def case_inputs(name):
    kind, k = name.split("_")[0], int(name.split("_")[1][1:])
    rng = np.random.RandomState(7000 + k)
    if kind == "dilated":
        from theboss_b200.boson_sampling_utilities.boson_sampling_utilities import prepare_interferometer_matrix_in_expanded_space
        _, U_lossy, _ = workloads.c5_lossy(30, 60)
        U = np.ascontiguousarray(prepare_interferometer_matrix_in_expanded_space(U_lossy))
        s = np.array([1] * k + [0] * (120 - k), dtype=np.int32)
        # a POSSIBLE outcome: a loss mode (rows 60 ..) couples to exactly one input mode, so only the loss modes of
        # occupied inputs may hold a particle (any other placement has permanent 0): 11 lost particles + 18 detected ones
        t = np.zeros(120, dtype=np.int32)
        for i in rng.choice(k, 11, replace=False):
            t[60 + int(np.argmax(np.abs(U[60:, i])))] += 1
        t[:60:6] += _occ(rng, 10, k - 1 - 11)          # bunched detections (every sixth mode): a ~2^24-term walk for the oracle
        return U, s, t
    m = 2 * k
    U = workloads.haar(m, 900 + k)
    s = np.array([1] * k + [0] * (m - k), dtype=np.int32)
    if kind == "cf":
        t = np.zeros(m, dtype=np.int32)
        t[rng.choice(m, k - 1, replace=False)] = 1
    elif kind == "bunched":
        t = _occ(rng, m, k - 1)
    elif kind == "bunchedin":
        s = _occ(rng, m, k)
        t = _occ(rng, m, k - 1)
        if k > 28:   # k = 31: outputs bunched on 14 modes, so that the 80-bit Chin-Huh walk over them stays at ~1e7 terms per single
            t = np.zeros(m, dtype=np.int32)
            t[::4][:14] = _occ(rng, 14, k - 1)
    else:
        raise ValueError(name)
    return U, s, t


CASES = ["cf_k25", "bunched_k25", "cf_k26", "bunched_k26", "bunchedin_k26", "bunchedin_k31", "bunched_k28", "cf_k28", "dilated_k30"]


# Ryser-form sums lose about one bit per particle: the 80-bit sub-Ryser sweep is good to ~2^k * 5e-20 (1.5e-11 at k = 28,
# 1.2e-10 at k = 31).  The cases beyond k = 28 therefore take every minor from its definition -- the single permanent with one
# input particle removed -- in the Chin-Huh form (chin_huh_permanent_calculator.py:38-59, no cancellation growth), walked
# over the bunched OUTPUT occupation through perm(U; s, t) = perm(U^T; t, s).
CH_SINGLES = {"bunchedin_k31", "dilated_k30"}


def single(args):
    """One minor of a CH_SINGLES case: the single permanent with one particle of input mode i removed."""
    from oracle import pyoracle as orc
    name, i = args
    U, s, t = case_inputs(name)
    si = s.copy()
    si[i] -= 1
    return name, int(i), orc.guan_permanent(np.ascontiguousarray(U.T), t, si, orc.CHIN_HUH, "ld")


def run(name):
    from oracle import pyoracle as orc
    U, s, t = case_inputs(name)
    t0 = time.time()
    minors = orc.submatrices(U, s, t, orc.RYSER, "ld")
    return name, minors, time.time() - t0


def main():
    names = sys.argv[1:] or CASES
    path = os.path.join(HERE, "minors_large.npz")
    out = dict(np.load(path)) if os.path.exists(path) else {}

    def store(name, minors, secs):
        U, s, t = case_inputs(name)
        out[f"{name}_s"], out[f"{name}_t"], out[f"{name}_minors"] = s, t, minors
        if name.startswith("dilated"):
            out[f"{name}_U"] = U
        print(name, f"{secs:.0f} s", "max |minor|", np.abs(minors).max(), flush=True)
        out["names"] = np.array(sorted(set(list(out.get("names", [])) + [name])))
        np.savez_compressed(path, **out)

    with ProcessPoolExecutor(max_workers=max(1, (os.cpu_count() or 2) - 1)) as ex:
        for name in [n for n in names if n in CH_SINGLES]:        # one process per minor
            t0 = time.time()
            s = case_inputs(name)[1]
            minors = np.zeros(len(s), dtype=np.complex128)
            for _, i, val in ex.map(single, [(name, int(i)) for i in np.nonzero(s)[0]]):
                minors[i] = val
            store(name, minors, time.time() - t0)
        for name, minors, secs in ex.map(run, [n for n in names if n not in CH_SINGLES]):
            store(name, minors, secs)


if __name__ == "__main__":
    main()
