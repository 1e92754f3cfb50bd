"""Helper of tests/test_gpu_dispatch.py: runs fixed sampling jobs in a fresh process (the K3 dispatch knobs
BP_K3_* are read once per process) and stores the samples in the .npz given as argv[1]."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import workloads  # noqa: E402
from theboss_b200 import _native  # noqa: E402


def jobs():
    rng = np.random.RandomState(2025)
    U = workloads.haar(20, 14)
    s = np.array([1] * 14 + [0] * 6, dtype=np.int32)
    tape = rng.random_sample((40, 1 + 2 * 14))
    return U, s, tape


if __name__ == "__main__":
    h = _native.default_handle(0)
    U, s, tape = jobs()
    out = {
        "plain": h.gccb_simulate(U, s, tape.shape[0], tape=tape),
        "lossy": h.gccb_simulate(U, s, tape.shape[0], eta=0.8, tape=tape),
        "philox": h.gccb_simulate(U, s, 2000, seed=77),
        "philox_lossy": h.gccb_simulate(U, s, 2000, eta=0.6, seed=78),
    }
    U2, s2, t2 = workloads.c3_step(12, 24)
    pmf, minors = h.gccb_pmf(U2, s2, t2, want_minors=True)
    out["pmf"], out["minors"] = pmf, minors
    np.savez(sys.argv[1], **out)
