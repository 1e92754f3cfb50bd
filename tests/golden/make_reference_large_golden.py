"""Outputs of the UNMODIFIED reference at the largest sizes its Python loops finish in minutes, on the BASELINE workloads:
config 2 items (n = 20, m = 40, repeated rows and columns) through ChinHuhPermanentCalculator and
GlynnGrayPermanentCalculator, and config 3 steps (all minors, k = 12 .. 16) through BSCCRyserSubmatricesPermanentCalculator
(k <= 14 also through the 5x slower BSCCCHSubmatricesPermanentCalculator).

Run in the build container only (the GPU box has no /root/reference); about 2 minutes on 8 cores:

    python tests/golden/make_reference_large_golden.py

Result: tests/golden/reference_large.json.  tests/test_oracle_golden.py pins the oracle to it, and
tests/test_gpu_zz_reference_runs.py compares kernels K2 / K3 with it at BASELINE's tolerance of 1e-10.
(The reference's Ryser single-permanent calculator is left out on purpose: its own float64 error is 2e-10 at n = 20,
SURVEY.md Appendix C.)
"""
import json
import multiprocessing
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("THEBOSS_REFERENCE", "/root/reference")
sys.path[:0] = [REPO, REF, os.path.join(REPO, "oracle", "refshim")]

from tests import workloads  # noqa: E402

C2_ITEMS = 6
C3_STEPS = [(12, False), (12, True), (14, False), (14, True), (16, False), (16, True)]   # (k, collision-free outputs)


def job(spec):
    from theboss.boson_sampling_utilities.permanent_calculators import (
        bs_cc_ch_submatrices_permanent_calculator as sub_ch, bs_cc_ryser_submatrices_permanent_calculator as sub_ryser,
        chin_huh_permanent_calculator as ch, glynn_gray_permanent_calculator as glynn)
    kind, t0 = spec[0], time.time()
    if kind == "c2":
        _, item, which = spec
        U, S, T = workloads.c2_batch(20, 40, C2_ITEMS)
        cls = ch.ChinHuhPermanentCalculator if which == "chin_huh" else glynn.GlynnGrayPermanentCalculator
        value = cls(U, [int(x) for x in S[item]], [int(x) for x in T[item]]).compute_permanent()
        return spec, [value.real, value.imag], time.time() - t0
    _, k, free, which = spec
    U, s, t = workloads.c3_step(k, 2 * k, collision_free=free)
    cls = sub_ryser.BSCCRyserSubmatricesPermanentCalculator if which == "ryser" else sub_ch.BSCCCHSubmatricesPermanentCalculator
    values = cls(U, [int(x) for x in s], [int(x) for x in t]).compute_permanents()
    return spec, [[v.real, v.imag] for v in values], time.time() - t0


def main():
    specs = [("c2", i, which) for i in range(C2_ITEMS) for which in ("chin_huh", "glynn")]
    specs += [("c3", k, free, "ryser") for k, free in C3_STEPS]
    specs += [("c3", k, free, "chin_huh") for k, free in C3_STEPS if k <= 14]
    out = {"generator": "tests/golden/make_reference_large_golden.py", "c2": {"n": 20, "m": 40, "items": C2_ITEMS, "chin_huh": {}, "glynn": {}},
           "c3": {}}
    with multiprocessing.get_context("spawn").Pool(min(8, os.cpu_count() or 1)) as pool:
        for spec, value, seconds in pool.imap_unordered(job, specs):
            print(spec, f"{seconds:.1f} s", flush=True)
            if spec[0] == "c2":
                out["c2"][spec[2]][str(spec[1])] = value
            else:
                out["c3"].setdefault(f"k{spec[1]}_{'free' if spec[2] else 'bunched'}", {})[spec[3]] = value
    with open(os.path.join(HERE, "reference_large.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
