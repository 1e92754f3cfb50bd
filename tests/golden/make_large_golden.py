"""Ground-truth values for the headline sizes, from the 80-bit long-double CPU oracle.

The Python reference cannot run these sizes (~2.3 h for one n=30 permanent),
so the fixtures come from oracle/bossperm_oracle.c (`_ld` variant), which tests/test_oracle_golden.py
pins to the reference at small n.  Inputs are rebuilt from seeds by tests/workloads.py.

    python tests/golden/make_large_golden.py
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import pyoracle as orc  # noqa: E402
from tests import workloads  # noqa: E402

out = {}
for n in (20, 24, 26, 30):
    A = workloads.c4_matrix(n)
    t0 = time.time()
    v = orc.glynn_matrix(A, "ld", nthreads=os.cpu_count() or 1, nchunks=256)
    out[f"glynn_n{n}"] = {"re": v.real, "im": v.imag, "seconds": round(time.time() - t0, 2),
                          "precision": "x87 long double, 256 chunks summed in order"}
    print(n, v, out[f"glynn_n{n}"]["seconds"], "s", flush=True)
with open(os.path.join(HERE, "large_permanents.json"), "w") as f:
    json.dump(out, f, indent=1)
