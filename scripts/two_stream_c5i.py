"""Experiment: a uniform-loss run (C5(i), n = 30, m = 60, eta = 0.5) as ONE request on one handle vs two halves on two handles (two
streams, two host threads; Philox keyed by the global sample index, so the samples are the same)."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import workloads
from theboss_b200 import _native

U, _, s = workloads.c5_lossy(30, 60)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
h0, h1, h2, h3 = _native.Handle(0), _native.Handle(0), _native.Handle(0), _native.Handle(0)


def one():
    return h0.gccb_simulate(U, s, S, eta=0.5, seed=5)


def split(handles):
    k = len(handles)
    out = [None] * k
    bounds = [S * i // k for i in range(k + 1)]

    def work(i):
        out[i] = handles[i].gccb_simulate(U, s, bounds[i + 1] - bounds[i], eta=0.5, seed=5, first_sample=bounds[i])
    th = [threading.Thread(target=work, args=(i,)) for i in range(k)]
    [t.start() for t in th]
    [t.join() for t in th]
    return np.concatenate(out)


ref = one()
for name, fn in (("one handle", one), ("two handles", lambda: split([h0, h1])), ("four handles", lambda: split([h0, h1, h2, h3]))):
    fn()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); out = fn(); ts.append(time.perf_counter() - t0)
    print(f"{name}: {min(ts) * 1e3:.3f} ms for {S} samples, identical {np.array_equal(out, ref)}", flush=True)
