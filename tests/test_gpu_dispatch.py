"""K3 dispatch (GPU): every block shape / work split the launcher can choose must give the same answers.

The knobs (BP_K3_WARP_MAX_K: one-warp blocks for small steps, BP_K3_TREE_MAX_C: columns per lane, i.e. from which k two and
four lanes share a term stream, BP_K3_TPG / BP_K3_CAP: terms per lane group and chunk blocks per sample) are read once per process, so every setting runs the
same jobs in a fresh interpreter (tests/_knob_job.py).  The default setting is checked against the oracle's sampling
loop on the same decision tape; the others against the default.  Tiny BP_K3_TPG values force many chunk blocks per
sample at small k, which exercises the multi-sample 512-thread chunk reduction of the finish kernel and the
launch-slot indirection of ragged (lossy) runs."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(tmp_path, name, **env):
    path = os.path.join(str(tmp_path), name + ".npz")
    e = {k: v for k, v in os.environ.items() if not k.startswith("BP_K3_")}
    e.update({k: str(v) for k, v in env.items()})
    subprocess.run([sys.executable, os.path.join(REPO, "tests", "_knob_job.py"), path], check=True, env=e, cwd=REPO, timeout=600)
    return np.load(path)


@pytest.fixture(scope="module")
def default_run(tmp_path_factory):
    return _run(tmp_path_factory.mktemp("knobs"), "default")


def test_default_dispatch_matches_the_oracle(default_run):
    from oracle import pyoracle as orc
    from tests._knob_job import jobs
    U, s, tape = jobs()
    assert np.array_equal(default_run["plain"], np.array(orc.gccb_simulate(U, s, tape)))
    assert np.array_equal(default_run["lossy"], np.array(orc.gccb_uniform_losses_simulate(U, s, 0.8, tape)))
    assert np.all(default_run["philox"].sum(axis=1) == 14)
    lost = 14 - default_run["philox_lossy"].sum(axis=1)
    assert lost.min() >= 0 and 3.0 < lost.mean() < 8.5       # Binomial(14, 0.4) losses: mean 5.6


@pytest.mark.parametrize("env", [
    {"BP_K3_WARP_MAX_K": 0},                          # 128-thread blocks everywhere
    {"BP_K3_WARP_MAX_K": 8},
    {"BP_K3_TPG": 2},                                 # up to hundreds of chunk blocks per sample
    {"BP_K3_TPG": 3, "BP_K3_WARP_MAX_K": 0, "BP_K3_CAP": 4},
    {"BP_K3_TREE_MAX_C": 8},                          # two lanes per term stream from k = 9
    {"BP_K3_TREE_MAX_C": 6, "BP_K3_WARP_MAX_K": 0},   # two lanes from k = 7, four from k = 13
    {"BP_K3_TREE_MAX_C": 6, "BP_K3_TPG": 5},
    {"BP_K3_TREE_MAX_C": 7, "BP_K3_TPG": 1},          # every lane group gets a single short period: Guan step at every boundary
], ids=lambda e: ",".join(f"{k[6:]}={v}" for k, v in e.items()))
def test_every_block_shape_gives_the_same_samples(tmp_path, default_run, env):
    got = _run(tmp_path, "variant", **env)
    for key in ("plain", "lossy", "philox", "philox_lossy"):
        assert np.array_equal(got[key], default_run[key]), key
    assert np.abs(got["pmf"] - default_run["pmf"]).max() <= 1e-13
    scale = np.abs(default_run["minors"]).max()
    assert np.abs(got["minors"] - default_run["minors"]).max() <= 1e-12 * scale
