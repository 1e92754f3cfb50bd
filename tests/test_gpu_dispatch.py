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
    {"BP_K3_WARP2_MIN_K": 99},                        # one lane per term stream in every one-warp block
    {"BP_K3_WARP2_MIN_K": 11},
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


@pytest.mark.parametrize("env", [
    {"BP_K3_TREE_MAX_C": 6, "BP_K3_WARP_MAX_K": 0},   # two lanes from k = 7, FOUR lanes (k3_minors_kernel<4, C>) from k = 13
    {"BP_K3_TREE_MAX_C": 8, "BP_K3_WARP_MAX_K": 6},   # two lanes from k = 9
], ids=lambda e: ",".join(f"{k[6:]}={v}" for k, v in e.items()))
def test_lane_split_variants_reproduce_the_seeded_reference_runs_through_the_strategies(env):
    """The seeded runs of the UNMODIFIED reference (GCC-B, its uniform-loss variant and the lossy-network wrapper at n = 12 .. 16,
    tests/golden/gccb_seeded_samples.npz) replayed through the drop-in strategy classes in a process whose dispatch knobs send the
    steps k >= 13 to the four-lane kernels -- the layouts BASELINE config 5(ii) reaches at k >= 31, which no reference run can:
    every sample must still equal the reference's bit for bit (lossy_networks_generalized_cliffords_simulation_strategy.py:53-83,
    bs_cc_ryser_submatrices_permanent_calculator.py:80-119)."""
    code = ("import os, sys; sys.path.insert(0, %r); from tests.test_host_logic import _check_seeded_gccb_fixture; "
            "_check_seeded_gccb_fixture(os.path.join(%r, 'tests', 'golden')); print('seeded fixtures reproduced')" % (REPO, REPO))
    e = {k: v for k, v in os.environ.items() if not k.startswith("BP_K3_")}
    e.update({k: str(v) for k, v in env.items()})
    run = subprocess.run([sys.executable, "-c", code], env=e, cwd=REPO, capture_output=True, text=True, timeout=900)
    assert run.returncode == 0 and "seeded fixtures reproduced" in run.stdout, run.stdout[-2000:] + run.stderr[-2000:]
