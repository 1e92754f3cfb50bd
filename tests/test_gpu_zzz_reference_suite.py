"""The reference's OWN acceptance suite against the CUDA kernels (GPU box).

`scripts/install_reference.sh` puts the unmodified reference package and its tests/ directory under baseline/_ref
(git-ignored, shipped to the GPU box with the working tree).  oracle/reference_suite_plugin.py registers every module
that theboss_b200 rebuilds under its ``theboss.*`` name; with BOSSPERM_SUITE_HANDLE=cuda the drop-in classes keep their real
handle, so the 76 tests of the reference (calculators vs the O(n!) classic one, minors vs singles, the known-answer
distribution, TVD acceptance of every GCC / lossy / BOBS strategy through the factories) exercise kernels K1 .. K4.
The CPU twin of this test (tests/test_reference_suite_cpu.py) runs the same suite over the oracle stand-in.
"""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(REPO, "baseline", "_ref")


def test_reference_test_suite_passes_against_the_cuda_kernels(tmp_path):
    assert os.path.isdir(os.path.join(REF, "tests")) and os.path.isdir(os.path.join(REF, "theboss")), (
        "baseline/_ref is missing: run scripts/install_reference.sh in the build container (it ships with the working tree)")
    env = dict(os.environ, PYTHONPATH=REPO, THEBOSS_REFERENCE=REF, BOSSPERM_SUITE_HANDLE="cuda")
    run = subprocess.run([sys.executable, "-m", "pytest", "-p", "oracle.reference_suite_plugin", os.path.join(REF, "tests"),
                          "-q", "-p", "no:cacheprovider"], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=1500)
    tail = run.stdout[-4000:] + run.stderr[-1000:]
    assert run.returncode == 0, tail
    summary = re.search(r"(\d+) passed(?:, (\d+) skipped)?", run.stdout)
    assert summary and int(summary.group(1)) >= 76, tail        # the count the suite reaches on the reference itself
    aliased = re.search(r"theboss -> theboss_b200 for (\d+) modules", run.stdout)
    assert aliased and int(aliased.group(1)) >= 25, tail
    launched = re.search(r"CUDA handle launched (\d+) kernels", run.stdout)
    assert launched and int(launched.group(1)) > 1000, tail      # the kernels were the ones exercised
    print(run.stdout[-1500:])
