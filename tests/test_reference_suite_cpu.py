"""The reference's OWN test suite, run against the drop-in package (build container only).

oracle/reference_suite_plugin.py registers every module that theboss_b200 rebuilds under its ``theboss.*`` name, so the
unmodified test modules of /root/reference/tests import the drop-in calculators, strategies, factories, distribution
calculators and utilities; modules outside the permanent hot path stay the reference's.  The container has no GPU: the
arithmetic under the C-ABI boundary comes from the CPU oracle (oracle/handle_standin.py), i.e. this run pins the API
surface, argument handling, RNG usage and host logic of the drop-in to what the reference's tests expect -- the kernels
themselves are pinned by the `-m gpu` tests.  Skipped where the reference checkout is absent (the GPU box).
"""
import os
import re
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("THEBOSS_REFERENCE", "/root/reference")


def test_reference_test_suite_passes_against_the_drop_in(tmp_path):
    if not os.path.isdir(os.path.join(REF, "tests")):
        pytest.skip("reference checkout not available on this machine")
    env = dict(os.environ, PYTHONPATH=REPO, THEBOSS_REFERENCE=REF)
    run = subprocess.run([sys.executable, "-m", "pytest", "-p", "oracle.reference_suite_plugin", os.path.join(REF, "tests"),
                          "-q", "-p", "no:cacheprovider"], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=900)
    tail = run.stdout[-3000:]
    assert run.returncode == 0, tail
    summary = re.search(r"(\d+) passed(?:, (\d+) skipped)?", run.stdout)
    assert summary and int(summary.group(1)) >= 76, tail        # the count the suite reaches on the reference itself
    assert "failed" not in run.stdout.splitlines()[-1], tail
    aliased = re.search(r"theboss -> theboss_b200 for (\d+) modules", run.stdout)
    assert aliased and int(aliased.group(1)) >= 25, tail
    calls = re.search(r"oracle-backed handle served (\d+) calls", run.stdout)
    assert calls and int(calls.group(1)) > 1000, tail              # the drop-in classes were the ones exercised
