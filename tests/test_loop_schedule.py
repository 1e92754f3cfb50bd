"""Build-time guard of the K3 term loop's instruction schedule (no GPU needed).

The loop is bound by register-file reads (one 64-bit register operand per cycle per SMSP, so a DFMA with three
fresh sources costs 3 cycles), and how many DFMAs ptxas pairs through the operand reuse cache changes with edits that do not touch
the loop at all: between two commits of round 2 the k = 24 step lost 3.6 % to such a reshuffle.  scripts/sass_rf.py evaluates the
register-read cost of the loop on the SASS of the built object; this test keeps the variants that carry the n = 24 run at the level
the committed profiles were measured at."""
import os
import re
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(REPO, "theboss_b200", "csrc", "build", "minors_kernel.o")

# (lanes per term stream, columns per lane): modelled cycles per term per warp must not exceed the bound (measured builds: 378 / 348 / 311)
BOUNDS = {(2, 12): 382, (2, 11): 352, (2, 10): 315}


@pytest.mark.skipif(not os.path.exists(OBJ), reason="library objects not built (python -c 'import __graft_entry__ as g; g.build()')")
@pytest.mark.parametrize("lpg,c", sorted(BOUNDS))
def test_term_loop_register_read_cost(lpg, c):
    out = subprocess.run([sys.executable, os.path.join(REPO, "scripts", "sass_rf.py"), OBJ, f"k3_minors_kernelILi{lpg}ELi{c}ELi128E"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    m = re.search(r"per instruction: (\d+) cycles", out.stdout)
    fp64 = re.search(r"FP64 (\d+) ", out.stdout)
    assert m and fp64, out.stdout
    assert int(fp64.group(1)) == 14 * c - 8, out.stdout        # 90 DFMA + 45 DMUL + 25 DADD at C = 12: the loop itself has not changed
    assert int(m.group(1)) <= BOUNDS[(lpg, c)], (
        f"ptxas scheduled the term loop of k3_minors_kernel<{lpg}, {c}, 128> at {m.group(1)} register-read cycles per term "
        f"(bound {BOUNDS[(lpg, c)]}): try the other spellings of k3_cmul / cmul_acc and re-measure with scripts/k3_dev.sh")
