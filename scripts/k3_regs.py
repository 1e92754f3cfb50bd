#!/usr/bin/env python
"""Registers / spills of every k3_minors_kernel instantiation from the ptxas log (no GPU needed).
usage: k3_regs.py [ptxas-log]"""
import re, subprocess, sys
log = open(sys.argv[1] if len(sys.argv) > 1 else "theboss_b200/csrc/build/minors_kernel.ptxas.log").read()
rows = []
for e in re.split(r"ptxas info\s+: Compiling entry function '", log)[1:]:
    name = e.split("'")[0]
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    m = re.search(r"k3_minors_kernel<(\d+), (\d+), (\d+)>", dem)
    if not m:
        continue
    regs = int(re.search(r"Used (\d+) registers", e).group(1))
    sp = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", e)
    rows.append((int(m.group(1)), int(m.group(3)), int(m.group(2)), regs, int(sp.group(1)), int(sp.group(2))))
for r in sorted(rows):
    print("LPG%d T%3d C%2d regs %3d spill st %4d ld %4d" % r)
