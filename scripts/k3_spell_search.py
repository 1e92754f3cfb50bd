#!/usr/bin/env python
"""Offline search over equivalent spellings of K3's tree arithmetic (compile-time knob K3_SPELL of minors_kernel.cu): compiles ONE
instantiation per spelling and ranks them by the register-read model of scripts/sass_rf.py.  No GPU needed.
usage: k3_spell_search.py LPG C THREADS [first last]"""
import os, re, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "..", "theboss_b200", "csrc")
lpg, c, t = sys.argv[1:4]
first, last = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (0, 1024)
os.makedirs("/tmp/k3s", exist_ok=True)

def one(spell):
    out = f"/tmp/k3s/k3_{lpg}_{c}_{t}_{spell}.o"
    r = subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-ccbin", "/usr/bin/g++",
                        "-Xcompiler", "-fPIC,-O2", f"-DK3_DEV_LPG={lpg}", f"-DK3_DEV_C={c}", f"-DK3_DEV_THREADS={t}", f"-DK3_SPELL={spell}", "-Xptxas", "-v",
                        "-c", "minors_kernel.cu", "-o", out], cwd=SRC, capture_output=True, text=True)
    regs = spill = -1
    m = re.search(r"k3_minors_kernel.*?\n.*?\n.*?Used (\d+) registers", r.stderr, re.S)
    blocks = r.stderr.split("Compiling entry function")
    for b in blocks:
        if "k3_minors_kernel" in b:
            mm = re.search(r"Used (\d+) registers", b); regs = int(mm.group(1)) if mm else -1
            ms = re.search(r"(\d+) bytes spill stores", b); spill = int(ms.group(1)) if ms else -1
    rf = subprocess.run([sys.executable, os.path.join(HERE, "sass_rf.py"), out, f"k3_minors_kernelILi{lpg}ELi{c}ELi{t}E"], capture_output=True, text=True).stdout
    m = re.search(r"FP64 (\d+).*?register reads (\d+), max\(pipe, reads\) per instruction: (\d+)", rf, re.S)
    os.remove(out)
    if not m: return spell, None
    return spell, (int(m.group(2)), int(m.group(3)), int(m.group(1)), regs, spill)

with ThreadPoolExecutor(8) as ex:
    res = list(ex.map(one, range(first, last)))
res = [r for r in res if r[1]]
res.sort(key=lambda r: (r[1][0], r[1][1]))
for spell, (reads, both, fp, regs, spill) in res:
    print(f"spell {spell:4d} (0x{spell:03x}): reads {reads}  max-model {both}  FP64 {fp}  regs {regs}  spill {spill}")
