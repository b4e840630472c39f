#!/usr/bin/env python
"""Per-kernel SASS evidence of the shipped library: which kernels use TMA bulk copies (UBLKCP), mbarriers (SYNCS),
read-only / cache-global loads, peer-capable stores, and that no tensor-core instruction is present (north_star: the
path is bandwidth-bound, no tensor cores).  Usage: python tools/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "aocl-sparse_b200", "libaoclsparse_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pats = collections.OrderedDict([
    ("UBLKCP", r"\bUBLKCP"), ("UTMALDG", r"\bUTMALDG"), ("SYNCS(mbarrier)", r"\bSYNCS"), ("LDG.CONSTANT(nc)", r"\bLDG\.E\S*CONSTANT"),
    ("LDG.STRONG.GPU(cg)", r"\bLDG\.E\S*STRONG"), ("LDG", r"\bLDG"), ("STG", r"\bSTG"), ("LDS", r"\bLDS"), ("STS", r"\bSTS"),
    ("ATOMG/RED", r"\b(ATOMG|RED)\b"), ("ATOMS", r"\bATOMS"), ("BAR", r"\bBAR\."), ("ACQBULK/PDL", r"\bACQBULK|\bPREEXIT"),
    ("DFMA", r"\bDFMA"), ("FFMA", r"\bFFMA"), ("UTC*MMA", r"\bUTC\w*MMA"), ("HMMA/IMMA/DMMA", r"\b[HID]MMA"), ("LDTM/STTM", r"\b(LDTM|STTM)"),
])
kern = None
counts = collections.OrderedDict()
arch = None
for ln in out.splitlines():
    m = re.match(r"\s*arch = (\S+)", ln)
    if m:
        arch = m.group(1)
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = m.group(1) + " [" + str(arch) + "]"
        counts[kern] = collections.Counter()
        continue
    if kern and "/*" in ln:
        for name, pat in pats.items():
            if re.search(pat, ln):
                counts[kern][name] += 1
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)} -- instruction counts per kernel ({len(counts)} kernels)")
print("# columns: " + " | ".join(pats))
try:
    dem = subprocess.run(["c++filt"], input="\n".join(k.split(" [")[0] for k in counts), capture_output=True, text=True).stdout.splitlines()
except Exception:
    dem = [k for k in counts]
tot = collections.Counter()
for (k, c), d in zip(counts.items(), dem):
    tot.update(c)
    short = re.sub(r"\((?!anonymous namespace\)).*", "", d).replace("(anonymous namespace)::", "")
    print(f"{short[:110]:110s} " + " ".join(f"{c.get(n, 0):5d}" for n in pats) + "  " + k.split(" [")[1].rstrip("]"))
print("TOTAL".ljust(110) + " " + " ".join(f"{tot.get(n, 0):5d}" for n in pats))
print("# tensor-core instructions (UTC*MMA / HMMA / LDTM): %d -- none expected: SpMV / SpMM here is HBM- and L2-bound"
      % (tot.get("UTC*MMA", 0) + tot.get("HMMA/IMMA/DMMA", 0) + tot.get("LDTM/STTM", 0)))
