#!/usr/bin/env python
"""profiles/<tag>_sass_tma_excerpt.txt: the TMA / mbarrier instructions of the built kernels (cuobjdump -sass of the in-tree
libiss_b200.so), with two lines of context each -- the evidence that the bulk-copy engine is what moves the quality rows
in (UBLKCP.S.G + SYNCS.ARRIVE.TRANS64 / SYNCS.PHASECHK.TRANS64.TRYWAIT) and the FASTQ images out (UBLKCP.G.S)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
so = os.path.join(ROOT, "insilicoseq_b200", "csrc", "libiss_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.splitlines()
out = ["cuobjdump -sass insilicoseq_b200/csrc/libiss_b200.so (sm_100a), instructions matching UBLKCP|SYNCS|UTMA, +-2 lines", ""]
fn = None
count = {}
keep = set()
for i, ln in enumerate(sass):
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        fn = m.group(1)
    mo = re.search(r"(UBLKCP[.\w]*|SYNCS[.\w]*|UTMA[.\w]*)", ln)
    if mo and "/*" in ln and not ln.strip().startswith("/* 0x"):
        count.setdefault(fn, {})
        op = mo.group(1)
        count[fn][op] = count[fn].get(op, 0) + 1
        for j in range(i - 4, i + 5):
            keep.add((fn, j))
for fn, ops in count.items():
    out.append("%s" % fn)
    for op, n in sorted(ops.items()):
        out.append("    %-44s x %d" % (op, n))
out.append("")
last_fn, last_j = None, -10
shown = {}
for fn, j in sorted(keep, key=lambda t: (t[0] or "", t[1])):
    ln = sass[j]
    if ln.strip().startswith("/* 0x") or "/*" not in ln:
        continue
    if shown.get(fn, 0) > 60:  # a page per kernel is enough
        continue
    if fn != last_fn:
        out.append("---- %s" % fn)
    elif j > last_j + 2:
        out.append("        ...")
    out.append(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ln.rstrip()))
    shown[fn] = shown.get(fn, 0) + 1
    last_fn, last_j = fn, j
open(os.path.join(ROOT, "profiles", "%s_sass_tma_excerpt.txt" % tag), "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))
