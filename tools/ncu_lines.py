#!/usr/bin/env python
"""Attribute ncu per-SASS metrics to source lines: tools/ncu_lines.py <report.ncu-rep> [top]
(uses nvdisasm -g line info of the in-tree .so; prints the busiest source lines by stall samples).
ISS_KERNEL_FILE (default iss_kernels) and ISS_KERNEL_SYM (default the 256-thread generate kernel) choose another kernel."""
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("ISS_SO", os.path.join(ROOT, "insilicoseq_b200/csrc/libiss_b200.so"))], cwd=tmp,
               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
KFILE = os.environ.get("ISS_KERNEL_FILE", "iss_kernels")
cub = [f for f in os.listdir(tmp) if f.startswith(KFILE + ".") and f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
SYM = os.environ.get("ISS_KERNEL_SYM", "_Z19iss_generate_kernelILi%sE" % os.environ.get("ISS_KERNEL_THREADS", "256"))
start = [i for i, l in enumerate(sass) if l.startswith(".text." + SYM)][0]
ins, cur, inl = [], None, None
for l in sass[start + 1:]:
    if (l.startswith(".text.") or l.startswith("\t.section")) and ins:
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        own = m.group(1).endswith(KFILE + ".cu")
        cur = int(m.group(2)) if own else "H:%s:%s" % (m.group(1).split("/")[-1], m.group(2))
        mm = re.search(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        inl = int(mm.group(2)) if mm and mm.group(1).endswith(KFILE + ".cu") else None
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), cur, inl, m.group(2)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split("\n")))
hdr = rows[1]
idx = {h: k for k, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)][: len(ins)]
base = int(data[0][0], 16)


def f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


agg = {}
byline_inl = {}
for (off, line, inl_, text), r in zip(ins, data):
    assert int(r[0], 16) - base == off
    key = line
    a = agg.setdefault(key, [0.0] * 6)
    for j, name in enumerate(["Instructions Executed", "# Samples", "stall_long_sb", "stall_wait", "stall_barrier", "stall_short_sb"]):
        a[j] += f(r[idx[name]])
    if isinstance(line, str) and inl_:
        b = byline_inl.setdefault(inl_, [0.0, 0.0])
        b[0] += f(r[idx["Instructions Executed"]])
        b[1] += f(r[idx["# Samples"]])
ti = sum(a[0] for a in agg.values())
ts = sum(a[1] for a in agg.values())
src = open(os.path.join(ROOT, "insilicoseq_b200/csrc/%s.cu" % KFILE)).read().split("\n")
print("warp-instructions %.4g  samples %d" % (ti, ts))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    text = src[key - 1].strip()[:100] if isinstance(key, int) else str(key)
    print("%-28s inst %5.1f%% samp %5.1f%% long_sb %6.0f wait %6.0f barrier %6.0f short_sb %5.0f | %s"
          % (key, 100 * a[0] / ti, 100 * a[1] / ts, a[2], a[3], a[4], a[5], text))
print("-- header-inlined code by call site (line: inst%, samp%)")
for key, b in sorted(byline_inl.items(), key=lambda kv: -kv[1][0])[:12]:
    print("   %5d inst %5.1f%% samp %5.1f%% | %s" % (key, 100 * b[0] / ti, 100 * b[1] / ts, src[key - 1].strip()[:100]))
