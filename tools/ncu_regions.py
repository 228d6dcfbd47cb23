#!/usr/bin/env python
"""Instruction and stall-sample shares of an ncu report by code REGION: tools/ncu_regions.py <report.ncu-rep> <line,line,...>
The SASS of the kernel is walked in address order; an instruction belongs to the region of the last boundary line (a line of
the kernel's own .cu file from the list) whose code was seen before it, so code inlined from headers (Philox) is counted
where it is used.  Same environment variables as tools/ncu_lines.py."""
import csv, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("ISS_SO", os.path.join(ROOT, "insilicoseq_b200/csrc/libiss_b200.so"))], cwd=tmp,
               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
KFILE = os.environ.get("ISS_KERNEL_FILE", "iss_kernels")
cub = [f for f in os.listdir(tmp) if f.startswith(KFILE + ".") and f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
SYM = os.environ.get("ISS_KERNEL_SYM", "_Z19iss_generate_kernelILi%sE" % os.environ.get("ISS_KERNEL_THREADS", "256"))
start = [i for i, l in enumerate(sass) if l.startswith(".text." + SYM)][0]
ins, cur = [], None
for l in sass[start + 1:]:
    if (l.startswith(".text.") or l.startswith("\t.section")) and ins:
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        # outermost call site inside the kernel's own file, if the line info carries one
        chain = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        own = [int(n) for f_, n in chain if f_.endswith(KFILE + ".cu")]
        if m.group(1).endswith(KFILE + ".cu"):
            own.insert(0, int(m.group(2)))
        own = [x for x in own if x >= int(os.environ.get('ISS_KERNEL_FIRST_LINE', '0'))]  # helpers above the kernel: carried too
        cur = own[-1] if own else cur   # header code without a call chain: the region it sits in, by address
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), cur, m.group(2)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split("\n")))
hdr = rows[1]
idx = {h: k for k, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)][: len(ins)]
def f(x):
    try: return float(x)
    except ValueError: return 0.0
bounds = sorted(int(x) for x in sys.argv[2].split(",")) if len(sys.argv) > 2 else None
agg = {}
for (off, line, text), r in zip(ins, data):
    if line is None:
        key = "?"
    elif bounds:
        b = [x for x in bounds if x <= line]
        key = b[-1] if b else 0
    else:
        key = line
    a = agg.setdefault(key, [0.0, 0.0])
    a[0] += f(r[idx["Instructions Executed"]]); a[1] += f(r[idx["# Samples"]])
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
src = open(os.path.join(ROOT, "insilicoseq_b200/csrc/%s.cu" % KFILE)).read().split("\n")
print("warp-instructions %.4g  samples %d" % (ti, ts))
for key, a in sorted(agg.items(), key=lambda kv: (str(type(kv[0])), kv[0])):
    text = src[key - 1].strip()[:90] if isinstance(key, int) and key > 0 else ""
    if a[0] / ti > 0.002 or a[1] / ts > 0.002:
        print("%-6s inst %5.1f%% samp %5.1f%% | %s" % (key, 100 * a[0] / ti, 100 * a[1] / ts, text))
