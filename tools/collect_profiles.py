#!/usr/bin/env python
"""Copy the evidence of one GPU session (gpurun_out/<tag>/, scratch) into profiles/ (tracked) under a round prefix, and
refresh profiles/latest.json -- the ncu DRAM traffic of the dominant kernel together with the hash of the kernel sources
it was captured for (bench.py reports `roofline.traffic` only while that hash matches the built sources).

    python tools/collect_profiles.py gpurun_out/r2c r2c
"""
import csv
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, tag = sys.argv[1], sys.argv[2]
dst = os.path.join(ROOT, "profiles")
names = {"bench_n1.json": "bench_n1.json", "launches.csv": "launches.csv", "ncu_details.txt": "ncu_details.txt", "ncu_raw.csv": "ncu_raw.csv",
         "ncu_by_source_line.txt": "ncu_by_source_line.txt", "sanitizer_memcheck.txt": "sanitizer_memcheck.txt",
         "d2h_probe_n1.json": "d2h_probe_n1.json", "summary.txt": "session_summary.txt"}
for a, b in names.items():
    p = os.path.join(src, a)
    if os.path.exists(p) and os.path.getsize(p):
        shutil.copyfile(p, os.path.join(dst, "%s_%s" % (tag, b)))


def kernel_source_hash():
    h = hashlib.sha256()
    for f in ("iss_kernels.cu", "iss_device.cuh"):
        h.update(open(os.path.join(ROOT, "insilicoseq_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


raw = os.path.join(src, "ncu_raw.csv")
if os.path.exists(raw):
    rows = [r for r in csv.reader(open(raw)) if r]
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}

    def num(name):
        v, u = float(vals[col[name]]), units[col[name]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

    info = {"round": tag, "kernel": vals[col["Kernel Name"]] if "Kernel Name" in col else "iss_generate_kernel",
            "kernel_source_sha16": kernel_source_hash(), "config": 1,
            "dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
            "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
            "duration_ms_under_ncu": float(vals[col["gpu__time_duration.sum"]]) * {"ms": 1, "us": 1e-3, "ns": 1e-6, "msecond": 1, "usecond": 1e-3, "nsecond": 1e-6}.get(units[col["gpu__time_duration.sum"]].lower(), 1),
            "warp_instructions": float(vals[col["smsp__inst_executed.sum"]]),
            "source": "ncu --set full --clock-control none, one 2.5 M-pair launch of bench.py --device-only (%s_ncu_raw.csv)" % tag}
    json.dump(info, open(os.path.join(dst, "latest.json"), "w"), indent=1)
    print(json.dumps(info, indent=1))
