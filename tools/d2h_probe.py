#!/usr/bin/env python
"""Host side of the multi-GPU end-to-end figure: sustained D2H rate per GPU when 1, 4 or 8 GPUs copy at once, for each way
of obtaining the pinned ring (iss_probe_d2h: cudaMallocHost / mmap+mbind+register / write-combined), ring size and
thread placement.  One thread per GPU (ctypes releases the GIL); every thread starts behind a barrier.

    python tools/d2h_probe.py [--seconds 1.5] [--out gpurun_out/d2h_probe.json]
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from insilicoseq_b200 import _lib, parallel  # noqa: E402

MODES = {"pinned": 0, "numa": 1, "wc": 2}


def run_config(lib, devices, mode, ring_mb, chunk_mb, seconds, bind):
    res = {}
    bar = threading.Barrier(len(devices))
    all_cpus = os.sched_getaffinity(0)

    def work(d):
        os.sched_setaffinity(0, all_cpus)
        if bind:
            parallel.bind_to_device_numa(d)
        g, node = ctypes.c_double(0), ctypes.c_int(-1)
        bar.wait()
        rc = lib.iss_probe_d2h(d, MODES[mode], ring_mb << 20, chunk_mb << 20, seconds, ctypes.byref(g), ctypes.byref(node))
        res[d] = (rc, g.value, node.value)

    ts = [threading.Thread(target=work, args=(d,)) for d in devices]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    rates = [res[d][1] for d in devices]
    return {"devices": devices, "alloc": mode, "ring_MB": ring_mb, "chunk_MB": chunk_mb, "bind_thread": bind,
            "GBps_per_gpu": [round(r, 2) for r in rates], "GBps_total": round(sum(rates), 1),
            "ring_node": [res[d][2] for d in devices], "rc": [res[d][0] for d in devices]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=1.5)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "d2h_probe.json"))
    args = ap.parse_args()
    lib = _lib.load()
    n = ctypes.c_int(0)
    _lib.check(lib.iss_device_count(ctypes.byref(n)))
    n = n.value
    info = {"n_gpus": n, "logical_cpus": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
    for cmd, key in ((["nvidia-smi", "topo", "-m"], "topo"), (["lscpu"], "lscpu"), (["numactl", "-H"], "numactl"),
                     (["cat", "/sys/fs/cgroup/cpu.max"], "cpu_max")):
        try:
            info[key] = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout
        except Exception as e:  # noqa: BLE001
            info[key] = "unavailable: %s" % e
    rows = []
    sets = [[0]]
    if n >= 4:
        sets.append(list(range(4)))
    if n >= 8:
        sets += [list(range(4, 8)), list(range(8)), [0, 1, 4, 5]]
    elif n > 1:
        sets.append(list(range(n)))
    for devs in sets:
        for mode, ring, chunk, bind in (("pinned", 512, 85, False), ("pinned", 512, 85, True), ("numa", 512, 85, False),
                                        ("wc", 512, 85, False), ("numa", 32, 4, False), ("numa", 128, 16, False),
                                        ("numa", 8, 1, False)):
            row = run_config(lib, devs, mode, ring, chunk, args.seconds, bind)
            rows.append(row)
            print(json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"host": info, "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
