#!/usr/bin/env python
"""bench.py -- read-pair generation throughput of the B200 engine on BASELINE.json configs[1]:
5 Mbp synthetic genome, MiSeq KDE model truncated to 150 positions, 10 M 150 bp pairs per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over the whole workload (10 M pairs per GPU, weak scaling:
rank r generates chunk/cpu_number r with Philox key seed + r, exactly like `--cpus N` would).
  value  : pairs/s with everything resident in HBM and FASTQ left in HBM (device-only), CUDA-event timed
  e2e    : pairs/s through the public API (Engine.generate), FASTQ copied into the pinned host ring,
           host<->device copies inside the timed region
  roofline: algorithmic bytes (2-bit genome windows read + FASTQ bytes written) / kernel time vs measured HBM peak
  cpu_baseline: the C oracle port of the reference path on the host cores (bounded sample), N=1 only
--impl reference times that same CPU port as the reference arm (the reference itself is pure Python at
~1 k pairs/s/core and is not present on the GPU box).
Multi-GPU: launched by torchrun, one rank per GPU; torch.distributed (gloo) carries only the control
plane (NCCL unique id, barriers, max-over-ranks); the genome + model travel by one ncclBroadcast inside the library.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GENOME_LEN = 5_000_000
GENOME_NAME = "syn5M"
GENOME_SEED = 20260924
READ_LEN = 150
PAIRS_PER_GPU = 10_000_000
SEED = 42
MODEL_FILE = "miSeq_0.npz"  # `--model miseq` (iss/generator.py:380), truncated to 150 positions


def make_genome():
    rng = np.random.default_rng(GENOME_SEED)
    return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=GENOME_LEN, dtype=np.uint8)].tobytes()


def measured_peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def usable_cores():
    """CPUs this process may actually use: the affinity mask capped by the cgroup CPU quota (a GPU box can show 128
    logical CPUs while the container is limited to 16: oversubscribing it makes the CPU arm slower, not faster)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]))))
            else:
                quota = int(txt[0])
                period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if quota > 0:
                    n = min(n, max(1, quota // period))
            break
        except (OSError, ValueError, IndexError):
            continue
    return n


# ---- CPU arm: the C oracle port of the reference path (test infrastructure, used here only as the baseline) ------
def _oracle_chunk(args):
    npz, i0, i1 = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle

    g = _oracle_chunk.genome
    om = oracle.OracleModel(npz)
    t = time.perf_counter()
    r1, r2 = oracle.generate(om, g, GENOME_NAME, oracle.make_key(SEED, 0), 0, 0, i0, i1)
    return time.perf_counter() - t, len(r1)


def cpu_port_throughput(pairs_per_core, cores=None, repeats=1):
    """pairs/s of the oracle port using a process pool (the reference parallelises the same way: mp.Pool)."""
    import multiprocessing as mp

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle

    oracle.build()
    cores = cores or usable_cores()
    tmp = tempfile.mkdtemp(prefix="iss_bench_")
    npz = os.path.join(tmp, "miseq_150.npz")
    oracle.truncate_profile(os.path.join(ROOT, "insilicoseq_b200", "profiles", MODEL_FILE), npz, READ_LEN)
    _oracle_chunk.genome = make_genome()
    ctx = mp.get_context("fork")
    out = []
    with ctx.Pool(cores) as pool:
        pool.map(_oracle_chunk, [(npz, 0, 64)] * cores)  # warm the workers
        for _ in range(repeats):
            t = time.perf_counter()
            pool.map(_oracle_chunk, [(npz, k * pairs_per_core, (k + 1) * pairs_per_core) for k in range(cores)])
            out.append(cores * pairs_per_core / (time.perf_counter() - t))
    return out, cores


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    per_core = 100_000
    cores = usable_cores()
    vals, cores = cpu_port_throughput(per_core, cores, repeats=max(args.warmup, 0) + args.steps)
    vals = vals[-args.steps:]
    v = statistics.mean(vals)
    sample = "%d pairs per step (%d per core x %d processes) of the same workload" % (per_core * cores, per_core, cores)
    line = {
        "impl": "reference", "metric": "150bp_read_pairs_per_sec", "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * per_core * cores / v,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(n):
    return {"workload": "configs[1]: 5 Mbp synthetic genome (iid ACGT, 1 record), MiSeq KDE model (miSeq_0.npz) truncated to 150 "
                        "positions, 10M 150 bp pairs per GPU, seed 42",
            "pairs_per_gpu": PAIRS_PER_GPU, "read_length": READ_LEN, "genome_bp": GENOME_LEN,
            "parallelism": "reads sharded by count: chunk r -> GPU r (cpu_number r, key seed+r); 1 ncclBroadcast of genome+model at init",
            "l2": "each step streams ~6.5 GB of FASTQ per GPU through L2 (126 MB); the 1.25 MB packed genome is L2-resident by design"}


def philox_blocks_per_pair(model, L):
    """Philox4x32-10 calls per pair under RNG contract v2, substitution-choice blocks excluded:
    3 PAIR blocks + ceil(L/2) QUAL blocks per read + one block per group of 4 indel sites that can fire."""
    arrays = model.device_tables().arrays
    ins = np.asarray(arrays["ins_nz"]).reshape(2, L)
    dele = np.asarray(arrays["del_nz"]).reshape(2, L)
    groups = 0
    for o in range(2):
        n_ins = sum(bin(int(x) & 15).count("1") for x in ins[o, :L - 1])
        n_del = int(np.count_nonzero(dele[o, :L - 1] & 15))
        groups += (n_ins + 3) // 4 + (n_del + 3) // 4
    return 3 + 2 * ((L + 1) // 2) + groups


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_GPU, help="pairs per GPU per step (default: the BASELINE workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--device-only", action="store_true", help="kernel A/B runs: only the device-resident leg")
    args = ap.parse_args()
    from insilicoseq_b200 import parallel

    rank, world, local = parallel.env_rank()
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"  # stdout carries exactly one JSON line (NCCL's version banner goes there)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    dist = parallel.init_control_plane(rank, world)  # gloo, control plane only: ids, barriers, max over ranks

    from insilicoseq_b200 import engine as eng_mod
    from insilicoseq_b200 import error_models

    numa_node = parallel.bind_to_device_numa(local)  # before the pinned ring is allocated
    eng = eng_mod.Engine(local)
    if rank == 0:
        model = error_models.KDErrorModel(os.path.join(error_models.PROFILE_DIR, MODEL_FILE)).truncate(READ_LEN)
        eng.set_genome([(GENOME_NAME, make_genome())])
        eng.set_model(model)
        philox_blocks = philox_blocks_per_pair(model, READ_LEN)
    if world > 1:
        uid = parallel.share_from_root(dist, eng_mod.Engine.unique_id() if rank == 0 else None, rank)
        eng.comm_init_rank(world, rank, uid)
        eng.bcast(0, read_length=READ_LEN, record_names=[GENOME_NAME], record_lengths=[GENOME_LEN])
    L = READ_LEN
    items = eng_mod.worker_items([(0, args.pairs)], rank, SEED, [GENOME_LEN], L)

    def barrier():
        if dist is not None:
            dist.barrier()
        eng.sync()
        try:
            import torch

            if torch.cuda.is_available():
                torch.cuda.synchronize()
        except Exception:  # noqa: BLE001
            pass

    def reduce_max(x):
        return parallel.max_over_ranks(dist, x)

    def step_device():
        for _ in eng.generate(items, device_only=True, batch_pairs=2_500_000, n_slots=1):
            pass

    def step_e2e():
        tail = 0
        for b in eng.generate(items, batch_pairs=1 << 18, n_slots=3):
            tail ^= int(b.r1[-1]) ^ int(b.r2[0])  # touch the host copy: proves the bytes arrived
        return tail

    # ---- device-resident throughput ("value") --------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    clocks = ClockSampler(local)
    barrier()
    s0 = eng.stats()
    if rank == 0:
        clocks.start()
    eng.mark(0)
    for _ in range(args.steps):
        step_device()
    eng.mark(1)
    ms = eng.elapsed_ms(0, 1)
    barrier()
    s1 = eng.stats()
    clk = clocks.stop() if rank == 0 else None
    ms = reduce_max(ms)
    pairs_step = args.pairs * world
    value = pairs_step * args.steps / (ms * 1e-3)
    launches = s1["kernel_launches"] - s0["kernel_launches"]
    k_ms = s1["kernel_ms"] - s0["kernel_ms"]
    fq = s1["fastq_bytes"] - s0["fastq_bytes"]
    gen_pairs = s1["pairs"] - s0["pairs"]
    algo_bytes = fq + 2 * ((L + 3) // 4) * gen_pairs
    peak, peak_src = measured_peaks()
    achieved = algo_bytes / (k_ms * 1e-3) / 1e9

    if args.device_only:
        if rank == 0:
            print(json.dumps({"metric": "150bp_read_pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world,
                              "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                              "roofline": {"avg_launch_ms": k_ms / max(launches, 1), "achieved": achieved, "frac": achieved / peak},
                              "partial": "device-only A/B run"}))
        eng.close()
        return
    # ---- end to end through the public API, host ring included ("e2e") -------------------------------------
    step_e2e()
    barrier()
    e0 = eng.stats()
    eng.mark(2)
    for _ in range(args.steps):
        step_e2e()
    eng.mark(3)
    ms_e = reduce_max(eng.elapsed_ms(2, 3))
    barrier()
    e1 = eng.stats()
    e2e_value = pairs_step * args.steps / (ms_e * 1e-3)

    # ---- the same with --compress: gzip members produced on the GPU, only compressed bytes cross PCIe ------------
    def step_gz():
        n = 0
        for b in eng.generate(items, batch_pairs=1 << 18, compress=True):
            n += b.nbytes + b.nbytes2
        return n

    step_gz()
    barrier()
    eng.mark(0)
    gz_bytes = 0
    for _ in range(args.steps):
        gz_bytes += step_gz()
    eng.mark(1)
    ms_z = reduce_max(eng.elapsed_ms(0, 1))
    barrier()

    if rank == 0:
        traffic = None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "latest.json")))
            traffic = prof.get("dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            pass
        line = {
            "metric": "150bp_read_pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 (integer thresholds on Philox4x32 words; bytes out)", "data": "synthetic",
            "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": "pairs/s", "ms_per_step": ms_e / args.steps,
                    "h2d_bytes_per_step": world * (e1["h2d_bytes"] - e0["h2d_bytes"]) // args.steps,
                    "d2h_bytes_per_step": world * (e1["d2h_bytes"] - e0["d2h_bytes"]) // args.steps,
                    "d2h_GBps_per_gpu": (e1["d2h_bytes"] - e0["d2h_bytes"]) / (ms_e * 1e-3) / 1e9},
            "e2e_gzip": {"value": pairs_step * args.steps / (ms_z * 1e-3), "unit": "pairs/s", "ms_per_step": ms_z / args.steps,
                         "d2h_bytes_per_step": world * gz_bytes // args.steps,
                         "compression_ratio": (e1["d2h_bytes"] - e0["d2h_bytes"]) / max(gz_bytes, 1),
                         "note": "iss generate --compress: gzip (RFC 1952) members built on the GPU, decompresses to the same FASTQ"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "iss_generate_kernel",
                         "algorithmic_bytes_per_launch": algo_bytes / max(launches, 1),
                         "avg_launch_ms": k_ms / max(launches, 1),
                         "note": "the kernel is integer-issue bound, not HBM bound (SURVEY.md 8d); see DESIGN.md section 6",
                         "philox_blocks_per_pair": philox_blocks, "philox_blocks_per_s": philox_blocks * value},
            "clocks": clk,
            "host": {"numa_node_rank0": numa_node, "logical_cpus": os.cpu_count(), "usable_cpus": usable_cores()},
        }
        if world == 1 and not args.no_cpu_baseline:
            per_core = 200_000
            vals, cores = cpu_port_throughput(per_core, repeats=1)
            line["cpu_baseline"] = {"value": vals[0], "unit": "pairs/s", "cores": cores, "kind": "port",
                                    "sample": "%d pairs (%d per core) of the same workload, oracle/iss_oracle.c in %d processes"
                                              % (per_core * cores, per_core, cores)}
        print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
