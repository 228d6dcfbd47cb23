#!/usr/bin/env python
"""bench.py -- read-pair generation throughput of the B200 engine on the BASELINE.json workloads.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1|2|3|4] [--impl reference]

Default (what the driver runs): --config 1 = BASELINE.json configs[1]: 5 Mbp synthetic genome, MiSeq KDE model truncated
to 150 positions, 10 M 150 bp pairs per GPU.  The other configs are the remaining BASELINE.json entries:
  2  configs[2]: 100-record synthetic metagenome (~0.4 Gbp, lognormal abundance), NovaSeq (150 positions), 100 M pairs per GPU
  3  configs[3]: the 5 Mbp genome, HiSeq (native L = 126), 10^9 pairs in 8 chunks -> 1.25e8 pairs per GPU
  4  configs[4]: 1000-record 3 Gbp metagenome (packed genome 750 MB > L2: HBM-random gathers), NovaSeq, 1.25e8 pairs per GPU
One "step" = one pass of the hot path over the rank's share of the workload (weak scaling: rank r generates chunk /
cpu_number r with Philox key seed + r, exactly like `--cpus N` would).
  value     : pairs/s with everything resident in HBM and FASTQ left in HBM (device-only), CUDA-event timed
  e2e       : pairs/s through the public API (Engine.generate), FASTQ copied into the pinned host ring, host<->device
              copies inside the timed region
  e2e_gzip  : the same with `--compress` (gzip members built on the GPU; only compressed bytes cross PCIe)
  e2e_files : pairs/s of generator.worker_iterator writing real FASTQ files (native pwrite threads from the ring) on a
              bounded sample; host wall clock (file system work is host work)
  roofline  : algorithmic bytes (2-bit genome windows read + FASTQ bytes written) / kernel time vs measured HBM peak
  cpu_baseline: the reference's own CPU path on the host cores (bounded sample), N=1 only -- the UNMODIFIED reference
              (`python -m iss generate`, installed under baseline/_ref, Biopython stand-in from oracle/bio_standin) when it
              is present, with the C port of the same path (oracle/iss_oracle.c) reported beside it
--impl reference times the reference's CPU path the same way as the reference arm.
Multi-GPU: launched by torchrun, one rank per GPU; torch.distributed (gloo) carries only the control plane (NCCL unique
id, barriers, max-over-ranks); the genome + model travel by one ncclBroadcast inside the library.
"""
import argparse
import hashlib
import json
import os
import shutil
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

GENOME_SEED = 20260924
SEED = 42
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
BIO_STANDIN = os.path.join(ROOT, "oracle", "bio_standin")

CONFIGS = {
    1: dict(name="configs[1]", genome="syn5M", model="miSeq_0.npz", read_length=150, pairs_per_gpu=10_000_000,
            text="5 Mbp synthetic genome (iid ACGT, 1 record), MiSeq KDE model (miSeq_0.npz) truncated to 150 positions, "
                 "10M 150 bp pairs per GPU, seed 42"),
    2: dict(name="configs[2]", genome="meta100", model="NovaSeq", read_length=150, pairs_per_gpu=100_000_000,
            text="100-record synthetic metagenome (lengths log-uniform 1-10 Mbp, iid ACGT), lognormal abundance "
                 "(np.random.seed(42)), NovaSeq KDE model truncated to 150 positions, 100M pairs per GPU, seed 42"),
    3: dict(name="configs[3]", genome="syn5M", model="HiSeq", read_length=126, pairs_per_gpu=125_000_000,
            text="5 Mbp synthetic genome, HiSeq KDE model at its native read length 126 (it has no 150 bp form), 1e9 pairs "
                 "in 8 chunks -> 1.25e8 pairs per GPU (at N < 8 the first N chunks), seed 42"),
    4: dict(name="configs[4]", genome="meta1000", model="NovaSeq", read_length=150, pairs_per_gpu=125_000_000,
            text="1000-record 3 Gbp synthetic metagenome (packed genome 750 MB > L2), lognormal abundance, NovaSeq KDE model "
                 "truncated to 150 positions, 1e9 pairs in 8 chunks -> 1.25e8 pairs per GPU (at N < 8 the first N chunks), "
                 "FASTQ spilled to the pinned host ring, seed 42"),
}


# ---- synthetic genomes ----------------------------------------------------------------------------------------------
class Rec(object):
    """What generate_work_divider and Engine.register_records need of a record: .id, .seq (bytes-like or a length stub)."""

    __slots__ = ("id", "seq")

    def __init__(self, rid, seq):
        self.id = rid
        self.seq = seq


def record_lengths(genome):
    if genome == "syn5M":
        return ["syn5M"], [5_000_000]
    rng = np.random.default_rng(GENOME_SEED)
    if genome == "meta100":
        lens = np.exp(rng.uniform(np.log(1e6), np.log(1e7), size=100)).astype(np.int64)
    elif genome == "meta1000":
        lens = np.exp(rng.uniform(np.log(1e6), np.log(1e7), size=1000))
        lens = (lens * (3.0e9 / lens.sum())).astype(np.int64)
    else:
        raise ValueError(genome)
    return ["g%04d" % k for k in range(len(lens))], [int(x) for x in lens]


def make_records(genome):
    """Records with sequences (rank 0 only): iid uniform ACGT, one generator stream per workload."""
    names, lens = record_lengths(genome)
    rng = np.random.default_rng(GENOME_SEED if genome == "syn5M" else GENOME_SEED + 1)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    recs = []
    for name, n in zip(names, lens):
        recs.append(Rec(name, lut[rng.integers(0, 4, size=n, dtype=np.uint8)].tobytes()))
    return recs


def make_genome():
    """The configs[1] genome as bytes (also used by the CPU arms)."""
    rng = np.random.default_rng(GENOME_SEED)
    return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=5_000_000, dtype=np.uint8)].tobytes()


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


# ---- CPU arms (test infrastructure used here only as the measured baseline) -----------------------------------------
def _oracle_chunk(args):
    npz, i0, i1 = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle

    g = _oracle_chunk.genome
    om = oracle.OracleModel(npz)
    t = time.perf_counter()
    r1, r2 = oracle.generate(om, g, "syn5M", oracle.make_key(SEED, 0), 0, 0, i0, i1)
    return time.perf_counter() - t, len(r1)


def _bench_model_npz(tmp, cfg):
    """The benchmark's model as an .npz file the CPU arms can load (bundled profile truncated to the read length)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle

    src = os.path.join(ROOT, "insilicoseq_b200", "profiles", cfg["model"])
    if int(np.load(src, allow_pickle=True)["read_length"]) == cfg["read_length"]:
        return src
    npz = os.path.join(tmp, "bench_model_%d.npz" % cfg["read_length"])
    oracle.truncate_profile(src, npz, cfg["read_length"])
    return npz


def cpu_port_throughput(cfg, pairs_per_core, cores=None, repeats=1):
    """pairs/s of the C port of the reference path (oracle/iss_oracle.c) in a process pool (the reference parallelises
    the same way: mp.Pool), configs[1]-shaped work (one 5 Mbp record)."""
    import multiprocessing as mp

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle

    oracle.build()
    cores = cores or usable_cores()
    tmp = tempfile.mkdtemp(prefix="iss_bench_")
    npz = _bench_model_npz(tmp, cfg)
    _oracle_chunk.genome = make_genome()
    ctx = mp.get_context("fork")
    out = []
    with ctx.Pool(cores) as pool:
        pool.map(_oracle_chunk, [(npz, 0, 64)] * cores)  # warm the workers
        for _ in range(repeats):
            t = time.perf_counter()
            pool.map(_oracle_chunk, [(npz, k * pairs_per_core, (k + 1) * pairs_per_core) for k in range(cores)])
            out.append(cores * pairs_per_core / (time.perf_counter() - t))
    shutil.rmtree(tmp, ignore_errors=True)
    return out, cores


def reference_available():
    return os.path.isdir(os.path.join(REF_DIR, "iss")) and os.path.isdir(BIO_STANDIN)


def reference_cli_throughput(cfg, cores, pairs_per_core=(1500, 6000), repeats=1):
    """pairs/s of the UNMODIFIED reference: `python -m iss generate --cpus <cores>` from baseline/_ref (its own mp.Pool,
    iss/app.py:99-106), the Biopython stand-in on PYTHONPATH (biopython is not installed; it supplies containers and
    FASTQ text only -- SURVEY.md 8c).  The command is run at two sizes and the rate is the slope, which removes the
    start-up (imports, model load, FASTA parse, abundance file) from a sample this small; the whole-command rate of the
    larger run is returned as well."""
    tmp = tempfile.mkdtemp(prefix="iss_ref_")
    try:
        npz = _bench_model_npz(tmp, cfg)
        fasta = os.path.join(tmp, "genome.fasta")
        g = make_genome().decode()
        with open(fasta, "w") as fh:
            fh.write(">syn5M\n")
            fh.write("\n".join(g[i:i + 80] for i in range(0, len(g), 80)))
            fh.write("\n")
        env = dict(os.environ)
        env["PYTHONPATH"] = os.pathsep.join([REF_DIR, BIO_STANDIN] + ([env["PYTHONPATH"]] if env.get("PYTHONPATH") else []))
        slopes, whole = [], []
        for _ in range(repeats):
            times = []
            for ppc in pairs_per_core:
                n_pairs = ppc * cores
                out = os.path.join(tmp, "out%d" % ppc)
                cmd = [sys.executable, "-m", "iss", "generate", "--genomes", fasta, "--model", npz, "-n", str(2 * n_pairs),
                       "--cpus", str(cores), "--seed", str(SEED), "--quiet", "-o", out]
                t = time.perf_counter()
                subprocess.run(cmd, check=True, env=env, cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                times.append((n_pairs, time.perf_counter() - t))
                got = sum(1 for _ in open(out + "_R1.fastq", "rb")) // 4
                if got != n_pairs:
                    raise RuntimeError("reference wrote %d pairs, expected %d" % (got, n_pairs))
                for suffix in ("_R1.fastq", "_R2.fastq", "_abundance.txt"):
                    if os.path.exists(out + suffix):
                        os.remove(out + suffix)
            (n1, t1), (n2, t2) = times
            slopes.append((n2 - n1) / max(t2 - t1, 1e-9))
            whole.append(n2 / t2)
        return slopes, whole, {"pairs": [p * cores for p in pairs_per_core]}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def cpu_baseline_block(cfg, cores, repeats=1):
    """(value, dict) -- the reference arm's number and its description."""
    port_vals, cores = cpu_port_throughput(cfg, 100_000, cores, repeats=repeats)
    port = statistics.mean(port_vals)
    if reference_available():
        slopes, whole, info = reference_cli_throughput(cfg, cores, repeats=repeats)
        v = statistics.mean(slopes)
        blk = {"value": v, "unit": "pairs/s", "cores": cores, "kind": "reference",
               "sample": "unmodified InSilicoSeq 2.0.1 (baseline/_ref, Biopython stand-in), `python -m iss generate --cpus %d` on the "
                         "configs[1] genome + model at %d and %d pairs; value = slope between the two runs (start-up removed)"
                         % (cores, info["pairs"][0], info["pairs"][1]),
               "whole_command_value": statistics.mean(whole), "step_pairs": info["pairs"][1],
               "port_value": port, "port_note": "oracle/iss_oracle.c (C restatement of the same path) in %d processes" % cores}
        return v, blk
    blk = {"value": port, "unit": "pairs/s", "cores": cores, "kind": "port",
           "sample": "%d pairs (%d per core) of the configs[1] workload, oracle/iss_oracle.c in %d processes (baseline/_ref absent)"
                     % (100_000 * cores, 100_000, cores)}
    return port, blk


def run_reference_arm(args, rank, world, cfg):
    if rank != 0:
        return
    cores = usable_cores()
    reps = max(1, min(args.steps, 3))
    v, blk = cpu_baseline_block(cfg, cores, repeats=reps)
    line = {
        "impl": "reference", "metric": metric_name(cfg), "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * blk.get("step_pairs", 100_000 * cores) / v,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(cfg, args.gpus),
        "cpu_baseline": blk,
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "bounded sample repeated %d time(s); the reference's own multiprocessing CPU path on this box's host cores" % reps,
    }
    print(json.dumps(line))


def metric_name(cfg):
    return "%dbp_read_pairs_per_sec" % cfg["read_length"]


def workload_config(cfg, n):
    return {"workload": "%s: %s" % (cfg["name"], cfg["text"]),
            "pairs_per_gpu": cfg["pairs_per_gpu"], "read_length": cfg["read_length"],
            "parallelism": "reads sharded by count: chunk r -> GPU r (cpu_number r, key seed+r); 1 ncclBroadcast of genome+model at init",
            "l2": "each step streams GBs of FASTQ per GPU through L2 (126 MB), far more than L2 holds; the packed genome is "
                  "L2-resident for the 5 Mbp workloads by design and HBM-resident (750 MB) for configs[4]"}


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


def kernel_source_hash():
    h = hashlib.sha256()
    for f in ("iss_kernels.cu", "iss_device.cuh"):
        h.update(open(os.path.join(ROOT, "insilicoseq_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def rank_items(cfg, eng_mod, generator, abundance, names, lens, rank, world, L):
    """Engine work items of this rank: chunk `rank` of the `--cpus world` division of world * pairs_per_gpu pairs."""
    per_gpu = cfg["pairs_per_gpu"]
    if len(names) == 1:
        return eng_mod.worker_items([(0, per_gpu)], rank, SEED, lens, L), per_gpu
    stubs = [Rec(nm, range(n)) for nm, n in zip(names, lens)]
    np.random.seed(42)
    ab = abundance.lognormal(names)

    class M(object):
        read_length = L

    n_reads = 2 * per_gpu * world
    index = {id(r): k for k, r in enumerate(stubs)}
    chunks = list(generator.generate_work_divider(stubs, None, ab, n_reads, None, None, M, "bench", per_gpu))
    work = [(index[id(r)], n) for r, n, _ in chunks[rank]]
    return eng_mod.worker_items(work, rank, SEED, lens, L), sum(n for _, n in work)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS))
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU per step (default: the BASELINE workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--device-only", action="store_true", help="kernel A/B runs: only the device-resident leg")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.pairs:
        cfg["pairs_per_gpu"] = args.pairs
    from insilicoseq_b200 import parallel

    rank, world, local = parallel.env_rank()
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"  # stdout carries exactly one JSON line (NCCL's version banner goes there)
    if args.impl == "reference":
        run_reference_arm(args, rank, world, CONFIGS[1] if args.config != 3 else cfg)
        return
    if args.warmup < 3:
        args.warmup = 3
    host_cpus = usable_cores()  # before the NUMA binding below narrows the affinity mask

    dist = parallel.init_control_plane(rank, world)  # gloo, control plane only: ids, barriers, max over ranks

    from insilicoseq_b200 import abundance, generator
    from insilicoseq_b200 import engine as eng_mod
    from insilicoseq_b200 import error_models

    numa_node = parallel.bind_to_device_numa(local)  # before the pinned ring is allocated
    eng = eng_mod.Engine(local)
    L = cfg["read_length"]
    names, lens = record_lengths(cfg["genome"])
    philox_blocks = None
    t_setup = time.perf_counter()
    if rank == 0:
        model = error_models.KDErrorModel(os.path.join(error_models.PROFILE_DIR, cfg["model"]))
        if int(model.read_length) != L:
            model = model.truncate(L)
        recs = make_records(cfg["genome"])
        eng.set_genome([("warm", b"ACGT" * 64)])  # the first launch of a process loads the module: keep that out of the timing
        t_pack = time.perf_counter()
        eng.register_records(recs)
        t_pack = time.perf_counter() - t_pack
        del recs
        eng.set_model(model)
        philox_blocks = philox_blocks_per_pair(model, L)
    else:
        t_pack = 0.0
    if world > 1:
        uid = parallel.share_from_root(dist, eng_mod.Engine.unique_id() if rank == 0 else None, rank)
        eng.comm_init_rank(world, rank, uid)
        eng.bcast(0, read_length=L, record_names=names, record_lengths=lens)
    t_setup = time.perf_counter() - t_setup
    items, my_pairs = rank_items(cfg, eng_mod, generator, abundance, names, lens, rank, world, L)

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

    def reduce_sum(x):
        return parallel.sum_over_ranks(dist, x)

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
    pairs_step = reduce_sum(my_pairs)
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
            print(json.dumps({"metric": metric_name(cfg), "value": value, "unit": "pairs/s", "n_gpus": world,
                              "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                              "roofline": {"avg_launch_ms": k_ms / max(launches, 1), "achieved": achieved, "frac": achieved / peak},
                              "config": {"workload": cfg["name"]}, "partial": "device-only A/B run"}))
        eng.close()
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
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
    d2h_step = reduce_sum(e1["d2h_bytes"] - e0["d2h_bytes"]) / args.steps
    h2d_step = reduce_sum(e1["h2d_bytes"] - e0["h2d_bytes"]) / args.steps

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
    gz_step = reduce_sum(gz_bytes) / args.steps

    # ---- files written: generator.worker_iterator (the reference's worker API) on a bounded sample -----------------
    files = files_leg(cfg, eng, eng_mod, generator, items, my_pairs, rank, world, barrier, reduce_max, reduce_sum,
                      min(args.steps, 3), names, lens, L)

    if rank == 0:
        traffic, traffic_note = None, "no ncu capture recorded in profiles/latest.json"
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "latest.json")))
            if prof.get("kernel_source_sha16") == kernel_source_hash() and args.config == prof.get("config", 1):
                traffic = prof.get("dram_bytes_per_launch")
                traffic_note = "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of one launch (%s)" % prof.get("round")
            else:
                traffic_note = ("profiles/latest.json was captured for other kernel sources or another workload (%s, sha %s): "
                                "not reported" % (prof.get("round"), prof.get("kernel_source_sha16")))
        except Exception:  # noqa: BLE001
            pass
        st = eng.stats()
        line = {
            "metric": metric_name(cfg), "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 (integer thresholds on Philox4x32 words; bytes out)", "data": "synthetic",
            "config": workload_config(cfg, world),
            "e2e": {"value": e2e_value, "unit": "pairs/s", "ms_per_step": ms_e / args.steps,
                    "h2d_bytes_per_step": int(h2d_step), "d2h_bytes_per_step": int(d2h_step),
                    "d2h_GBps_per_gpu": d2h_step / world * args.steps / (ms_e * 1e-3) / 1e9},
            "e2e_gzip": {"value": pairs_step * args.steps / (ms_z * 1e-3), "unit": "pairs/s", "ms_per_step": ms_z / args.steps,
                         "d2h_bytes_per_step": int(gz_step),
                         "compression_ratio": d2h_step / max(gz_step, 1),
                         "note": "iss generate --compress: gzip (RFC 1952) members built on the GPU, decompresses to the same FASTQ"},
            "e2e_files": files,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src, "kernel": "iss_generate_kernel",
                         "algorithmic_bytes_per_launch": algo_bytes / max(launches, 1),
                         "avg_launch_ms": k_ms / max(launches, 1),
                         "note": "the kernel is integer-issue bound, not HBM bound (SURVEY.md 8d); see DESIGN.md section 6",
                         "philox_blocks_per_pair": philox_blocks, "philox_blocks_per_s": philox_blocks * value},
            "clocks": clk,
            "host": {"numa_node_rank0": numa_node, "ring_numa_node_rank0": st["ring_numa_node"], "ring_alloc_mode": st["ring_alloc_mode"],
                     "logical_cpus": os.cpu_count(), "usable_cpus": host_cpus,
                     "setup_s": round(t_setup, 2), "set_genome_s": round(t_pack, 3)},
        }
        if world == 1 and not args.no_cpu_baseline and args.config == 1:
            _, line["cpu_baseline"] = cpu_baseline_block(cfg, host_cpus, repeats=1)
        print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def files_leg(cfg, eng, eng_mod, generator, items, my_pairs, rank, world, barrier, reduce_max, reduce_sum, steps, names, lens, L):
    """generator.worker_iterator -> {prefix}_R1.fastq/_R2.fastq written by the native writer threads; tmpfs when it has the
    room (the FS cost is then a memcpy into the page cache), else the temp dir.  A bounded sample: <= 4 M pairs per GPU."""
    sample = min(my_pairs, 4_000_000)
    need = int(sample * (2 * L + 40) * 2 * 1.05)
    target = None
    for cand in ("/dev/shm", tempfile.gettempdir()):
        try:
            if shutil.disk_usage(cand).free > need * world * 1.5 + (1 << 30):
                target = cand
                break
        except OSError:
            continue
    if target is None:
        return {"value": None, "note": "no file system with %d MB free" % (need * world >> 20)}
    # the first `sample` pairs of the rank's items
    sub, left = [], sample
    for it in items:
        if left <= 0:
            break
        take = min(it[2], left)
        sub.append((it[0], it[1], take, it[3], it[4], it[5]))
        left -= take
    d = tempfile.mkdtemp(prefix="iss_files_r%d_" % rank, dir=target)
    prefix = os.path.join(d, "w")
    try:
        wt = max(2, min(16, usable_cores()))

        def run(null=False, gz=False):
            t = time.perf_counter()
            if null:
                st = eng.write_files(sub, None, None, None)
            elif gz:
                st = eng.write_files(sub, prefix + "_R1.fastq.gz", prefix + "_R2.fastq.gz", None, append=True, truncate=True,
                                     compress=True, writer_threads=wt)
            else:
                st = eng.write_files(sub, prefix + "_R1.fastq", prefix + "_R2.fastq", prefix + ".vcf", truncate=True, writer_threads=wt)
            return time.perf_counter() - t, st

        run()
        barrier()
        t0 = time.perf_counter()
        tot = None
        for _ in range(steps):
            _, tot = run()
        dt = reduce_max(time.perf_counter() - t0)
        barrier()
        ok = os.path.getsize(prefix + "_R1.fastq") == tot["bytes_r1"]
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            run(null=True)
        dt_null = reduce_max(time.perf_counter() - t0)
        barrier()
        run(gz=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            _, tot_gz = run(gz=True)
        dt_gz = reduce_max(time.perf_counter() - t0)
        barrier()
        pairs = reduce_sum(sample)
        nbytes = reduce_sum(tot["bytes_r1"] + tot["bytes_r2"])
        return {"value": pairs * steps / dt, "unit": "pairs/s", "clock": "host wall clock, max over ranks",
                "target": "tmpfs %s" % target if target == "/dev/shm" else target, "pairs_per_step": int(pairs),
                "bytes_per_step": int(nbytes), "GBps_written": nbytes * steps / dt / 1e9, "files_ok": bool(ok),
                "discard_value": pairs * steps / dt_null, "gzip_value": pairs * steps / dt_gz,
                "gzip_bytes_per_step": int(reduce_sum(tot_gz["bytes_r1"] + tot_gz["bytes_r2"])), "writer_threads": wt,
                "note": "Engine.write_files (what generator.worker_iterator calls): native pwrite threads from the pinned ring; "
                        "discard_value = same call with no output paths (copy to the host, no file system); gzip_value = --compress members "
                        "appended to .fastq.gz files"}
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
