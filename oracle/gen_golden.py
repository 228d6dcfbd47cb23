#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference driven by the shared
Philox stream (oracle/patched_reference.py).  Runs only where /root/reference exists (this container);
the fixtures it writes are committed so that the GPU box needs neither the reference nor this script.

    python oracle/gen_golden.py            # regenerate everything
    python oracle/gen_golden.py --check    # also run the C oracle on every case and compare (default on)

Each case is a directory tests/golden/<name>/ holding
    case.json   inputs: model file, records (literal or synthetic-by-seed), work list, seed, cpu, options
    R1.fastq, R2.fastq[, out.vcf]   what the reference wrote (worker_iterator, iss/generator.py:223-251)
    or, for the larger cases, sha256 digests of those files inside case.json.
"""
import argparse
import hashlib
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
MODELS = os.path.join(GOLDEN, "models")
sys.path.insert(0, HERE)
import philox  # noqa: E402
import patched_reference as pr  # noqa: E402
import oracle  # noqa: E402
from fixtures import synth_genome, sprinkle, materialise_records, model_path  # noqa: E402,F401

REF_PROFILES = os.path.join(pr.REFERENCE_ROOT, "iss", "profiles")


def write_stress_model(path, base_npz, seed, dense):
    """A KDE profile with adversarial indel tables (p=1.0 entries, several insert lanes at one position,
    consecutive deletions) to exercise the sequential list semantics of introduce_indels."""
    d = np.load(base_npz, allow_pickle=True)
    L = int(d["read_length"])
    rng = np.random.RandomState(seed)
    out = {k: d[k] for k in d.files}
    for orient in ("forward", "reverse"):
        ins = np.empty(L, dtype=object)
        dele = np.empty(L, dtype=object)
        for pos in range(L):
            pi = {b: 0.0 for b in "ATCG"}
            pd = {b: 0.0 for b in "ATCG"}
            if rng.rand() < dense:
                for b in "ATCG":
                    r = rng.rand()
                    pi[b] = float(1.0 if r < 0.03 else (rng.rand() * 0.5 if r < 0.3 else 0.0))
            if rng.rand() < dense:
                for b in "ATCG":
                    r = rng.rand()
                    pd[b] = float(1.0 if r < 0.05 else (rng.rand() * 0.6 if r < 0.5 else 0.0))
            ins[pos] = pi
            dele[pos] = pd
        out["ins_" + orient] = ins
        out["del_" + orient] = dele
    with open(path, "wb") as fh:
        np.savez_compressed(fh, **out)


def write_full_model(path, base_npz, seed):
    """Every position carries all four insertion lanes and a deletion probability for every base (small values, so reads
    still differ): the indel site lists are as long as they can get -- 4 (L-1) insertion sites and L-1 deletion sites."""
    d = np.load(base_npz, allow_pickle=True)
    L = int(d["read_length"])
    rng = np.random.RandomState(seed)
    out = {k: d[k] for k in d.files}
    for orient in ("forward", "reverse"):
        ins = np.empty(L, dtype=object)
        dele = np.empty(L, dtype=object)
        for pos in range(L):
            ins[pos] = {b: float(rng.rand() * 0.02) for b in "ATCG"}
            dele[pos] = {b: float(rng.rand() * 0.05) for b in "ATCG"}
        out["ins_" + orient] = ins
        out["del_" + orient] = dele
    with open(path, "wb") as fh:
        np.savez_compressed(fh, **out)


def write_tie_model(path, base_npz, seed):
    """Quality CDF rows whose 32-bit thresholds collide in their top 16 bits: every row is a few clusters of CDF values
    2**-26 .. 2**-19 apart (plus plateaus of equal values), so a draw whose top 16 bits hit a cluster must be resolved by the
    exact 32-bit compare, several entries deep.  The kernel searches quality rows on 16 bits first (iss_kernels.cu)."""
    d = np.load(base_npz, allow_pickle=True)
    L = int(d["read_length"])
    rng = np.random.RandomState(seed)
    out = {k: d[k] for k in d.files}
    for orient in ("forward", "reverse"):
        hist = np.empty(4, dtype=object)
        for b in range(4):
            rows = []
            for _pos in range(L):
                centers = np.sort(rng.rand(6))
                vals = []
                for c in centers:
                    k = int(rng.randint(3, 9))
                    step = 2.0 ** -int(rng.randint(19, 27))
                    base = np.floor(c * 65536) / 65536 + rng.randint(0, 200) * 2.0 ** -32
                    inc = np.cumsum(rng.randint(0, 3, size=k)) * step  # zeros -> equal neighbours (plateaus)
                    vals.extend(base + inc)
                vals = np.sort(np.clip(np.array(vals), 0.0, 1.0))[:40]
                row = np.concatenate([vals, np.ones(41 - len(vals))])
                rows.append(row.astype(np.float64))
            hist[b] = rows
        out["quality_hist_" + orient] = hist
        out["mean_count_" + orient] = np.array([5, 20, 40, 35], dtype=np.int64)
    with open(path, "wb") as fh:
        np.savez_compressed(fh, **out)


def write_flat_quality_model(path, base_npz, levels, lo=0):
    """Every quality row is the CDF of a uniform distribution over `levels` phred values lo .. lo + levels - 1.
    levels=60: 60 distinct 16-bit prefixes per row, so long reads overflow the kernel's tie queue now and then;
    lo=2, levels=3: phred 2-4, i.e. about every other base a substitution -- the substitution list of a tile overflows and
    the kernel's exact pass takes over."""
    d = np.load(base_npz, allow_pickle=True)
    L = int(d["read_length"])
    out = {k: d[k] for k in d.files}
    row = np.concatenate([np.zeros(lo), np.arange(1, levels + 1) / float(levels)])
    row = np.concatenate([row, np.ones(max(0, 41 - len(row)))])
    for orient in ("forward", "reverse"):
        hist = np.empty(4, dtype=object)
        for b in range(4):
            hist[b] = [row.copy() for _ in range(L)]
        out["quality_hist_" + orient] = hist
        out["mean_count_" + orient] = np.array([1, 2, 3, 4], dtype=np.int64)
    with open(path, "wb") as fh:
        np.savez_compressed(fh, **out)


def write_long_model(path, base_npz, reps):
    """The per-position arrays of a bundled profile repeated `reps` times: a read length beyond ~860, where a record no
    longer fits a 64-pair tile and the kernel falls back to 32-pair tiles."""
    d = np.load(base_npz, allow_pickle=True)
    L = int(d["read_length"])
    out = {k: d[k] for k in d.files}
    out["read_length"] = np.int64(L * reps)
    for orient in ("forward", "reverse"):
        q = d["quality_hist_" + orient]
        hist = np.empty(len(q), dtype=object)
        for b in range(len(q)):
            hist[b] = list(q[b]) * reps if len(q[b]) else []
        out["quality_hist_" + orient] = hist
        for key in ("subst_choices_", "ins_", "del_"):
            arr = d[key + orient]
            rep = np.empty(len(arr) * reps, dtype=object)
            for k in range(len(rep)):
                rep[k] = arr[k % len(arr)]
            out[key + orient] = rep
    with open(path, "wb") as fh:
        np.savez_compressed(fh, **out)


def sha(b):
    return hashlib.sha256(b).hexdigest()


def build_cases():
    os.makedirs(MODELS, exist_ok=True)
    # model files used by the fixtures (data, not source): small test profile + derived stress/truncated profiles
    shutil.copyfile(os.path.join(pr.REFERENCE_ROOT, "data", "ecoli.npz"), os.path.join(MODELS, "ecoli.npz"))
    write_stress_model(os.path.join(MODELS, "stress20.npz"), os.path.join(MODELS, "ecoli.npz"), 5, dense=0.5)
    write_stress_model(os.path.join(MODELS, "stress20_sparse.npz"), os.path.join(MODELS, "ecoli.npz"), 6, dense=0.12)
    oracle.truncate_profile(os.path.join(REF_PROFILES, "miSeq_0.npz"), os.path.join(MODELS, "miseq_150.npz"), 150)
    write_full_model(os.path.join(MODELS, "full20.npz"), os.path.join(MODELS, "ecoli.npz"), 7)
    oracle.truncate_profile(os.path.join(MODELS, "stress20.npz"), os.path.join(MODELS, "stress3.npz"), 3)
    for n in (2, 16, 17):  # around the 16-base chunk of the gather and the smallest length the indel loop accepts
        oracle.truncate_profile(os.path.join(MODELS, "stress20.npz"), os.path.join(MODELS, "stress%d.npz" % n), n)

    write_tie_model(os.path.join(MODELS, "ties20.npz"), os.path.join(MODELS, "ecoli.npz"), 11)
    write_long_model(os.path.join(MODELS, "novaseq_x6.npz"), os.path.join(REF_PROFILES, "NovaSeq"), 6)
    write_flat_quality_model(os.path.join(MODELS, "flat60_x6.npz"), os.path.join(MODELS, "novaseq_x6.npz"), 60)
    write_flat_quality_model(os.path.join(MODELS, "lowq20.npz"), os.path.join(MODELS, "ecoli.npz"), 3, lo=2)

    cases = []

    def rec(name, length=None, seed=None, literal=None, **sp):
        r = {"id": name}
        if literal is not None:
            r["literal"] = literal
        else:
            r.update({"synth_seed": seed, "length": length})
        if sp:
            r["sprinkle"] = sp
        return r

    cases.append(dict(name="novaseq_single", model="profile:NovaSeq", seed=42, cpu=0,
                      records=[rec("syn20k", 20000, 11)], work=[[0, 40]], inline=True))
    cases.append(dict(name="hiseq_multi_short", model="profile:HiSeq", seed=7, cpu=3,
                      records=[rec("contig_1", 3000, 21), rec("NC_000913.3", 5000, 22), rec("tiny", 800, 23),
                               rec("too_short", 100, 24), rec("c5", 1500, 25)],
                      work=[[0, 25], [1, 30], [2, 12], [3, 5], [4, 9]], inline=True))
    cases.append(dict(name="miseq_legacy_indels", model="profile:MiSeq", seed=123, cpu=1,
                      records=[rec("gen10k", 10000, 31)], work=[[0, 60]], inline=True))
    cases.append(dict(name="miseq150_bench_model", model="models/miseq_150.npz", seed=42, cpu=0,
                      records=[rec("syn30k", 30000, 41)], work=[[0, 120]], inline=True))
    cases.append(dict(name="stress_indels_dense", model="models/stress20.npz", seed=99, cpu=2,
                      records=[rec("s1", 400, 51), rec("s2", 90, 52)], work=[[0, 150], [1, 60]], inline=True))
    cases.append(dict(name="stress_indels_sparse_iupac", model="models/stress20_sparse.npz", seed=1234567, cpu=0,
                      records=[rec("mixed", 600, 61, frac_lower=0.3, n_runs=6, iupac="RYKMSWBDHVn")],
                      work=[[0, 200]], inline=True))
    cases.append(dict(name="ecoli_kde_vcf", model="models/stress20_sparse.npz", seed=5, cpu=0, store_mutations=True,
                      records=[rec("vcfrec", 500, 71)], work=[[0, 80]], inline=True))
    cases.append(dict(name="nextseq_large_digest", model="profile:nextSeq.npz", seed=2024, cpu=11,
                      records=[rec("big", 200000, 81)], work=[[0, 1200]], inline=False))
    cases.append(dict(name="novaseq_metagenome_digest", model="profile:NovaSeq", seed=31337, cpu=0,
                      records=[rec("g%d" % k, 4000 + 977 * k, 90 + k) for k in range(12)],
                      work=[[k, 20 + 7 * k] for k in range(12)], inline=False))
    cases.append(dict(name="amplicon_miseq", model="profile:MiSeq", seed=42, cpu=0, sequence_type="amplicon",
                      records=[rec("my_amplicon", 800, 101)], work=[[0, 25]], inline=True))
    cases.append(dict(name="gc_bias_hiseq", model="profile:HiSeq", seed=8, cpu=1, gc_bias=True,
                      records=[rec("gcrec", 6000, 111)], work=[[0, 120]], inline=True))
    cases.append(dict(name="fragment_length_novaseq", model="profile:NovaSeq", seed=77, cpu=0,
                      fragment_length=420, fragment_sd=35,
                      records=[rec("fragrec", 9000, 121), rec("shortfrag", 430, 122)], work=[[0, 60], [1, 30]],
                      inline=True))
    cases.append(dict(name="basic_mode", model="basic", seed=42, cpu=0, store_mutations=True,
                      records=[rec("ecoli_like", 1330, 131), rec("b2", 700, 132)], work=[[0, 150], [1, 40]], inline=True))
    cases.append(dict(name="basic_mode_fragment", model="basic", seed=3, cpu=1, fragment_length=450, fragment_sd=0,
                      records=[rec("my_genome", 1000, 133)], work=[[0, 60]], inline=True))
    cases.append(dict(name="perfect_mode", model="perfect", seed=9, cpu=0,
                      records=[rec("pg", 2500, 134)], work=[[0, 400]], inline=True))
    cases.append(dict(name="full_indel_tables", model="models/full20.npz", seed=2718, cpu=5, store_mutations=True,
                      records=[rec("f1", 700, 141), rec("f2", 64, 142)], work=[[0, 160], [1, 40]], inline=True))
    cases.append(dict(name="read_length_3", model="models/stress3.npz", seed=31, cpu=0,
                      records=[rec("r3", 300, 143, frac_lower=0.2, n_runs=2, iupac="RN")], work=[[0, 300]], inline=True))
    for n in (2, 16, 17):
        cases.append(dict(name="read_length_%d" % n, model="models/stress%d.npz" % n, seed=100 + n, cpu=1, store_mutations=(n == 17),
                          records=[rec("q%d" % n, 200 + n, 150 + n, frac_lower=0.1, n_runs=1, iupac="N"), rec("pure%d" % n, 90, 160 + n)],
                          work=[[0, 90], [1, 70]], inline=True))
    cases.append(dict(name="quality_ties_16bit", model="models/ties20.npz", seed=606, cpu=0,
                      records=[rec("tie_rec", 3000, 170)], work=[[0, 2500]], inline=False))
    cases.append(dict(name="long_reads_906", model="models/novaseq_x6.npz", seed=907, cpu=2,
                      records=[rec("long_a", 9000, 171), rec("long_b", 2100, 172)], work=[[0, 70], [1, 31]], inline=False))
    cases.append(dict(name="low_quality_overflow", model="models/lowq20.npz", seed=808, cpu=1, store_mutations=True,
                      records=[rec("lq", 2000, 173)], work=[[0, 400]], inline=False))
    cases.append(dict(name="long_names_cpu_digits", model="profile:NovaSeq", seed=4242, cpu=40950,
                      records=[rec("NZ_" + "contig.with.a.long.name|" * 3 + "end", 2600, 144), rec("x", 1900, 145)],
                      work=[[0, 33], [1, 70]], inline=True))
    return cases


def run_reference(case, tmp):
    recs = materialise_records(case)
    fl, fsd = case.get("fragment_length"), case.get("fragment_sd")
    if case["model"] in ("basic", "perfect"):
        em = pr.load_simple_model(case["model"], fl, fsd, bool(case.get("store_mutations")))
    else:
        em = pr.load_kde_model(model_path(case["model"]), fl, fsd, bool(case.get("store_mutations")))
    work = [(pr.make_record(recs[ri][1], recs[ri][0]), n, "default") for ri, n in case["work"]]
    prefix = os.path.join(tmp, "w")
    pr.run_worker(work, em, case["cpu"], prefix, case["seed"], case.get("sequence_type", "metagenomics"),
                  bool(case.get("gc_bias")))
    r1 = open(prefix + "_R1.fastq", "rb").read()
    r2 = open(prefix + "_R2.fastq", "rb").read()
    vcf = open(prefix + ".vcf", "rb").read()
    return r1, r2, vcf


def run_c_oracle(case):
    recs = materialise_records(case)
    fl, fsd = case.get("fragment_length"), case.get("fragment_sd")
    ft = pr.fragment_table(fl, fsd) if fl is not None else None
    if case["model"] in ("basic", "perfect"):
        om = oracle.OracleModel(case["model"], fl, fsd, bool(case.get("store_mutations")), frag_table=ft,
                                basic_table=pr.basic_phred_table(30))
    else:
        om = oracle.OracleModel(model_path(case["model"]), fl, fsd, bool(case.get("store_mutations")), frag_table=ft)
    work = [(recs[ri][0], recs[ri][1].encode(), n) for ri, n in case["work"]]
    return oracle.run_worker(om, work, case["cpu"], case["seed"],
                             amplicon=case.get("sequence_type") == "amplicon", gc_bias=bool(case.get("gc_bias")),
                             vcf=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    assert pr.reference_available(), "needs /root/reference"
    cases = build_cases()
    ok = True
    for case in cases:
        if args.only and case["name"] != args.only:
            continue
        cdir = os.path.join(GOLDEN, case["name"])
        os.makedirs(cdir, exist_ok=True)
        with tempfile.TemporaryDirectory() as tmp:
            r1, r2, vcf = run_reference(case, tmp)
        meta = {k: v for k, v in case.items() if k != "inline"}
        if case["model"].startswith("profile:"):
            meta["model"] = case["model"]
        meta["sha256"] = {"R1": sha(r1), "R2": sha(r2), "vcf": sha(vcf)}
        meta["n_bytes"] = {"R1": len(r1), "R2": len(r2), "vcf": len(vcf)}
        if case["inline"]:
            open(os.path.join(cdir, "R1.fastq"), "wb").write(r1)
            open(os.path.join(cdir, "R2.fastq"), "wb").write(r2)
            if case.get("store_mutations"):
                open(os.path.join(cdir, "out.vcf"), "wb").write(vcf)
        json.dump(meta, open(os.path.join(cdir, "case.json"), "w"), indent=1, sort_keys=True)
        msg = "%-32s R1 %7d B  vcf %6d B" % (case["name"], len(r1), len(vcf))
        if not args.no_check:
            c1, c2, cv = run_c_oracle(case)
            same = (c1 == r1, c2 == r2, (cv == vcf) if case.get("store_mutations") else True)
            msg += "   C-oracle: R1 %s R2 %s VCF %s" % same
            ok = ok and all(same)
        print(msg, flush=True)
    if not ok:
        sys.exit("C oracle differs from the Philox-driven reference")


if __name__ == "__main__":
    main()
