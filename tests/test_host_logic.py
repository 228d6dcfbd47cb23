"""CPU tests of the host layer around the hot path: FASTA parsing, work division (against the live reference when
it is mounted, and against a committed expectation otherwise), abundance draws, CLI parsing, and the
world_size-2 sharding logic over gloo with the oracle standing in for the GPU."""
import io
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import fixtures
from insilicoseq_b200 import abundance, app, fasta, generator, parallel, util


class _Model:
    read_length = 126


def _records():
    return [fasta.Record(fixtures.synth_genome(300 + k, 3000 + 1700 * k).encode(), id="g%d" % k) for k in range(7)]


def _chunks(n_reads, cpus, ab):
    recs = _records()
    chunk_size = -((n_reads // 2) // -cpus)
    return [[(r.id, n) for r, n, _ in c]
            for c in generator.generate_work_divider(recs, None, ab, n_reads, None, None, _Model(), "x", chunk_size)]


def test_fasta_parse_roundtrip(tmp_path):
    p = tmp_path / "a.fa"
    p.write_text(">id1 some description\nACGT\nacgtNN\n\n>id2\nTTTT\n>\nGG\n")
    recs = list(fasta.parse(str(p)))
    assert [(r.id, r.seq) for r in recs] == [("id1", b"ACGTacgtNN"), ("id2", b"TTTT"), ("", b"GG")]
    assert recs[0].description == "id1 some description"
    out = io.StringIO()
    fasta.write(recs[:2], out)
    assert out.getvalue() == ">id1 some description\nACGTacgtNN\n>id2\nTTTT\n"
    assert util.count_records(str(p)) == ["id1", "id2", ""]


def test_convert_n_reads():
    assert util.convert_n_reads("1000000") == 1000000
    assert util.convert_n_reads("0.5M") == 500000
    assert util.convert_n_reads("2k") == 2000 and util.convert_n_reads("1G") == 10 ** 9
    with pytest.raises(SystemExit):
        util.convert_n_reads("12x")


def test_work_divider_expected():
    np.random.seed(42)
    ab = abundance.lognormal(["g%d" % k for k in range(7)])
    got = _chunks(5000, 3, ab)
    exp_path = os.path.join(fixtures.GOLDEN, "work_divider_expected.json")
    exp = json.load(open(exp_path))
    assert [[list(x) for x in c] for c in got] == exp["chunks"]
    assert sum(n for c in got for _, n in c) in (2500, 2501)


@pytest.mark.reference
def test_work_divider_and_abundance_match_live_reference():
    import patched_reference as pr

    if not pr.reference_available():
        pytest.skip("reference not mounted")
    iss = pr.import_reference()
    from iss import abundance as rab
    from iss import generator as rgen

    ids = ["g%d" % k for k in range(7)]
    for name in ("uniform", "halfnormal", "exponential", "lognormal", "zero_inflated_lognormal"):
        np.random.seed(7)
        a = getattr(abundance, name)(ids)
        np.random.seed(7)
        b = getattr(rab, name)(ids)
        assert a == b, name
    np.random.seed(42)
    ab = abundance.lognormal(ids)
    recs = _records()
    ref_recs = [pr.make_record(r.seq.decode(), r.id) for r in recs]
    for n_reads, cpus in [(5000, 3), (1001, 2), (77, 8)]:
        chunk_size = -((n_reads // 2) // -cpus)
        ref = [[(r.id, n) for r, n, _ in c]
               for c in rgen.generate_work_divider(ref_recs, None, ab, n_reads, None, None, _Model(), "x", chunk_size)]
        assert _chunks(n_reads, cpus, ab) == ref


def test_cli_accepts_the_reference_flags():
    p = app.build_parser()
    ns = p.parse_args(["generate", "--genomes", "a.fa", "b.fa", "--model", "NovaSeq", "-n", "1M", "--seed", "42",
                       "--cpus", "8", "-o", "out", "--gc_bias", "-z", "--sequence_type", "amplicon", "--store_mutations",
                       "--fragment-length", "400", "--fragment-length-sd", "20", "--abundance", "uniform"])
    assert ns.genomes == ["a.fa", "b.fa"] and ns.cpus == 8 and ns.gc_bias and ns.compress and ns.store_mutations
    assert ns.fragment_length == 400 and ns.fragment_length_sd == 20 and ns.sequence_type == "amplicon"
    with pytest.raises(SystemExit):  # mutually exclusive abundance group, as in the reference
        p.parse_args(["generate", "-o", "x", "--abundance_file", "a", "--coverage_file", "b"])


def test_chunk_to_gpu_mapping_is_a_partition():
    jobs = [(c, "w%d" % c) for c in range(11)]
    for world in (1, 2, 3, 8):
        parts = [parallel.rank_jobs(jobs, world, r) for r in range(world)]
        assert sorted(j for p in parts for j in p) == jobs
        merged = parallel.merge_in_chunk_order([[(c, ("%d," % c).encode()) for c, _ in p] for p in parts])
        assert merged == b"".join(("%d," % c).encode() for c in range(11))


GLOO_WORKER = r'''
import os, sys, json, hashlib
root = sys.argv[1]
sys.path[:0] = [root, os.path.join(root, "oracle"), os.path.join(root, "tests")]
import fixtures, oracle
from insilicoseq_b200 import parallel, fasta, generator, abundance
rank, world, _ = parallel.env_rank()
dist = parallel.init_control_plane(rank, world)
uid = parallel.share_from_root(dist, b"\x07" * 128 if rank == 0 else None, rank)   # stands in for the NCCL unique id
assert uid == b"\x07" * 128
class M: read_length = 126
recs = [fasta.Record(fixtures.synth_genome(300 + k, 3000 + 1700 * k).encode(), id="g%d" % k) for k in range(4)]
ab = abundance.uniform([r.id for r in recs])
cpus, n_reads = 4, 600
chunks = list(generator.generate_work_divider(recs, None, ab, n_reads, None, None, M(), "x", -((n_reads // 2) // -cpus)))
jobs = list(enumerate(chunks))
om = oracle.OracleModel(fixtures.model_path("profile:HiSeq"))
mine = []
for cpu, work in parallel.rank_jobs(jobs, world, rank):      # the oracle stands in for the GPU of this rank
    r1, r2 = oracle.run_worker(om, [(r.id, r.seq, n) for r, n, _ in work], cpu, 11)
    mine.append((cpu, r1))
allparts = parallel.gather_to_root(dist, mine, rank, world)
t = parallel.max_over_ranks(dist, 1.0 + rank)
assert t == float(world)
if rank == 0:
    merged = parallel.merge_in_chunk_order(allparts)
    single = b"".join(oracle.run_worker(om, [(r.id, r.seq, n) for r, n, _ in work], cpu, 11)[0] for cpu, work in jobs)
    assert merged == single
    print("GLOO_OK", hashlib.sha256(merged).hexdigest()[:12], len(merged))
if dist is not None:
    dist.barrier(); dist.destroy_process_group()
'''


def test_world_size_2_sharding_over_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(GLOO_WORKER)
    import socket

    for _ in range(2):  # a rendezvous port can be taken between probing and binding; one retry on a fresh port
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", str(port), str(script), fixtures.ROOT]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
        if out.returncode == 0:
            break
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert "GLOO_OK" in out.stdout


@pytest.mark.reference
def test_draft_and_coverage_paths_match_live_reference(tmp_path):
    """--draft / --coverage set-up (iss/abundance.py:185-317) against the reference's own functions."""
    import patched_reference as pr

    if not pr.reference_available():
        pytest.skip("reference not mounted")
    pr.import_reference()
    from iss import abundance as rab

    genomes = tmp_path / "g.fasta"
    draft1 = tmp_path / "draft1.fasta"
    with open(genomes, "w") as fh:
        for k in range(3):
            fh.write(">full%d\n%s\n" % (k, fixtures.synth_genome(600 + k, 2000 + 500 * k)))
    with open(draft1, "w") as fh:
        for k in range(4):
            fh.write(">contig%d desc\n%s\n" % (k, fixtures.synth_genome(700 + k, 300 + 211 * k)))
    ab = {"full0": 0.2, "full1": 0.3, "full2": 0.1, str(draft1): 0.4}
    for mode in ("abundance", "coverage"):
        assert abundance.expand_draft_abundance(dict(ab), [str(draft1)], mode) == rab.expand_draft_abundance(dict(ab), [str(draft1)], mode)
    cov = {"full0": 3.0, "full1": 1.5, "full2": 0.25}
    a = abundance.coverage_scaling(10000, dict(cov), str(genomes), 126)
    b = rab.coverage_scaling(10000, dict(cov), str(genomes), 126)
    assert a == b
    ids = ["full0", "full1", "full2", "contig0", "contig1", "contig2", "contig3"]
    np.random.seed(5)
    x = abundance.draft(list(ids), [str(draft1)], abundance.lognormal, str(tmp_path / "o1"))
    np.random.seed(5)
    y = rab.draft(list(ids), [str(draft1)], rab.lognormal, str(tmp_path / "o2"))
    assert set(x) == set(y) and all(abs(x[k] - y[k]) < 1e-15 for k in x)  # set() order inside draft() is not stable
    assert abundance.to_coverage(1000, 0.5, 126, 5000) == rab.to_coverage(1000, 0.5, 126, 5000)


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver times beside the GPU arm) needs no GPU: one JSON line with the
    contract's keys; `cpu_baseline.kind` is "reference" when the unmodified reference is staged under baseline/_ref (then the
    C port's rate is reported beside it), else "port"."""
    out = subprocess.run([sys.executable, os.path.join(fixtures.ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "150bp_read_pairs_per_sec" and d["unit"] == "pairs/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1
    staged = os.path.isdir(os.path.join(fixtures.ROOT, "baseline", "_ref", "iss"))
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port") and d["cpu_baseline"]["cores"] >= 1
    if staged:  # the pure-Python reference is orders of magnitude slower than its C restatement
        assert d["cpu_baseline"]["port_value"] > 5 * d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_fasta_reader_matches_line_by_line_definition(tmp_path):
    """fasta.parse cuts a memory-mapped file at line-initial '>' and strips white space with one translate per record; the
    definition it must agree with is the line-by-line one (Bio.SeqIO: id = first token of the title, sequence = the
    following lines without white space), on paths, binary handles and text handles."""
    import random

    def by_lines(data):
        title, chunks, out = None, [], []
        for line in io.BytesIO(data):
            if line.startswith(b">"):
                if title is not None:
                    out.append((title, b"".join(chunks)))
                title, chunks = line[1:].rstrip(), []
            elif title is not None:
                chunks.append(b"".join(line.split()))
        if title is not None:
            out.append((title, b"".join(chunks)))
        return [((t.decode().split(None, 1) or [""])[0], t.decode(), s) for t, s in out]

    rng = random.Random(5)
    pieces = [b">", b"\n", b"\r\n", b"ACGT", b"acgtn", b" ", b"\t", b">id desc here", b"\n>", b"\n\n", b"N" * 7, b">x", b"AC>GT",
              b"\x0b", b"\x0c", b"\r"]
    path = tmp_path / "f.fa"
    for _ in range(1500):
        data = b"".join(rng.choice(pieces) for _ in range(rng.randint(0, 14)))
        want = by_lines(data)
        path.write_bytes(data)
        for src in (str(path), io.BytesIO(data), io.StringIO(data.decode("ascii"))):
            assert [(r.id, r.description, r.seq) for r in fasta.parse(src)] == want, data
