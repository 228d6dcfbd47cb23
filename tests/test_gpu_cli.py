"""`iss generate` drop-in, end to end on the GPU: files written by the CLI vs the oracle run over the same work
chunks; GPU-count invariance (the multi-GPU path with its NCCL broadcast) when >= 2 devices are visible."""
import ctypes
import gzip
import os

import numpy as np
import pytest

import fixtures
import oracle
from insilicoseq_b200 import _lib, abundance, app, fasta, generator

pytestmark = pytest.mark.gpu


def _ndev():
    n = ctypes.c_int(0)
    _lib.check(_lib.load().iss_device_count(ctypes.byref(n)))
    return n.value


def _write_inputs(tmp_path):
    recs = [("chr%d" % k, fixtures.synth_genome(500 + k, 6000 + 2500 * k)) for k in range(5)]
    recs.append(("tiny_contig", fixtures.synth_genome(99, 90)))  # shorter than the read length: skipped with a warning
    fa = tmp_path / "genomes.fasta"
    with open(fa, "w") as fh:
        for name, seq in recs:
            fh.write(">%s some description\n" % name)
            for i in range(0, len(seq), 70):
                fh.write(seq[i:i + 70] + "\n")
    ab = tmp_path / "abundance.txt"
    w = [0.3, 0.25, 0.2, 0.15, 0.09, 0.01]
    with open(ab, "w") as fh:
        for (name, _), x in zip(recs, w):
            fh.write("%s\t%s\n" % (name, x))
    return str(fa), str(ab), recs, dict(zip([n for n, _ in recs], w))


def _expected(recs, ab, model_spec, n_reads, cpus, seed, **kw):
    class M:
        read_length = oracle.OracleModel(fixtures.model_path(model_spec)).read_length

    records = [fasta.Record(seq.encode(), id=name) for name, seq in recs]
    chunk_size = -((n_reads // 2) // -cpus)
    chunks = list(generator.generate_work_divider(records, None, ab, n_reads, None, None, M(), "x", chunk_size))
    om = oracle.OracleModel(fixtures.model_path(model_spec))
    r1, r2 = [], []
    for cpu, work in list(enumerate(chunks))[:cpus]:
        a, b = oracle.run_worker(om, [(r.id, r.seq, n) for r, n, _ in work], cpu, seed, **kw)
        r1.append(a)
        r2.append(b)
    return b"".join(r1), b"".join(r2)


def test_cli_matches_oracle(tmp_path):
    fa, abf, recs, ab = _write_inputs(tmp_path)
    out = str(tmp_path / "run1")
    app.main(["generate", "--genomes", fa, "--abundance_file", abf, "--model", "hiseq", "-n", "4000", "--seed", "42",
              "--cpus", "3", "--gpus", "1", "-o", out, "-q"])
    e1, e2 = _expected(recs, ab, "profile:HiSeq", 4000, 3, 42)
    assert open(out + "_R1.fastq", "rb").read() == e1
    assert open(out + "_R2.fastq", "rb").read() == e2
    assert not os.path.exists(out + ".iss.tmp.genomes.fasta") and not os.path.exists(out + ".iss.tmp.0_R1.fastq")


def test_cli_gc_bias_amplicon_compress(tmp_path):
    fa, abf, recs, ab = _write_inputs(tmp_path)
    out = str(tmp_path / "run2")
    app.main(["generate", "-g", fa, "-b", abf, "--model", "novaseq", "-n", "1500", "--seed", "7", "--cpus", "2", "--gpus", "1",
              "-o", out, "-q", "--gc_bias", "-z", "--sequence_type", "amplicon"])
    e1, e2 = _expected(recs, ab, "profile:NovaSeq", 1500, 2, 7, amplicon=True, gc_bias=True)
    assert gzip.open(out + "_R1.fastq.gz", "rb").read() == e1
    assert gzip.open(out + "_R2.fastq.gz", "rb").read() == e2


def test_cli_lognormal_abundance_is_the_references_draw(tmp_path):
    fa, abf, recs, ab = _write_inputs(tmp_path)
    out = str(tmp_path / "run3")
    app.main(["generate", "-g", fa, "--model", "hiseq", "-n", "2000", "--seed", "11", "--cpus", "2", "--gpus", "1", "-o", out, "-q"])
    np.random.seed(11)  # load_error_model seeds; abundance.lognormal is the first consumer (generator.py:397-400,589)
    exp = abundance.lognormal([n for n, _ in recs])
    got = dict((ln.split()[0], float(ln.split()[1])) for ln in open(out + "_abundance.txt"))
    assert got == {k: float(v) for k, v in exp.items()}
    e1, _ = _expected(recs, exp, "profile:HiSeq", 2000, 2, 11)
    assert open(out + "_R1.fastq", "rb").read() == e1


def test_gpu_count_invariance(tmp_path):
    if _ndev() < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    fa, abf, recs, ab = _write_inputs(tmp_path)
    outs = []
    for g in sorted({1, 2, min(_ndev(), 8)}):
        out = str(tmp_path / ("g%d" % g))
        app.main(["generate", "-g", fa, "-b", abf, "--model", "miseq", "-n", "6000", "--seed", "5", "--cpus", "8",
                  "--gpus", str(g), "-o", out, "-q"])
        outs.append((open(out + "_R1.fastq", "rb").read(), open(out + "_R2.fastq", "rb").read()))
    assert all(o == outs[0] for o in outs)
    e1, e2 = _expected(recs, ab, "profile:miSeq_0.npz", 6000, 8, 5)
    assert outs[-1] == (e1, e2)


def test_cli_baseline_config0_basic_mode(tmp_path):
    """BASELINE.json configs[0]: `iss generate --genomes ecoli.fasta --mode basic -n 1000 --seed 42 --cpus 1`
    (1 record of 1330 bp, like the reference's data/ecoli.fasta) -- 500 pairs, L=125, insert 200."""
    import patched_reference as pr

    seq = fixtures.synth_genome(4242, 1330)
    fa = tmp_path / "ecoli.fasta"
    fa.write_text(">NC_fake.1 Escherichia coli fragment\n" + "\n".join(seq[i:i + 70] for i in range(0, 1330, 70)) + "\n")
    out = str(tmp_path / "c0")
    app.main(["generate", "--genomes", str(fa), "--mode", "basic", "-n", "1000", "--seed", "42", "--cpus", "1",
              "--gpus", "1", "-o", out, "-q"])
    om = oracle.OracleModel("basic", basic_table=pr.basic_phred_table(30))
    e1, e2 = oracle.run_worker(om, [("NC_fake.1", seq.encode(), 500)], 0, 42)
    assert open(out + "_R1.fastq", "rb").read() == e1
    assert open(out + "_R2.fastq", "rb").read() == e2
    assert e1.count(b"\n") == 2000


def test_cli_store_mutations_vcf(tmp_path):
    fa, abf, recs, ab = _write_inputs(tmp_path)
    out = str(tmp_path / "vcfrun")
    app.main(["generate", "-g", fa, "-b", abf, "--model", "miseq", "-n", "3000", "--seed", "21", "--cpus", "2", "--gpus", "1",
              "-o", out, "-q", "--store_mutations"])
    vcf = open(out + ".vcf", "rb").read()
    assert vcf.startswith(b"##fileformat=VCFv4.1\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")

    class M:
        read_length = 301

    records = [fasta.Record(seq.encode(), id=name) for name, seq in recs]
    chunks = list(generator.generate_work_divider(records, None, ab, 3000, None, None, M(), "x", -((1500) // -2)))
    om = oracle.OracleModel(fixtures.model_path("profile:miSeq_0.npz"), store_mutations=True)
    body = b"".join(oracle.run_worker(om, [(r.id, r.seq, n) for r, n, _ in work], cpu, 21, vcf=True)[2]
                    for cpu, work in enumerate(chunks))
    assert vcf.split(b"\n", 2)[2] == body and len(body) > 500


def test_cli_compress_with_vcf(tmp_path):
    fa, abf, recs, ab = _write_inputs(tmp_path)
    out = str(tmp_path / "gz")
    app.main(["generate", "-g", fa, "-b", abf, "--model", "novaseq", "-n", "20000", "--seed", "3", "--cpus", "3", "--gpus", "1",
              "-o", out, "-q", "--compress", "--store_mutations"])
    e1, e2 = _expected(recs, ab, "profile:NovaSeq", 20000, 3, 3)
    raw1 = open(out + "_R1.fastq.gz", "rb").read()
    assert gzip.decompress(raw1) == e1 and gzip.open(out + "_R2.fastq.gz", "rb").read() == e2
    assert len(raw1) * 2.5 < len(e1)
    assert gzip.open(out + ".vcf.gz", "rb").read().startswith(b"##fileformat=VCFv4.1")
    assert not os.path.exists(out + "_R1.fastq") and not os.path.exists(out + ".iss.tmp.0_R1.fastq.gz")
