"""BASELINE-size runs on the GPU, checked through size-independent properties and sampled pairs:
  * configs[1] (5 Mbp genome, miseq-150 model, 10 M pairs): exact byte count (closed form), record framing at sampled
    offsets, sampled pairs byte-identical to the oracle, checksum-of-checksums invariant to the batch size;
  * a configs[2]-like metagenome (many records, lognormal abundance via the reference-compatible work divider):
    whole output identical to the oracle at a size it finishes in seconds."""
import hashlib
import os

import numpy as np
import pytest

import fixtures
import oracle
from insilicoseq_b200 import abundance, engine as eng_mod, error_models, fasta, generator

pytestmark = pytest.mark.gpu
L = 150
N_PAIRS = 10_000_000
GENOME_LEN = 5_000_000


def _genome():
    rng = np.random.default_rng(20260924)
    return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=GENOME_LEN, dtype=np.uint8)].tobytes()


def _model():
    return error_models.KDErrorModel(os.path.join(error_models.PROFILE_DIR, "miSeq_0.npz")).truncate(L)


def test_configs1_full_size_properties(tmp_path):
    genome = _genome()
    name = "syn5M"
    seed, cpu = 42, 0
    want_bytes = oracle.fastq_bytes(len(name), cpu, L, 0, N_PAIRS)
    rng = np.random.RandomState(5)
    sample = np.unique(np.concatenate([rng.randint(0, N_PAIRS, size=1500), [0, 9, 10, 99, 100, 999_999, 1_000_000,
                                                                            N_PAIRS - 1]]))
    sample_off = [oracle.fastq_bytes(len(name), cpu, L, 0, int(i)) for i in sample]
    digests = {}
    picked = {}
    with eng_mod.Engine(0) as e:
        e.set_genome([(name, genome)])
        e.set_model(_model())
        items = eng_mod.worker_items([(0, N_PAIRS)], cpu, seed, e.record_lengths, e.read_length)
        for bp in (1 << 18, 777_777):
            h1, h2 = hashlib.blake2b(digest_size=16), hashlib.blake2b(digest_size=16)
            total = 0
            for b in e.generate(items, batch_pairs=bp):
                lo, hi = total, total + b.nbytes
                h1.update(memoryview(b.r1))
                h2.update(memoryview(b.r2))
                assert b.r1[-1] == 10 and b.r2[-1] == 10 and b.r1[0] == 64 and b.r2[0] == 64
                if bp == 1 << 18:
                    for i, off in zip(sample, sample_off):
                        if lo <= off < hi:
                            n = 2 * L + 6 + len("%s_%d_%d/1" % (name, i, cpu))
                            picked[int(i)] = (b.r1[off - lo:off - lo + n].tobytes(), b.r2[off - lo:off - lo + n].tobytes())
                total = hi
            assert total == want_bytes
            digests[bp] = (h1.hexdigest(), h2.hexdigest())
    assert digests[1 << 18] == digests[777_777]  # batch-size invariance over the whole 6.4 GB
    assert len(picked) == len(sample)
    # sampled pairs against the oracle (it can generate any pair range on its own: draws are counter addressed)
    npz = str(tmp_path / "miseq_150.npz")
    oracle.truncate_profile(os.path.join(error_models.PROFILE_DIR, "miSeq_0.npz"), npz, L)
    om = oracle.OracleModel(npz)
    key = oracle.make_key(seed, cpu)
    for i, (g1, g2) in picked.items():
        o1, o2 = oracle.generate(om, genome, name, key, 0, cpu, i, i + 1)
        assert g1 == o1 and g2 == o2, i
        assert g1.startswith(("@%s_%d_%d/1\n" % (name, i, cpu)).encode())


def test_metagenome_many_records_matches_oracle():
    n_rec = 400
    rng = np.random.RandomState(3)
    lens = np.exp(rng.uniform(np.log(400), np.log(60000), size=n_rec)).astype(int)
    lens[7] = 120  # shorter than the read length: skipped
    recs = [fasta.Record(fixtures.synth_genome(1000 + k, int(n)).encode(), id="contig_%04d" % k) for k, n in enumerate(lens)]
    np.random.seed(42)
    ab = abundance.lognormal([r.id for r in recs])
    model = error_models.KDErrorModel(os.path.join(error_models.PROFILE_DIR, "NovaSeq"))
    n_reads, cpus = 120_000, 3
    chunk = -((n_reads // 2) // -cpus)
    chunks = list(generator.generate_work_divider(recs, None, ab, n_reads, None, None, model, "x", chunk))
    om = oracle.OracleModel(os.path.join(error_models.PROFILE_DIR, "NovaSeq"))
    with eng_mod.Engine(0) as e:
        generator.register_records(e, recs)
        e.set_model(model)
        for cpu, work in enumerate(chunks[:cpus]):
            items = eng_mod.worker_items([(e.record_index(r), n) for r, n, _ in work], cpu, 7, e.record_lengths, e.read_length)
            got1 = b"".join(b.r1.tobytes() for b in e.generate(items, batch_pairs=5000))
            exp1, _ = oracle.run_worker(om, [(r.id, r.seq, n) for r, n, _ in work], cpu, 7)
            assert hashlib.sha256(got1).digest() == hashlib.sha256(exp1).digest()
            assert len(work) > 50  # many segments per batch


def test_thousands_of_tiny_items_with_long_names():
    """Segment planner stress: 3000 contigs, 1-3 pairs each, record ids of 40-180 characters."""
    rng = np.random.RandomState(17)
    recs = []
    for k in range(3000):
        name = "NZ_CONTIG%05d.%s" % (k, "x" * int(rng.randint(25, 165)))
        recs.append(fasta.Record(fixtures.synth_genome(5000 + k, int(rng.randint(200, 700))).encode(), id=name))
    work = [(k, int(rng.randint(1, 4))) for k in range(3000)]
    model = error_models.KDErrorModel(os.path.join(error_models.PROFILE_DIR, "HiSeq"))
    om = oracle.OracleModel(os.path.join(error_models.PROFILE_DIR, "HiSeq"))
    with eng_mod.Engine(0) as e:
        generator.register_records(e, recs)
        e.set_model(model)
        items = eng_mod.worker_items(work, 12, 99, e.record_lengths, e.read_length)
        for bp in (0, 500):
            got1, got2 = b"", b""
            for b in e.generate(items, batch_pairs=bp):
                got1 += b.r1.tobytes()
                got2 += b.r2.tobytes()
            exp1, exp2 = oracle.run_worker(om, [(recs[k].id, recs[k].seq, n) for k, n in work], 12, 99)
            assert got1 == exp1 and got2 == exp2
