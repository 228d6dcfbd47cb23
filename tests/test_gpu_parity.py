"""Parity of the CUDA path (through the C ABI) with the oracle and the reference-written golden fixtures.
Bit-exact: this is byte/integer work, the bar is identical FASTQ bytes."""
import hashlib
import os

import numpy as np
import pytest

import fixtures
import oracle
import philox
from gpu_helpers import first_diff, product_model, run_product

pytestmark = pytest.mark.gpu
CASES = fixtures.list_cases()


def test_device_philox_known_answers():
    from insilicoseq_b200.engine import Engine

    with Engine(0) as e:
        assert e.philox((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
        assert e.philox((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
        ctr, key = (0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)
        assert e.philox(ctr, key) == philox.philox4x32_10(ctr, key)


@pytest.mark.parametrize("name", CASES)
def test_golden_fixture(name):
    case, exp = fixtures.load_case(name)
    r1, r2, stats = run_product(case)
    if exp["R1"] is not None:
        assert first_diff(r1, exp["R1"]) is None
        assert first_diff(r2, exp["R2"]) is None
    assert hashlib.sha256(r1).hexdigest() == case["sha256"]["R1"]
    assert hashlib.sha256(r2).hexdigest() == case["sha256"]["R2"]
    if case.get("store_mutations"):
        if exp["vcf"] is not None:  # the larger cases keep only the digest
            assert first_diff(stats["vcf"], exp["vcf"]) is None
        assert hashlib.sha256(stats["vcf"]).hexdigest() == case["sha256"]["vcf"]


@pytest.mark.parametrize("spec,npairs,bp", [("profile:MiSeq", 4000, 0), ("models/stress20.npz", 20000, 257),
                                           ("models/stress20_sparse.npz", 20000, 0), ("models/miseq_150.npz", 8000, 1000)])
def test_store_mutations_vcf_against_oracle(spec, npairs, bp):
    """--store_mutations: the VCF lines (iss/generator.py:598-620) in the reference's order, incl. indel records."""
    recs_spec = [dict(id="vcf_a", synth_seed=17, length=60000), dict(id="vcf_b", synth_seed=18, length=900)]
    if "sparse" in spec:
        recs_spec[0]["sprinkle"] = dict(frac_lower=0.2, n_runs=20, iupac="RYKMSWBDHVN")
    case = dict(model=spec, seed=99, cpu=2, store_mutations=True, records=recs_spec, work=[[0, npairs], [1, 200]])
    r1, r2, stats = run_product(case, batch_pairs=bp)
    om = oracle.OracleModel(fixtures.model_path(spec), store_mutations=True)
    recs = fixtures.materialise_records(case)
    work = [(recs[ri][0], recs[ri][1].encode(), n) for ri, n in case["work"]]
    o1, o2, ov = oracle.run_worker(om, work, case["cpu"], case["seed"], vcf=True)
    assert first_diff(r1, o1) is None and first_diff(r2, o2) is None
    assert first_diff(stats["vcf"], ov) is None
    assert len(ov) > 1000


@pytest.mark.parametrize("name", ["novaseq_single", "hiseq_multi_short", "miseq_legacy_indels", "miseq150_bench_model",
                                  "amplicon_miseq", "nextseq_large_digest", "basic_mode"])
def test_tma_staged_gather_path(name):
    """opts.tma_gather: genome windows staged into shared memory by cp.async.bulk + mbarrier (2-bit records, L >= 96)."""
    case, exp = fixtures.load_case(name)
    r1, r2, _ = run_product(case, tma_gather=True, batch_pairs=97)
    assert hashlib.sha256(r1).hexdigest() == case["sha256"]["R1"]
    assert hashlib.sha256(r2).hexdigest() == case["sha256"]["R2"]


@pytest.mark.parametrize("batch_pairs", [1, 7, 64, 100])
@pytest.mark.parametrize("name", ["hiseq_multi_short", "stress_indels_dense"])
def test_batch_size_invariance(name, batch_pairs):
    """Small batches run concurrently on the ring's streams: results must not depend on the batch size."""
    case, exp = fixtures.load_case(name)
    r1, r2, _ = run_product(case, batch_pairs=batch_pairs)
    assert r1 == exp["R1"] and r2 == exp["R2"]


@pytest.mark.parametrize("spec,npairs", [("models/miseq_150.npz", 20000), ("profile:MiSeq", 6000),
                                        ("profile:HiSeq", 20000), ("models/stress20.npz", 30000),
                                        ("models/ties20.npz", 120000),     # thresholds colliding in their top 16 bits
                                        ("models/novaseq_x6.npz", 1500),   # L = 906: 32-pair tiles
                                        ("models/flat60_x6.npz", 1500),    # 60 distinct prefixes x 58 k draws per tile: tie queue overflows
                                        ("models/lowq20.npz", 30000)])     # phred 2-4: substitution list overflows, exact pass
def test_against_oracle_at_scale(spec, npairs):
    """Sizes the C oracle finishes in seconds; byte-identical streams."""
    case = dict(model=spec, seed=20260924, cpu=5, records=[dict(id="scale_a", synth_seed=7, length=150000),
                                                             dict(id="scale_b", synth_seed=8, length=40000)],
                work=[[0, npairs], [1, npairs // 3]])
    r1, r2, stats = run_product(case)
    om = oracle.OracleModel(fixtures.model_path(spec))
    recs = fixtures.materialise_records(case)
    work = [(recs[ri][0], recs[ri][1].encode(), n) for ri, n in case["work"]]
    o1, o2 = oracle.run_worker(om, work, case["cpu"], case["seed"])
    assert first_diff(r1, o1) is None
    assert first_diff(r2, o2) is None
    if spec in ("profile:MiSeq", "models/stress20.npz"):
        assert stats["slow_reads"] > 0  # the sequential indel replay really ran


@pytest.mark.parametrize("name", ["miseq150_bench_model", "stress_indels_dense", "hiseq_multi_short", "quality_ties_16bit",
                                  "nextseq_large_digest", "read_length_2"])
@pytest.mark.parametrize("tile", ["64", "128"])
def test_tile_geometries(name, tile, monkeypatch):
    """The default for short KDE reads is 128-pair tiles with two Philox blocks per quality-row stage (the <512, 2, 2>
    build); ISS_TILE=64 forces 64-pair tiles (256-thread CTAs, the <256, 4, 1> build), ISS_TILE=128 the 128-pair tiles with
    one block per stage (<512, 2, 1>, four warps per position stream)."""
    monkeypatch.setenv("ISS_TILE", tile)
    case, exp = fixtures.load_case(name)
    r1, r2, _ = run_product(case, batch_pairs=300)
    assert hashlib.sha256(r1).hexdigest() == case["sha256"]["R1"]
    assert hashlib.sha256(r2).hexdigest() == case["sha256"]["R2"]


@pytest.mark.parametrize("rows", ["short", "64"])
@pytest.mark.parametrize("name", ["novaseq_single", "gc_bias_hiseq", "hiseq_multi_short", "nextseq_large_digest", "fragment_length_novaseq",
                                  "ecoli_kde_vcf"])
def test_short_quality_rows(name, rows, monkeypatch):
    """Models whose quality rows have at most 8 distinct 16-bit thresholds (binned qualities: NovaSeq, NextSeq, HiSeq) are
    searched on 8-entry rows + a count byte per entry (three steps instead of six; the <512, 2, *, 8> builds); ISS_QROWS=64 keeps such a model on the full rows.  All must give the golden bytes -- ties and draws
    with a zero top half, which go to the exact pass, included."""
    if rows == "64":
        monkeypatch.setenv("ISS_QROWS", "64")
    case, exp = fixtures.load_case(name)
    r1, r2, _ = run_product(case, batch_pairs=700)
    assert hashlib.sha256(r1).hexdigest() == case["sha256"]["R1"]
    assert hashlib.sha256(r2).hexdigest() == case["sha256"]["R2"]


def test_native_file_writer(tmp_path):
    """Engine.write_files (iss_ctx_write_files): the FASTQ / VCF bytes land in files through the library's pwrite threads,
    at given offsets of a shared final file (two workers, plain text) or appended (gzip members)."""
    import gzip

    from insilicoseq_b200 import engine as eng_mod
    from insilicoseq_b200 import error_models

    case, exp = fixtures.load_case("ecoli_kde_vcf")
    recs = fixtures.materialise_records(case)
    with eng_mod.Engine(0) as e:
        e.set_genome([(n, s.encode()) for n, s in recs])
        e.set_model(error_models.KDErrorModel(fixtures.model_path(case["model"]), store_mutations=True))
        work = case["work"]
        items0 = eng_mod.worker_items(work, case["cpu"], case["seed"], e.record_lengths, e.read_length)
        items1 = eng_mod.worker_items(work, case["cpu"] + 1, case["seed"], e.record_lengths, e.read_length)
        n0, n1 = e.fastq_bytes(items0), e.fastq_bytes(items1)
        assert n0 == len(exp["R1"])
        f1, f2, v = str(tmp_path / "final_R1.fastq"), str(tmp_path / "final_R2.fastq"), str(tmp_path / "w.vcf")
        for f in (f1, f2):
            with open(f, "wb") as fh:
                fh.truncate(n0 + n1)
        st1 = e.write_files(items1, f1, f2, None, offset_r1=n0, offset_r2=n0, batch_pairs=17)  # second worker first
        st0 = e.write_files(items0, f1, f2, v, offset_r1=0, offset_r2=0, batch_pairs=23)
        assert st0["bytes_r1"] == n0 and st1["bytes_r1"] == n1 and st0["n_pairs"] == sum(n for _, n in work)
        got1, got2 = open(f1, "rb").read(), open(f2, "rb").read()
        assert got1[:n0] == exp["R1"] and got2[:n0] == exp["R2"]
        again = b"".join(b.r1.tobytes() for b in e.generate(items1))
        assert got1[n0:] == again
        assert open(v, "rb").read() == exp["vcf"]
        # appended gzip members
        z1, z2 = str(tmp_path / "a_R1.fastq.gz"), str(tmp_path / "a_R2.fastq.gz")
        e.write_files(items0, z1, z2, None, append=True, truncate=True, compress=True, batch_pairs=31)
        e.write_files(items1, z1, z2, None, append=True, compress=True, batch_pairs=31)
        assert gzip.decompress(open(z1, "rb").read()) == got1 and gzip.decompress(open(z2, "rb").read()) == got2
        # not a regular file: the write(2) path; and the same through ISS_WRITER=pwrite on a regular file
        st = e.write_files(items0, "/dev/null", "/dev/null", None)
        assert st["bytes_r1"] == n0
        os.environ["ISS_WRITER"] = "pwrite"
        try:
            p1, p2 = str(tmp_path / "p_R1.fastq"), str(tmp_path / "p_R2.fastq")
            e.write_files(items0, p1, p2, None, truncate=True, batch_pairs=19)
            assert open(p1, "rb").read() == exp["R1"] and open(p2, "rb").read() == exp["R2"]
        finally:
            del os.environ["ISS_WRITER"]
        # a path that cannot be opened: ISS_E_IO, and the engine stays usable
        from insilicoseq_b200 import _lib

        with pytest.raises(_lib.NativeError) as err:
            e.write_files(items0, str(tmp_path / "no_such_dir" / "x"), f2)
        assert err.value.code == _lib.ISS_E_IO
        assert b"".join(b.r1.tobytes() for b in e.generate(items0)) == exp["R1"]


def test_abort_leaves_a_usable_context():
    """iss_ctx_abort drops a half-consumed submit; the next submit starts from a clean ring."""
    from insilicoseq_b200 import engine as eng_mod
    from insilicoseq_b200 import error_models

    case, exp = fixtures.load_case("novaseq_single")
    recs = fixtures.materialise_records(case)
    with eng_mod.Engine(0) as e:
        e.set_genome([(n, s.encode()) for n, s in recs])
        e.set_model(error_models.KDErrorModel(fixtures.model_path(case["model"])))
        items = eng_mod.worker_items(case["work"], case["cpu"], case["seed"], e.record_lengths, e.read_length)
        e.submit(items, batch_pairs=5)
        e.abort()
        assert b"".join(b.r1.tobytes() for b in e.generate(items, batch_pairs=7)) == exp["R1"]


def test_error_conventions_at_the_c_abi():
    """ISS_E_CHAR for characters outside util.rev_comp's dictionary (the reference raises KeyError mid-run),
    ISS_E_SHORT for a record not longer than the read length (reference: AssertionError, generator.py:130),
    ISS_E_STATE for calls out of order."""
    from insilicoseq_b200 import _lib, error_models
    from insilicoseq_b200.engine import Engine, philox_key

    with Engine(0) as e:
        with pytest.raises(_lib.NativeError) as err:
            e.set_genome([("bad", b"ACGTACGTXACGT" * 20)])
        assert err.value.code == _lib.ISS_E_CHAR and "0x58" in str(err.value)
        with pytest.raises(_lib.NativeError) as err:
            e.submit([(0, 0, 10, 0, 0, 1)])
        assert err.value.code == _lib.ISS_E_STATE
        e.set_genome([("ok", fixtures.synth_genome(1, 500)), ("short", fixtures.synth_genome(2, 126)), ("u", b"ACGU" * 100)][:2])
        e.set_model(error_models.KDErrorModel(fixtures.model_path("profile:HiSeq")))
        with pytest.raises(_lib.NativeError) as err:
            e.submit([(1, 0, 10, 0, 0, philox_key(1, 0))])  # len == L: `assert read_length < len(record.seq)` fails
        assert err.value.code == _lib.ISS_E_SHORT
        with pytest.raises(RuntimeError):
            e.submit([(0, 0, 10, 0, 0, 1)], sequence_type="nanopore")
        # the engine is still usable after the rejected calls
        out = b"".join(b.r1.tobytes() for b in e.generate([(0, 0, 5, 0, 0, philox_key(1, 0))]))
        assert out.count(b"\n") == 20


def test_uracil_is_rejected_like_the_reference_keyerror():
    from insilicoseq_b200 import _lib
    from insilicoseq_b200.engine import Engine

    with Engine(0) as e:
        with pytest.raises(_lib.NativeError) as err:
            e.set_genome([("rna", b"ACGU" * 100)])
        assert err.value.code == _lib.ISS_E_CHAR


def test_forced_indels_like_the_reference_tests():
    """iss/test/test_error_model.py:59-105 force events with probability 1.0 in the model's dicts (ins at position 1,
    del at positions 0/1) and check the list semantics of introduce_indels / adjust_seq_length.  Same forcing here,
    through the plugin surface (dict attributes of the model), against the oracle; every read takes the replay path."""
    from insilicoseq_b200 import error_models
    from insilicoseq_b200 import engine as eng_mod

    m = error_models.KDErrorModel(fixtures.model_path("models/ecoli.npz"))
    m.ins_for[1]["G"] = 1.0
    m.del_for[0]["A"] = 1.0
    m.del_rev[0]["C"] = 1.0
    m.del_rev[1]["G"] = 1.0
    om = oracle.OracleModel(fixtures.model_path("models/ecoli.npz"))
    ins_p, del_p = om.table("ins_p").copy(), om.table("del_p").copy()
    ins_p[0, 1, list(m.ins_for[1].keys()).index("G")] = 1.0
    del_p[0, 0, "ACGT".index("A")] = 1.0
    del_p[1, 0, "ACGT".index("C")] = 1.0
    del_p[1, 1, "ACGT".index("G")] = 1.0
    om.set_table("ins_p", ins_p)
    om.set_table("del_p", del_p)
    genome = ("GG" + "GTACC" * 100 + "GG" + "ATATA" * 100).encode()
    with eng_mod.Engine(0) as e:
        e.set_genome([("ref_genome", genome)])
        e.set_model(m)
        items = eng_mod.worker_items([(0, 3000)], 0, 87, e.record_lengths, e.read_length)
        r1, r2 = b"", b""
        for b in e.generate(items):
            r1 += b.r1.tobytes()
            r2 += b.r2.tobytes()
        st = e.stats()
    o1, o2 = oracle.run_worker(om, [("ref_genome", genome, 3000)], 0, 87)
    assert first_diff(r1, o1) is None and first_diff(r2, o2) is None
    assert st["slow_reads"] >= 3000  # every forward read has the forced insertion
    seqs = r1.split(b"\n")[1::4]
    assert all(len(s) == 20 for s in seqs)


def test_work_item_split_across_submits():
    """An item may be generated in pieces (i0 > 0) -- by another submit or another GPU: draws are counter addressed,
    so the pieces concatenate to the whole."""
    from insilicoseq_b200 import error_models
    from insilicoseq_b200.engine import Engine, philox_key

    genome = fixtures.synth_genome(77, 40000).encode()
    key = philox_key(9, 4)
    with Engine(0) as e:
        e.set_genome([("split_rec", genome)])
        e.set_model(error_models.KDErrorModel(fixtures.model_path("profile:HiSeq")))
        whole = b"".join(b.r2.tobytes() for b in e.generate([(0, 4, 5000, 0, 123, key)]))
        parts = b""
        for i0, n in ((0, 1), (1, 998), (999, 1), (1000, 3000), (4000, 1000)):
            parts += b"".join(b.r2.tobytes() for b in e.generate([(0, 4, n, i0, 123, key)], batch_pairs=700))
    assert parts == whole
    om = oracle.OracleModel(fixtures.model_path("profile:HiSeq"))
    _, o2 = oracle.generate(om, genome, "split_rec", key, 123, 4, 0, 5000)
    assert whole == o2


def test_gpu_gzip_roundtrip_arbitrary_bytes():
    """The GPU gzip is checked by decompression with Python's gzip: FASTQ text, empty input, a single
    symbol, all byte values, incompressible data (stored-block fallback), a skewed distribution that needs the 15-bit
    length limit, sizes around the 64 KB member boundary, and a long non-uniform stream."""
    import gzip

    from insilicoseq_b200.engine import Engine

    case, exp = fixtures.load_case("nextseq_large_digest")
    rng = np.random.RandomState(11)
    fib = [1, 2]
    while len(fib) < 21:
        fib.append(fib[-1] + fib[-2])
    skew = b"".join(bytes([65 + i]) * f for i, f in enumerate(fib))
    samples = [b"", b"A", b"AAAA" * 5000, bytes(range(256)) * 300, rng.bytes(200_000), skew, skew[:65536] + skew,
               b"@r_1/1\nACGT\n+\nIIII\n" * 10000, rng.bytes(65535), rng.bytes(65536), rng.bytes(65537), b"x" * 65536, b"y" * 65537]
    fq = open(os.path.join(fixtures.GOLDEN, "miseq_legacy_indels", "R1.fastq"), "rb").read() * 40
    samples.append(fq)
    # the code of a call comes from a strided sample (> 4 MB: not every 16 KB piece is looked at); a stream that changes
    # character half way must still round-trip, whatever the sample saw
    samples.append(b"ACGT" * 1_500_000 + rng.bytes(3_000_000) + b"\x00\xff" * 70_001)
    with Engine(0) as e:
        for data in samples:
            z = e.gzip(data)
            assert gzip.decompress(z) == data, len(data)
        ratio = len(fq) / len(e.gzip(fq))
    assert ratio > 2.0, ratio  # Huffman-only: ~2.4x on L=301 MiSeq text, ~3x on 150 bp reads (gzip -6: ~3.8x)


def test_gpu_gzip_name_line_matches():
    """The LZ77 part of the GPU gzip: one match per FASTQ record, its name line against the previous record's.  Cases that
    stress the parallel planning: reads so short that a 128-byte slice holds several records, names longer than a slice
    and longer than the 258-byte match limit, quality lines that begin with '@', records 40 KB apart (beyond the 32 KB
    window), "\n@" patterns in arbitrary bytes, and streams whose members end inside a name line."""
    import gzip

    from insilicoseq_b200.engine import Engine

    rng = np.random.RandomState(5)

    def fastq(n, L, name, qual0=None):
        out = []
        for i in range(n):
            seq = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), L))
            q = bytes(rng.randint(33, 75, L).astype(np.uint8))
            if qual0 is not None:
                q = qual0 + q[1:]
            out.append(b"@%s_%d_0/1\n%s\n+\n%s\n" % (name, i, seq, q))
        return b"".join(out)

    samples = [fastq(30000, 8, b"s"), fastq(20000, 20, b"contig_7"), fastq(3000, 150, b"x" * 140), fastq(1500, 150, b"n" * 300),
               fastq(4000, 150, b"syn5M", qual0=b"@"), fastq(40, 20000, b"far_apart"), fastq(5000, 151, b"NZ_CP0123.1|plasmid"),
               b"\n@".join(rng.bytes(200) for _ in range(3000)), (b"@a\nAC\n+\n@@\n" * 30000)[3:]]
    with Engine(0) as e:
        for data in samples:
            z = e.gzip(data)
            assert gzip.decompress(z) == data, len(data)
        # the benchmark's text: MiSeq model at 150 positions.  Literals alone reach 2.82 on it; the name match must show
        from insilicoseq_b200 import engine as eng_mod
        from insilicoseq_b200 import error_models

        e.set_genome([("syn5M", fixtures.synth_genome(3, 400000).encode())])
        e.set_model(error_models.KDErrorModel(fixtures.model_path("models/miseq_150.npz")))
        items = eng_mod.worker_items([(0, 30000)], 0, 42, e.record_lengths, e.read_length)
        plain = b"".join(b.r1.tobytes() for b in e.generate(items))
        z = b"".join(b.r1.tobytes() for b in e.generate(items, compress=True))
        assert gzip.decompress(z) == plain
        assert len(plain) / len(z) > 2.95, len(plain) / len(z)


def test_genome_file_as_text_matches_the_record_upload(tmp_path):
    """Engine.load_fasta (iss_fasta_index + iss_ctx_set_genome_fasta: title lines found on the host, white space squeezed
    out and bases packed on the GPU) against fasta.parse + register_records: same ids, lengths and -- the proof that the
    device blob is the same -- the same reads.  The file has a preamble, wrapped and unwrapped lines, CRLF line ends, empty
    lines, blanks inside lines, a lower-case / IUPAC record (kept as bytes on the device), records far smaller and larger
    than the 4 KB blocks of the device scan, and a last line without a line end."""
    from insilicoseq_b200 import engine as eng_mod
    from insilicoseq_b200 import error_models, fasta

    rng = np.random.RandomState(23)

    def bases(n, alphabet=b"ACGT"):
        return bytes(rng.choice(np.frombuffer(alphabet, dtype=np.uint8), n))

    def wrap(seq, width, eol=b"\n"):
        return eol.join(seq[i:i + width] for i in range(0, len(seq), width)) + eol

    parts = [b"# a comment in front of the first title\n\n",
             b">chr1 the first one\n" + wrap(bases(70001), 60),
             b">small\n" + wrap(bases(400), 80),
             b">crlf_rec windows line ends\r\n" + wrap(bases(9000), 70, b"\r\n"),
             b">iupac lower case and N\n" + wrap(bases(5000, b"ACGTNacgtRYKM"), 61),
             b">gaps\n" + wrap(bases(3000), 50).replace(b"\n", b" \n\n", 7),
             b">oneline\n" + bases(20000) + b"\n",
             b">tiny1\n" + bases(350) + b"\n>tiny2\n" + bases(333) + b"\n>tiny3\n" + bases(360) + b"\n",
             b">last no line end\n" + bases(1234)]
    path = tmp_path / "genome.fasta"
    path.write_bytes(b"".join(parts))
    model = error_models.KDErrorModel(fixtures.model_path("models/miseq_150.npz"))
    with eng_mod.Engine(0) as a, eng_mod.Engine(0) as b:
        recs_a = a.register_records(list(fasta.parse(str(path))))
        recs_b = b.load_fasta(str(path))
        assert [r.id for r in recs_b] == [r.id for r in recs_a]
        assert [r.description for r in recs_b] == [r.description for r in recs_a]
        assert [len(r) for r in recs_b] == [len(r) for r in recs_a] == b.record_lengths
        assert all(bytes(rb.seq) == ra.seq for ra, rb in zip(recs_a, recs_b))
        outs = []
        for e, recs in ((a, recs_a), (b, recs_b)):
            e.set_model(model)
            work = [(e.record_index(r), 900) for r in recs]
            items = eng_mod.worker_items(work, 1, 77, e.record_lengths, e.read_length)
            outs.append([(bt.r1.tobytes(), bt.r2.tobytes()) for bt in e.generate(items, batch_pairs=2000)])
        assert outs[0] == outs[1]
        assert sum(len(x) for x, _ in outs[0]) > 9 * 900 * 300
        # an empty file, and a file without a title line: no records
        empty = tmp_path / "empty.fasta"
        empty.write_bytes(b"")
        assert b.load_fasta(str(empty)) == []
        empty.write_bytes(b"ACGT\nACGT\n")
        assert b.load_fasta(str(empty)) == []
        # a character outside the reference's alphabet is reported with its record and position
        bad = tmp_path / "bad.fasta"
        bad.write_bytes(b">ok\nACGT\n>bad\nAC\nG!T\n")
        with pytest.raises(RuntimeError, match="record 1: character 0x21 at position 3"):
            b.load_fasta(str(bad))


def test_gpu_gzip_distance_one_runs():
    """The second kind of LZ77 token of the GPU gzip: runs (>= 4 bytes equal to the byte in front of them, distance 1), which
    is most of a binned-quality line.  Runs of every length around the thresholds (3, 4, 5, 63..66, 127..130, 258+), runs
    that cross the 64-byte half, the 128-byte slice and the 64 KB member, runs that touch a name line on either side
    ("\n\n\n\n\n@", a quality line ending in the run byte), a member that is one run, and the NovaSeq fixture, whose
    quality lines must now cost well under a bit per base."""
    import gzip

    from insilicoseq_b200.engine import Engine

    rng = np.random.RandomState(17)

    def binned_fastq(n, L, name):
        out = []
        for i in range(n):
            seq = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), L))
            q = np.full(L, ord("F"), dtype=np.uint8)
            q[rng.rand(L) < 0.06] = ord(":")
            q[rng.rand(L) < 0.02] = ord(",")
            out.append(b"@%s_%d_0/1\n%s\n+\n%s\n" % (name, i, seq, q.tobytes()))
        return b"".join(out)

    ladders = b"".join(bytes([97 + (k % 7)]) * k + b"|" for k in list(range(1, 12)) + [62, 63, 64, 65, 66, 126, 127, 128, 129, 130, 257, 258, 259, 600])
    samples = [binned_fastq(4000, 151, b"nova"), binned_fastq(20000, 12, b"s"), ladders * 40, b"\0" * 200_001,
               (b"ab" * 40 + b"c" * 50) * 3000, b"z" * 65535 + b"y" * 3 + b"z" * 70000, (b"\n\n\n\n\n\n@q_1/1\nAC\n+\nFFFFFFFFFF\n") * 9000,
               bytes(rng.choice(np.frombuffer(b"AAAAAAAB", dtype=np.uint8), 300_000)), b"Q" * 65536 * 3,
               (b"@r_%d/1\n" % 7) + b"F" * 500 + b"\n"]
    with Engine(0) as e:
        for data in samples:
            z = e.gzip(data)
            assert gzip.decompress(z) == data, len(data)
        assert len(samples[0]) / len(e.gzip(samples[0])) > 3.4  # 2.9 with literals and name lines only
        fq = open(os.path.join(fixtures.GOLDEN, "novaseq_single", "R1.fastq"), "rb").read() * 30
        z = e.gzip(fq)
        assert gzip.decompress(z) == fq
        assert len(fq) / len(z) > 3.5, len(fq) / len(z)  # 3.15 without the runs (host prototype: 3.94 with them)


def test_compressed_batches_decompress_to_the_plain_output():
    import gzip

    case, exp = fixtures.load_case("hiseq_multi_short")
    from insilicoseq_b200 import engine as eng_mod

    recs = fixtures.materialise_records(case)
    with eng_mod.Engine(0) as e:
        e.set_genome([(n, s.encode()) for n, s in recs])
        e.set_model(product_model(case))
        items = eng_mod.worker_items(case["work"], case["cpu"], case["seed"], e.record_lengths, e.read_length)
        for bp in (0, 13):
            z1 = b"".join(b.r1.tobytes() for b in e.generate(items, compress=True, batch_pairs=bp))
            z2 = b"".join(b.r2.tobytes() for b in e.generate(items, compress=True, batch_pairs=bp))
            assert gzip.decompress(z1) == exp["R1"] and gzip.decompress(z2) == exp["R2"]
