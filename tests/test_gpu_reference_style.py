"""Tests written the way the reference's own tests are (iss/test/test_generator.py), against the same entry points:
generator.simulate_read / simulate_reads on a record + error model.  The reference pins literal strings under
MT19937; here the pin is the oracle under the shared Philox stream."""
import io

import pytest

import fixtures
import oracle
from insilicoseq_b200 import generator
from insilicoseq_b200.error_models import BasicErrorModel, KDErrorModel
from insilicoseq_b200.fasta import Record

pytestmark = pytest.mark.gpu
ECOLI = fixtures.model_path("models/ecoli.npz")


def test_small_input():  # iss/test/test_generator.py:62-67
    with pytest.raises(AssertionError):
        err_mod = KDErrorModel(ECOLI)
        ref_genome = Record("AAAAACCCCC", id="my_genome", description="test genome")
        generator.simulate_read(ref_genome, err_mod, 1, 0, "metagenomics")


def test_kde():  # iss/test/test_generator.py:78-83
    err_mod = KDErrorModel(ECOLI)
    ref_genome = Record("CGTTTCAACC" * 400, id="my_genome", description="test genome")
    fwd, rev, _ = generator.simulate_read(ref_genome, err_mod, 1, 0, "metagenomics", seed=42)
    om = oracle.OracleModel(ECOLI)
    o1, o2 = oracle.generate(om, ref_genome.seq, "my_genome", oracle.make_key(42, 0), 0, 0, 1, 2)
    assert fwd.format() == o1.decode() and rev.format() == o2.decode()
    assert fwd.id == "my_genome_1_0/1" and rev.id == "my_genome_1_0/2"
    assert len(fwd.seq) == 20 and len(fwd.letter_annotations["phred_quality"]) == 20


def test_kde_short():  # iss/test/test_generator.py:86-91: genome shorter than the template -> the fallback branches
    from insilicoseq_b200.error_models import fragment_table

    err_mod = KDErrorModel(ECOLI, 1000, 10)
    ref_genome = Record("AAACC" * 100, id="my_genome", description="test genome")
    fwd, rev, _ = generator.simulate_read(ref_genome, err_mod, 1, 0, "metagenomics", seed=42)
    om = oracle.OracleModel(ECOLI, 1000, 10, frag_table=fragment_table(1000, 10))
    o1, o2 = oracle.generate(om, ref_genome.seq, "my_genome", oracle.make_key(42, 0), 0, 0, 1, 2)
    assert fwd.format() == o1.decode() and rev.format() == o2.decode()


def test_amp():  # iss/test/test_generator.py:94-116
    err_mod = KDErrorModel(fixtures.model_path("profile:MiSeq"), 1000, 10)
    seq = "TTT" + "A" * 194 + "GGCC" + "A" * 196 + "TTT"
    ref_genome = Record(seq, id="my_amplicon", description="test amplicon")
    fwd, rev, _ = generator.simulate_read(ref_genome, err_mod, 1, 0, "amplicon", seed=42)
    from insilicoseq_b200.error_models import fragment_table

    assert len(fwd.seq) == 301 and len(rev.seq) == 301
    assert fwd.seq.startswith("TTTAAAAAAA") and rev.seq.startswith("AAATTTTTTT")  # reads start at both ends of the template
    om = oracle.OracleModel(fixtures.model_path("profile:MiSeq"), 1000, 10, frag_table=fragment_table(1000, 10))
    o1, o2 = oracle.generate(om, ref_genome.seq, "my_amplicon", oracle.make_key(42, 0), 0, 0, 1, 2, amplicon=True)
    assert fwd.format() == o1.decode() and rev.format() == o2.decode()


def test_unknown_sequence_type():
    err_mod = KDErrorModel(ECOLI)
    ref_genome = Record("CGTTTCAACC" * 400, id="my_genome")
    with pytest.raises(RuntimeError):
        generator.simulate_read(ref_genome, err_mod, 1, 0, "nanopore")


def test_simulate_and_save():  # iss/test/test_generator.py:40-49 (basic model, gc_bias)
    err_mod = BasicErrorModel(451, 0)
    ref_genome = Record("AAAAACCCCC" * 100, id="my_genome", description="test genome")
    f, r, m = io.BytesIO(), io.StringIO(), io.BytesIO()
    generator.simulate_reads(ref_genome, err_mod, 1000, 0, f, r, m, "metagenomics", gc_bias=True, seed=3)
    assert f.getvalue().count(b"\n") == 4000 and r.getvalue().count("\n") == 4000
    assert f.getvalue().startswith(b"@my_genome_0_0/1\n")


def test_simulate_and_save_short():  # iss/test/test_generator.py:51-59: record shorter than the template
    err_mod = BasicErrorModel()
    ref_genome = Record("AACCC" * 100, id="my_genome", description="test genome")
    f, r, m = io.BytesIO(), io.BytesIO(), io.BytesIO()
    generator.simulate_reads(ref_genome, err_mod, 1000, 0, f, r, m, "metagenomics", gc_bias=True)
    assert f.getvalue().count(b"\n") == 4000
