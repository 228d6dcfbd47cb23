"""Parity of the CUDA path (through the C ABI) with the oracle and the reference-written golden fixtures.
Bit-exact: this is byte/integer work, the bar is identical FASTQ bytes."""
import hashlib

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


@pytest.mark.parametrize("batch_pairs", [1, 7, 64, 100])
@pytest.mark.parametrize("name", ["hiseq_multi_short", "stress_indels_dense"])
def test_batch_size_invariance(name, batch_pairs):
    """Small batches run concurrently on the ring's streams: results must not depend on the batch size."""
    case, exp = fixtures.load_case(name)
    r1, r2, _ = run_product(case, batch_pairs=batch_pairs)
    assert r1 == exp["R1"] and r2 == exp["R2"]


@pytest.mark.parametrize("spec,npairs", [("models/miseq_150.npz", 20000), ("profile:MiSeq", 6000),
                                        ("profile:HiSeq", 20000), ("models/stress20.npz", 30000)])
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
