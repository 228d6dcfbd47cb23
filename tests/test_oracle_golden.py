"""CPU tests of the test infrastructure itself: the C oracle (oracle/iss_oracle.c) against the golden
fixtures written by the Philox-driven UNMODIFIED reference (oracle/gen_golden.py), Philox known-answer
vectors, and -- when /root/reference is present -- a live re-run of the reference for one case."""
import hashlib
import os

import numpy as np
import pytest

import fixtures
import oracle
import philox

CASES = fixtures.list_cases()


def _run_oracle(case):
    recs = fixtures.materialise_records(case)
    fl, fsd = case.get("fragment_length"), case.get("fragment_sd")
    ft = None
    if fl is not None:
        import patched_reference as pr

        ft = pr.fragment_table(fl, fsd)
    if case["model"] in ("basic", "perfect"):
        import patched_reference as pr

        om = oracle.OracleModel(case["model"], fl, fsd, bool(case.get("store_mutations")), frag_table=ft,
                                basic_table=pr.basic_phred_table(30))
    else:
        om = oracle.OracleModel(fixtures.model_path(case["model"]), fl, fsd, bool(case.get("store_mutations")),
                                frag_table=ft)
    work = [(recs[ri][0], recs[ri][1].encode(), n) for ri, n in case["work"]]
    return oracle.run_worker(om, work, case["cpu"], case["seed"], amplicon=case.get("sequence_type") == "amplicon",
                             gc_bias=bool(case.get("gc_bias")), vcf=True)


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    kats = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, want in kats:
        assert philox.philox4x32_10(ctr, key) == want
        assert oracle.philox(ctr, key) == want
        got = philox.philox4x32_10_np(*[[c] for c in ctr], *[[k] for k in key])
        assert tuple(int(g[0]) for g in got) == want


def test_bundled_profiles_are_the_reference_files():
    ref = "/root/reference/iss/profiles"
    if not os.path.isdir(ref):
        pytest.skip("reference not mounted")
    for f in os.listdir(ref):
        a = open(os.path.join(ref, f), "rb").read()
        b = open(os.path.join(fixtures.PROFILES, f), "rb").read()
        assert a == b, f


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_matches_reference_fixture(name):
    case, exp = fixtures.load_case(name)
    r1, r2, vcf = _run_oracle(case)
    assert hashlib.sha256(r1).hexdigest() == case["sha256"]["R1"]
    assert hashlib.sha256(r2).hexdigest() == case["sha256"]["R2"]
    if case.get("store_mutations"):
        assert hashlib.sha256(vcf).hexdigest() == case["sha256"]["vcf"]
    if exp["R1"] is not None:
        assert r1 == exp["R1"] and r2 == exp["R2"]
    if exp["vcf"] is not None:
        assert vcf == exp["vcf"]


def test_fastq_bytes_closed_form():
    for hdr, cpu, L, i0, i1 in [(5, 0, 150, 0, 1234), (11, 17, 301, 95, 10050), (1, 123456, 20, 999990, 1000020)]:
        want = sum(len("@%s_%d_%d/1\n%s\n+\n%s\n" % ("h" * hdr, i, cpu, "A" * L, "I" * L)) for i in range(i0, i1))
        assert oracle.fastq_bytes(hdr, cpu, L, i0, i1) == want


def test_short_record_is_assertion_error():
    om = oracle.OracleModel(fixtures.model_path("models/ecoli.npz"))
    with pytest.raises(oracle.OracleError) as e:
        oracle.generate(om, b"AAAAACCCCC", "my_genome", 1, 0, 0, 1, 2)  # iss/test/test_generator.py:62-67
    assert e.value.code == oracle.E_SHORT


@pytest.mark.reference
def test_live_reference_agrees_with_fixture():
    import patched_reference as pr

    if not pr.reference_available():
        pytest.skip("reference not mounted")
    import gen_golden
    import tempfile

    case, exp = fixtures.load_case("stress_indels_sparse_iupac")
    with tempfile.TemporaryDirectory() as tmp:
        r1, r2, vcf = gen_golden.run_reference(case, tmp)
    assert r1 == exp["R1"] and r2 == exp["R2"]


@pytest.mark.reference
@pytest.mark.parametrize("seed", range(8))
def test_c_oracle_matches_live_reference_on_random_models(seed, tmp_path):
    """Fuzz pin of the C restatement: random read length, random indel tables (sparse to full, p = 1.0 entries), random
    genome with lower case / IUPAC runs and a record too short for the model; the UNMODIFIED reference under the shared
    Philox stream and oracle/iss_oracle.c must write the same FASTQ and VCF bytes."""
    import patched_reference as pr

    if not pr.reference_available():
        pytest.skip("reference not mounted")
    import gen_golden

    rng = np.random.RandomState(1000 + seed)
    L = int(rng.randint(2, 21))
    stress = str(tmp_path / "stress.npz")
    gen_golden.write_stress_model(stress, fixtures.model_path("models/ecoli.npz"), 50 + seed, dense=float(rng.choice([0.05, 0.3, 0.7, 1.0])))
    model = str(tmp_path / "m.npz")
    oracle.truncate_profile(stress, model, L)
    case = dict(name="fuzz%d" % seed, model=model, seed=int(rng.randint(0, 1 << 30)), cpu=int(rng.randint(0, 300)),
                store_mutations=bool(seed & 1), gc_bias=bool(seed == 5), sequence_type="amplicon" if seed == 6 else "metagenomics",
                records=[{"id": "fz%d" % seed, "synth_seed": 900 + seed, "length": int(rng.randint(max(3 * L + 5, 60), 600)),
                          "sprinkle": {"frac_lower": 0.2, "n_runs": 2, "iupac": "RYSWN"}},
                         {"id": "short", "synth_seed": 950 + seed, "length": max(1, L - 1)},
                         {"id": "plain", "synth_seed": 970 + seed, "length": 4 * L + 11}],
                work=[[0, 70], [1, 3], [2, 40]])
    r1, r2, vcf = gen_golden.run_reference(case, str(tmp_path))
    c1, c2, cv = gen_golden.run_c_oracle(case)
    assert len(r1) > 0 and c1 == r1 and c2 == r2
    if case["store_mutations"]:
        assert cv == vcf
