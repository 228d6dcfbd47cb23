"""CPU tests of the host logic: exactness of the f64 -> integer-threshold lowering, the closed-form FASTQ
offsets, the work bookkeeping, and that the C-ABI library loads and exports every declared symbol."""
import os
import re

import numpy as np
import pytest

import fixtures
import oracle
from insilicoseq_b200 import _lib, error_models
from insilicoseq_b200 import engine as eng_mod

TWO32 = 2.0 ** 32


def _words_around(values):
    """32-bit words at and next to every threshold, plus random ones."""
    rng = np.random.RandomState(1)
    v = np.asarray(values, dtype=np.float64).ravel()
    near = np.concatenate([np.floor(v), np.floor(v) + 1, np.floor(v) - 1, np.ceil(v), np.ceil(v) + 1])
    near = near[(near >= 0) & (near < TWO32)]
    return np.unique(np.concatenate([near, rng.randint(0, 2 ** 32, size=2000, dtype=np.uint64).astype(np.float64),
                                     [0, 1, TWO32 - 1]])).astype(np.uint64)


@pytest.mark.parametrize("spec", ["profile:HiSeq", "models/miseq_150.npz", "models/stress20.npz"])
def test_threshold_lowering_is_exact(spec):
    path = fixtures.model_path(spec)
    m = error_models.KDErrorModel(path)
    t = m.device_tables().arrays
    om = oracle.OracleModel(path)  # the f64 tables exactly as the reference holds them
    L = int(m.read_length)
    # insert size: searchsorted(cdf, u) 'left'
    cdf = om.table("insert_cdf")
    for w in _words_around(cdf * TWO32)[::7]:
        u = float(w) * 2.0 ** -32
        assert int(np.searchsorted(cdf, u)) == int((np.uint64(w) > t["insert_thr"].astype(np.uint64)).sum())
    # quality rows (sampled) and the error test
    qc = om.table("qual_cdf")
    pp = om.table("phred_prob")
    rng = np.random.RandomState(2)
    for _ in range(60):
        o, b, pos = rng.randint(2), rng.randint(4), rng.randint(L)
        if not t["bin_valid"][o, b]:
            continue
        row = qc[o, b, pos]
        for w in _words_around(row * TWO32)[::11]:
            u = float(w) * 2.0 ** -32
            assert int(np.searchsorted(row, u)) == int((np.uint64(w) > t["qual_thr"][o, b, pos].astype(np.uint64)).sum())
    for q in range(om.nq + 1):
        for w in _words_around([pp[q] * TWO32]):
            assert (float(w) * 2.0 ** -32 > pp[q]) == (int(w) > int(t["err_thr"][q]))
    # bins and substitution choices: 'right'
    for o in range(2):
        cdfb = om.table("bin_cdf")[o]
        for w in _words_around(cdfb * TWO32):
            u = float(w) * 2.0 ** -32
            assert int(np.searchsorted(cdfb, u, side="right")) == int((np.uint64(w) >= t["bin_thr"][o]).sum())
    # indel Bernoulli tests
    ins_p = om.table("ins_p")
    for o in range(2):
        for pos in range(L):
            for lane in range(4):
                p = ins_p[o, pos, lane]
                nz = (t["ins_nz"][o, pos] >> lane) & 1
                for w in _words_around([p * TWO32])[:12]:
                    assert (float(w) * 2.0 ** -32 < p) == bool(nz and int(w) <= int(t["ins_thr"][o, pos, lane]))


def test_base_index_and_subst_alt_order():
    m = error_models.KDErrorModel(fixtures.model_path("profile:NovaSeq"))
    t = m.device_tables().arrays
    for c, base in enumerate(error_models.BASES):
        assert (ord(base) >> 1) & 3 == c and (ord(base.lower()) >> 1) & 3 == c
        alts, probs = m.subst_choices_for[10][base]
        assert [chr(x) for x in t["subst_alt"][0, 10, c, :3]] == list(alts)


def test_all_bundled_profiles_lower():
    for f in sorted(os.listdir(error_models.PROFILE_DIR)):
        m = error_models.KDErrorModel(os.path.join(error_models.PROFILE_DIR, f))
        t = m.device_tables()
        assert t.scalars["read_length"] == int(m.read_length)
        assert t.arrays["qual_thr"].shape == (2, 4, int(m.read_length), 64)


def test_fragment_table_matches_oracle_side():
    import patched_reference as pr

    for mu, sd in [(420, 35), (1000, 10), (450, 0)]:
        lo_a, t_a = error_models.fragment_table(mu, sd)
        lo_b, t_b = pr.fragment_table(mu, sd)
        assert lo_a == lo_b and np.array_equal(t_a, t_b)


def test_basic_and_perfect_lowering():
    import patched_reference as pr

    thr_a, val_a = error_models.basic_phred_table(30)
    thr_b, val_b = pr.basic_phred_table(30)
    assert np.array_equal(thr_a, thr_b) and np.array_equal(val_a, val_b)
    assert val_a[-1] == 40 and (np.diff(val_a) == 1).all() and (np.diff(thr_a.astype(np.int64)) >= 0).all()
    b = error_models.BasicErrorModel().device_tables()
    assert b.scalars["qual_mode"] == 2 and b.scalars["fixed_insert"] == 200 and b.scalars["read_length"] == 125
    assert not b.arrays["ins_nz"].any() and not b.arrays["del_nz"].any()
    pm = error_models.PerfectErrorModel().device_tables()
    assert pm.scalars["qual_mode"] == 1 and pm.scalars["const_phred"] == 40
    # identity 'substitution': choice over [1,0,0,0] etc. always returns the original base
    for c, base in enumerate(error_models.BASES):
        thr = pm.arrays["subst_thr"][0, 5, c]
        alts = pm.arrays["subst_alt"][0, 5, c]
        for w in (0, 1, 2 ** 31, 2 ** 32 - 1):
            assert chr(alts[int((np.uint64(w) >= thr).sum())]) == base


def test_worker_items_bookkeeping():
    items = eng_mod.worker_items([(0, 5), (1, 3), (2, 0), (0, 4)], cpu_number=3, seed=7,
                                 record_lengths=[1000, 100, 1000], read_length=126)
    # record 1 is too short: skipped, but the pair ordinal still advances (RNG contract v2)
    assert items == [(0, 3, 5, 0, 0, 10), (0, 3, 4, 0, 8, 10)]
    assert eng_mod.philox_key(-5, 2) == (2 ** 64 - 3)


def test_native_library_exports_every_declared_symbol():
    L = _lib.load()  # builds in-tree if needed; loading needs no GPU
    header = open(os.path.join(fixtures.ROOT, "include", "iss_b200.h")).read()
    declared = set(re.findall(r"\b(iss_[a-z0-9_]+)\s*\(", header))
    declared -= {"iss_ctx"}
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name)
    assert L.iss_abi_version() == 1
    for args in [(5, 0, 150, 0, 1234), (11, 17, 301, 95, 10050), (1, 123456, 20, 999990, 1000020)]:
        assert L.iss_fastq_bytes(*args) == oracle.fastq_bytes(*args)


def test_product_fails_loudly_without_gpu():
    import ctypes

    L = _lib.load()
    n = ctypes.c_int(0)
    rc = L.iss_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.NativeError):
        eng_mod.Engine(0)


def test_bad_model_files_exit_like_the_reference(tmp_path):
    """iss/test/test_error_model.py:108-110 (test_bad_err_mod): an unreadable profile logs and exits."""
    empty = tmp_path / "empty_file"
    empty.write_bytes(b"")
    with pytest.raises(SystemExit):
        error_models.KDErrorModel(str(empty))
    from insilicoseq_b200 import generator

    with pytest.raises(SystemExit):
        generator.load_error_model("kde", None, None, None, None, False)  # --model is required in --mode kde
    with pytest.raises(SystemExit):
        generator.load_error_model("kde", None, "hiseq", 400, None, False)  # fragment length without sd


def test_lowering_rejects_what_the_reference_could_not_sample():
    m = error_models.KDErrorModel(fixtures.model_path("models/ecoli.npz"))
    m.subst_choices_for[3]["A"] = (["T", "C", "G"], [0.5, 0.6, 0.2])  # np.random.choice: probabilities do not sum to 1
    with pytest.raises(ValueError):
        m.device_tables()
    m = error_models.KDErrorModel(fixtures.model_path("models/ecoli.npz"))
    m.ins_for[2]["G"] = 1.0       # forced events as in iss/test/test_error_model.py:64-65
    m.del_for[0]["A"] = float("nan")  # `u < nan` is False: never fires
    t = m.device_tables().arrays
    lane = list(m.ins_for[2].keys()).index("G")
    assert (t["ins_nz"][0, 2] >> lane) & 1 and t["ins_thr"][0, 2, lane] == 0xFFFFFFFF
    assert not (t["del_nz"][0, 0] >> error_models.BASES.index("A")) & 1


# ---- plug-in surface ---------------------------------------------------------------------------------
def _tables_equal(a, b):
    assert a.scalars == b.scalars
    assert set(a.arrays) == set(b.arrays)
    for k in a.arrays:
        assert np.array_equal(np.asarray(a.arrays[k]), np.asarray(b.arrays[k])), k


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir("/root/reference/iss"), reason="needs /root/reference")
def test_reference_side_model_instances_lower_without_a_class_swap():
    """error_models.lower takes the reference's OWN model objects (iss.error_models.kde.KDErrorModel, Basic, Perfect):
    the tables are identical to those of this package's classes built from the same inputs."""
    import patched_reference as pr

    pr.import_reference()
    from iss.error_models import basic, kde, perfect

    npz = fixtures.model_path("profile:HiSeq")
    _tables_equal(error_models.lower(kde.KDErrorModel(npz, 400, 30, True)), error_models.lower(error_models.KDErrorModel(npz, 400, 30, True)))
    _tables_equal(error_models.lower(basic.BasicErrorModel(None, None, True)), error_models.lower(error_models.BasicErrorModel(None, None, True)))
    _tables_equal(error_models.lower(perfect.PerfectErrorModel()), error_models.lower(error_models.PerfectErrorModel()))
    assert error_models.sampling_family(kde.KDErrorModel(npz)) == "kde"


def test_models_that_override_a_hot_path_method_are_refused():
    """A subclass that replaces gen_phred_scores / mut_sequence / ... with Python code cannot run on the GPU: lowering must
    say so instead of silently ignoring the override.  Attribute-only plug-ins stay welcome."""
    npz = fixtures.model_path("profile:HiSeq")

    class Louder(error_models.KDErrorModel):
        def gen_phred_scores(self, cdfs, orientation):  # what a third-party plug-in of the reference might do
            return [40] * int(self.read_length)

    with pytest.raises(TypeError, match="gen_phred_scores"):
        error_models.lower(Louder(npz))

    class Shorter(error_models.KDErrorModel):
        def random_insert_size(self):
            return 10

    with pytest.raises(TypeError, match="random_insert_size"):
        Shorter(npz).device_tables()
    m = error_models.KDErrorModel(npz)
    m.mut_sequence = lambda record, orientation: record  # instance-level override
    with pytest.raises(TypeError, match="mut_sequence"):
        error_models.lower(m)

    class TablesOnly(error_models.KDErrorModel):  # changes data, not code: fine
        def __init__(self, path):
            super().__init__(path)
            self.i_size_cdf = np.linspace(0.0, 1.0, 500)

    t = error_models.lower(TablesOnly(npz))
    assert t.scalars["n_insert"] == 500

    class Duck(object):  # not derived from anything: the KDE attribute set alone decides
        pass

    src = error_models.KDErrorModel(npz)
    d = Duck()
    for k, v in vars(src).items():
        setattr(d, k, v)
    _tables_equal(error_models.lower(d), error_models.lower(src))
