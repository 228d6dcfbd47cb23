"""ctypes front-end of the C oracle (oracle/iss_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs; never by the product package.  Needs only numpy (no reference, no Biopython), so it also
runs on the GPU box.

`OracleModel` unpacks an InSilicoSeq KDE profile (.npz written by iss/bam.py:82-97, attributes as
iss/error_models/kde.py:24-50 names them) into the flat f64 tables the C code reads.  The tables keep
the reference's floating-point values; derived arrays are computed with the same numpy expressions the
reference evaluates at run time:
  * quality-bin cdf: `norm_mean = mean / sum(mean)` (kde.py:72) then numpy legacy choice's
    `cdf = p.cumsum(); cdf /= cdf[-1]`
  * substitution choice cdf: same cumsum/normalise of the stored probabilities
  * phred_prob[q] = 1 - 10 ** (-np.int64(q) / 10)   (iss/util.py:27-28 on the np.int64 searchsorted returns)
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libiss_oracle.so")
MAX_ALT = 4
BASES = "ACGT"


def build(force=False):
    src = os.path.join(HERE, "iss_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=gnu11", "-Wall", "-o", LIB_PATH, src])
    return LIB_PATH


class _CModel(ctypes.Structure):
    _fields_ = [
        ("read_length", ctypes.c_int32),
        ("nq", ctypes.c_int32),
        ("n_insert", ctypes.c_int32),
        ("frag_mode", ctypes.c_int32),
        ("frag_lo", ctypes.c_int64),
        ("n_frag", ctypes.c_int32),
        ("store_mutations", ctypes.c_int32),
        ("frag_thr", ctypes.c_void_p),
        ("insert_cdf", ctypes.c_void_p),
        ("bin_cdf", ctypes.c_void_p),
        ("bin_rows", ctypes.c_void_p),
        ("qual_cdf", ctypes.c_void_p),
        ("phred_prob", ctypes.c_void_p),
        ("subst_n", ctypes.c_void_p),
        ("subst_cdf", ctypes.c_void_p),
        ("subst_alt", ctypes.c_void_p),
        ("ins_p", ctypes.c_void_p),
        ("ins_base", ctypes.c_void_p),
        ("del_p", ctypes.c_void_p),
        ("ins_rank", ctypes.c_void_p),
        ("del_rank", ctypes.c_void_p),
        ("fixed_insert", ctypes.c_int32),
        ("qual_mode", ctypes.c_int32),
        ("const_phred", ctypes.c_int32),
        ("n_basic", ctypes.c_int32),
        ("basic_thr", ctypes.c_void_p),
        ("basic_val", ctypes.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        L.iss_oracle_generate.restype = ctypes.c_int
        L.iss_oracle_generate.argtypes = [
            ctypes.POINTER(_CModel), ctypes.c_void_p, ctypes.c_int64, ctypes.c_char_p, ctypes.c_int32,
            ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int32, ctypes.c_int32,
            ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64,
            ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64),
            ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64),
        ]
        L.iss_oracle_fastq_bytes.restype = ctypes.c_uint64
        L.iss_oracle_fastq_bytes.argtypes = [ctypes.c_int32, ctypes.c_uint64, ctypes.c_int32, ctypes.c_uint64,
                                             ctypes.c_uint64]
        L.iss_oracle_philox.restype = None
        L.iss_oracle_philox.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib = L
    return _lib


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().iss_oracle_philox(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return tuple(int(x) for x in out)


def make_key(seed, cpu_number):
    return (int(seed) + int(cpu_number)) & 0xFFFFFFFFFFFFFFFF


def _choice_cdf(p):
    p = np.asarray(p, dtype=np.float64)
    cdf = p.cumsum()
    cdf /= cdf[-1]
    return cdf


def phred_prob_table(nq):
    return np.array([1 - 10 ** (-np.int64(q) / 10) for q in range(nq + 1)], dtype=np.float64)


def truncate_profile(npz_in, npz_out, read_length):
    """The '150 bp' benchmark models: read_length=N and every per-position array sliced [:N]
    (BASELINE.md section 1).  Everything else is kept verbatim."""
    d = np.load(npz_in, allow_pickle=True)
    out = {}
    for k in d.files:
        v = d[k]
        if k == "read_length":
            out[k] = np.int64(read_length)
        elif k in ("quality_hist_forward", "quality_hist_reverse"):
            o = np.empty(len(v), dtype=object)
            for b in range(len(v)):
                o[b] = list(v[b])[:read_length]
            out[k] = o
        elif k in ("subst_choices_forward", "subst_choices_reverse", "ins_forward", "ins_reverse",
                   "del_forward", "del_reverse"):
            out[k] = v[:read_length]
        else:
            out[k] = v
    with open(npz_out, "wb") as fh:
        np.savez_compressed(fh, **out)
    return npz_out


class OracleModel:
    """Flat f64 view of an error model for the C oracle: a KDE profile (.npz), or the reference's
    BasicErrorModel / PerfectErrorModel attribute sets (npz_path = "basic" / "perfect")."""

    def __init__(self, npz_path, fragment_length=None, fragment_sd=None, store_mutations=False, frag_table=None,
                 basic_table=None):
        if npz_path in ("basic", "perfect"):
            self._init_simple(npz_path, fragment_length, fragment_sd, store_mutations, frag_table, basic_table)
            return
        prof = np.load(npz_path, allow_pickle=True)
        assert str(prof["model"]) == "kde", prof["model"]
        L = int(prof["read_length"])
        self.read_length = L
        self.npz_path = npz_path
        self.store_mutations = bool(store_mutations)
        ic = np.ascontiguousarray(prof["insert_size"], dtype=np.float64)
        means = [np.asarray(prof["mean_count_forward"]), np.asarray(prof["mean_count_reverse"])]
        quals = [prof["quality_hist_forward"], prof["quality_hist_reverse"]]
        subs = [prof["subst_choices_forward"], prof["subst_choices_reverse"]]
        ins = [prof["ins_forward"], prof["ins_reverse"]]
        dels = [prof["del_forward"], prof["del_reverse"]]

        nq = None
        for o in range(2):
            for b in range(4):
                rows = quals[o][b]
                if len(rows):
                    nq = len(rows[0]) if nq is None else nq
                    assert all(len(r) == nq for r in rows)
        self.nq = nq
        bin_cdf = np.zeros((2, 4), dtype=np.float64)
        bin_rows = np.zeros((2, 4), dtype=np.int32)
        qual_cdf = np.zeros((2, 4, L, nq), dtype=np.float64)
        for o in range(2):
            mean = means[o]
            norm_mean = mean / sum(mean)  # kde.py:72
            bin_cdf[o] = _choice_cdf(norm_mean)
            for b in range(4):
                rows = quals[o][b]
                bin_rows[o, b] = len(rows)
                n = min(len(rows), L)
                if n:
                    qual_cdf[o, b, :n] = np.asarray(rows[:n], dtype=np.float64)
        subst_n = np.zeros((2, L, 4), dtype=np.int32)
        subst_cdf = np.ones((2, L, 4, MAX_ALT), dtype=np.float64)
        subst_alt = np.full((2, L, 4, MAX_ALT), ord("N"), dtype=np.uint8)
        ins_p = np.zeros((2, L, 4), dtype=np.float64)
        ins_base = np.full((2, L, 4), ord("A"), dtype=np.uint8)
        del_p = np.zeros((2, L, 4), dtype=np.float64)
        for o in range(2):
            for pos in range(L):
                for c, base in enumerate(BASES):
                    alts, probs = subs[o][pos][base]
                    na = len(alts)
                    subst_n[o, pos, c] = na
                    subst_cdf[o, pos, c, :na] = _choice_cdf(probs)
                    subst_alt[o, pos, c, :na] = [ord(a) for a in alts]
                    del_p[o, pos, c] = dels[o][pos][base]
                for lane, (base, p) in enumerate(ins[o][pos].items()):
                    ins_p[o, pos, lane] = p
                    ins_base[o, pos, lane] = ord(base)
        self._arrays = dict(insert_cdf=ic, bin_cdf=bin_cdf, bin_rows=bin_rows, qual_cdf=qual_cdf,
                            phred_prob=phred_prob_table(nq), subst_n=subst_n, subst_cdf=subst_cdf,
                            subst_alt=subst_alt, ins_p=ins_p, ins_base=ins_base, del_p=del_p)
        self.frag_lo = 0
        frag_thr = np.zeros(1, dtype=np.uint64)
        self.frag_mode = 0
        self.n_frag = 0
        if fragment_length is not None and fragment_sd is not None:
            assert frag_table is not None, "pass the (lo, thresholds) table the product built"
            self.frag_mode = 1
            self.frag_lo = int(frag_table[0])
            t = np.ascontiguousarray(frag_table[1], dtype=np.uint64)
            self.n_frag = len(t)
            frag_thr = t if len(t) else frag_thr
        self._arrays["frag_thr"] = frag_thr
        self._arrays["basic_thr"] = np.zeros(1, dtype=np.uint32)
        self._arrays["basic_val"] = np.zeros(1, dtype=np.uint8)
        self.c = _CModel()
        self._fill()

    def _init_simple(self, kind, fragment_length, fragment_sd, store_mutations, frag_table, basic_table):
        """iss/error_models/basic.py:18-39 and perfect.py:16-33: L=125, insert 200, no indels; basic substitutes
        uniformly among the 3 other bases, perfect 'substitutes' a base by itself."""
        L = 125
        self.read_length = L
        self.npz_path = kind
        self.store_mutations = bool(store_mutations) and kind == "basic"
        self.nq = 41
        self.kind = kind
        if kind == "basic":
            choices = {"A": ("TCG", [1 / 3] * 3), "T": ("ACG", [1 / 3] * 3), "C": ("ATG", [1 / 3] * 3), "G": ("ATC", [1 / 3] * 3)}
        else:
            choices = {"A": ("ATCG", [1, 0, 0, 0]), "T": ("ATCG", [0, 1, 0, 0]), "C": ("ATCG", [0, 0, 1, 0]), "G": ("ATCG", [0, 0, 0, 1])}
        subst_n = np.zeros((2, L, 4), dtype=np.int32)
        subst_cdf = np.ones((2, L, 4, MAX_ALT), dtype=np.float64)
        subst_alt = np.full((2, L, 4, MAX_ALT), ord("N"), dtype=np.uint8)
        for c, base in enumerate(BASES):
            alts, probs = choices[base]
            subst_n[:, :, c] = len(alts)
            subst_cdf[:, :, c, :len(alts)] = _choice_cdf(probs)
            subst_alt[:, :, c, :len(alts)] = [ord(a) for a in alts]
        ins_base = np.zeros((2, L, 4), dtype=np.uint8)
        ins_base[:] = [ord(b) for b in "ATCG"]
        # phred_to_prob on the python ints prob_to_phred returns (identical values to the np.int64 form)
        self._arrays = dict(insert_cdf=np.zeros(1), bin_cdf=np.ones((2, 4)), bin_rows=np.zeros((2, 4), dtype=np.int32),
                            qual_cdf=np.zeros(1), phred_prob=np.array([1 - 10 ** (-q / 10) for q in range(self.nq + 1)]),
                            subst_n=subst_n, subst_cdf=subst_cdf, subst_alt=subst_alt, ins_p=np.zeros((2, L, 4)),
                            ins_base=ins_base, del_p=np.zeros((2, L, 4)))
        self.frag_lo, self.frag_mode, self.n_frag = 0, 0, 0
        frag_thr = np.zeros(1, dtype=np.uint64)
        if fragment_length is not None and fragment_sd is not None:
            self.frag_mode = 1
            self.frag_lo = int(frag_table[0])
            t = np.ascontiguousarray(frag_table[1], dtype=np.uint64)
            self.n_frag = len(t)
            frag_thr = t if len(t) else frag_thr
        self._arrays["frag_thr"] = frag_thr
        if kind == "basic":
            thr, val = basic_table
            self._arrays["basic_thr"] = np.ascontiguousarray(thr, dtype=np.uint32) if len(thr) else np.zeros(1, dtype=np.uint32)
            self._arrays["basic_val"] = np.ascontiguousarray(val, dtype=np.uint8)
            self._n_basic = len(thr)
        else:
            self._arrays["basic_thr"] = np.zeros(1, dtype=np.uint32)
            self._arrays["basic_val"] = np.zeros(1, dtype=np.uint8)
            self._n_basic = 0
        self.c = _CModel()
        self._fill()

    def _fill(self):
        c = self.c
        c.read_length = self.read_length
        c.nq = self.nq
        c.n_insert = len(self._arrays["insert_cdf"])
        c.frag_mode = self.frag_mode
        c.frag_lo = self.frag_lo
        c.n_frag = self.n_frag
        c.store_mutations = int(self.store_mutations)
        # RNG contract v2: dense slot ranks of the non-zero indel sites over the loop positions 0..L-2
        L = self.read_length
        ins_p, del_p = self._arrays["ins_p"], self._arrays["del_p"]
        ins_rank = np.full((2, L, 4), -1, dtype=np.int32)
        del_rank = np.full((2, L), -1, dtype=np.int32)
        for o in range(2):
            ri = rd = 0
            for pos in range(L - 1):
                for lane in range(4):
                    if ins_p[o, pos, lane] > 0:
                        ins_rank[o, pos, lane] = ri
                        ri += 1
                if (del_p[o, pos] > 0).any():
                    del_rank[o, pos] = rd
                    rd += 1
        self._arrays["ins_rank"] = ins_rank
        self._arrays["del_rank"] = del_rank
        kind = getattr(self, "kind", "kde")
        c.fixed_insert = -1 if kind == "kde" else 200
        c.qual_mode = {"kde": 0, "perfect": 1, "basic": 2}[kind]
        c.const_phred = 40 if kind == "perfect" else 0
        c.n_basic = getattr(self, "_n_basic", 0)
        for name, arr in self._arrays.items():
            arr = np.ascontiguousarray(arr)
            self._arrays[name] = arr
            setattr(c, name, arr.ctypes.data)

    def set_table(self, name, arr):
        """Overwrite one table (tests force thresholds like the reference's tests do)."""
        self._arrays[name] = np.ascontiguousarray(arr, dtype=self._arrays[name].dtype)
        self._fill()

    def table(self, name):
        return self._arrays[name]


def fastq_bytes(hdrlen, cpu, L, i0, i1):
    return int(lib().iss_oracle_fastq_bytes(hdrlen, cpu, L, i0, i1))


class OracleError(RuntimeError):
    def __init__(self, code):
        RuntimeError.__init__(self, "oracle error %d" % code)
        self.code = code


E_SHORT = 1


def generate(model, record, header, key, pair_base, cpu, i0, i1, amplicon=False, gc_bias=False, vcf=False):
    """Pairs i0..i1-1 of one work item -> (R1 bytes, R2 bytes[, VCF bytes]).  record: bytes (ASCII bases)."""
    rec = np.frombuffer(record, dtype=np.uint8)
    hdr = header.encode() if isinstance(header, str) else header
    cap = fastq_bytes(len(hdr), cpu, model.read_length, i0, i1)
    o1 = np.empty(max(cap, 1), dtype=np.uint8)
    o2 = np.empty(max(cap, 1), dtype=np.uint8)
    l1 = ctypes.c_uint64(0)
    l2 = ctypes.c_uint64(0)
    vcap = 256 * 64 * max(1, i1 - i0) if vcf else 0
    vb = np.empty(max(vcap, 1), dtype=np.uint8)
    vl = ctypes.c_uint64(0)
    rc = lib().iss_oracle_generate(ctypes.byref(model.c), rec.ctypes.data, len(rec), hdr, len(hdr), key, pair_base,
                                   cpu, int(amplicon), int(gc_bias), i0, i1, o1.ctypes.data, o2.ctypes.data, cap,
                                   ctypes.byref(l1), ctypes.byref(l2), vb.ctypes.data if vcf else None, vcap,
                                   ctypes.byref(vl))
    if rc:
        raise OracleError(rc)
    r1 = o1[: l1.value].tobytes()
    r2 = o2[: l2.value].tobytes()
    if vcf:
        return r1, r2, vb[: vl.value].tobytes()
    return r1, r2


def run_worker(model, work, cpu, seed, amplicon=False, gc_bias=False, vcf=False):
    """Reference worker_iterator semantics (generator.py:223-251) on the C oracle.
    work: list of (header str, record bytes, n_pairs).  Returns (R1, R2[, VCF]) bytes of the worker."""
    key = make_key(seed, cpu)
    base = 0
    parts1, parts2, partsv = [], [], []
    for header, record, n_pairs in work:
        try:
            res = generate(model, record, header, key, base, cpu, 0, n_pairs, amplicon, gc_bias, vcf)
            parts1.append(res[0])
            parts2.append(res[1])
            if vcf:
                partsv.append(res[2])
        except OracleError as e:
            if e.code != E_SHORT:
                raise
        base += n_pairs
    if vcf:
        return b"".join(parts1), b"".join(parts2), b"".join(partsv)
    return b"".join(parts1), b"".join(parts2)
