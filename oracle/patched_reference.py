"""The reference CPU path driven by the shared Philox stream ("oracle mode (ii)").

TEST INFRASTRUCTURE ONLY.  Runs only where /root/reference exists (this container); it is the
generator of the golden fixtures under tests/golden/ and the pin for oracle/iss_oracle.c.

What it does: imports the UNMODIFIED reference package (`/root/reference/iss`) with the
stand-in `Bio` ahead of it on sys.path, then replaces only the five RNG entry points the hot
path calls -- random.random, random.randrange, numpy.random.rand, numpy.random.choice,
numpy.random.normal -- with functions that return the Philox word of the slot belonging to the
calling site (oracle/philox.py, RNG contract v2).  The calling site is recovered from the
caller's frame (function name, source line text, and its locals `position`, `orientation`,
`nucl_to_insert`, `phred_list`).  Every line of arithmetic that decides a read -- CDF searches,
Bernoulli tests, list insert/pop semantics, length adjustment, reverse complement, FASTQ text --
is the reference's own code.

Sites (reference file:line):
  iss/error_models/kde.py:97            np.random.rand     -> PAIR blk0 w0  (insert size)
  iss/error_models/kde.py:74            np.random.choice   -> PAIR blk0 w1/w2 (quality bin)
  iss/error_models/kde.py:84            np.random.rand     -> QUAL(o) phred word of len(phred_list)
  iss/generator.py:135 / 144 / 175      random.randrange   -> PAIR blk1 (w0,w1) / blk1 (w2,w3) / blk2 (w0,w1)
  iss/generator.py:88                   np.random.rand     -> PAIR blk0 w3 (gc_bias)
  iss/generator.py:122                  np.random.normal   -> PAIR blk2 (w2,w3) via integer fragment table
  iss/error_models/__init__.py:194      random.random      -> INS(o) word of the site's rank among the non-zero (position, lane) entries
  iss/error_models/__init__.py:209      random.random      -> DEL(o) word of the position's rank among the positions with a non-zero entry
  iss/error_models/__init__.py:94       random.random      -> QUAL(o) test word of position
  iss/error_models/__init__.py:95-97    np.random.choice   -> SUBC(o) word of position
"""
import linecache
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("ISS_REFERENCE_ROOT", "/root/reference")

if HERE not in sys.path:
    sys.path.insert(0, HERE)
import philox  # noqa: E402


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "iss"))


def import_reference():
    """Import the unmodified reference with the Bio stand-in ahead of it. Returns the `iss` package."""
    standin = os.path.join(HERE, "bio_standin")
    for p in (REFERENCE_ROOT, standin):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REFERENCE_ROOT)
    sys.path.insert(0, standin)
    import iss  # noqa: F401
    import iss.generator  # noqa: F401
    import iss.error_models.kde  # noqa: F401

    assert os.path.realpath(iss.__file__).startswith(os.path.realpath(REFERENCE_ROOT)), iss.__file__
    return iss


def legacy_choice_index(p, u):
    """Index numpy's legacy RandomState.choice(a, p=p) returns for uniform u (numpy/random/mtrand.pyx:
    cdf = p.cumsum(); cdf /= cdf[-1]; idx = cdf.searchsorted(u, side='right'))."""
    p = np.asarray(p, dtype=np.float64)
    cdf = p.cumsum()
    cdf /= cdf[-1]
    return int(cdf.searchsorted(u, side="right"))


def fragment_table(mu, sd):
    """Integer fragment-length distribution for `int(np.random.normal(mu, sd))` (generator.py:122).

    The Philox contract defines that draw by inversion on integers: value k has probability mass
    P(int(X)=k), X~N(mu,sd), tabulated on [mu-8sd-1, mu+8sd+1] as 64-bit cumulative thresholds
    T_k = floor(cdf_k * 2**64); the draw is lo + #{k : T_k <= r64}.  Built by ONE function shared by
    product and oracle callers (the table is data handed to both), so no floating point runs on the
    device.  sd == 0 gives the constant int(mu).  Returns (lo, thresholds uint64[n])."""
    from scipy.stats import norm

    mu = float(mu)
    sd = float(sd)
    if sd <= 0:
        return int(mu), np.zeros(0, dtype=np.uint64)
    lo = int(np.floor(mu - 8 * sd)) - 1
    hi = int(np.ceil(mu + 8 * sd)) + 1
    ks = np.arange(lo, hi + 1)
    # int() truncates toward zero: for k>0 value k <- [k,k+1); k<0 <- (k-1,k]; k==0 <- (-1,1)
    # int() truncates toward zero: int(X) <= k  <=>  X < k+1 (k >= 0)  or  X <= k (k < 0)
    upper = np.where(ks >= 0, ks + 1.0, ks.astype(np.float64))
    cdf = np.clip(norm.cdf(upper, loc=mu, scale=sd), 0.0, 1.0)
    thr = np.floor(cdf[:-1] * 2.0 ** 64)
    thr = np.minimum(thr, float(2 ** 64 - 1025)).astype(np.uint64)
    return lo, thr


def basic_phred_table(mean_quality, sd=0.01, cap=0.9999):
    """Oracle-side twin of insilicoseq_b200.error_models.basic_phred_table (same definition, separate code):
    BasicErrorModel's per-position normal draw (iss/error_models/basic.py:52) by inversion on the phred integer."""
    import math

    from scipy.stats import norm

    mu = float(1 - 10 ** (-mean_quality / 10))
    vmax = int(round(-10 * np.log10(1 - cap)))
    rows = []
    v = vmax
    while v > -1000:
        t = float(norm.cdf(1.0 - 10.0 ** (-(v - 0.5) / 10.0), loc=mu, scale=sd))
        if t * 2.0 ** 32 < 1.0:
            break
        rows.append((v, t))
        v -= 1
    rows.sort()
    thr = np.array([min(math.ceil(t * 2.0 ** 32), 2 ** 32 - 1) for _, t in rows], dtype=np.uint32)
    val = np.array([v] + [k for k, _ in rows], dtype=np.int64)
    return thr, val


class PhiloxDriver:
    """Context manager: while active, the five RNG entry points return slot-addressed Philox draws."""

    def __init__(self, seed, cpu_number=0, trace=None):
        self.key = philox.make_key(seed, cpu_number)
        self.item_base = 0  # pairs of preceding work items of this worker
        self.cur_item_pairs = 0
        self.i = None
        self.attempt = 0
        self.trace = trace  # optional list collecting (P, attempt, stream, block, word)
        self._saved = None
        self._frag_cache = {}

    # ---- addressing -------------------------------------------------------------------
    def _word(self, stream, blk, w):
        P = self.item_base + self.i
        if self.trace is not None:
            self.trace.append((P, self.attempt, stream, blk, w))
        return philox.block(self.key, P, stream, blk, self.attempt)[w]

    def _u(self, stream, blk, w):
        return self._word(stream, blk, w) * 2.0 ** -32

    def _r64(self, stream, blk, w):
        lo = self._word(stream, blk, w)
        hi = self._word(stream, blk, w + 1)
        return lo | (hi << 32)

    @staticmethod
    def _orient(frame):
        o = frame.f_locals["orientation"]
        return {"forward": 0, "reverse": 1}[o]

    @staticmethod
    def _src(frame):
        return linecache.getline(frame.f_code.co_filename, frame.f_lineno)

    def _ranks(self, tables, kind):
        """Dense slot ranks of the non-zero indel sites (RNG contract v2), over loop positions 0..L-2."""
        key = (id(tables), kind)
        if key not in self._frag_cache:
            ranks = {}
            n = len(tables) - 1  # the loop runs over range(read_length - 1)
            for pos in range(n):
                if kind == "ins":
                    for lane, p in enumerate(tables[pos].values()):
                        if p > 0:
                            ranks[(pos, lane)] = len(ranks)
                elif any(p > 0 for p in tables[pos].values()):
                    ranks[pos] = len(ranks)
            self._frag_cache[key] = ranks
        return self._frag_cache[key]

    # ---- replacements -----------------------------------------------------------------
    def _random(self):
        f = sys._getframe(1)
        name = f.f_code.co_name
        if name == "introduce_indels":
            o = self._orient(f)
            pos = f.f_locals["position"]
            src = self._src(f)
            if "< prob" in src:
                tables = f.f_locals["insertions"]
                lane = list(tables[pos].keys()).index(f.f_locals["nucl_to_insert"])
                r = self._ranks(tables, "ins").get((pos, lane))
                if r is None:
                    return 1.0  # probability not > 0: no slot, the test is false for any draw
                return self._u(philox.read_stream(o, philox.K_INS), r >> 2, r & 3)
            if "deletions[" in src:
                r = self._ranks(f.f_locals["deletions"], "del").get(pos)
                if r is None:
                    return 1.0
                return self._u(philox.read_stream(o, philox.K_DEL), r >> 2, r & 3)
            raise RuntimeError("unmapped random.random() site in introduce_indels: %r" % src)
        if name == "mut_sequence":
            o = self._orient(f)
            pos = f.f_locals["position"]
            return self._u(philox.read_stream(o, philox.K_QUAL), pos >> 1, 2 * (pos & 1) + 1)
        raise RuntimeError("unmapped random.random() caller %s" % name)

    def _randrange(self, start, stop=None):
        f = sys._getframe(1)
        if f.f_code.co_name != "simulate_read":
            raise RuntimeError("unmapped random.randrange() caller %s" % f.f_code.co_name)
        if stop is None:
            start, stop = 0, start
        src = self._src(f)
        if "len(record.seq) - fragment_length" in src:
            blk, w = 1, 0
        elif "len(record.seq) - read_length" in src:
            blk, w = 1, 2
        elif "randrange(read_length, len(record.seq))" in src:
            blk, w = 2, 0
        else:
            raise RuntimeError("unmapped random.randrange() site: %r" % src)
        width = int(stop) - int(start)
        if width <= 0:
            raise ValueError("empty range for randrange() (%d, %d, %d)" % (start, stop, width))
        return int(start) + ((self._r64(philox.S_PAIR, blk, w) * width) >> 64)

    def _rand(self, *shape):
        if shape:
            raise RuntimeError("np.random.rand with a shape is not on the hot path")
        f = sys._getframe(1)
        name = f.f_code.co_name
        if name == "random_insert_size":
            return self._u(philox.S_PAIR, 0, 0)
        if name == "gen_phred_scores":
            o = self._orient(f)
            pos = len(f.f_locals["phred_list"])
            return self._u(philox.read_stream(o, philox.K_QUAL), pos >> 1, 2 * (pos & 1))
        if name == "reads_generator":
            return self._u(philox.S_PAIR, 0, 3)
        raise RuntimeError("unmapped np.random.rand() caller %s" % name)

    def _choice(self, a, size=None, replace=True, p=None):
        if size is not None or p is None:
            raise RuntimeError("np.random.choice form not on the hot path")
        f = sys._getframe(1)
        name = f.f_code.co_name
        if name == "gen_phred_scores":
            o = self._orient(f)
            u = self._u(philox.S_PAIR, 0, 1 + o)
        elif name == "mut_sequence":
            o = self._orient(f)
            pos = f.f_locals["position"]
            u = self._u(philox.read_stream(o, philox.K_SUBC), pos >> 2, pos & 3)
        else:
            raise RuntimeError("unmapped np.random.choice() caller %s" % name)
        pa = np.asarray(p, dtype=np.float64)
        if np.isnan(pa).any():
            raise ValueError("probabilities contain NaN")
        if abs(float(pa.sum()) - 1.0) > 1e-8:
            raise ValueError("probabilities do not sum to 1")
        idx = legacy_choice_index(pa, u)
        seq = list(a) if not isinstance(a, (int, np.integer)) else list(range(int(a)))
        return seq[idx] if not isinstance(seq[idx], str) else np.str_(seq[idx])

    def _normal(self, loc=0.0, scale=1.0, size=None):
        f = sys._getframe(1)
        if f.f_code.co_name == "gen_phred_scores" and size is not None:
            # BasicErrorModel (basic.py:52): position p takes the QUAL phred word of p; the value returned is the
            # probability whose phred is exactly the tabulated integer (prob_to_phred(min(x, .9999)) maps it back)
            o = self._orient(f)
            mq = f.f_locals["mean_quality"]
            if ("basic", mq, scale) not in self._frag_cache:
                self._frag_cache[("basic", mq, scale)] = basic_phred_table(mq, scale)
            thr, val = self._frag_cache[("basic", mq, scale)]
            out = np.empty(int(size), dtype=np.float64)
            for pos in range(int(size)):
                w = self._word(philox.read_stream(o, philox.K_QUAL), pos >> 1, 2 * (pos & 1))
                v = int(val[int(np.searchsorted(thr, np.uint32(w), side="right"))])
                out[pos] = 1.0 - 10.0 ** (-v / 10.0)
            return out
        if size is not None or f.f_code.co_name != "simulate_read":
            raise RuntimeError("np.random.normal form not on the hot path")
        keyt = (float(loc), float(scale))
        if keyt not in self._frag_cache:
            self._frag_cache[keyt] = fragment_table(loc, scale)
        lo, thr = self._frag_cache[keyt]
        r64 = self._r64(philox.S_PAIR, 2, 2)
        k = lo + int(np.searchsorted(thr, np.uint64(r64), side="right"))
        return float(k)

    # ---- context ----------------------------------------------------------------------
    def __enter__(self):
        iss = import_reference()
        gen = iss.generator
        self._gen = gen
        self._saved = (random.random, random.randrange, np.random.rand, np.random.choice, np.random.normal,
                       gen.simulate_read, gen.simulate_reads)
        random.random = self._random
        random.randrange = self._randrange
        np.random.rand = self._rand
        np.random.choice = self._choice
        np.random.normal = self._normal
        orig_read = self._saved[5]
        orig_reads = self._saved[6]
        drv = self

        def simulate_read(record, error_model, i, cpu_number, sequence_type):
            if drv.i == i and drv._same_item:
                drv.attempt += 1
            else:
                drv.attempt = 0
            drv.i = i
            drv._same_item = True
            return orig_read(record, error_model, i, cpu_number, sequence_type)

        def simulate_reads(*args, **kwargs):
            n_pairs = kwargs["n_pairs"] if "n_pairs" in kwargs else args[2]
            drv._same_item = False
            drv.i = None
            try:
                return orig_reads(*args, **kwargs)
            finally:
                drv.item_base += int(n_pairs)

        gen.simulate_read = simulate_read
        gen.simulate_reads = simulate_reads
        self._same_item = False
        return self

    def __exit__(self, *exc):
        (random.random, random.randrange, np.random.rand, np.random.choice, np.random.normal,
         self._gen.simulate_read, self._gen.simulate_reads) = self._saved
        return False

    def set_pair(self, i, attempt=0, item_base=None):
        """For driving simulate_read / error-model methods directly (unit-level fixtures)."""
        self.i = i
        self.attempt = attempt
        self._same_item = False
        if item_base is not None:
            self.item_base = item_base


def run_worker(work, error_model, cpu_number, worker_prefix, seed, sequence_type="metagenomics", gc_bias=False):
    """Reference `worker_iterator` (iss/generator.py:223) under the Philox stream.

    work: list of (SeqRecord, n_pairs, 'default').  Writes {worker_prefix}_R1.fastq/_R2.fastq/.vcf
    exactly as the reference does."""
    with PhiloxDriver(seed, cpu_number) as drv:
        drv._gen.worker_iterator(work, error_model, cpu_number, worker_prefix, seed, sequence_type, gc_bias)


def make_record(seq, rec_id):
    import_reference()
    from Bio.Seq import Seq
    from Bio.SeqRecord import SeqRecord

    return SeqRecord(Seq(seq), id=rec_id, description="")


def load_simple_model(kind, fragment_length=None, fragment_sd=None, store_mutations=False):
    """BasicErrorModel / PerfectErrorModel of the reference.  The reference's PerfectErrorModel never sets
    `store_mutations` (perfect.py:16-33) although mut_sequence reads it (error_models/__init__.py:98) -- it raises
    AttributeError on the first Q40 'error' (~1e-4 per base).  The attribute is injected on the INSTANCE here so the
    evident intent (identity substitution) can be pinned; no reference code is changed."""
    import_reference()
    from iss.error_models import basic, perfect

    if kind == "basic":
        return basic.BasicErrorModel(fragment_length, fragment_sd, store_mutations)
    m = perfect.PerfectErrorModel(fragment_length, fragment_sd)
    m.store_mutations = False
    return m


def load_kde_model(npz_path, fragment_length=None, fragment_sd=None, store_mutations=False):
    iss = import_reference()
    from iss.error_models import kde

    return kde.KDErrorModel(npz_path, fragment_length, fragment_sd, store_mutations)
