"""ErrorModel plugin surface of InSilicoSeq, lowered to the B200 engine's integer-threshold tables.

Mirrors the reference's classes and attribute names (iss/error_models/__init__.py:15, kde.py:9,
basic.py:10, perfect.py:7): a model is an object with `read_length`, `quality_forward/reverse`,
`subst_choices_for/rev`, `ins_for/rev`, `del_for/rev`, `fragment_length`, `fragment_sd`,
`store_mutations` (+ KDE: `i_size_cdf`, `mean_forward/reverse`).  Any subclass that provides those
attributes can be handed to the engine: `device_tables()` turns the f64 probabilities into exact
integer thresholds on 32-bit Philox words (include/iss_b200.h), after which the device path has no
floating point.  The per-read methods of the reference (introduce_indels, gen_phred_scores,
mut_sequence ...) are what the CUDA kernels implement for whole batches; they are not re-offered as
per-record Python loops (there is no CPU path in this package).  What the engine runs is therefore
fixed by the model's DATA: `lower(model)` accepts an instance of this package's classes or of the
reference's own (iss.error_models.kde.KDErrorModel, ...basic.BasicErrorModel, ...perfect.PerfectErrorModel
-- no class swap needed) and REFUSES a model whose class overrides one of the hot-path methods with
code the GPU cannot know about (`check_plugin_surface`), instead of silently ignoring the override.

Lowering, with w a 32-bit word and u = w * 2**-32 (scaling an f64 by 2**32 is exact):
    searchsorted(cdf, u, 'left')  = #{j : cdf[j] <  u} = #{j : w >  floor(cdf[j] * 2**32)}
    choice(a, p) (side='right')   = #{j : cdf[j] <= u} = #{j : w >= ceil(cdf[j] * 2**32)}
    random() <  p                 = w <  ceil(p * 2**32)
    random() >  P                 = w >  floor(P * 2**32)
"""
import logging
import math
import os
import sys

import numpy as np

from . import _lib

TWO32 = 4294967296.0
BASES = "ACTG"  # device base index: (ascii >> 1) & 3  ->  A=0 C=1 T=2 G=3
QROW = _lib.ISS_QROW
MAX_ALT = _lib.ISS_MAX_ALT
PROFILE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles")


# ---- exact threshold lowering ---------------------------------------------------------------------
def thr_left(cdf):
    """G with  #{j: cdf[j] < u} == #{j: w > G[j]}."""
    c = np.asarray(cdf, dtype=np.float64)
    if np.isnan(c).any():
        raise ValueError("CDF contains NaN")
    g = np.floor(c * TWO32)
    return np.clip(g, 0.0, TWO32 - 1).astype(np.uint32)


def thr_right(cdf):
    """H (uint64) with  #{j: cdf[j] <= u} == #{j: w >= H[j]}."""
    c = np.asarray(cdf, dtype=np.float64)
    if np.isnan(c).any():
        raise ValueError("CDF contains NaN")
    h = np.ceil(c * TWO32)
    return np.clip(h, 0.0, TWO32).astype(np.uint64)


def thr_bernoulli(p):
    """(D uint32, nz bool) with  (u < p) == (nz and w <= D)."""
    a = np.asarray(p, dtype=np.float64)
    a = np.where(np.isnan(a), 0.0, a)  # `u < nan` is False
    c = np.clip(np.ceil(np.clip(a, 0.0, 1.0) * TWO32), 0.0, TWO32)
    nz = c > 0
    d = np.where(nz, c - 1.0, 0.0).astype(np.uint32)
    return d, nz


def thr_greater(P):
    """G with (u > P) == (w > G)."""
    a = np.asarray(P, dtype=np.float64)
    return np.clip(np.floor(a * TWO32), 0.0, TWO32 - 1).astype(np.uint32)


def choice_cdf(p):
    """The cdf numpy's legacy RandomState.choice builds from p (mtrand.pyx: cumsum, then / last)."""
    pa = np.asarray(p, dtype=np.float64)
    if np.isnan(pa).any():
        raise ValueError("probabilities contain NaN")
    if (pa < 0).any():
        raise ValueError("probabilities are not non-negative")
    if abs(math.fsum(pa) - 1.0) > math.sqrt(np.finfo(np.float64).eps):
        raise ValueError("probabilities do not sum to 1")
    cdf = pa.cumsum()
    cdf /= cdf[-1]
    return cdf


def phred_to_prob_table(n=QROW):
    """phred_to_prob (iss/util.py:16-29) on the np.int64 scores searchsorted returns."""
    return np.array([1 - 10 ** (-np.int64(q) / 10) for q in range(n)], dtype=np.float64)


def fragment_table(mu, sd):
    """`int(np.random.normal(mu, sd))` (iss/generator.py:122) as an integer inversion table.

    RNG contract v2 defines the draw on the 64-bit word r of PAIR block 2 (w2, w3):
        fragment_length = lo + #{k : r >= T[k]},   T[k] = floor(P(int(X) <= lo + k) * 2**64),  X ~ N(mu, sd)
    tabulated on mu +- 8 sd.  sd == 0 degenerates to the constant int(mu)."""
    from scipy.stats import norm

    mu = float(mu)
    sd = float(sd)
    if sd <= 0:
        return int(mu), np.zeros(0, dtype=np.uint64)
    lo = int(np.floor(mu - 8 * sd)) - 1
    hi = int(np.ceil(mu + 8 * sd)) + 1
    ks = np.arange(lo, hi + 1)
    # int() truncates toward zero: int(X) <= k  <=>  X < k + 1 (k >= 0)  or  X <= k (k < 0)
    upper = np.where(ks >= 0, ks + 1.0, ks.astype(np.float64))
    cdf = np.clip(norm.cdf(upper, loc=mu, scale=sd), 0.0, 1.0)
    thr = np.floor(cdf[:-1] * 2.0 ** 64)
    thr = np.minimum(thr, float(2 ** 64 - 1025)).astype(np.uint64)
    return lo, thr


class DeviceTables:
    """Flat arrays in the layout of `iss_model` (kept alive here while the C call copies them)."""

    def __init__(self):
        self.arrays = {}
        self.scalars = {}

    def c_struct(self):
        m = _lib.IssModel()
        for k, v in self.scalars.items():
            setattr(m, k, v)
        for k, arr in self.arrays.items():
            arr = np.ascontiguousarray(arr)
            self.arrays[k] = arr
            setattr(m, k, arr.ctypes.data if arr.size else None)
        return m


HOT_PATH_METHODS = ("gen_phred_scores", "random_insert_size", "introduce_error_scores", "mut_sequence", "introduce_indels",
                    "adjust_seq_length")  # iss/error_models/__init__.py:52,69,114,158, kde.py:52,88
_KNOWN_MODULES = ("insilicoseq_b200.error_models", "iss.error_models", "iss.error_models.kde", "iss.error_models.basic",
                  "iss.error_models.perfect")
_FAMILY_OF_CLASS = {"KDErrorModel": "kde", "BasicErrorModel": "basic", "PerfectErrorModel": "perfect"}


def check_plugin_surface(model):
    """The engine replaces the reference's per-read methods with CUDA kernels driven by the model's tables.  A class
    (or instance) that overrides one of them with its own Python code would be ignored -- refuse it instead."""
    for name in HOT_PATH_METHODS:
        if name in vars(model):
            raise TypeError("%s.%s is overridden on the instance: the B200 engine implements the reference's %s on the GPU from "
                            "the model's tables and cannot call Python code per read" % (type(model).__name__, name, name))
        for cls in type(model).__mro__:
            if name in vars(cls):
                if cls.__module__ not in _KNOWN_MODULES:
                    raise TypeError("%s.%s overrides %s: the B200 engine implements the reference's samplers on the GPU from the "
                                    "model's tables and cannot call Python code per read; express the change in the tables "
                                    "(quality_forward/reverse, subst_choices_*, ins_*/del_*, i_size_cdf) instead"
                                    % (cls.__module__, cls.__qualname__, name))
                break


def sampling_family(model):
    """'kde' | 'basic' | 'perfect': which of the reference's three samplers the model's methods denote."""
    for cls in type(model).__mro__:
        if cls.__module__ in _KNOWN_MODULES and cls.__name__ in _FAMILY_OF_CLASS:
            return _FAMILY_OF_CLASS[cls.__name__]
    if hasattr(model, "i_size_cdf") and hasattr(model, "mean_forward"):
        return "kde"  # a conforming attribute set on some other class (the plugin surface is the attributes)
    raise TypeError("cannot tell which sampler %s stands for: derive from KDErrorModel / BasicErrorModel / PerfectErrorModel "
                    "or provide the KDE attribute set (i_size_cdf, mean_forward, ...)" % type(model).__name__)


def lower(model):
    """DeviceTables of any conforming model object -- this package's classes or the reference's own instances."""
    check_plugin_surface(model)
    family = sampling_family(model)
    L = int(model.read_length)
    t = DeviceTables()
    t.scalars.update(read_length=L, store_mutations=int(bool(getattr(model, "store_mutations", False))),
                     n_insert=0, fixed_insert=0, frag_mode=0, n_frag=0, frag_lo=0, const_phred=0, n_basic=0,
                     qual_mode=0)
    # fragment length (generator.py:121-126)
    frag_thr = np.zeros(0, dtype=np.uint64)
    if getattr(model, "fragment_length", None) is not None and getattr(model, "fragment_sd", None) is not None:
        lo, frag_thr = fragment_table(model.fragment_length, model.fragment_sd)
        t.scalars.update(frag_mode=1, n_frag=len(frag_thr), frag_lo=lo)
    t.arrays["frag_thr"] = frag_thr
    t.arrays["err_thr"] = thr_greater(phred_to_prob_table(QROW))
    {"kde": _lower_quality_kde, "basic": _lower_quality_basic, "perfect": _lower_quality_perfect}[family](model, t)
    # substitutions (error_models/__init__.py:83-97): nucl_choices[position][base] = (alts, probs)
    subs = (model.subst_choices_for, model.subst_choices_rev)
    subst_n = np.zeros((2, L, 4), dtype=np.uint8)
    subst_thr = np.full((2, L, 4, MAX_ALT), 1 << 32, dtype=np.uint64)
    subst_alt = np.full((2, L, 4, MAX_ALT), ord("N"), dtype=np.uint8)
    for o in range(2):
        for pos in range(L):
            entry = subs[o][pos]
            for c, base in enumerate(BASES):
                alts, probs = entry[base]
                na = len(alts)
                if not 1 <= na <= MAX_ALT:
                    raise ValueError("substitution table with %d alternatives" % na)
                subst_n[o, pos, c] = na
                subst_thr[o, pos, c, :na] = thr_right(choice_cdf(probs))
                subst_alt[o, pos, c, :na] = [ord(str(a)) for a in alts]
    t.arrays.update(subst_n=subst_n, subst_thr=subst_thr, subst_alt=subst_alt)
    # indels (error_models/__init__.py:176-210): dict base -> probability, insertion lanes in dict order
    ins = (model.ins_for, model.ins_rev)
    dels = (model.del_for, model.del_rev)
    ins_thr = np.zeros((2, L, 4), dtype=np.uint32)
    ins_nz = np.zeros((2, L), dtype=np.uint8)
    ins_base = np.full((2, L, 4), ord("A"), dtype=np.uint8)
    del_thr = np.zeros((2, L, 4), dtype=np.uint32)
    del_nz = np.zeros((2, L), dtype=np.uint8)
    for o in range(2):
        for pos in range(L):
            lanes = list(ins[o][pos].items())
            if len(lanes) > 4:
                raise ValueError("insertion table with more than 4 lanes")
            d, nz = thr_bernoulli([p for _, p in lanes])
            for lane, (base, _) in enumerate(lanes):
                ins_thr[o, pos, lane] = d[lane]
                ins_base[o, pos, lane] = ord(str(base))
                ins_nz[o, pos] |= int(nz[lane]) << lane
            d, nz = thr_bernoulli([dels[o][pos][b] for b in BASES])
            del_thr[o, pos] = d
            for c in range(4):
                del_nz[o, pos] |= int(nz[c]) << c
    t.arrays.update(ins_thr=ins_thr, ins_nz=ins_nz, ins_base=ins_base, del_thr=del_thr, del_nz=del_nz)
    for k in ("insert_thr", "bin_thr", "bin_valid", "qual_thr", "basic_thr", "basic_val"):
        t.arrays.setdefault(k, np.zeros(0, dtype=np.uint32))
    return t


def _lower_quality_kde(model, t):
    """kde.py:52-98: insert size = CDF index, quality bin by np.random.choice, per-position searchsorted."""
    L = int(model.read_length)
    ic = np.asarray(model.i_size_cdf, dtype=np.float64)
    if (np.diff(ic) < 0).any():
        raise ValueError("insert size CDF is not sorted")
    t.arrays["insert_thr"] = thr_left(ic)
    t.scalars.update(qual_mode=0, n_insert=len(ic))
    bin_thr = np.zeros((2, 4), dtype=np.uint64)
    bin_valid = np.zeros((2, 4), dtype=np.uint8)
    qual_thr = np.full((2, 4, L, QROW), 0xFFFFFFFF, dtype=np.uint32)
    for o, (mean, quals) in enumerate(((model.mean_forward, model.quality_forward),
                                       (model.mean_reverse, model.quality_reverse))):
        if len(mean) != 4:
            raise ValueError("expected 4 mean-quality bins")
        norm_mean = mean / sum(mean)  # kde.py:72
        bin_thr[o] = thr_right(choice_cdf(norm_mean))  # np.random.choice(range(4), p=norm_mean), kde.py:74
        for b in range(4):
            rows = quals[b]
            if len(rows) >= L:
                arr = np.asarray(list(rows[:L]), dtype=np.float64)
                if arr.shape[1] > QROW:
                    raise ValueError("quality CDF rows longer than %d" % QROW)
                if (np.diff(arr, axis=1) < 0).any():
                    raise ValueError("quality CDF row is not sorted")
                qual_thr[o, b, :, : arr.shape[1]] = thr_left(arr)
                bin_valid[o, b] = 1
            elif norm_mean[b] > 0:
                # the reference would build a short phred list and fail inside SeqRecord (kde.py:80-86)
                raise ValueError("quality bin %d has probability %g but only %d CDF rows" % (b, norm_mean[b], len(rows)))
    t.arrays.update(bin_thr=bin_thr, bin_valid=bin_valid, qual_thr=qual_thr)


def _lower_quality_basic(model, t):
    """basic.py:41-58: phred ~ normal around quality_forward by inversion on the integer table, constant insert size."""
    thr, val = basic_phred_table(model.quality_forward)
    if val.min() < 0 or val.max() >= QROW:
        raise ValueError("phred table out of range")
    t.arrays["basic_thr"] = thr
    t.arrays["basic_val"] = val.astype(np.uint8)
    t.scalars.update(qual_mode=2, n_basic=len(thr), n_insert=0, fixed_insert=int(model.insert_size))


def _lower_quality_perfect(model, t):
    """perfect.py:36-52: constant phred, constant insert size."""
    t.scalars.update(qual_mode=1, const_phred=int(model.quality_forward), n_insert=0, fixed_insert=int(model.insert_size))


class ErrorModel(object):
    """Base class: the attribute contract + lowering to device tables."""

    fixed_insert = None  # basic / perfect: constant insert size (basic.py:19, perfect.py:18)

    @property
    def logger(self):
        component = "{}.{}".format(type(self).__module__, type(self).__name__)
        return logging.getLogger(component)

    def load_npz(self, npz_path, model):
        """Same behaviour as the reference (error_models/__init__.py:27-50): log + exit(1) on failure."""
        import _pickle

        try:
            error_profile = np.load(npz_path, allow_pickle=True)
            assert error_profile["model"] == model
        except (OSError, IOError, EOFError, _pickle.UnpicklingError) as e:
            self.logger.error("Failed to read ErrorModel file: %s" % e)
            sys.exit(1)
        except AssertionError:
            self.logger.error("Trying to load a %s ErrorModel in %s mode" % (error_profile["model"], model))
            sys.exit(1)
        else:
            self.logger.debug("Loaded ErrorProfile: %s" % npz_path)
        return error_profile

    def device_tables(self):
        return lower(self)


class KDErrorModel(ErrorModel):
    """KDE error model from an .npz profile (iss/error_models/kde.py:9-50) -- same constructor and attributes."""

    def __init__(self, npz_path, fragment_length=None, fragment_sd=None, store_mutations=False):
        super().__init__()
        self.npz_path = npz_path
        self.error_profile = self.load_npz(npz_path, "kde")
        self.store_mutations = store_mutations
        self.read_length = self.error_profile["read_length"]
        self.i_size_cdf = self.error_profile["insert_size"]
        self.fragment_length = fragment_length
        self.fragment_sd = fragment_sd
        self.mean_forward = self.error_profile["mean_count_forward"]
        self.mean_reverse = self.error_profile["mean_count_reverse"]
        self.quality_forward = self.error_profile["quality_hist_forward"]
        self.quality_reverse = self.error_profile["quality_hist_reverse"]
        self.subst_choices_for = self.error_profile["subst_choices_forward"]
        self.subst_choices_rev = self.error_profile["subst_choices_reverse"]
        self.ins_for = self.error_profile["ins_forward"]
        self.ins_rev = self.error_profile["ins_reverse"]
        self.del_for = self.error_profile["del_forward"]
        self.del_rev = self.error_profile["del_reverse"]
        self.error_profile = ""  # discard the error profile after reading

    def truncate(self, read_length):
        """Shorten the model to its first `read_length` positions (the '150 bp' benchmark models of
        BASELINE.md: read_length=N and every per-position array sliced [:N]; everything else verbatim)."""
        n = int(read_length)
        if not 2 <= n <= int(self.read_length):
            raise ValueError("cannot truncate a %d bp model to %d" % (int(self.read_length), n))
        self.read_length = np.int64(n)
        for attr in ("quality_forward", "quality_reverse"):
            q = getattr(self, attr)
            out = np.empty(len(q), dtype=object)
            for b in range(len(q)):
                out[b] = list(q[b])[:n]
            setattr(self, attr, out)
        for attr in ("subst_choices_for", "subst_choices_rev", "ins_for", "ins_rev", "del_for", "del_rev"):
            setattr(self, attr, getattr(self, attr)[:n])
        return self


def basic_phred_table(mean_quality, sd=0.01, cap=0.9999):
    """BasicErrorModel.gen_phred_scores (basic.py:52-53) as an integer inversion table.

    The reference draws x ~ N(phred_to_prob(mean_quality), sd) per position and maps it through
    phred(x) = prob_to_phred(min(x, cap)) = int(round(-10 log10(1 - min(x, cap)))) (util.py:32-45), which is
    non-decreasing in x.  RNG contract v2 defines the normal draw by inversion on the phred integer:
        phred = val[#{k : w >= thr[k]}],   thr[k] = ceil(P(phred < val[k+1]) * 2**32)
    with P(phred >= v) = P(x >= 1 - 10**(-(v - 0.5)/10)).  Returns (thr uint32[n] ascending, val int64[n+1])."""
    from scipy.stats import norm

    mu = float(1 - 10 ** (-mean_quality / 10))
    vmax = int(round(-10 * np.log10(1 - cap)))
    ts = {}
    v = vmax
    while v > -1000:
        x_v = 1.0 - 10.0 ** (-(v - 0.5) / 10.0)   # smallest probability whose phred is >= v
        t = float(norm.cdf(x_v, loc=mu, scale=sd))  # u >= t  <=>  phred >= v
        if t * TWO32 < 1.0:
            break
        ts[v] = t
        v -= 1
    vmin = v
    order = sorted(ts)
    thr = np.array([min(math.ceil(ts[k] * TWO32), 2 ** 32 - 1) for k in order], dtype=np.uint32)
    val = np.array([vmin] + order, dtype=np.int64)
    return thr, val


class BasicErrorModel(ErrorModel):
    """basic.py:10-63 -- fixed L=125, insert 200, phred ~ normal around Q30, uniform substitutions, no indels."""

    def __init__(self, fragment_length=None, fragment_sd=None, store_mutations=False):
        super().__init__()
        self.read_length = 125
        self.insert_size = 200
        self.fragment_length = fragment_length
        self.fragment_sd = fragment_sd
        self.store_mutations = store_mutations
        self.quality_forward = self.quality_reverse = 30
        self.subst_choices_for = self.subst_choices_rev = [
            {
                "A": (["T", "C", "G"], [1 / 3, 1 / 3, 1 / 3]),
                "T": (["A", "C", "G"], [1 / 3, 1 / 3, 1 / 3]),
                "C": (["A", "T", "G"], [1 / 3, 1 / 3, 1 / 3]),
                "G": (["A", "T", "C"], [1 / 3, 1 / 3, 1 / 3]),
            }
            for _ in range(self.read_length)
        ]
        self.ins_for = self.ins_rev = self.del_for = self.del_rev = [
            {"A": 0.0, "T": 0.0, "C": 0.0, "G": 0.0} for _ in range(125)
        ]


class PerfectErrorModel(ErrorModel):
    """perfect.py:7-52 -- Q40 everywhere, identity 'substitutions', no indels."""

    def __init__(self, fragment_length=None, fragment_sd=None):
        super().__init__()
        self.read_length = 125
        self.insert_size = 200
        self.fragment_length = fragment_length
        self.fragment_sd = fragment_sd
        self.store_mutations = False
        self.quality_forward = self.quality_reverse = 40
        self.subst_choices_for = self.subst_choices_rev = [
            {
                "A": (["A", "T", "C", "G"], [1, 0, 0, 0]),
                "T": (["A", "T", "C", "G"], [0, 1, 0, 0]),
                "C": (["A", "T", "C", "G"], [0, 0, 1, 0]),
                "G": (["A", "T", "C", "G"], [0, 0, 0, 1]),
            }
            for _ in range(self.read_length)
        ]
        self.ins_for = self.ins_rev = self.del_for = self.del_rev = [
            {"A": 0.0, "T": 0.0, "C": 0.0, "G": 0.0} for _ in range(self.read_length)
        ]


PRECOMPUTED = {  # iss/generator.py:377-387
    "hiseq": "HiSeq", "novaseq": "NovaSeq", "miseq": "miSeq_0.npz", "miseq-20": "miSeq_20.npz",
    "miseq-24": "miSeq_24.npz", "miseq-28": "miSeq_28.npz", "miseq-32": "miSeq_32.npz",
    "miseq-36": "miSeq_36.npz", "nextseq": "nextSeq.npz",
}


def resolve_model_path(model):
    if model.lower() in PRECOMPUTED:
        return os.path.join(PROFILE_DIR, PRECOMPUTED[model.lower()])
    return model
