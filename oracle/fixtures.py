"""Fixture helpers shared by oracle/gen_golden.py (writer) and tests/ (readers).

TEST INFRASTRUCTURE ONLY.  Deterministic synthetic records and the case.json -> inputs mapping.
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
PROFILES = os.path.join(ROOT, "insilicoseq_b200", "profiles")  # byte-identical copies of iss/profiles/*

import sys
if HERE not in sys.path:
    sys.path.insert(0, HERE)
import philox  # noqa: E402


def synth_genome(seed, n, alphabet="ACGT"):
    """Deterministic synthetic genome: 2 bits per base from Philox4x32-10 words (key=(seed,0x67656e6f),
    counter=(j,0,0,0)), base = alphabet[(word >> 2k) & 3].  Version-independent (no numpy RNG)."""
    nblk = (n + 63) // 64
    j = np.arange(nblk, dtype=np.uint64)
    z = np.zeros(nblk, dtype=np.uint64)
    w = philox.philox4x32_10_np(j, z, z, z, np.full(nblk, seed & 0xFFFFFFFF, dtype=np.uint64),
                                np.full(nblk, 0x67656E6F, dtype=np.uint64))
    words = np.stack(w, axis=1).reshape(-1)  # 4 words per block, 16 bases per word
    shifts = (2 * np.arange(16, dtype=np.uint64))[None, :]
    codes = ((words[:, None] >> shifts) & np.uint64(3)).reshape(-1)[:n].astype(np.int64)
    lut = np.frombuffer(alphabet.encode(), dtype=np.uint8)
    return lut[codes].tobytes().decode()


def sprinkle(seq, seed, frac_lower=0.0, n_runs=0, iupac=""):
    """Add lowercase stretches, N runs and isolated IUPAC codes (deterministic)."""
    s = bytearray(seq.encode())
    rng = np.random.RandomState(seed)  # legacy MT19937 stream: frozen by numpy's compatibility policy
    n = len(s)
    if frac_lower > 0:
        for _ in range(max(1, int(n * frac_lower / 50))):
            a = int(rng.randint(0, n - 50))
            s[a:a + 50] = bytes(s[a:a + 50]).lower()
    for _ in range(n_runs):
        a = int(rng.randint(0, n - 30))
        ln = int(rng.randint(1, 25))
        s[a:a + ln] = b"N" * ln
    for ch in iupac:
        for _ in range(3):
            s[int(rng.randint(0, n))] = ord(ch)
    return s.decode()



def model_path(spec):
    """'profile:<file>' -> bundled profile; anything else is relative to tests/golden/."""
    if spec.startswith("profile:"):
        return os.path.join(PROFILES, spec.split(":", 1)[1])
    return os.path.join(GOLDEN, spec)


def materialise_records(case):
    out = []
    for r in case["records"]:
        if "literal" in r:
            s = r["literal"]
        else:
            s = synth_genome(r["synth_seed"], r["length"])
        if "sprinkle" in r:
            s = sprinkle(s, r["synth_seed"], **r["sprinkle"])
        out.append((r["id"], s))
    return out


def list_cases():
    return sorted(d for d in os.listdir(GOLDEN) if os.path.exists(os.path.join(GOLDEN, d, "case.json")))


def load_case(name):
    cdir = os.path.join(GOLDEN, name)
    case = json.load(open(os.path.join(cdir, "case.json")))
    exp = {}
    for key, fn in (("R1", "R1.fastq"), ("R2", "R2.fastq"), ("vcf", "out.vcf")):
        p = os.path.join(cdir, fn)
        exp[key] = open(p, "rb").read() if os.path.exists(p) else None
    return case, exp
