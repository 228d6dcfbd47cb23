"""Summary statistics of a paired FASTQ sample against its source genome (test infrastructure): used to compare the
GPU output with the UNMODIFIED reference driven by its own MT19937 streams (no Philox patch) -- distributions, not bytes."""
import numpy as np


def fastq_records(buf):
    lines = buf.split(b"\n")
    return lines[1::4], lines[3::4]


def revcomp(s):
    return s.translate(bytes.maketrans(b"ACGT", b"TGCA"))[::-1]


def stats(r1, r2, genome, L):
    s1, q1 = fastq_records(r1)
    s2, q2 = fastq_records(r2)
    n = min(len(s1), len(s2))
    s1, q1, s2, q2 = s1[:n], q1[:n], s2[:n], q2[:n]
    out = {"n_pairs": n}
    for tag, qs in (("fwd", q1), ("rev", q2)):
        a = np.frombuffer(b"".join(qs), dtype=np.uint8).reshape(n, L).astype(np.int64) - 33
        out["mean_q_" + tag] = a.mean(axis=0).tolist()
        out["q_hist_" + tag] = np.bincount(a.ravel(), minlength=42)[:42].tolist()
        out["frac_q_lt20_" + tag] = float((a < 20).mean())
    exact1 = exact2 = 0
    inserts = []
    for a, b in zip(s1, s2):
        p1 = genome.find(a)
        p2 = genome.find(revcomp(b))
        exact1 += p1 >= 0
        exact2 += p2 >= 0
        if p1 >= 0 and p2 >= 0:
            inserts.append(p2 - (p1 + L))
    out["frac_exact_fwd"] = exact1 / n
    out["frac_exact_rev"] = exact2 / n
    ins = np.array(inserts)
    out["n_insert"] = int(len(ins))
    out["insert_mean"] = float(ins.mean())
    out["insert_sd"] = float(ins.std())
    out["insert_quantiles"] = np.quantile(ins, [0.05, 0.25, 0.5, 0.75, 0.95]).tolist()
    comp = np.bincount(np.frombuffer(b"".join(s1), dtype=np.uint8), minlength=128)
    out["base_comp_fwd"] = {c: int(comp[ord(c)]) for c in "ACGT"}
    return out
