"""Genome file -> device, two ways: python tools/fasta_load_bench.py [n_records] [bases_per_record]
(a) fasta.parse (mmap + bytes.translate per record) + Engine.register_records, (b) Engine.load_fasta (title lines found by a
memchr scan, the text copied to HBM as it is, white space squeezed out and bases packed by kernels).  Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from insilicoseq_b200 import engine as eng_mod
from insilicoseq_b200 import fasta

n_rec = int(sys.argv[1]) if len(sys.argv) > 1 else 330
per = int(sys.argv[2]) if len(sys.argv) > 2 else 3_000_000
path = "/dev/shm/iss_fasta_bench.fasta" if os.path.isdir("/dev/shm") else "/tmp/iss_fasta_bench.fasta"
rng = np.random.default_rng(5)
t0 = time.time()
with open(path, "wb") as fh:
    for r in range(n_rec):
        seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=per, dtype=np.uint8)]
        lines = np.full((per // 60, 61), 10, dtype=np.uint8)  # 60 columns + '\n', like Bio.SeqIO.write
        lines[:, :60] = seq[:per // 60 * 60].reshape(-1, 60)
        fh.write(b">contig_%d synthetic record %d\n" % (r, r))
        fh.write(lines.tobytes())
        fh.write(seq[per // 60 * 60:].tobytes() + b"\n")
t_write = time.time() - t0
size = os.path.getsize(path)
out = {"file_bytes": size, "records": n_rec, "bases": n_rec * per, "write_s": round(t_write, 3)}
try:
    with eng_mod.Engine(0) as e:
        e.set_genome([("warm", b"ACGT" * 1000)])  # first-launch module load out of the timings
        t = time.time()
        recs = e.load_fasta(path)
        out["load_fasta_s"] = round(time.time() - t, 3)
        out["lengths_ok"] = all(len(r) == per for r in recs) and len(recs) == n_rec
        t = time.time()
        recs = e.load_fasta(path)
        out["load_fasta_again_s"] = round(time.time() - t, 3)
    with eng_mod.Engine(0) as e:
        e.set_genome([("warm", b"ACGT" * 1000)])
        t = time.time()
        recs = list(fasta.parse(path))
        out["python_parse_s"] = round(time.time() - t, 3)
        t = time.time()
        e.register_records(recs)
        out["register_records_s"] = round(time.time() - t, 3)
finally:
    os.unlink(path)
print(json.dumps(out))
