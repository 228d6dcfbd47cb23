"""Round trip of the GPU gzip on a fixture's FASTQ and on a compressed generate (small; meant to run under compute-sanitizer)."""
import gzip
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import fixtures  # noqa: E402
from gpu_helpers import run_product  # noqa: E402
from insilicoseq_b200.engine import Engine  # noqa: E402

case, exp = fixtures.load_case("miseq150_bench_model")
fq = exp["R1"] * 3 + b"x" * 70000 + bytes(range(256)) * 300
with Engine(0) as e:
    z = e.gzip(fq)
assert gzip.decompress(z) == fq
r1, r2, st = run_product(case, compress=True)
assert gzip.decompress(r1) == exp["R1"] and gzip.decompress(r2) == exp["R2"]
print("gz_check ok: %d -> %d bytes; generate --compress %d + %d bytes" % (len(fq), len(z), len(r1), len(r2)))
