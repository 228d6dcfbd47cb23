"""Per-case parity report on the GPU box (not a test: prints every case, never stops at the first failure)."""
import hashlib
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import fixtures  # noqa: E402
from gpu_helpers import first_diff, run_product  # noqa: E402

only = sys.argv[1:] or fixtures.list_cases()
for name in only:
    case, exp = fixtures.load_case(name)
    t = time.time()
    try:
        r1, r2, st = run_product(case)
    except Exception as e:  # noqa: BLE001
        print("%-30s EXC %r" % (name, e))
        traceback.print_exc()
        continue
    ok1 = hashlib.sha256(r1).hexdigest() == case["sha256"]["R1"]
    ok2 = hashlib.sha256(r2).hexdigest() == case["sha256"]["R2"]
    print("%-30s R1 %s R2 %s  (%d B, slow=%d, subst=%d, %.2fs)" % (name, ok1, ok2, len(r1), st["slow_reads"], st["subst_errors"], time.time() - t))
    for tag, got, want in (("R1", r1, exp["R1"]), ("R2", r2, exp["R2"]), ("VCF", st["vcf"], exp["vcf"])):
        if want is not None and got != want:
            d = first_diff(got, want)
            print("   %s first diff at byte %d (len got %d want %d)\n     got : %r\n     want: %r" % (tag, d[0], len(got), len(want), d[1], d[2]))
