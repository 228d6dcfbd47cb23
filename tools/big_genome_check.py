#!/usr/bin/env python
"""configs[4]-shaped check on one GPU: a multi-record genome larger than 2^32 bases (64-bit offsets) whose packed form
(> 1 GB) does not fit L2, NovaSeq model; sampled pairs are compared with the oracle and the device-resident and
end-to-end throughputs are printed.  Not part of the test suite (needs ~10 GB of host RAM and ~1 min)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import oracle  # noqa: E402
from insilicoseq_b200 import engine as eng_mod, error_models  # noqa: E402

n_rec = int(sys.argv[1]) if len(sys.argv) > 1 else 9
rec_len = int(sys.argv[2]) if len(sys.argv) > 2 else 500_000_000
pairs = int(sys.argv[3]) if len(sys.argv) > 3 else 20_000_000
t0 = time.time()
rng = np.random.default_rng(7)
lut = np.frombuffer(b"ACGT", dtype=np.uint8)
recs = [("chr%d" % k, lut[rng.integers(0, 4, size=rec_len, dtype=np.uint8)].tobytes()) for k in range(n_rec)]
print("genome: %d records x %d bp = %.2f Gbp generated in %.1fs" % (n_rec, rec_len, n_rec * rec_len / 1e9, time.time() - t0), flush=True)
model_path = os.path.join(error_models.PROFILE_DIR, "NovaSeq")
with eng_mod.Engine(0) as e:
    t0 = time.time()
    e.set_genome(recs)
    print("set_genome (classify + 2-bit pack + upload): %.1fs, device blob %.2f GB" % (time.time() - t0, e.stats()["genome_bytes_device"] / 1e9), flush=True)
    e.set_model(error_models.KDErrorModel(model_path))
    per = pairs // n_rec
    work = [(k, per) for k in range(n_rec)]
    items = eng_mod.worker_items(work, 3, 1234, e.record_lengths, e.read_length)
    # sampled parity: first 300 pairs of the first, middle and last record (last record sits beyond base 2^32)
    om = oracle.OracleModel(model_path)
    key = oracle.make_key(1234, 3)
    got = {}
    off = 0
    pos = 0
    want_idx = {0: 0, n_rec // 2: (n_rec // 2) * per, n_rec - 1: (n_rec - 1) * per}
    buf = []
    for b in e.generate(items, batch_pairs=1 << 18):
        buf.append(b.r1.tobytes() if pos < 3 * per or True else b"")
    r1 = b"".join(buf)
    L = e.read_length
    ok = True
    base_bytes = 0
    for k in range(n_rec):
        nbytes = oracle.fastq_bytes(len(recs[k][0]), 3, L, 0, per)
        if k in want_idx:
            exp, _ = oracle.generate(om, recs[k][1], recs[k][0], key, k * per, 3, 0, 300)
            same = r1[base_bytes:base_bytes + len(exp)] == exp
            print("record %d (bases %.2f-%.2f G): first 300 pairs identical to the oracle: %s" % (k, k * rec_len / 1e9, (k + 1) * rec_len / 1e9, same), flush=True)
            ok = ok and same
        base_bytes += nbytes
    assert base_bytes == len(r1)
    del r1, buf
    for mode, kw in (("device-resident", dict(device_only=True, batch_pairs=2_500_000, n_slots=1)), ("end-to-end", dict(batch_pairs=1 << 18, n_slots=3))):
        for rep in range(2):
            e.sync()
            e.mark(0)
            for b in e.generate(items, **kw):
                pass
            e.mark(1)
            ms = e.elapsed_ms(0, 1)
        print("%s: %.3g pairs/s (%.1f ms for %d pairs)" % (mode, per * n_rec / ms * 1e3, ms, per * n_rec), flush=True)
    print(json.dumps({"ok": ok, "stats": e.stats()}))
sys.exit(0 if ok else 1)
