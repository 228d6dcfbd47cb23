"""Device timeline of the --compress path (ISS_TRACE=1): python tools/gz_trace.py  -> stderr lines per batch."""
import os, sys, time
os.environ["ISS_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from insilicoseq_b200 import engine as eng_mod, error_models
rng = np.random.default_rng(20260924)
genome = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=5_000_000, dtype=np.uint8)].tobytes()
with eng_mod.Engine(0) as e:
    e.set_genome([("syn5M", genome)])
    e.set_model(error_models.KDErrorModel(os.path.join(error_models.PROFILE_DIR, "miSeq_0.npz")).truncate(150))
    items = eng_mod.worker_items([(0, 24 << 18)], 0, 42, e.record_lengths, e.read_length)
    for dev_only in (True, False):
        for rep in range(2):
            e.mark(0)
            t = time.time()
            n = sum(b.nbytes + b.nbytes2 for b in e.generate(items, batch_pairs=1 << 18, compress=True, device_only=dev_only))
            e.mark(1)
            print("device_only", dev_only, "rep", rep, "ms", e.elapsed_ms(0, 1), "host s", time.time() - t, "bytes", n, file=sys.stderr)
