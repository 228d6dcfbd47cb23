import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from insilicoseq_b200 import engine as eng_mod, error_models
rng = np.random.default_rng(20260924)
genome = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=5_000_000, dtype=np.uint8)].tobytes()
with eng_mod.Engine(0) as e:
    e.set_genome([("syn5M", genome)])
    e.set_model(error_models.KDErrorModel(os.path.join(error_models.PROFILE_DIR, "miSeq_0.npz")).truncate(150))
    items = eng_mod.worker_items([(0, 1 << 20)], 0, 42, e.record_lengths, e.read_length)
    for rep in range(3):
        n = sum(b.nbytes + b.nbytes2 for b in e.generate(items, batch_pairs=1 << 18, compress=True))
    print("compressed bytes", n)
