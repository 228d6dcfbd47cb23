#!/usr/bin/env python
"""A few device-resident launches of one model, for ncu:  python tools/profile_model.py <profile file or bundled name> [pairs] [truncate]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from insilicoseq_b200 import engine as eng_mod, error_models  # noqa: E402

name = sys.argv[1]
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
trunc = int(sys.argv[3]) if len(sys.argv) > 3 else 0
path = os.path.join(error_models.PROFILE_DIR, name) if os.path.exists(os.path.join(error_models.PROFILE_DIR, name)) else error_models.resolve_model_path(name)
rng = np.random.default_rng(20260924)
genome = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=5_000_000, dtype=np.uint8)].tobytes()
with eng_mod.Engine(0) as e:
    e.set_genome([("syn5M", genome)])
    m = error_models.KDErrorModel(path)
    if trunc:
        m.truncate(trunc)
    e.set_model(m)
    items = eng_mod.worker_items([(0, pairs)], 0, 42, e.record_lengths, e.read_length)
    for rep in range(4):
        e.mark(0)
        for _ in e.generate(items, device_only=True, batch_pairs=pairs, n_slots=1):
            pass
        e.mark(1)
        print("rep %d: %.3g pairs/s" % (rep, pairs / e.elapsed_ms(0, 1) * 1e3), flush=True)
    print(e.stats())
