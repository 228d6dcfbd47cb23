#!/usr/bin/env python
"""Device-resident and end-to-end throughput of every bundled error model at its native read length
(5 Mbp synthetic genome, 4 M pairs per step) -- the 'native L' rows next to the 150 bp benchmark model."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from insilicoseq_b200 import engine as eng_mod, error_models  # noqa: E402

rng = np.random.default_rng(20260924)
genome = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=5_000_000, dtype=np.uint8)].tobytes()
PAIRS = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
rows = []
with eng_mod.Engine(0) as e:
    e.set_genome([("syn5M", genome)])
    models = [("hiseq", None), ("novaseq", None), ("miseq", None), ("miseq", 150), ("nextseq", None), ("MiSeq(legacy file)", None),
              ("basic", None), ("perfect", None)]
    for name, trunc in models:
        if name == "basic":
            m = error_models.BasicErrorModel()
        elif name == "perfect":
            m = error_models.PerfectErrorModel()
        elif name.startswith("MiSeq(legacy"):
            m = error_models.KDErrorModel(os.path.join(error_models.PROFILE_DIR, "MiSeq"))
        else:
            m = error_models.KDErrorModel(error_models.resolve_model_path(name))
        if trunc:
            m.truncate(trunc)
        e.set_model(m)
        items = eng_mod.worker_items([(0, PAIRS)], 0, 42, e.record_lengths, e.read_length)
        res = {}
        for mode, kw in (("device", dict(device_only=True, batch_pairs=2_000_000, n_slots=1)), ("e2e", dict(batch_pairs=1 << 18, n_slots=3))):
            for rep in range(2):
                s0 = e.stats()
                e.sync()
                e.mark(0)
                for b in e.generate(items, **kw):
                    pass
                e.mark(1)
                ms = e.elapsed_ms(0, 1)
                s1 = e.stats()
            res[mode] = PAIRS / ms * 1e3
            res[mode + "_slow_reads"] = s1["slow_reads"] - s0["slow_reads"]
        row = {"model": name + ("[:%d]" % trunc if trunc else ""), "L": int(m.read_length), "device_pairs_per_s": res["device"],
               "e2e_pairs_per_s": res["e2e"], "slow_reads_per_Mpair": res["device_slow_reads"] / PAIRS * 1e6}
        rows.append(row)
        print(json.dumps(row), flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "model_sweep.json"), "w"), indent=1)
