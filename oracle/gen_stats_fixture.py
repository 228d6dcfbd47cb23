#!/usr/bin/env python
"""tests/golden/stats_<model>.json: statistics of reads produced by the UNMODIFIED reference with its own MT19937
streams (random.seed / np.random.seed as iss/generator.py:234-236 does) -- no Philox patch involved."""
import json
import os
import random
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fixtures  # noqa: E402
import patched_reference as pr  # noqa: E402
import read_stats  # noqa: E402

N_PAIRS = int(sys.argv[1]) if len(sys.argv) > 1 else 12000
ONLY = sys.argv[2].split(",") if len(sys.argv) > 2 else None
# name -> (model spec, fragment_length, fragment_sd).  `basic` and `fragment` exist because those two branches draw from
# np.random.normal in the reference (iss/error_models/basic.py:41-53, iss/generator.py:121-123): the Philox contract defines
# them by inversion on integer tables (error_models.basic_phred_table / fragment_table), a definition shared by the oracle
# and the product -- these fixtures pin it to the UNMODIFIED reference under its own MT19937 streams.
CASES = (("novaseq", "profile:NovaSeq", None, None), ("hiseq", "profile:HiSeq", None, None),
         ("basic", "basic", None, None), ("fragment", "profile:NovaSeq", 400, 30))
for name, spec, frag, frag_sd in CASES:
    if ONLY and name not in ONLY:
        continue
    iss = pr.import_reference()
    from iss import generator
    from iss.error_models import basic, kde

    genome = fixtures.synth_genome(4001, 60000)
    if spec == "basic":
        em = basic.BasicErrorModel(frag, frag_sd, False)
    else:
        em = kde.KDErrorModel(fixtures.model_path(spec), frag, frag_sd, False)
    rec = pr.make_record(genome, "stat_rec")
    t = time.time()
    with tempfile.TemporaryDirectory() as tmp:
        prefix = os.path.join(tmp, "w")
        generator.worker_iterator([(rec, N_PAIRS, "default")], em, 0, prefix, 20260924, "metagenomics", False)  # MT19937
        r1 = open(prefix + "_R1.fastq", "rb").read()
        r2 = open(prefix + "_R2.fastq", "rb").read()
    st = read_stats.stats(r1, r2, genome.encode(), int(em.read_length))
    st.update({"model": spec, "fragment_length": frag, "fragment_sd": frag_sd, "genome_synth_seed": 4001, "genome_len": 60000,
               "rng": "MT19937 (unmodified reference)", "seconds": time.time() - t})
    json.dump(st, open(os.path.join(fixtures.GOLDEN, "stats_%s.json" % name), "w"), indent=1)
    print(name, "pairs", st["n_pairs"], "exact fwd/rev %.4f %.4f" % (st["frac_exact_fwd"], st["frac_exact_rev"]),
          "insert mean %.1f sd %.1f" % (st["insert_mean"], st["insert_sd"]), "%.0fs" % st["seconds"], flush=True)
