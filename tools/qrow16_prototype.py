"""Host prototype of the NEXT quality-row layout (not used by the product): search the top 16 bits of a row's thresholds,
fall back to the exact 32-bit row only on a 16-bit tie.  Checks on the bundled models that
  #{j : w > r[j]}  ==  #{j : (w >> 16) > (r[j] >> 16)}      whenever (w >> 16) differs from every r[j] >> 16 that decides,
measures how often the tie path is needed, and how many rows fit one 128-byte line (<= 56 live entries).

    python tools/qrow16_prototype.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from insilicoseq_b200 import error_models  # noqa: E402

rng = np.random.default_rng(7)
for name in sorted(os.listdir(error_models.PROFILE_DIR)):
    try:
        model = error_models.KDErrorModel(os.path.join(error_models.PROFILE_DIR, name))
    except SystemExit:
        continue
    rows = np.asarray(model.device_tables().arrays["qual_thr"], dtype=np.uint64).reshape(-1, 64)
    live = rows[(rows != 0xFFFFFFFF).any(axis=1)]
    n_live = (live != 0xFFFFFFFF).sum(axis=1)
    hi = live >> np.uint64(16)
    w = rng.integers(0, 1 << 32, size=(live.shape[0], 64), dtype=np.uint64)
    exact = (w[:, :, None] > live[:, None, :]).sum(axis=2)
    c16 = ((w >> np.uint64(16))[:, :, None] > hi[:, None, :]).sum(axis=2)
    # a tie matters when w16 equals the 16-bit value at the boundary position c16 (rows are sorted)
    at = np.take_along_axis(np.concatenate([hi, np.full((hi.shape[0], 1), 1 << 20, dtype=np.uint64)], axis=1), c16, axis=1)
    tie = at == (w >> np.uint64(16))
    assert ((exact == c16) | tie).all(), name
    wrong_without_fallback = (exact != c16).mean()
    print("%-14s rows %5d  live entries max %2d  one-line rows %5.1f %%  tie rate %.5f  (would be wrong without the fallback: %.5f)"
          % (name, live.shape[0], n_live.max(), 100.0 * (n_live <= 56).mean(), tie.mean(), wrong_without_fallback))
