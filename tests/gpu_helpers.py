"""Shared helpers of the GPU parity tests: run one fixture case through the product (Engine -> C ABI -> CUDA)."""
import fixtures
from insilicoseq_b200 import engine as eng_mod
from insilicoseq_b200 import error_models


def product_model(case):
    fl, fsd = case.get("fragment_length"), case.get("fragment_sd")
    if case["model"] == "basic":
        return error_models.BasicErrorModel(fl, fsd, bool(case.get("store_mutations")))
    if case["model"] == "perfect":
        return error_models.PerfectErrorModel(fl, fsd)
    return error_models.KDErrorModel(fixtures.model_path(case["model"]), case.get("fragment_length"),
                                     case.get("fragment_sd"), bool(case.get("store_mutations")))


def run_product(case, device=0, batch_pairs=0, engine=None, model=None, tma_gather=False, compress=False):
    """Returns (R1 bytes, R2 bytes, stats) of one worker, reference worker_iterator semantics."""
    recs = fixtures.materialise_records(case)
    own = engine is None
    e = engine or eng_mod.Engine(device)
    try:
        e.set_genome([(name, seq.encode()) for name, seq in recs])
        m = model or product_model(case)
        e.set_model(m)
        items = eng_mod.worker_items(case["work"], case["cpu"], case["seed"], e.record_lengths, e.read_length)
        r1, r2, vcf = [], [], []
        for b in e.generate(items, sequence_type=case.get("sequence_type", "metagenomics"),
                            gc_bias=bool(case.get("gc_bias")), batch_pairs=batch_pairs, tma_gather=tma_gather, compress=compress):
            r1.append(b.r1.tobytes())
            r2.append(b.r2.tobytes())
            if b.vcf is not None:
                vcf.append(b.vcf.tobytes())
        st = e.stats()
        st["vcf"] = b"".join(vcf)
        return b"".join(r1), b"".join(r2), st
    finally:
        if own:
            e.close()


def first_diff(a, b):
    n = min(len(a), len(b))
    for k in range(n):
        if a[k] != b[k]:
            lo = a.rfind(b"@", 0, k + 1)
            return k, a[max(lo, 0):k + 80], b[max(lo, 0):k + 80]
    if len(a) != len(b):
        return n, a[n - 40:n + 40], b[n - 40:n + 40]
    return None
