"""Distributional fidelity against the UNMODIFIED reference under its own MT19937 streams (fixtures written by
oracle/gen_stats_fixture.py; no Philox patch involved on that side).  Byte parity is defined under the shared Philox
stream; this checks that moving to that stream changed nothing statistically: per-position mean quality, quality
histogram, fraction of error-free reads (substitution + indel rates), insert-size distribution, base composition."""
import json
import math
import os

import numpy as np
import pytest

import fixtures
import oracle
import read_stats


def _check(st, ref, L):
    n, m = st["n_pairs"], ref["n_pairs"]
    for tag in ("fwd", "rev"):
        d = np.abs(np.array(st["mean_q_" + tag]) - np.array(ref["mean_q_" + tag]))
        assert d.max() < 0.6, (tag, float(d.max()))
        p = np.array(st["q_hist_" + tag], dtype=float)
        q = np.array(ref["q_hist_" + tag], dtype=float)
        tv = 0.5 * np.abs(p / p.sum() - q / q.sum()).sum()
        assert tv < 0.006, (tag, tv)
        pe, qe = st["frac_exact_" + tag], ref["frac_exact_" + tag]
        sigma = math.sqrt(qe * (1 - qe) * (1.0 / n + 1.0 / m))
        assert abs(pe - qe) < 5 * sigma, (tag, pe, qe, sigma)
    sd = ref["insert_sd"]
    # (<=: the Basic model's insert size is a constant, sd == 0)
    assert abs(st["insert_mean"] - ref["insert_mean"]) <= 5 * sd * math.sqrt(1.0 / st["n_insert"] + 1.0 / ref["n_insert"])
    assert abs(st["insert_sd"] - sd) <= 0.06 * sd
    assert np.abs(np.array(st["insert_quantiles"]) - np.array(ref["insert_quantiles"])).max() < 30
    tot = sum(st["base_comp_fwd"].values())
    rtot = sum(ref["base_comp_fwd"].values())
    for c in "ACGT":
        assert abs(st["base_comp_fwd"][c] / tot - ref["base_comp_fwd"][c] / rtot) < 0.004


# `basic` and `fragment`: the two branches whose np.random.normal draws RNG contract v2 defines by inversion on integer tables
# (error_models.basic_phred_table / fragment_table, a definition the oracle shares with the product) -- pinned here to the
# unmodified reference under MT19937 (iss/error_models/basic.py:41-53, iss/generator.py:121-123).
STAT_CASES = ["novaseq", "hiseq", "basic", "fragment"]


def _oracle_model(ref):
    import patched_reference as pr

    fl, fsd = ref.get("fragment_length"), ref.get("fragment_sd")
    ft = pr.fragment_table(fl, fsd) if fl is not None else None
    if ref["model"] == "basic":
        return oracle.OracleModel("basic", fl, fsd, False, frag_table=ft, basic_table=pr.basic_phred_table(30))
    return oracle.OracleModel(fixtures.model_path(ref["model"]), fl, fsd, False, frag_table=ft)


@pytest.mark.parametrize("name", STAT_CASES)
def test_cpu_port_matches_unmodified_reference_statistically(name):
    ref = json.load(open(os.path.join(fixtures.GOLDEN, "stats_%s.json" % name)))
    genome = fixtures.synth_genome(ref["genome_synth_seed"], ref["genome_len"]).encode()
    om = _oracle_model(ref)
    r1, r2 = oracle.generate(om, genome, "stat_rec", oracle.make_key(777, 0), 0, 0, 0, 30000)
    _check(read_stats.stats(r1, r2, genome, om.read_length), ref, om.read_length)


@pytest.mark.gpu
@pytest.mark.parametrize("name", STAT_CASES)
def test_gpu_matches_unmodified_reference_statistically(name):
    from insilicoseq_b200 import engine as eng_mod, error_models

    ref = json.load(open(os.path.join(fixtures.GOLDEN, "stats_%s.json" % name)))
    genome = fixtures.synth_genome(ref["genome_synth_seed"], ref["genome_len"]).encode()
    with eng_mod.Engine(0) as e:
        e.set_genome([("stat_rec", genome)])
        if ref["model"] == "basic":
            m = error_models.BasicErrorModel(ref.get("fragment_length"), ref.get("fragment_sd"))
        else:
            m = error_models.KDErrorModel(fixtures.model_path(ref["model"]), ref.get("fragment_length"), ref.get("fragment_sd"))
        e.set_model(m)
        items = eng_mod.worker_items([(0, 30000)], 1, 31337, e.record_lengths, e.read_length)
        r1 = r2 = b""
        for b in e.generate(items):
            r1 += b.r1.tobytes()
            r2 += b.r2.tobytes()
    _check(read_stats.stats(r1, r2, genome, int(m.read_length)), ref, int(m.read_length))
