"""FASTA text -> device (the genome file as data, iss/generator.py:270,468-490).  The host part -- finding the
title lines -- is checked here on the CPU against this package's reader (fasta.parse, Bio.SeqIO's semantics); the device
part against register_records in test_gpu_parity.py."""
import ctypes
import io

import numpy as np
import pytest

from insilicoseq_b200 import _lib, fasta

WS = b" \t\n\r\x0b\x0c"

TEXTS = [
    b"",
    b"no title at all\nACGT\n",
    b">a\nACGT\n",
    b">a",
    b">a\n",
    b">only_title_then_another\n>b desc here\nAC\nGT\n",
    b"preamble line\nmore\n>r1 first record\nACGT\nAC GT\n\n>r2\tsecond\r\nNNNN\r\nacgt\r\n>r3\n",
    b">   spaced id   trailing   \nAAAA\n> \nCC\n>x>y z\nG>G\nTT\n",
    b">a\nAC>GT\n>b\n>c\nA\n",
    (b">long\n" + b"ACGT" * 20000 + b"\n") * 3,
    b"\n\n>late\nAC\n",
]


def c_index(text):
    L = _lib.load()
    buf = np.frombuffer(text, dtype=np.uint8) if text else np.zeros(0, dtype=np.uint8)
    recs = ctypes.POINTER(_lib.IssFastaRec)()
    n = ctypes.c_uint32(0)
    _lib.check(L.iss_fasta_index(buf.ctypes.data if len(buf) else None, len(buf), ctypes.byref(recs), ctypes.byref(n)))
    try:
        return [(text[recs[k].title_off:recs[k].title_off + recs[k].title_len], text[recs[k].id_off:recs[k].id_off + recs[k].id_len],
                 text[recs[k].body_off:recs[k].body_off + recs[k].body_len], recs[k].gt_pos) for k in range(n.value)]
    finally:
        L.iss_fasta_free(recs)


@pytest.mark.parametrize("k", range(len(TEXTS)))
def test_index_matches_the_python_reader(k):
    text = TEXTS[k]
    got = c_index(text)
    exp = list(fasta.parse(io.BytesIO(text)))
    assert len(got) == len(exp)
    for (title, rid, body, gt), rec in zip(got, exp):
        assert text[gt:gt + 1] == b">" and (gt == 0 or text[gt - 1:gt] == b"\n")
        assert title.decode() == rec.description
        assert rid.decode() == rec.id
        assert body.translate(None, WS) == rec.seq


def test_index_on_random_text():
    rng = np.random.RandomState(3)
    alphabet = np.frombuffer(b"ACGTNacgt>\n\n\r \t", dtype=np.uint8)
    for _ in range(200):
        text = bytes(rng.choice(alphabet, rng.randint(0, 400)))
        got = c_index(text)
        exp = list(fasta.parse(io.BytesIO(text)))
        assert [(t.decode("ascii", "replace"), b.translate(None, WS)) for t, _, b, _ in got] == [(r.description, r.seq) for r in exp], text


def test_text_sequences_read_the_mapping_on_demand():
    """fasta.TextSeq: what a record's .seq is after Engine.load_fasta -- the base count the device reported, and the bytes
    cut out of the text (white space removed) only when asked for."""
    text = b">a x\nAC GT\r\nNN\n\n>b\nacgt\n"
    (t1, i1, b1, _), (t2, i2, b2, _) = c_index(text)
    s1 = fasta.TextSeq(text, text.index(b1), len(b1), 6)
    s2 = fasta.TextSeq(memoryview(text), text.index(b2), len(b2), 4)
    assert len(s1) == 6 and bytes(s1) == b"ACGTNN" and str(s1) == "ACGTNN" and s1 == "ACGTNN" and s1 == b"ACGTNN"
    assert len(s2) == 4 and bytes(s2) == b"acgt" and s2[1:3] == b"cg"
    rec = fasta.Record(s1, id=i1.decode(), description=t1.decode())
    assert len(rec) == 6 and rec.id == "a" and rec.description == "a x"
    out = io.StringIO()
    assert fasta.write([fasta.Record(bytes(s1), id="a", description="a x")], out) == 1
    assert out.getvalue() == ">a x\nACGTNN\n"
