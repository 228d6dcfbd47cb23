"""Minimal pure-Python stand-in for the handful of Biopython names that
InSilicoSeq imports (Seq, MutableSeq, SeqRecord, SeqIO.parse/write,
SeqUtils.gc_fraction, Entrez).

TEST INFRASTRUCTURE ONLY.  Biopython is not installed in this image and there
is no network; the reference (`/root/reference/iss`) imports it at module
import time (iss/generator.py:11-14, iss/error_models/__init__.py:10), so the
oracle tooling under `oracle/` puts this directory on sys.path *ahead of* the
unmodified reference.  Biopython supplies only containers and text I/O on the
hot path -- all arithmetic lives in the reference's own code -- so this
stand-in implements container/IO semantics only (Biopython 1.81 behaviour for
the methods the reference calls).  Never imported by the product package.
"""
__version__ = "0.0-standin"
