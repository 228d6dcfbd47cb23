"""insilicoseq_b200 -- InSilicoSeq's per-read-pair generation hot path on B200s (sm_100a CUDA behind a C ABI),
with the reference's `iss generate` command line and ErrorModel plugin surface around it."""
from .version import __version__  # noqa: F401
