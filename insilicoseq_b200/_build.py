"""Build the native library in-tree: nvcc -> insilicoseq_b200/csrc/libiss_b200.so (sm_100a only)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libiss_b200.so")
SOURCES = ["iss_kernels.cu", "iss_deflate.cu", "iss_runtime.cu", "iss_hostmem.cu", "iss_io.cu", "iss_genome.cu", "iss_fasta.cu"]
HEADERS = ["iss_device.cuh", "iss_kernels.h", "iss_deflate.h", "iss_hostmem.h", "iss_genome.h", "iss_fasta.h", os.path.join("..", "..", "include", "iss_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build_native(force=False, extra_flags=(), verbose=False):
    if not force and not is_stale():
        return LIB
    extra_flags = list(extra_flags) + os.environ.get("ISS_NVCC_EXTRA", "").split()
    cmd = [nvcc_path()] + NVCC_FLAGS + list(extra_flags) + ["-o", LIB] + SOURCES + ["-lnccl"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB
