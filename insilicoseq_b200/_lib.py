"""ctypes binding of libiss_b200.so (include/iss_b200.h).  No torch, no fallback: if the native
library is missing or cannot be loaded this module raises -- there is no CPU path in the product."""
import ctypes
import os

from . import _build

u8p = ctypes.POINTER(ctypes.c_uint8)
ISS_QROW = 64
ISS_MAX_ALT = 4
ISS_UNIQUE_ID_BYTES = 128

ISS_OK, ISS_E_CUDA, ISS_E_ARG, ISS_E_STATE, ISS_E_CHAR, ISS_E_SHORT, ISS_E_DEVICE, ISS_E_NOMEM, ISS_E_IO = range(9)


class IssRecord(ctypes.Structure):
    _fields_ = [("seq", ctypes.c_void_p), ("len", ctypes.c_uint64), ("name", ctypes.c_char_p),
                ("name_len", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


class IssFastaRec(ctypes.Structure):
    _fields_ = [("gt_pos", ctypes.c_uint64), ("title_off", ctypes.c_uint64), ("title_len", ctypes.c_uint64),
                ("id_off", ctypes.c_uint64), ("id_len", ctypes.c_uint32), ("reserved", ctypes.c_uint32),
                ("body_off", ctypes.c_uint64), ("body_len", ctypes.c_uint64), ("seq_len", ctypes.c_uint64)]


class IssModel(ctypes.Structure):
    _fields_ = [
        ("read_length", ctypes.c_uint32), ("qual_mode", ctypes.c_uint32), ("n_insert", ctypes.c_uint32),
        ("fixed_insert", ctypes.c_int32), ("frag_mode", ctypes.c_uint32), ("n_frag", ctypes.c_uint32),
        ("frag_lo", ctypes.c_int64), ("const_phred", ctypes.c_uint32), ("n_basic", ctypes.c_uint32),
        ("store_mutations", ctypes.c_uint32), ("reserved", ctypes.c_uint32),
        ("frag_thr", ctypes.c_void_p), ("insert_thr", ctypes.c_void_p), ("bin_thr", ctypes.c_void_p),
        ("bin_valid", ctypes.c_void_p), ("qual_thr", ctypes.c_void_p), ("err_thr", ctypes.c_void_p),
        ("subst_n", ctypes.c_void_p), ("subst_thr", ctypes.c_void_p), ("subst_alt", ctypes.c_void_p),
        ("ins_thr", ctypes.c_void_p), ("ins_nz", ctypes.c_void_p), ("ins_base", ctypes.c_void_p),
        ("del_thr", ctypes.c_void_p), ("del_nz", ctypes.c_void_p), ("basic_thr", ctypes.c_void_p),
        ("basic_val", ctypes.c_void_p),
    ]


class IssWorkItem(ctypes.Structure):
    _fields_ = [("record", ctypes.c_uint32), ("cpu_number", ctypes.c_uint32), ("n_pairs", ctypes.c_uint64),
                ("i0", ctypes.c_uint64), ("item_base", ctypes.c_uint64), ("key", ctypes.c_uint64)]


class IssGenOpts(ctypes.Structure):
    _fields_ = [("sequence_type", ctypes.c_uint32), ("gc_bias", ctypes.c_uint32), ("batch_pairs", ctypes.c_uint64),
                ("n_slots", ctypes.c_uint32), ("device_only", ctypes.c_uint32), ("tma_gather", ctypes.c_uint32),
                ("compress", ctypes.c_uint32)]


class IssBatch(ctypes.Structure):
    _fields_ = [("r1", ctypes.c_void_p), ("r2", ctypes.c_void_p), ("nbytes", ctypes.c_uint64), ("nbytes2", ctypes.c_uint64),
                ("vcf", ctypes.c_void_p), ("vcf_bytes", ctypes.c_uint64), ("n_pairs", ctypes.c_uint64),
                ("first_item", ctypes.c_uint64), ("batch_index", ctypes.c_uint64), ("slot", ctypes.c_uint32),
                ("kernel_ms", ctypes.c_float)]


class IssStats(ctypes.Structure):
    _fields_ = [("pairs", ctypes.c_uint64), ("fastq_bytes", ctypes.c_uint64), ("d2h_bytes", ctypes.c_uint64),
                ("h2d_bytes", ctypes.c_uint64), ("kernel_launches", ctypes.c_uint64), ("kernel_ms", ctypes.c_double),
                ("slow_reads", ctypes.c_uint64), ("subst_errors", ctypes.c_uint64),
                ("genome_bytes_device", ctypes.c_uint64), ("model_bytes_device", ctypes.c_uint64),
                ("ring_numa_node", ctypes.c_int32), ("ring_alloc_mode", ctypes.c_uint32)]


class IssFileSink(ctypes.Structure):
    _fields_ = [("path_r1", ctypes.c_char_p), ("path_r2", ctypes.c_char_p), ("path_vcf", ctypes.c_char_p),
                ("offset_r1", ctypes.c_uint64), ("offset_r2", ctypes.c_uint64), ("append", ctypes.c_uint32),
                ("truncate", ctypes.c_uint32), ("writer_threads", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


class IssWriteStats(ctypes.Structure):
    _fields_ = [("bytes_r1", ctypes.c_uint64), ("bytes_r2", ctypes.c_uint64), ("bytes_vcf", ctypes.c_uint64),
                ("n_pairs", ctypes.c_uint64), ("n_batches", ctypes.c_uint64), ("wall_ms", ctypes.c_double),
                ("write_ms", ctypes.c_double)]


EXPORTS = {
    "iss_abi_version": (ctypes.c_int, []),
    "iss_last_error": (ctypes.c_char_p, []),
    "iss_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "iss_device_pci_bus_id": (ctypes.c_int, [ctypes.c_int, ctypes.c_char_p, ctypes.c_int]),
    "iss_ctx_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "iss_ctx_destroy": (None, [ctypes.c_void_p]),
    "iss_ctx_set_genome": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(IssRecord), ctypes.c_uint32]),
    "iss_fasta_index": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(ctypes.POINTER(IssFastaRec)),
                                       ctypes.POINTER(ctypes.c_uint32)]),
    "iss_fasta_free": (None, [ctypes.POINTER(IssFastaRec)]),
    "iss_ctx_set_genome_fasta": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(IssFastaRec),
                                                ctypes.c_uint32]),
    "iss_ctx_set_model": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(IssModel)]),
    "iss_comm_unique_id": (ctypes.c_int, [ctypes.c_void_p]),
    "iss_comm_init_rank": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "iss_comm_init_all": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]),
    "iss_ctx_bcast": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "iss_ctx_bcast_all": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int]),
    "iss_ctx_submit": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(IssWorkItem), ctypes.c_uint64,
                                      ctypes.POINTER(IssGenOpts)]),
    "iss_ctx_next": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(IssBatch), ctypes.POINTER(ctypes.c_int)]),
    "iss_ctx_release": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(IssBatch)]),
    "iss_ctx_abort": (ctypes.c_int, [ctypes.c_void_p]),
    "iss_ctx_write_files": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(IssWorkItem), ctypes.c_uint64,
                                           ctypes.POINTER(IssGenOpts), ctypes.POINTER(IssFileSink), ctypes.POINTER(IssWriteStats)]),
    "iss_probe_d2h": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_double,
                                     ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)]),
    "iss_ctx_stats": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(IssStats)]),
    "iss_ctx_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "iss_ctx_mark": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "iss_ctx_elapsed_ms": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]),
    "iss_fastq_bytes": (ctypes.c_uint64, [ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64,
                                          ctypes.c_uint64]),
    "iss_ctx_gzip": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64,
                                    ctypes.POINTER(ctypes.c_uint64)]),
    "iss_ctx_philox": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
}

_lib = None


class NativeError(RuntimeError):
    def __init__(self, code, text):
        RuntimeError.__init__(self, "libiss_b200 error %d: %s" % (code, text))
        self.code = code


def lib_path():
    return _build.LIB


def load(build_if_missing=True):
    """Load the native library (building it in-tree with nvcc when absent and a compiler exists)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path) or (build_if_missing and _build.is_stale() and _have_nvcc()):
        if not build_if_missing:
            raise RuntimeError("native library %s is missing; run __graft_entry__.build()" % path)
        _build.build_native()
    L = ctypes.CDLL(path)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(L, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if L.iss_abi_version() != 1:
        raise RuntimeError("libiss_b200 ABI mismatch")
    _lib = L
    return L


def _have_nvcc():
    try:
        _build.nvcc_path()
        return True
    except RuntimeError:
        return False


def check(rc):
    if rc != ISS_OK:
        raise NativeError(rc, (load().iss_last_error() or b"").decode(errors="replace"))
