"""Python face of the native engine: one `Engine` == one `iss_ctx` == one GPU.

    eng = Engine(device=0)
    eng.set_genome([(record_id, sequence_bytes), ...])
    eng.set_model(KDErrorModel(path))
    for batch in eng.generate(items):          # finished FASTQ text, in submit order
        f1.write(batch.r1); f2.write(batch.r2)  # views into the pinned ring slot, valid until the next step
"""
import ctypes
import logging
import os

import numpy as np

from . import _lib

MASK64 = 0xFFFFFFFFFFFFFFFF


def philox_key(seed, cpu_number):
    """Philox key of a worker: the reference seeds worker c with seed + c (iss/generator.py:234-236)."""
    return (int(seed) + int(cpu_number)) & MASK64


class Batch(object):
    __slots__ = ("r1", "r2", "vcf", "n_pairs", "nbytes", "nbytes2", "first_item", "index", "kernel_ms", "_c")

    def __init__(self, c, device_only):
        self._c = c
        self.nbytes = int(c.nbytes)
        self.nbytes2 = int(c.nbytes2)
        self.n_pairs = int(c.n_pairs)
        self.first_item = int(c.first_item)
        self.index = int(c.batch_index)
        self.kernel_ms = float(c.kernel_ms)
        self.vcf = None
        if device_only or not self.nbytes:
            self.r1 = self.r2 = None
        else:
            self.r1 = np.ctypeslib.as_array(ctypes.cast(c.r1, _lib.u8p), shape=(self.nbytes,))
            self.r2 = np.ctypeslib.as_array(ctypes.cast(c.r2, _lib.u8p), shape=(self.nbytes2,))
        if c.vcf and c.vcf_bytes:
            self.vcf = np.ctypeslib.as_array(ctypes.cast(c.vcf, _lib.u8p), shape=(int(c.vcf_bytes),))


class Engine(object):
    def __init__(self, device=0):
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(self._lib.iss_ctx_create(int(device), ctypes.byref(h)))
        self._h = h
        self.device = int(device)
        self.read_length = None
        self.record_names = []
        self.record_lengths = []
        self._records = {}  # id(record object) -> (device index, the object itself: the strong reference keeps the id unique)
        self.logger = logging.getLogger(__name__)

    # -- lifecycle ------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.iss_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    # -- inputs ---------------------------------------------------------------------------------------
    def set_genome(self, records):
        """records: iterable of (record_id, sequence) with sequence as bytes/str (ASCII, case preserved)."""
        recs = []
        keep = []
        for name, seq in records:
            if not isinstance(seq, (str, bytes, bytearray, memoryview, np.ndarray)):
                seq = str(seq)  # e.g. Bio.Seq.Seq of a reference-side SeqRecord
            if isinstance(seq, str):
                seq = seq.encode("ascii")
            if isinstance(name, str):
                name = name.encode()
            buf = np.frombuffer(seq, dtype=np.uint8)
            keep.append((buf, name))
            recs.append((buf.ctypes.data if len(buf) else None, len(buf), name, len(name)))
        arr = (_lib.IssRecord * max(len(recs), 1))()
        for k, (ptr, ln, name, nl) in enumerate(recs):
            arr[k].seq = ptr
            arr[k].len = ln
            arr[k].name = name
            arr[k].name_len = nl
        _lib.check(self._lib.iss_ctx_set_genome(self._h, arr, len(recs)))
        self.record_names = [n.decode() for _, n in keep]
        self.record_lengths = [len(b) for b, _ in keep]
        self._records = {}

    def load_fasta(self, path):
        """Genome file -> device, the text itself: the host finds the title lines (iss_fasta_index, a memchr scan), the
        GPU blanks them, squeezes the white space out and packs the bases (iss_ctx_set_genome_fasta).  Returns the
        records as fasta.Record objects (id, description, len()) whose .seq reads the mapped file only on demand, and
        registers them like register_records does.  Replaces list(SeqIO.parse(genome_file, "fasta")) + the upload."""
        import mmap

        from . import fasta

        with open(path, "rb") as fh:
            try:
                mm = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)
            except ValueError:  # empty file
                mm = None
        if mm is None:
            self.set_genome([])
            self._records = {}
            return []
        text = np.frombuffer(mm, dtype=np.uint8)
        recs_p = ctypes.POINTER(_lib.IssFastaRec)()
        n = ctypes.c_uint32(0)
        _lib.check(self._lib.iss_fasta_index(text.ctypes.data, len(text), ctypes.byref(recs_p), ctypes.byref(n)))
        try:
            _lib.check(self._lib.iss_ctx_set_genome_fasta(self._h, text.ctypes.data, len(text), recs_p, n.value))
            records = []
            for k in range(n.value):
                f = recs_p[k]
                title = bytes(mm[f.title_off:f.title_off + f.title_len]).decode("ascii", errors="replace")
                rid = bytes(mm[f.id_off:f.id_off + f.id_len]).decode("ascii", errors="replace")
                records.append(fasta.Record(fasta.TextSeq(mm, f.body_off, f.body_len, f.seq_len), id=rid, description=title))
        finally:
            self._lib.iss_fasta_free(recs_p)
        self._fasta_map = (mm, text)  # keeps the mapping alive for the lazy .seq of the records
        self.record_names = [r.id for r in records]
        self.record_lengths = [len(r) for r in records]
        self._records = {id(r): (k, r) for k, r in enumerate(records)}
        return records

    def register_records(self, records):
        """Upload record objects (anything with .id and .seq) and remember each one's device index INSIDE the engine:
        caller-owned records are not modified, and the same record may be registered with several engines."""
        records = list(records)
        self.set_genome([(r.id, r.seq) for r in records])
        self._records = {id(r): (k, r) for k, r in enumerate(records)}
        return records

    def adopt_records(self, other):
        """Share another engine's record map (the genome arrived by ncclBroadcast, same indices)."""
        self._records = dict(other._records)

    def record_index(self, record):
        """Device index of a record passed to register_records (an int is taken as the index itself)."""
        if isinstance(record, (int, np.integer)):
            return int(record)
        hit = self._records.get(id(record))
        if hit is None or hit[1] is not record:
            raise KeyError("record %r was not registered with this engine" % getattr(record, "id", record))
        return hit[0]

    def set_model(self, model):
        """model: an ErrorModel of this package, or a reference-side instance (iss.error_models.*) -- lowered from its
        attributes by error_models.lower, which refuses classes that override the hot-path methods."""
        from . import error_models

        tables = error_models.lower(model)
        cm = tables.c_struct()
        _lib.check(self._lib.iss_ctx_set_model(self._h, ctypes.byref(cm)))
        self.read_length = int(model.read_length)
        self._tables = tables

    # -- multi-GPU --------------------------------------------------------------------------------------
    @staticmethod
    def unique_id():
        buf = (ctypes.c_uint8 * _lib.ISS_UNIQUE_ID_BYTES)()
        _lib.check(_lib.load().iss_comm_unique_id(buf))
        return bytes(buf)

    def comm_init_rank(self, nranks, rank, unique_id):
        buf = (ctypes.c_uint8 * _lib.ISS_UNIQUE_ID_BYTES).from_buffer_copy(unique_id)
        _lib.check(self._lib.iss_comm_init_rank(self._h, int(nranks), int(rank), buf))

    def bcast(self, root=0, read_length=None, record_names=None, record_lengths=None):
        """ncclBroadcast of the genome + model blobs from `root`.  Host-side metadata (names, lengths, L) that a
        non-root rank needs for planning travels with the caller (it is tiny) -- see parallel.py."""
        _lib.check(self._lib.iss_ctx_bcast(self._h, int(root)))
        if read_length is not None:
            self.read_length = int(read_length)
        if record_names is not None:
            self.record_names = list(record_names)
            self.record_lengths = list(record_lengths)

    @staticmethod
    def comm_init_all(engines):
        arr = (ctypes.c_void_p * len(engines))(*[e._h for e in engines])
        _lib.check(_lib.load().iss_comm_init_all(arr, len(engines)))

    @staticmethod
    def bcast_all(engines, root=0):
        arr = (ctypes.c_void_p * len(engines))(*[e._h for e in engines])
        _lib.check(_lib.load().iss_ctx_bcast_all(arr, len(engines), int(root)))
        src = engines[root]
        for e in engines:
            e.read_length = src.read_length
            e.record_names = list(src.record_names)
            e.record_lengths = list(src.record_lengths)

    # -- generation -------------------------------------------------------------------------------------
    @staticmethod
    def _c_items(items):
        items = list(items)
        arr = (_lib.IssWorkItem * max(len(items), 1))()
        for k, it in enumerate(items):
            rec, cpu, n, i0, base, key = it
            arr[k].record = rec
            arr[k].cpu_number = cpu
            arr[k].n_pairs = n
            arr[k].i0 = i0
            arr[k].item_base = base
            arr[k].key = key & MASK64
        return arr, len(items)

    @staticmethod
    def _c_opts(sequence_type, gc_bias, batch_pairs, n_slots, device_only, tma_gather, compress):
        st = {"metagenomics": 0, "amplicon": 1}
        if sequence_type not in st:
            # generator.py:139 raises RuntimeError for the forward read of an unknown sequence type
            raise RuntimeError("sequence type '%s' is not supported" % sequence_type)
        return _lib.IssGenOpts(st[sequence_type], int(bool(gc_bias)), int(batch_pairs), int(n_slots), int(bool(device_only)),
                               int(bool(tma_gather)), int(bool(compress)))

    def submit(self, items, sequence_type="metagenomics", gc_bias=False, batch_pairs=0, n_slots=0, device_only=False,
               tma_gather=False, compress=False):
        """items: iterable of tuples (record, cpu_number, n_pairs, i0, item_base, key)."""
        arr, n = self._c_items(items)
        opts = self._c_opts(sequence_type, gc_bias, batch_pairs, n_slots, device_only, tma_gather, compress)
        self._device_only = bool(device_only)
        _lib.check(self._lib.iss_ctx_submit(self._h, arr, n, ctypes.byref(opts)))

    def abort(self):
        """Drop the rest of the current submit (waits for the batches in flight); the engine stays usable."""
        _lib.check(self._lib.iss_ctx_abort(self._h))

    def write_files(self, items, path_r1, path_r2, path_vcf=None, offset_r1=0, offset_r2=0, append=False, truncate=False,
                    writer_threads=0, sequence_type="metagenomics", gc_bias=False, batch_pairs=0, n_slots=0,
                    compress=False):
        """Generate `items` and write the FASTQ (and VCF) bytes to files with the library's native writer threads:
        every batch goes from its pinned ring slot to the file by pwrite(2); no Python object holds read bytes and the
        GIL is released for the whole call.  append=False places this call's bytes at [offset_r1, ...) / [offset_r2, ...)
        so several workers can fill ONE final file at disjoint, precomputed offsets (`fastq_bytes`).  A path of None
        discards that stream after the copy to the host.  Returns the iss_write_stats fields as a dict."""
        arr, n = self._c_items(items)
        opts = self._c_opts(sequence_type, gc_bias, batch_pairs, n_slots, False, False, compress)
        enc = lambda q: None if q is None else os.fsencode(q)  # noqa: E731
        sink = _lib.IssFileSink(enc(path_r1), enc(path_r2), enc(path_vcf), int(offset_r1), int(offset_r2),
                                int(bool(append)), int(bool(truncate)), int(writer_threads), 0)
        st = _lib.IssWriteStats()
        _lib.check(self._lib.iss_ctx_write_files(self._h, arr, n, ctypes.byref(opts), ctypes.byref(sink), ctypes.byref(st)))
        return {k: getattr(st, k) for k, _ in _lib.IssWriteStats._fields_}

    def fastq_bytes(self, items):
        """Bytes of one FASTQ stream (R1 == R2) that `items` produce: the closed form the planner uses on the device."""
        total = 0
        for rec, cpu, n, i0, _base, _key in items:
            total += int(self._lib.iss_fastq_bytes(len(self.record_names[rec].encode()), int(cpu), int(self.read_length),
                                                   int(i0), int(i0) + int(n)))
        return total

    def batches(self):
        b = _lib.IssBatch()
        have = ctypes.c_int(0)
        while True:
            _lib.check(self._lib.iss_ctx_next(self._h, ctypes.byref(b), ctypes.byref(have)))
            if not have.value:
                return
            try:
                yield Batch(b, self._device_only)
            finally:
                _lib.check(self._lib.iss_ctx_release(self._h, ctypes.byref(b)))

    def generate(self, items, **kw):
        self.submit(items, **kw)
        return self.batches()

    def stats(self):
        s = _lib.IssStats()
        _lib.check(self._lib.iss_ctx_stats(self._h, ctypes.byref(s)))
        return {k: getattr(s, k) for k, _ in _lib.IssStats._fields_}

    def sync(self):
        _lib.check(self._lib.iss_ctx_sync(self._h))

    def mark(self, idx):
        _lib.check(self._lib.iss_ctx_mark(self._h, int(idx)))

    def elapsed_ms(self, a, b):
        ms = ctypes.c_float(0)
        _lib.check(self._lib.iss_ctx_elapsed_ms(self._h, int(a), int(b), ctypes.byref(ms)))
        return float(ms.value)

    def gzip(self, data):
        """gzip bytes on the GPU (RFC 1952; the member format of `compress=True` batches)."""
        data = bytes(data)
        cap = len(data) + len(data) // 2048 + 256
        out = np.empty(cap, dtype=np.uint8)
        n = ctypes.c_uint64(0)
        src = np.frombuffer(data, dtype=np.uint8) if data else np.zeros(1, dtype=np.uint8)
        _lib.check(self._lib.iss_ctx_gzip(self._h, src.ctypes.data, len(data), out.ctypes.data, cap, ctypes.byref(n)))
        return out[: n.value].tobytes()

    def philox(self, ctr, key):
        c = (ctypes.c_uint32 * 4)(*ctr)
        k = (ctypes.c_uint32 * 2)(*key)
        out = (ctypes.c_uint32 * 4)()
        _lib.check(self._lib.iss_ctx_philox(self._h, c, k, out))
        return tuple(int(x) for x in out)


def worker_items(work, cpu_number, seed, record_lengths, read_length, logger=None, record_names=None):
    """Work list of ONE worker [(record_index, n_pairs), ...] -> engine items, reproducing worker_iterator's
    bookkeeping (iss/generator.py:223-251): P = (pairs of the preceding items) + i; a record that is not longer
    than the read length is skipped with the reference's two warnings (generator.py:77-80)."""
    key = philox_key(seed, cpu_number)
    base = 0
    items = []
    for rec, n_pairs in work:
        n_pairs = int(n_pairs)
        if not (read_length < record_lengths[rec]):
            if logger is not None and n_pairs > 0:
                name = record_names[rec] if record_names else rec
                logger.warning("%s shorter than read length for this ErrorModel" % name)
                logger.warning("Skipping %s. You will have less reads than specified" % name)
        elif n_pairs > 0:
            items.append((int(rec), int(cpu_number), n_pairs, 0, base, key))
        base += n_pairs
    return items
