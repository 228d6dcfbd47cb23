"""Host side of `iss generate` around the GPU engine.

Same function names, arguments and error behaviour as the reference's iss/generator.py for everything
on either side of the hot path -- load_error_model (:359), load_genomes (:424),
load_readcount_or_abundance (:493), generate_work_divider (:254), worker_iterator (:223) -- but the
per-read loop (simulate_reads / simulate_read and the ErrorModel methods) is the CUDA engine.
"""
import logging
import os
import random
import sys

import numpy as np

from . import _lib, abundance, error_models, fasta, util
from .engine import Engine, worker_items


def load_error_model(mode, seed, model, fragment_length, fragment_length_sd, store_mutations):
    """iss/generator.py:359-421, including its seeding of the two global host RNGs (they drive the abundance draws)."""
    logger = logging.getLogger(__name__)
    logger.info("Using %s ErrorModel" % mode)
    if fragment_length is not None and fragment_length_sd is not None:
        logger.info(
            f"Using custom fragment length {fragment_length} and default fragment length sd {fragment_length_sd}"
        )
    elif bool(fragment_length) ^ bool(fragment_length_sd):
        logger.error("fragment_length and fragment_length_sd must be specified together")
        sys.exit(1)
    if seed:
        logger.info("Setting random seed to %i" % seed)
        random.seed(seed)
        np.random.seed(seed)
    if mode == "kde":
        if model is None:
            logger.error("--model is required in --mode kde")
            sys.exit(1)
        npz = error_models.resolve_model_path(model)
        return error_models.KDErrorModel(npz, fragment_length, fragment_length_sd, store_mutations)
    if mode == "basic":
        if model is not None:
            logger.warning("--model %s will be ignored in --mode %s" % (model, mode))
        return error_models.BasicErrorModel(fragment_length, fragment_length_sd, store_mutations)
    if mode == "perfect":
        if model is not None:
            logger.warning("--model %s will be ignored in --mode %s" % (model, mode))
        return error_models.PerfectErrorModel(fragment_length, fragment_length_sd)
    logger.error("unknown error model mode %r" % mode)
    sys.exit(1)


def load_genomes(genomes, draft, ncbi, n_genomes_ncbi, output, n_genomes):
    """iss/generator.py:424-490: concatenate the inputs into {output}.iss.tmp.genomes.fasta, optional reservoir pick."""
    logger = logging.getLogger(__name__)
    if not (genomes or draft or ncbi):
        logger.error("One of --genomes/-g, --draft, --ncbi/-k is required")
        sys.exit(1)
    genome_files = []
    if genomes:
        genome_files.extend(genomes)
    if draft:
        genome_files.extend(draft)
    if ncbi:
        # the reference downloads from NCBI here (iss/download.py); there is no network path in this engine
        logger.error("--ncbi needs network access and is not supported by the B200 engine; download the genomes first")
        sys.exit(1)
    genome_file = output + ".iss.tmp.genomes.fasta"
    util.concatenate(genome_files, output=genome_file)
    if n_genomes and not draft and not ncbi:
        genome_count = util.count_records(genome_file)
        picked = list(util.reservoir(fasta.parse(genome_file), genome_count, n_genomes))
        with open(genome_file, "w") as fh:
            fasta.write(picked, fh)
    if os.stat(genome_file).st_size == 0:
        logger.error("Genome(s) file seems empty: %s" % genome_file)
        sys.exit(1)
    try:
        genome_list = util.count_records(genome_file)
    except IOError as e:
        logger.error("Failed to open genome(s) file:%s" % e)
        sys.exit(1)
    return genome_list, genome_file


ABUNDANCE_DISPATCH = {
    "uniform": abundance.uniform,
    "halfnormal": abundance.halfnormal,
    "exponential": abundance.exponential,
    "lognormal": abundance.lognormal,
    "zero_inflated_lognormal": abundance.zero_inflated_lognormal,
}


def load_readcount_or_abundance(readcount_file, abundance_file, coverage_file, coverage, abundance_distribution, draft,
                                genome_list, genome_file, n_reads, output, error_model):
    """iss/generator.py:493-595."""
    logger = logging.getLogger(__name__)
    readcount_dic = None
    abundance_dic = None
    if readcount_file:
        logger.info("Using readcount file:%s" % readcount_file)
        logger.warning("--readcount_file disables --n_reads, n_reads will be calculated from the readcount file")
        if draft:
            raise RuntimeError("readcount_file is only supported using --genomes, not --draft")
        readcount_dic = abundance.parse_readcount_file(readcount_file)
    elif abundance_file:
        logger.info("Using abundance file:%s" % abundance_file)
        if draft:
            short = abundance.parse_abundance_file(abundance_file)
            complete = {k: v for k, v in short.items() if k not in draft}
            abundance_dic = {**complete, **abundance.expand_draft_abundance(short, draft)}
        else:
            abundance_dic = abundance.parse_abundance_file(abundance_file)
    elif coverage_file:
        logger.warning("--coverage_file is an experimental feature")
        logger.warning("--coverage_file disables --n_reads")
        logger.info("Using coverage file:%s" % coverage_file)
        if draft:
            cov = abundance.parse_abundance_file(coverage_file)
            complete = {k: v for k, v in cov.items() if k not in draft}
            abundance_dic = {**complete, **abundance.expand_draft_abundance(cov, draft, mode="coverage")}
        else:
            abundance_dic = abundance.parse_abundance_file(coverage_file)
    elif coverage in ABUNDANCE_DISPATCH:
        logger.info("Using %s coverage distribution" % coverage)
        if draft:
            abundance_dic = abundance.draft(genome_list, draft, ABUNDANCE_DISPATCH[abundance_distribution], output,
                                            mode="coverage")
        else:
            abundance_dic = ABUNDANCE_DISPATCH[coverage](genome_list)
        if n_reads:
            n_reads = util.convert_n_reads(n_reads)
            logger.info("scaling coverage to %s reads" % n_reads)
            abundance_dic = abundance.coverage_scaling(n_reads, abundance_dic, genome_file, error_model.read_length)
        abundance.to_file(abundance_dic, output, mode="coverage")
    elif abundance_distribution in ABUNDANCE_DISPATCH:
        logger.info("Using %s abundance distribution" % abundance_distribution)
        if draft:
            abundance_dic = abundance.draft(genome_list, draft, ABUNDANCE_DISPATCH[abundance_distribution], output)
        else:
            abundance_dic = ABUNDANCE_DISPATCH[abundance_distribution](genome_list)
            abundance.to_file(abundance_dic, output)
    else:
        logger.error("Could not get abundance, or coverage or readcount information")
        sys.exit(1)
    return readcount_dic, abundance_dic


def generate_work_divider(fasta_file, readcount_dic, abundance_dic, n_reads, coverage, coverage_file, error_model,
                          output, chunk_size):
    """iss/generator.py:254-356: per-record pair counts (with the rounding repair) split into chunks of
    `chunk_size` pairs.  Yields lists of (record, n_pairs, mode); `mode` is always 'default' here because
    nothing is pickled (the reference's 'memmap' mode works around a multiprocessing limit, :313-331)."""
    logger = logging.getLogger(__name__)
    current_chunk = 0
    total_reads_generated = 0
    total_reads_generated_unrounded = 0
    chunk_work = []
    for record in fasta_file:
        if readcount_dic is not None:
            if record.id not in readcount_dic:
                logger.warning(f"Record {record.id} not found in readcount file")
                continue
            n_pairs_unrounded = readcount_dic[record.id] / 2
        elif abundance_dic is not None:
            if record.id not in abundance_dic:
                logger.warning(f"Record {record.id} not found in abundance file")
                continue
            record_abundance = abundance_dic[record.id]
            genome_size = len(record.seq)
            if coverage or coverage_file:
                record_coverage = record_abundance
            else:
                record_coverage = abundance.to_coverage(n_reads, record_abundance, error_model.read_length, genome_size)
            n_pairs_unrounded = ((record_coverage * len(record.seq)) / error_model.read_length) / 2
        else:
            raise RuntimeError("No readcount or abundance file provided")
        n_pairs = round(n_pairs_unrounded)
        total_reads_generated_unrounded += n_pairs_unrounded
        total_reads_generated += n_pairs
        if round(total_reads_generated_unrounded) > total_reads_generated:
            logger.debug("Adding a pair to correct rounding error")
            n_pairs += 1
            total_reads_generated += 1
        logger.debug("Will generate %s read pairs for %s" % (n_pairs, record.id))
        if n_pairs == 0:
            continue
        n_pairs_remaining = n_pairs
        while n_pairs_remaining > 0:
            chunk_remaining = chunk_size - current_chunk
            n_pairs_added = n_pairs_remaining if n_pairs_remaining <= chunk_remaining else chunk_remaining
            chunk_work.append((record, n_pairs_added, "default"))
            n_pairs_remaining -= n_pairs_added
            current_chunk += n_pairs_added
            if current_chunk == chunk_size:
                yield chunk_work
                chunk_work = []
                current_chunk = 0
    if chunk_work:
        yield chunk_work


def register_records(engine, records):
    """Upload records to the engine (once).  The record -> device index map lives inside the engine
    (Engine.record_index); the caller's record objects are not touched."""
    return engine.register_records(records)


def worker_iterator(work, error_model, cpu_number, worker_prefix, seed, sequence_type, gc_bias, engine=None, compress=False,
                    sink=None, stop=None):
    """Drop-in for iss/generator.py:223-251: writes {worker_prefix}_R1.fastq, _R2.fastq and .vcf for one work chunk.

    work: list of (record, n_pairs, mode).  `engine` is an Engine that already holds the genome and the model; when
    omitted a private engine on GPU 0 is created for the call.  `compress=True` writes {worker_prefix}_R1.fastq.gz /
    _R2.fastq.gz instead: gzip members produced on the GPU (`--compress`).  The bytes are written by the library's
    native writer threads straight from the pinned ring (Engine.write_files): the reference's SeqIO.write loop
    (generator.py:64-66) has no Python counterpart here.

    sink (optional): dict(path_r1, path_r2, offset_r1, offset_r2) -- write this chunk's FASTQ at those offsets of the
    FINAL files instead of {worker_prefix}_R1/_R2 temp files (plain text only: the sizes are closed-form), which removes
    the parent's concatenation pass (iss/app.py:123-133).  The VCF always goes to {worker_prefix}.vcf.
    stop (optional): threading.Event checked before the chunk starts (cooperative cancellation from the CLI)."""
    logger = logging.getLogger(__name__)
    ext = ".fastq.gz" if compress else ".fastq"
    vcf_path = f"{worker_prefix}.vcf"
    if sink is None:
        paths = (f"{worker_prefix}_R1{ext}", f"{worker_prefix}_R2{ext}")
        offsets = (0, 0)
    else:
        if compress:
            raise ValueError("final-offset writes need plain FASTQ (compressed sizes are not known in advance)")
        paths = (sink["path_r1"], sink["path_r2"])
        offsets = (int(sink["offset_r1"]), int(sink["offset_r2"]))
    try:  # the reference opens its three files before anything else and exits on failure (generator.py:226-232)
        for q in (paths if sink is None else ()) + (vcf_path,):
            open(q, "wb").close()
    except PermissionError as e:
        logger.error("Failed to write temporary output file(s): %s" % e)
        sys.exit(1)
    if sequence_type not in ("metagenomics", "amplicon"):
        raise RuntimeError(f"sequence type '{sequence_type}' is not supported")
    if stop is not None and stop.is_set():
        return None
    own = engine is None
    if own:
        engine = Engine(0)
        uniq = []
        seen = set()
        for rec, _, _ in work:
            if id(rec) not in seen:
                seen.add(id(rec))
                uniq.append(rec)
        engine.register_records(uniq)
        engine.set_model(error_model)
    if seed is None:  # the reference leaves its RNGs unseeded: any stream will do
        seed = random.SystemRandom().getrandbits(63)
    try:
        for rec, n_pairs, _ in work:
            logger.debug("Cpu #%s: Generating %s read pairs" % (cpu_number, n_pairs))
        items = worker_items([(engine.record_index(rec), n) for rec, n, _ in work], cpu_number, seed, engine.record_lengths,
                             engine.read_length, logger=logger, record_names=engine.record_names)
        try:
            return engine.write_files(items, paths[0], paths[1], vcf_path, offset_r1=offsets[0], offset_r2=offsets[1],
                                      append=False, truncate=False, sequence_type=sequence_type, gc_bias=gc_bias,
                                      compress=compress)
        except _lib.NativeError as e:
            if e.code == _lib.ISS_E_IO:  # same convention as the reference for I/O trouble: log, exit(1)
                logger.error("Failed to write temporary output file(s): %s" % e)
                sys.exit(1)
            raise
    finally:
        if own:
            engine.close()


class Read(object):
    """What simulate_read hands back per read: the slice of Bio.SeqRecord the reference's callers use."""

    __slots__ = ("id", "seq", "description", "letter_annotations", "annotations")

    def __init__(self, rid, seq, quals):
        self.id = rid
        self.seq = seq
        self.description = ""
        self.letter_annotations = {"phred_quality": quals}
        self.annotations = {"mutations": []}

    def format(self, fmt="fastq-sanger"):
        return "@%s\n%s\n+\n%s\n" % (self.id, self.seq, "".join(chr(q + 33) for q in self.letter_annotations["phred_quality"]))


_private_engines = []  # [(record, error_model, device, engine)]: strong references, compared by identity


def _engine_for(record, error_model, device=0):
    """One cached engine per (record, model) pair for the single-pair API below."""
    for rec, em, dev, eng in _private_engines:
        if rec is record and em is error_model and dev == device:
            return eng
    if len(_private_engines) >= 4:
        _private_engines.pop(0)[3].close()
    eng = Engine(device)
    eng.register_records([record])
    eng.set_model(error_model)
    _private_engines.append((record, error_model, device, eng))
    return eng


def _parse_pair(r1, r2):
    a = bytes(r1).split(b"\n")
    b = bytes(r2).split(b"\n")
    mk = lambda x: Read(x[0][1:].decode(), x[1].decode(), [c - 33 for c in x[3]])  # noqa: E731
    return mk(a), mk(b)


def simulate_read(record, error_model, i, cpu_number, sequence_type, seed=0, item_base=0, engine=None):
    """iss/generator.py:98-192 for ONE pair, on the GPU: returns (forward, reverse, mutations).

    Same positional arguments as the reference.  The reference draws from the process-global MT19937 state; here the
    draws are the Philox slots of pair `item_base + i` under key `seed + cpu_number` (RNG contract v2), so the call
    is a pure function of its arguments.  A record that is not longer than the read length raises AssertionError
    (generator.py:130), an unknown sequence type RuntimeError (:139)."""
    from .engine import philox_key

    assert int(error_model.read_length) < len(record.seq)
    eng = engine or _engine_for(record, error_model)
    item = (eng.record_index(record), int(cpu_number), 1, int(i), int(item_base), philox_key(seed, cpu_number))
    out = None
    for batch in eng.generate([item], sequence_type=sequence_type):
        out = _parse_pair(batch.r1, batch.r2)
        vcf = bytes(batch.vcf) if batch.vcf is not None else b""
    mutations = []
    for line in vcf.decode().splitlines():
        f = line.split("\t")
        mutations.append({"id": f[0], "position": int(f[1]) - 1, "ref": f[3], "alt": f[4], "quality": f[5]})
    return out[0], out[1], mutations


def simulate_reads(record, error_model, n_pairs, cpu_number, forward_handle, reverse_handle, mutations_handle,
                   sequence_type, gc_bias=False, mode="default", seed=0, item_base=0, engine=None):
    """iss/generator.py:21-66: n_pairs pairs of one record written to the three handles (binary or text)."""
    from .engine import philox_key

    logger = logging.getLogger(__name__)
    logger.debug("Cpu #%s: Generating %s read pairs" % (cpu_number, n_pairs))
    if not int(error_model.read_length) < len(record.seq):
        logger.warning("%s shorter than read length for this ErrorModel" % record.id)
        logger.warning("Skipping %s. You will have less reads than specified" % record.id)
        return
    eng = engine or _engine_for(record, error_model)
    item = (eng.record_index(record), int(cpu_number), int(n_pairs), 0, int(item_base), philox_key(seed, cpu_number))

    def put(handle, view):
        data = bytes(view)
        try:
            handle.write(data)
        except TypeError:
            handle.write(data.decode())

    for batch in eng.generate([item], sequence_type=sequence_type, gc_bias=gc_bias):
        put(forward_handle, batch.r1)
        put(reverse_handle, batch.r2)
        if batch.vcf is not None:
            put(mutations_handle, batch.vcf)
