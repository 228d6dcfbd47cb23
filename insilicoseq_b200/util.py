"""Host-side helpers of `iss generate` around the engine.  What each one has to do is fixed by the reference's
observable behaviour (iss/util.py:95-268: messages, exit codes, file results); how it is done is this package's own.

File stitching in particular does not walk the data through Python: `concatenate` hands whole files to the kernel
(`os.copy_file_range`, falling back to `os.sendfile`, then to a plain read/write loop on file systems that support
neither), and for plain FASTQ the CLI does not call it at all -- the GPUs write the final files in place."""
import gzip
import logging
import os
import random
import sys

from . import fasta

_LOG = logging.getLogger(__name__)
_SUFFIX_EXPONENT = {"k": 3, "m": 6, "g": 9}


def _die(*lines):
    for ln in lines:
        _LOG.error(ln)
    sys.exit(1)


def count_records(fasta_file):
    """Ids of the records in a FASTA file; an empty list is fatal (util.py:95-116)."""
    ids = [rec.id for rec in fasta.parse(fasta_file)]
    if not ids:
        _die("Failed to find records in genome(s) file:%s" % fasta_file)
    return ids


def convert_n_reads(unit):
    """'1000' -> 1000, '0.5M' -> 500000, suffixes k/m/g in either case (util.py:137-161)."""
    text = str(unit)
    exponent = _SUFFIX_EXPONENT.get(text[-1:].lower())
    try:
        if exponent is None:
            return int(text)
        return int(float(text[:-1]) * 10 ** exponent)
    except ValueError:
        _die("%s is not a valid number of reads" % text)


def genome_file_exists(filename):
    """Refuse to overwrite an existing temporary genome file (util.py:164-176)."""
    if os.path.exists(filename):
        _die("%s already exists. Aborting." % filename, "Maybe use another --output prefix")


def reservoir(records, record_list, n=None):
    """Random subset of n records, yielded in file order (util.py:179-210).  Like the reference it samples the indices
    from range(total - 1), so the last record of the file is never part of a subset."""
    if n is None:
        yield from records
        return
    total = len(record_list)
    if n >= total:
        _die("-u should be strictly smaller than total number of records.")
    random.seed()
    keep = frozenset(random.sample(range(total - 1), n))
    for index, rec in enumerate(records):
        if index in keep:
            yield rec


def _splice(src_fd, dst_fd, length):
    """Move `length` bytes from src_fd to the current position of dst_fd inside the kernel where possible."""
    left = length
    mode = "cfr"
    while left > 0:
        try:
            if mode == "cfr":
                moved = os.copy_file_range(src_fd, dst_fd, min(left, 1 << 30))
            elif mode == "sendfile":
                moved = os.sendfile(dst_fd, src_fd, None, min(left, 1 << 30))
            else:
                chunk = os.read(src_fd, min(left, 8 << 20))
                moved = len(chunk)
                view = memoryview(chunk)
                while view:
                    view = view[os.write(dst_fd, view):]
        except (OSError, AttributeError):
            if mode == "plain":
                raise
            mode = "sendfile" if mode == "cfr" else "plain"
            continue
        if moved == 0:
            if mode == "plain":
                break  # source shorter than its stat size said
            mode = "sendfile" if mode == "cfr" else "plain"
            continue
        left -= moved


def concatenate(file_list, output, header=None):
    """output = [header line] + the files of file_list in order; None entries are skipped (util.py:216-238)."""
    _LOG.info("Stitching input files together")
    try:
        dst = os.open(output, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o666)
    except OSError as e:
        _die("Failed to open output file: %s" % e)
    try:
        if header is not None:
            os.write(dst, (header + "\n").encode())
        for name in file_list:
            if name is None:
                continue
            src = os.open(name, os.O_RDONLY)
            try:
                _splice(src, dst, os.fstat(src).st_size)
            finally:
                os.close(src)
    finally:
        os.close(dst)


def cleanup(file_list):
    """Delete temporary files; a file that cannot be removed is fatal (util.py:241-252)."""
    _LOG.info("Cleaning up")
    for name in file_list:
        if name is None:
            continue
        try:
            os.remove(name)
        except OSError:
            _die("Could not read temporary file: %s" % name, "You may have to remove temporary files manually")


def compress(filename, remove=True):
    """gzip a (small) host-side file such as the VCF; FASTQ is compressed on the GPU (util.py:255-268)."""
    _LOG.info("Compressing %s" % filename)
    target = filename + ".gz"
    with open(filename, "rb") as src, gzip.open(target, "wb") as dst:
        for block in iter(lambda: src.read(4 << 20), b""):
            dst.write(block)
    if remove:
        cleanup([filename])
    return target
