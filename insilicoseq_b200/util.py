"""Host-side helpers of `iss generate` (behaviour of iss/util.py:95-268): read counts with k/M/G suffixes,
record counting, reservoir sampling, file concatenation, cleanup, gzip."""
import gzip
import logging
import os
import random
import sys
from shutil import copyfileobj

from . import fasta


def count_records(fasta_file):
    """List of record ids; logs and exits when the file holds none (util.py:95-116)."""
    logger = logging.getLogger(__name__)
    ids = [r.id for r in fasta.parse(fasta_file)]
    if not ids:
        logger.error("Failed to find records in genome(s) file:%s" % fasta_file)
        sys.exit(1)
    return ids


def convert_n_reads(unit):
    """'0.5M' -> 500000 (util.py:137-161)."""
    logger = logging.getLogger(__name__)
    suffixes = {"k": 3, "m": 6, "g": 9}
    unit = str(unit)
    if unit[-1].isdigit():
        try:
            return int(unit)
        except ValueError:
            logger.error("%s is not a valid number of reads" % unit)
            sys.exit(1)
    if unit[-1].lower() in suffixes:
        return int(float(unit[:-1]) * 10 ** suffixes[unit[-1].lower()])
    logger.error("%s is not a valid number of reads" % unit)
    sys.exit(1)


def genome_file_exists(filename):
    logger = logging.getLogger(__name__)
    if os.path.exists(filename):
        logger.error("%s already exists. Aborting." % filename)
        logger.error("Maybe use another --output prefix")
        sys.exit(1)


def reservoir(records, record_list, n=None):
    """Pick n records in file order (util.py:179-210; note its range(0, total - 1) never draws the last record)."""
    logger = logging.getLogger(__name__)
    if n is None:
        for r in records:
            yield r
        return
    total = len(record_list)
    if not n < total:
        logger.error("-u should be strictly smaller than total number of records.")
        sys.exit(1)
    random.seed()
    wanted = set(random.sample(range(0, total - 1), n))
    for k, r in enumerate(records):
        if k in wanted:
            yield r


def concatenate(file_list, output, header=None):
    logger = logging.getLogger(__name__)
    logger.info("Stitching input files together")
    try:
        out_file = open(output, "wb")
    except (IOError, OSError) as e:
        logger.error("Failed to open output file: %s" % e)
        sys.exit(1)
    with out_file:
        if header is not None:
            out_file.write(str.encode(header + "\n"))
        for name in file_list:
            if name is not None:
                with open(name, "rb") as f:
                    copyfileobj(f, out_file)


def cleanup(file_list):
    logger = logging.getLogger(__name__)
    logger.info("Cleaning up")
    for temp_file in file_list:
        if temp_file is not None:
            try:
                os.remove(temp_file)
            except (IOError, OSError):
                logger.error("Could not read temporary file: %s" % temp_file)
                logger.error("You may have to remove temporary files manually")
                sys.exit(1)


def compress(filename, remove=True):
    logger = logging.getLogger(__name__)
    logger.info("Compressing %s" % filename)
    outfile = filename + ".gz"
    with open(filename, "rb") as i, gzip.open(outfile, "wb") as o:
        copyfileobj(i, o)
    if remove:
        cleanup([filename])
    return outfile
