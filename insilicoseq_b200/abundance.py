"""Abundance / coverage set-up of `iss generate` (behaviour of iss/abundance.py): O(#genomes) host work done once
per run.  The random distributions draw from numpy's legacy global stream exactly as the reference does, so a given
--seed yields the same abundance table."""
import logging
import os
import sys

import numpy as np

from . import fasta


def _parse_two_columns(path, cast):
    logger = logging.getLogger(__name__)
    try:
        if os.stat(path).st_size == 0:
            logger.error("File seems empty: %s" % path)
            sys.exit(1)
        f = open(path, "r")
    except (IOError, OSError) as e:
        logger.error("Failed to open file:%s" % e)
        sys.exit(1)
    out = {}
    with f:
        for line in f:
            fields = line.split()
            try:
                out[fields[0]] = cast(fields[1])
            except IndexError as e:
                logger.error("Failed to read file: %s" % e)
                sys.exit(1)
    return out


def parse_readcount_file(readcount_file):
    return _parse_two_columns(readcount_file, int)


def parse_abundance_file(abundance_file):
    d = _parse_two_columns(abundance_file, float)
    logging.getLogger(__name__).debug("Loaded abundance/coverage file: %s" % abundance_file)
    return d


def _scaled(record_list, dist):
    scaled = dist / sum(dist)
    return {rec: ab for rec, ab in zip(record_list, scaled)}


def uniform(record_list):
    n = len(record_list)
    return {rec: 1 / n for rec in record_list}


def halfnormal(record_list):
    from scipy import stats

    return _scaled(record_list, stats.halfnorm.rvs(loc=0.00, scale=1.00, size=len(record_list)))


def exponential(record_list):
    return _scaled(record_list, np.random.exponential(size=len(record_list)))


def lognormal(record_list):
    return _scaled(record_list, np.random.lognormal(size=len(record_list)))


def zero_inflated_lognormal(record_list):
    from scipy import stats

    n = len(record_list)
    zero_inflated = stats.bernoulli.rvs(p=0.2, size=n)
    return _scaled(record_list, (1 - zero_inflated) * np.random.lognormal(size=n))


def to_coverage(total_n_reads, species_abundance, read_length, genome_size):
    n_reads = total_n_reads * species_abundance
    return (n_reads * read_length) / genome_size


def coverage_scaling(total_n_reads, abundance_dic, genome_file, read_length):
    logger = logging.getLogger(__name__)
    total_reads = 0
    for record in fasta.parse(genome_file):
        if record.id not in abundance_dic:
            logger.error("Fasta record not found in abundance file: %r" % record.id)
            sys.exit(1)
        total_reads += abundance_dic[record.id] * len(record) / read_length / 2
    scale = total_n_reads / total_reads
    for key in abundance_dic:
        abundance_dic[key] *= scale
    return abundance_dic


def to_file(abundance_dic, output, mode="abundance"):
    logger = logging.getLogger(__name__)
    path = output + ("_abundance.txt" if mode == "abundance" else "_coverage.txt")
    try:
        f = open(path, "w")
    except PermissionError as e:
        logger.error("Failed to open output file: %s" % e)
        sys.exit(1)
    with f:
        for record, ab in abundance_dic.items():
            f.write("%s\t%s\n" % (record, ab))


def expand_draft_abundance(abundance_dic, draft, mode="abundance"):
    """A draft genome's abundance is split over its contigs by length; its coverage is copied to each contig."""
    out = {}
    for key, ab in abundance_dic.items():
        if key in draft:
            contigs = list(fasta.parse(key))
            total = sum(len(c) for c in contigs)
            for c in contigs:
                out[c.id] = ab * (len(c) / total) if mode == "abundance" else ab
    return out


def draft(genomes, draft_files, distribution, output, mode="abundance"):
    draft_records = []
    for d in draft_files:
        draft_records.extend(r.id for r in fasta.parse(d))
    genomes = list(set(genomes) - set(draft_records))
    abundance_dic = distribution(genomes + draft_files)
    complete = {k: v for k, v in abundance_dic.items() if k not in draft_files}
    to_file(abundance_dic, output)
    return {**complete, **expand_draft_abundance(abundance_dic, draft_files, mode)}
