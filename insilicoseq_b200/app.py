"""`iss generate` on B200s: the reference's command line (iss/app.py:172-460) with the process pool replaced by GPUs.

Every `generate` flag of the reference is accepted with the same meaning.  `--cpus N` keeps its role of
deciding the work chunks (chunk c carries cpu_number c in read ids and Philox key seed + c), so the
output for a given --cpus is independent of how many GPUs produce it; `--gpus` (new) chooses the devices.
`iss model` (BAM -> .npz, pysam) is a different sub-command and is not part of this engine.
"""
import argparse
import logging
import os
import random
import sys
import threading

from . import fasta, parallel, util
from .engine import Engine, worker_items
from .generator import (generate_work_divider, load_error_model, load_genomes, load_readcount_or_abundance,
                        register_records, worker_iterator)
from .version import __version__


def generate_reads(args):
    """iss/app.py:23-144 with `pool.starmap(worker_iterator, ...)` (:99-106) mapped onto GPU engines."""
    logger = logging.getLogger(__name__)
    logger.debug("iss-b200 version %s" % __version__)
    logger.info("Starting iss generate")
    error_model = load_error_model(args.mode, args.seed, args.model, args.fragment_length, args.fragment_length_sd,
                                   args.store_mutations)
    genome_list, genome_file = load_genomes(args.genomes, args.draft, args.ncbi, args.n_genomes_ncbi, args.output,
                                            args.n_genomes)
    readcount_dic, abundance_dic = load_readcount_or_abundance(
        args.readcount_file, args.abundance_file, args.coverage_file, args.coverage, args.abundance, args.draft,
        genome_list, genome_file, args.n_reads, args.output, error_model)
    if args.store_mutations:
        logger.info(f"Storing inserted sequence errors in {args.output}.vcf")
    logger.info("Using %s cpus for read generation" % args.cpus)
    if readcount_dic is not None:
        n_reads = sum(readcount_dic.values())
    else:
        n_reads = util.convert_n_reads(args.n_reads)
    logger.info("Generating %s reads" % n_reads)

    temp_file_list = [f"{args.output}.iss.tmp.{i}" for i in range(args.cpus)]
    fq = ".fastq.gz" if args.compress else ".fastq"  # --compress: gzip members come off the GPU, they concatenate
    out_r1, out_r2 = args.output + "_R1" + fq, args.output + "_R2" + fq
    engines = []
    threads = []
    stop = threading.Event()  # set on Ctrl-C or when a device thread fails: workers finish their current chunk and return
    in_place = not args.compress  # plain FASTQ: chunk sizes are closed-form, the GPUs write the final files at their offsets
    try:
        # the genome file goes to GPU 0 as text: the host finds the title lines, the device takes the white space out
        # and packs the bases (Engine.load_fasta) -- no per-record Python objects around the sequences
        engines.append(Engine(0))
        records = engines[0].load_fasta(genome_file)
        n_read_pairs = n_reads // 2
        chunk_size = -((n_read_pairs) // -args.cpus)
        logger.debug("Chunk size: %s" % chunk_size)
        work_chunks = list(generate_work_divider(records, readcount_dic, abundance_dic, n_reads, args.coverage,
                                                 args.coverage_file, error_model, args.output, chunk_size))
        # zip(work_chunks, temp_file_list) in the reference drops any chunk beyond --cpus (app.py:104)
        jobs = list(enumerate(zip(work_chunks, temp_file_list)))

        n_dev = _device_count()
        n_gpus = args.gpus if args.gpus else n_dev
        n_gpus = max(1, min(n_gpus, n_dev, max(len(jobs), 1)))
        logger.info("Using %s GPU(s) for read generation" % n_gpus)
        engines.extend(Engine(d) for d in range(1, n_gpus))
        engines[0].set_model(error_model)
        if n_gpus > 1:  # one NCCL broadcast of the packed genome + model tables, then no communication
            Engine.comm_init_all(engines)
            Engine.bcast_all(engines, 0)
            for e in engines[1:]:
                e.adopt_records(engines[0])
        seed = args.seed if args.seed is not None else random.SystemRandom().getrandbits(62)
        sinks = {}
        if in_place:  # byte offset of every chunk in the final files = FASTQ bytes of the chunks before it
            at = 0
            for cpu_number, (work, _prefix) in jobs:
                items = worker_items([(engines[0].record_index(rec), n) for rec, n, _ in work], cpu_number, seed,
                                     engines[0].record_lengths, engines[0].read_length)
                sinks[cpu_number] = dict(path_r1=out_r1, path_r2=out_r2, offset_r1=at, offset_r2=at)
                at += engines[0].fastq_bytes(items)
            for path in (out_r1, out_r2):
                with open(path, "wb") as fh:
                    fh.truncate(at)
        errors = []

        def run_device(d):
            try:
                parallel.bind_to_device_numa(d)
                for cpu_number, (work, prefix) in jobs:
                    if stop.is_set():
                        return
                    if parallel.chunk_owner(cpu_number, n_gpus) == d:
                        worker_iterator(work, error_model, cpu_number, prefix, seed, args.sequence_type, args.gc_bias,
                                        engine=engines[d], compress=args.compress, sink=sinks.get(cpu_number), stop=stop)
            except BaseException as e:  # noqa: BLE001 - re-raised in the parent thread
                errors.append(e)
                stop.set()

        threads = [threading.Thread(target=run_device, args=(d,)) for d in range(n_gpus)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        for cpu_number in range(len(jobs), args.cpus):  # workers that got no chunk still leave (empty) files
            for suffix in (() if in_place else ("_R1" + fq, "_R2" + fq)) + (".vcf",):
                open(temp_file_list[cpu_number] + suffix, "wb").close()
    except KeyboardInterrupt as e:
        logger.error("iss generate interrupted: %s" % e)
        stop.set()
        for t in threads:  # the device threads own their contexts until they return: wait before anything is torn down
            t.join()
        tmp = [t + s for t in temp_file_list for s in ("_R1.fastq", "_R2.fastq", "_R1.fastq.gz", "_R2.fastq.gz", ".vcf")]
        tmp.append(genome_file)
        util.cleanup([t for t in tmp if os.path.exists(t)])
        sys.exit(1)
    else:
        temp_mut = [t + ".vcf" for t in temp_file_list]
        temp_fq = []
        if not in_place:
            temp_R1 = [t + "_R1" + fq for t in temp_file_list]
            temp_R2 = [t + "_R2" + fq for t in temp_file_list]
            util.concatenate(temp_R1, out_r1)
            util.concatenate(temp_R2, out_r2)
            temp_fq = temp_R1 + temp_R2
        if args.store_mutations:
            util.concatenate(temp_mut, args.output + ".vcf",
                             "##fileformat=VCFv4.1\n" + "\t".join(["#CHROM", "POS", "ID", "REF", "ALT", "QUAL", "FILTER", "INFO"]))
        util.cleanup(temp_fq + temp_mut + [genome_file])
        if args.compress and args.store_mutations:  # the FASTQ files were compressed on the GPU (util.py:255-268 for the small VCF)
            util.compress(args.output + ".vcf")
        logger.info("Read generation complete")
    finally:
        stop.set()
        for t in threads:
            t.join()
        for e in engines:
            e.close()


def _device_count():
    import ctypes

    from . import _lib

    n = ctypes.c_int(0)
    _lib.check(_lib.load().iss_device_count(ctypes.byref(n)))
    return n.value


def build_parser():
    parser = argparse.ArgumentParser(prog="InSilicoSeq", usage="iss [subcommand] [options]",
                                     description="InSilicoSeq: A sequencing simulator (B200 engine)")
    parser.add_argument("-v", "--version", action="store_true", default=False, help="print software version and exit")
    sub = parser.add_subparsers(title="available subcommands", metavar="")
    g = sub.add_parser("generate", prog="iss generate", description="simulate reads from an error model",
                       help="simulate reads from an error model")
    log = g.add_mutually_exclusive_group()
    ab = g.add_mutually_exclusive_group()
    log.add_argument("--quiet", "-q", action="store_true", default=False, help="Disable info logging. (default: %(default)s).")
    log.add_argument("--debug", "-d", action="store_true", default=False, help="Enable debug logging. (default: %(default)s).")
    g.add_argument("--seed", type=int, metavar="<int>", default=None, help="Seed all the random number generators")
    g.add_argument("--cpus", "-p", default=2, type=int, metavar="<int>",
                   help="number of work chunks (the reference's worker processes). (default: %(default)s).")
    g.add_argument("--gpus", default=0, type=int, metavar="<int>", help="number of GPUs to use (default: all visible).")
    g.add_argument("--genomes", "-g", metavar="<genomes.fasta>", nargs="+", help="Input genome(s) from where the reads will originate")
    g.add_argument("--draft", metavar="<draft.fasta>", nargs="+", help="Input draft genome(s) from where the reads will originate")
    g.add_argument("--n_genomes", "-u", type=int, metavar="<int>",
                   help="How many genomes will be used for the simulation (random pick from the input multifasta)")
    g.add_argument("--ncbi", "-k", choices=["bacteria", "viruses", "archaea"], action="append", nargs="*", metavar="<str>",
                   help="Download input genomes from NCBI (not available in this engine: no network path)")
    g.add_argument("--n_genomes_ncbi", "-U", type=int, action="append", metavar="<int>", nargs="*",
                   help="How many genomes will be downloaded from NCBI.")
    dists = ["uniform", "halfnormal", "exponential", "lognormal", "zero_inflated_lognormal"]
    ab.add_argument("--abundance", "-a", choices=dists, metavar="<str>", default="lognormal",
                    help="abundance distribution (default: %(default)s).")
    ab.add_argument("--abundance_file", "-b", metavar="<abundance.txt>", help="abundance file for coverage calculations")
    ab.add_argument("--coverage", "-C", choices=dists, metavar="<str>", help="coverage distribution.")
    ab.add_argument("--coverage_file", "-D", metavar="<coverage.txt>", help="file containing coverage information")
    ab.add_argument("--readcount_file", "-R", metavar="<readcount.txt>", help="file containing read_count information")
    g.add_argument("--n_reads", "-n", metavar="<int>", default="1000000",
                   help="Number of reads to generate (default: %(default)s). Allows suffixes k, K, m, M, g and G (ex 0.5M for 500000).")
    g.add_argument("--mode", "-e", metavar="<str>", choices=["kde", "basic", "perfect"], default="kde",
                   help="Error model (default: %(default)s). Can be kde, basic or perfect")
    g.add_argument("--model", "-m", metavar="<npz>", default=None,
                   help="Error model file, or HiSeq, NextSeq, NovaSeq, MiSeq, Miseq-[20,24,28,32,36] for a bundled one (case insensitive).")
    g.add_argument("--gc_bias", "-c", action="store_true", default=False,
                   help="If set, may fail to sequence reads with abnormal GC content. (default: %(default)s)")
    g.add_argument("--compress", "-z", action="store_true", default=False, help="Compress the output in gzip format (default: %(default)s).")
    g.add_argument("--output", "-o", metavar="<fastq>", help="Output file path and prefix (Required)", required=True)
    g.add_argument("--sequence_type", "-t", choices=["metagenomics", "amplicon"], default="metagenomics", required=False,
                   help="Type of sequencing. Can be metagenomics or amplicon (default: %(default)s).")
    g.add_argument("--fragment-length", "-l", metavar="<int>", required=False, type=int,
                   help="Fragment length for metagenomics sequencing (default: %(default)s).")
    g.add_argument("--fragment-length-sd", "-s", metavar="<int>", required=False, type=int,
                   help="Fragment length standard deviation for metagenomics sequencing (default: %(default)s).")
    g.add_argument("--store_mutations", "-M", action="store_true", default=False,
                   help="Generates an additional VCF file with the mutations introduced in the reads")
    g._optionals.title = "arguments"
    g.set_defaults(func=generate_reads)
    return parser


def main(argv=None):
    parser = build_parser()
    args = parser.parse_args(argv)
    try:
        if args.version:
            print("iss-b200 version %s" % __version__)
            sys.exit(0)
        elif args.quiet:
            logging.basicConfig(level=logging.ERROR)
        elif args.debug:
            logging.basicConfig(level=logging.DEBUG)
        else:
            logging.basicConfig(level=logging.INFO)
        args.func(args)
        logging.shutdown()
    except AttributeError as e:
        logging.getLogger(__name__).debug(e)
        parser.print_help()
