/*
 * iss_b200.h -- C ABI of the B200 read-pair generation engine (libiss_b200.so).
 *
 * The reference (InSilicoSeq 2.0.1) is pure Python and has no FFI; its boundary for this path is
 * Python-level.  Each entry point below names the reference interface it replaces:
 *
 *   iss_ctx_set_genome   <- the SeqRecord objects pickled to every worker       iss/app.py:99-106,
 *                           (FASTA records parsed by SeqIO.parse)               iss/generator.py:270
 *   iss_fasta_index /    <- SeqIO.parse(genome_file, "fasta") + str(record.seq) iss/generator.py:270,468-490,
 *   iss_ctx_set_genome_fasta   per record: the genome FILE as text, title lines    iss/app.py:74-76
 *                           found on the host, everything else on the GPU
 *   iss_ctx_set_model    <- the ErrorModel object pickled to every worker       iss/error_models/kde.py:24-50
 *                           (attribute set of KDErrorModel / Basic / Perfect)   iss/error_models/basic.py, perfect.py
 *   iss_ctx_submit       <- pool.starmap(worker_iterator, work chunks)          iss/app.py:99-106
 *                           one iss_work_item == one (record, n_pairs, mode)     iss/generator.py:223-251,333-352
 *   iss_ctx_next/release <- the per-worker temp files {prefix}_R1.fastq,         iss/generator.py:227-229,
 *                           _R2.fastq, .vcf that the parent concatenates          iss/app.py:123-133
 *   iss_comm_* / bcast   <- multiprocessing's pickle scatter of genome + model   iss/app.py:99-106
 *
 * Everything is plain C: pointers + sizes, int status codes (0 = ok), iss_last_error() for text.
 * Host arrays passed in are copied before the call returns.  Device memory, streams, pinned ring
 * buffers and the NCCL communicator are owned by the context.  One host thread per context.
 *
 * Random numbers: Philox4x32-10, slot schedule "RNG contract v2" (DESIGN.md section 3).  All model
 * tables are integer thresholds on 32/64-bit Philox words, so the device path has no floating point.
 */
#ifndef ISS_B200_H
#define ISS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISS_ABI_VERSION 1

/* status codes */
#define ISS_OK 0
#define ISS_E_CUDA 1      /* a CUDA / NCCL call failed (text in iss_last_error) */
#define ISS_E_ARG 2       /* invalid argument */
#define ISS_E_STATE 3     /* call out of order (e.g. submit before set_model) */
#define ISS_E_CHAR 4      /* sequence character outside the reference's alphabet (ACGT + IUPAC, either case) */
#define ISS_E_SHORT 5     /* record not longer than the read length (reference: AssertionError, generator.py:130) */
#define ISS_E_DEVICE 6    /* device-side failure flag (negative window, empty quality bin ...) */
#define ISS_E_NOMEM 7
#define ISS_E_IO 8        /* an output file could not be opened or written (reference: logger.error + sys.exit(1), generator.py:230-232) */

typedef struct iss_ctx iss_ctx;

/* ---- context ------------------------------------------------------------------------------- */
int iss_abi_version(void);
const char *iss_last_error(void);
int iss_device_count(int *count);
/* PCI bus id ("0000:3b:00.0") of a device, for NUMA-aware placement of the host thread that owns its pinned ring. */
int iss_device_pci_bus_id(int device, char *buf, int len);
int iss_ctx_create(int device, iss_ctx **out);
void iss_ctx_destroy(iss_ctx *ctx);

/* ---- genome (replaces the pickled SeqRecords) ---------------------------------------------- */
typedef struct {
    const uint8_t *seq;  /* ASCII bases, as Bio.SeqIO.parse(...,"fasta") yields them (case preserved) */
    uint64_t len;
    const char *name;    /* record.id (first whitespace-delimited token of the FASTA title) */
    uint32_t name_len;
    uint32_t reserved;
} iss_record;

/* Copies + encodes the records: pure upper-case ACGT records are 2-bit packed, anything else
 * (lower case, IUPAC codes) is kept as bytes so case and ambiguity codes survive exactly as in
 * the reference.  Characters outside util.rev_comp's dictionary (iss/util.py:57-88) -> ISS_E_CHAR. */
int iss_ctx_set_genome(iss_ctx *ctx, const iss_record *records, uint32_t n_records);

/* FASTA text in, white space and title lines taken out on the device: replaces Bio.SeqIO.parse(genome_file, "fasta") and
 * str(record.seq) per record (iss/generator.py:270,468-490; iss/app.py:74-76).  iss_fasta_index is host-only (a memchr
 * scan for '>' at line starts; text in front of the first title is ignored, as Bio does); the records it returns are
 * malloc'ed and released with iss_fasta_free.  iss_ctx_set_genome_fasta copies the text to the GPU as it is, fills in
 * seq_len (bases per record = body bytes that are not white space) and builds the same genome blob as
 * iss_ctx_set_genome called with (id, sequence) pairs.  `text` must stay readable for the duration of the call only. */
typedef struct iss_fasta_rec {
    uint64_t gt_pos;     /* position of the title line's '>' */
    uint64_t title_off;  /* the title without '>' and without trailing white space: record.description */
    uint64_t title_len;
    uint64_t id_off;     /* its first white-space delimited token: record.id */
    uint32_t id_len;
    uint32_t reserved;
    uint64_t body_off;   /* the lines between this title and the next one (or the end of the text) */
    uint64_t body_len;
    uint64_t seq_len;    /* out: set by iss_ctx_set_genome_fasta */
} iss_fasta_rec;
int iss_fasta_index(const char *text, uint64_t n, iss_fasta_rec **records, uint32_t *n_records);
void iss_fasta_free(iss_fasta_rec *records);
int iss_ctx_set_genome_fasta(iss_ctx *ctx, const char *text, uint64_t n, iss_fasta_rec *records, uint32_t n_records);

/* ---- error model (replaces the pickled ErrorModel) ------------------------------------------ */
/* Integer-threshold lowering of the model's f64 tables.  With w a 32-bit Philox word, u = w * 2^-32:
 *   searchsorted(cdf, u, 'left')  == #{j : w >  G[j]},  G[j] = min(floor(cdf[j] * 2^32), 2^32 - 1)
 *   choice / searchsorted 'right' == #{j : w >= H[j]},  H[j] = ceil(cdf[j] * 2^32)        (uint64)
 *   u <  p                        == nz && w <= D,       C = ceil(p * 2^32), nz = C > 0, D = C - 1
 *   u >  P                        == w >  floor(P * 2^32)
 * Orientation index 0 = forward, 1 = reverse.  Base index: A=0, C=1, T=2, G=3 ((ascii >> 1) & 3). */
#define ISS_QROW 64  /* quality rows are padded to 64 thresholds with 0xFFFFFFFF */
#define ISS_MAX_ALT 4

#define ISS_QUAL_KDE 0    /* per-position CDF rows selected by a per-read quality bin   kde.py:52-86 */
#define ISS_QUAL_CONST 1  /* constant phred                                             perfect.py:35-42 */
#define ISS_QUAL_TABLE 2  /* position-independent table on the 32-bit draw              basic.py:41-53 */

typedef struct {
    uint32_t read_length;          /* L */
    uint32_t qual_mode;            /* ISS_QUAL_* */
    uint32_t n_insert;             /* entries of insert_thr; 0 => fixed_insert is used */
    int32_t fixed_insert;          /* basic.py:19 / perfect.py:18 insert_size */
    uint32_t frag_mode;            /* 1: --fragment-length/--fragment-length-sd given (generator.py:121-123) */
    uint32_t n_frag;
    int64_t frag_lo;               /* fragment length = frag_lo + #{k : r64 >= frag_thr[k]} */
    uint32_t const_phred;
    uint32_t n_basic;              /* ISS_QUAL_TABLE: phred = basic_val[#{k : w >= basic_thr[k]}] */
    uint32_t store_mutations;
    uint32_t reserved;
    const uint64_t *frag_thr;      /* [n_frag] */
    const uint32_t *insert_thr;    /* [n_insert]                G */
    const uint64_t *bin_thr;       /* [2][4]                    H */
    const uint8_t *bin_valid;      /* [2][4] 1 when the bin has >= L CDF rows */
    const uint32_t *qual_thr;      /* [2][4][L][ISS_QROW]       G, padded */
    const uint32_t *err_thr;       /* [ISS_QROW]                floor(phred_to_prob(q) * 2^32) */
    const uint8_t *subst_n;        /* [2][L][4] alternatives per original base */
    const uint64_t *subst_thr;     /* [2][L][4][ISS_MAX_ALT]    H */
    const uint8_t *subst_alt;      /* [2][L][4][ISS_MAX_ALT]    ASCII */
    const uint32_t *ins_thr;       /* [2][L][4]                 D, lanes in the model dict's key order */
    const uint8_t *ins_nz;         /* [2][L] bit lane set when that lane's probability is > 0 */
    const uint8_t *ins_base;       /* [2][L][4] ASCII */
    const uint32_t *del_thr;       /* [2][L][4]                 D, indexed by base */
    const uint8_t *del_nz;         /* [2][L] bit b set when base b's probability is > 0 */
    const uint32_t *basic_thr;     /* [n_basic] sorted */
    const uint8_t *basic_val;      /* [n_basic + 1] */
} iss_model;

int iss_ctx_set_model(iss_ctx *ctx, const iss_model *model);

/* ---- multi-GPU: genome + model are broadcast once, then ranks never talk again --------------- */
#define ISS_UNIQUE_ID_BYTES 128
int iss_comm_unique_id(uint8_t id[ISS_UNIQUE_ID_BYTES]);                 /* rank 0, then share out of band */
int iss_comm_init_rank(iss_ctx *ctx, int nranks, int rank, const uint8_t id[ISS_UNIQUE_ID_BYTES]);
int iss_comm_init_all(iss_ctx **ctxs, int n);                           /* one process, n devices */
/* ncclBroadcast of the packed genome blob and the model blob from `root`; non-root contexts need no
 * prior set_genome/set_model.  Rank mode: every rank calls it.  Group mode: pass all contexts. */
int iss_ctx_bcast(iss_ctx *ctx, int root);
int iss_ctx_bcast_all(iss_ctx **ctxs, int n, int root);

/* ---- work (replaces starmap(worker_iterator, chunks)) --------------------------------------- */
typedef struct {
    uint32_t record;      /* index into the records passed to iss_ctx_set_genome */
    uint32_t cpu_number;  /* chunk index: goes into read ids and the Philox key (generator.py:150,234-236) */
    uint64_t n_pairs;     /* pairs of this work item handled by THIS submit: i = i0 .. i0+n_pairs-1 */
    uint64_t i0;          /* first pair index (0 unless an item is split across submits / GPUs) */
    uint64_t item_base;   /* pairs of the preceding work items of the same worker (RNG contract: P = item_base + i) */
    uint64_t key;         /* Philox key = (seed + cpu_number) mod 2^64 */
} iss_work_item;

#define ISS_SEQ_METAGENOMICS 0
#define ISS_SEQ_AMPLICON 1

typedef struct {
    uint32_t sequence_type;   /* ISS_SEQ_* (generator.py:134-139,164-171) */
    uint32_t gc_bias;         /* generator.py:82-92 */
    uint64_t batch_pairs;     /* pairs per ring slot; 0 = default */
    uint32_t n_slots;         /* ring depth (>= 2); 0 = default (3; 6 with compress) */
    uint32_t device_only;     /* 1: keep FASTQ in HBM (no D2H) -- kernel-throughput measurements */
    uint32_t tma_gather;      /* 1: stage each read's packed genome window into shared memory with cp.async.bulk + mbarrier
                                 instead of LDG (needs L >= 96).  Measured 1-2 % slower on B200 for both an L2-resident and
                                 an HBM-random genome, hence off by default. */
    uint32_t compress;        /* 1: r1/r2 of every batch are gzip streams (RFC 1952 members of <= 64 KB each, dynamic-Huffman
                                 DEFLATE, compressed on the GPU) instead of plain FASTQ -- `--compress`, iss/util.py:255-268 */
} iss_gen_opts;

/* Starts asynchronous generation of all items (order preserved).  Records with len <= L are the
 * caller's to skip (reference logs a warning, generator.py:77-80); passing one -> ISS_E_SHORT. */
int iss_ctx_submit(iss_ctx *ctx, const iss_work_item *items, uint64_t n_items, const iss_gen_opts *opts);

typedef struct {
    const uint8_t *r1;        /* pinned host memory (device memory when device_only) */
    const uint8_t *r2;
    uint64_t nbytes;          /* bytes in r1 (plain FASTQ: == bytes in r2, the two records of a pair have equal length) */
    uint64_t nbytes2;         /* bytes in r2 (differs from nbytes only when opts.compress) */
    const uint8_t *vcf;       /* store_mutations: VCF lines of this batch in reference order, else NULL */
    uint64_t vcf_bytes;
    uint64_t n_pairs;
    uint64_t first_item;      /* index of the first work item contributing to this batch */
    uint64_t batch_index;
    uint32_t slot;
    float kernel_ms;          /* device time of this batch's kernels (CUDA events on the launch stream) */
} iss_batch;

/* Blocks until the next batch (in submit order) is complete in host memory.  Returns ISS_OK and
 * sets *have = 1, or *have = 0 when the submit is exhausted.  The batch stays valid until
 * iss_ctx_release, which also lets the engine launch the next batch into that slot. */
int iss_ctx_next(iss_ctx *ctx, iss_batch *batch, int *have);
int iss_ctx_release(iss_ctx *ctx, const iss_batch *batch);

/* Drops whatever is left of the current submit: waits for the batches in flight, frees every slot, and leaves the context
 * ready for a new set_genome / set_model / submit.  Also the way out after a failed iss_ctx_submit / iss_ctx_next. */
int iss_ctx_abort(iss_ctx *ctx);

/* ---- files (replaces worker_iterator's SeqIO.write loop and, for plain FASTQ, the parent's concatenate) ---------------
 * iss/generator.py:64-66,227-238 write {prefix}_R1.fastq, _R2.fastq, .vcf from Python objects; iss/app.py:123-133 then
 * copies every byte again into the final files.  iss_ctx_write_files runs submit + next/release itself and pwrite()s every
 * finished batch from the pinned ring slot with a pool of native threads.  With append == 0 the bytes of this call go to
 * [offset_r1, ...) of path_r1 (and likewise R2): plain FASTQ sizes are closed-form (iss_fastq_bytes), so several workers /
 * GPUs can write one final file at disjoint offsets and no concatenation pass is needed.  VCF text is always appended. */
typedef struct {
    const char *path_r1;      /* NULL: that stream is produced and copied to the host but not written */
    const char *path_r2;
    const char *path_vcf;
    uint64_t offset_r1;       /* append == 0: file offset of this call's first R1 / R2 byte */
    uint64_t offset_r2;
    uint32_t append;          /* 1: write at the end of the file (compressed streams: sizes are not known in advance) */
    uint32_t truncate;        /* 1: O_TRUNC on open (a worker's private file, like the reference's open(..., "w")) */
    uint32_t writer_threads;  /* pwrite threads; 0 = default (8) */
    uint32_t reserved;
} iss_file_sink;

typedef struct {
    uint64_t bytes_r1, bytes_r2, bytes_vcf;
    uint64_t n_pairs, n_batches;
    double wall_ms;           /* whole call, host clock */
    double write_ms;          /* part of it spent waiting for the writers */
} iss_write_stats;

int iss_ctx_write_files(iss_ctx *ctx, const iss_work_item *items, uint64_t n_items, const iss_gen_opts *opts,
                        const iss_file_sink *sink, iss_write_stats *out);

/* ---- measurement ------------------------------------------------------------------------------ */
/* Sustained device -> pinned host copy rate (GB/s, CUDA events) into a ring of ring_bytes allocated like the output ring
 * (alloc_mode: 0 cudaMallocHost, 1 mmap + mbind(GPU node) + cudaHostRegister, 2 write-combined), copies of chunk_bytes.
 * *node receives the NUMA node that backs the ring (-1 unknown).  Called concurrently for several devices it measures the
 * host side of the 8-GPU end-to-end figure (tools/d2h_probe.py). */
int iss_probe_d2h(int device, int alloc_mode, uint64_t ring_bytes, uint64_t chunk_bytes, double seconds, double *gbps, int *node);

typedef struct {
    uint64_t pairs;            /* pairs generated since context creation */
    uint64_t fastq_bytes;      /* R1 + R2 bytes produced */
    uint64_t d2h_bytes;
    uint64_t h2d_bytes;
    uint64_t kernel_launches;
    double kernel_ms;          /* sum of per-batch kernel device time */
    uint64_t slow_reads;       /* reads that took the sequential indel replay */
    uint64_t subst_errors;     /* substitution events */
    uint64_t genome_bytes_device;
    uint64_t model_bytes_device;
    int32_t ring_numa_node;    /* NUMA node backing the pinned output ring (-1: unknown / not bound) */
    uint32_t ring_alloc_mode;  /* 0 cudaMallocHost, 1 mmap + mbind(GPU node) + cudaHostRegister, 2 write-combined */
} iss_stats;
int iss_ctx_stats(iss_ctx *ctx, iss_stats *out);

/* Device-timeline marks for benchmarking (CUDA events): a mark is ordered after ALL work enqueued so far on
 * every stream of the context.  elapsed = ms between two marks (synchronises on the later one). */
#define ISS_N_MARKS 4
int iss_ctx_sync(iss_ctx *ctx);
int iss_ctx_mark(iss_ctx *ctx, int idx);
int iss_ctx_elapsed_ms(iss_ctx *ctx, int idx_from, int idx_to, float *ms);

/* FASTQ bytes of pairs i0..i1-1 of one read file: sum of len("@{name}_{i}_{cpu}/1\n{seq}\n+\n{qual}\n") */
uint64_t iss_fastq_bytes(uint32_t name_len, uint64_t cpu_number, uint32_t read_length, uint64_t i0, uint64_t i1);

/* gzip `n` host bytes on the GPU (same member format as opts.compress); *out_n receives the compressed size.
 * ISS_E_ARG when cap is too small (n + n/2048 + 64 always suffices). */
int iss_ctx_gzip(iss_ctx *ctx, const uint8_t *in, uint64_t n, uint8_t *out, uint64_t cap, uint64_t *out_n);

/* Philox4x32-10 on the device (one block), for known-answer tests of the device generator. */
int iss_ctx_philox(iss_ctx *ctx, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* ISS_B200_H */
