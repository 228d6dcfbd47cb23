// iss_device.cuh -- device-side data layout shared by the kernels and the host runtime.
//
// HBM layout (one allocation each, so that one ncclBroadcast moves each):
//   genome blob : [GenomeHeader][DevRecord x n][names][64 B pad][2-bit words ...][64 B pad][ascii bytes ...][64 B pad]
//   model blob  : [ModelHeader][tables ..., each 16 B aligned]
// Both blobs are position independent (offsets, not pointers), so a rank that receives them over
// NCCL rebuilds its pointer view with no host-side data.
#pragma once
#include <stdint.h>

#define ISS_QROW 64
#define ISS_MAX_ALT 4
#define ISS_SLOW_SLOTS 16
#define ISS_SURV_CAP 64  // queued indel-site survivors per tile before the scan falls back to a second walk
#define ISS_Q16_ROW_BYTES 132
#define ISS_Q8_ROW_BYTES 24   // 8 x u16 thresholds + 8 x u8 counts  // a 16-bit quality row: ISS_QROW u16 entries + 4 bytes, so that the rows of a position start in different banks

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. SC'11).  Slot schedule: DESIGN.md section 3 / oracle/philox.py.
// ---------------------------------------------------------------------------------------------
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        uint32_t hi0 = __umulhi(PHILOX_M0, c0), lo0 = PHILOX_M0 * c0;
        uint32_t hi1 = __umulhi(PHILOX_M1, c2), lo1 = PHILOX_M1 * c2;
#else
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0, p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += PHILOX_W0;
        k1 += PHILOX_W1;
    }
    Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// Same function with the ten round keys precomputed (rk[2r] = k0 + r*W0, rk[2r+1] = k1 + r*W1).  With rk in the
// kernel's constant bank a round is 2 IMAD.WIDE + 2 LOP3.
__host__ __device__ __forceinline__ Philox4 philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                             const uint32_t *__restrict__ rk) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        const uint32_t hi0 = __umulhi(PHILOX_M0, c0), lo0 = PHILOX_M0 * c0;
        const uint32_t hi1 = __umulhi(PHILOX_M1, c2), lo1 = PHILOX_M1 * c2;
#else
        const uint64_t p0 = (uint64_t)PHILOX_M0 * c0, p1 = (uint64_t)PHILOX_M1 * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        const uint32_t n0 = hi1 ^ c1 ^ rk[2 * r];
        const uint32_t n2 = hi0 ^ c3 ^ rk[2 * r + 1];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}
__host__ __device__ inline void philox_round_keys(uint64_t key, uint32_t rk[20]) {
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
    for (int r = 0; r < 10; ++r) { rk[2 * r] = k0; rk[2 * r + 1] = k1; k0 += PHILOX_W0; k1 += PHILOX_W1; }
}

// stream ids of RNG contract v2
#define ISS_S_PAIR 0u
#define ISS_K_QUAL 0u
#define ISS_K_SUBC 1u
#define ISS_K_DEL 2u
#define ISS_K_INS 3u
__host__ __device__ __forceinline__ uint32_t iss_read_stream(uint32_t o, uint32_t k) { return 1u + 4u * o + k; }

// ---------------------------------------------------------------------------------------------
// genome blob
// ---------------------------------------------------------------------------------------------
#define ISS_ENC_2BIT 0u
#define ISS_ENC_ASCII 1u

struct DevRecord {
    uint64_t data_off;  // 2-bit: index of the record's first base in the packed word array (multiple of 64)
                        // ascii: byte offset in the byte array
    uint64_t len;
    uint32_t enc;
    uint32_t name_off;
    uint32_t name_len;
    uint32_t pad;
};

struct GenomeHeader {
    uint64_t magic;
    uint64_t total_bytes;
    uint32_t n_records;
    uint32_t pad;
    uint64_t recs_off, names_off, names_bytes, packed_off, packed_bytes, ascii_off, ascii_bytes;
};

// ---------------------------------------------------------------------------------------------
// model blob
// ---------------------------------------------------------------------------------------------
struct ModelHeader {
    uint64_t magic;
    uint64_t total_bytes;
    uint32_t L, qual_mode, n_insert, frag_mode, n_frag, const_phred, n_basic, store_mutations;
    int32_t fixed_insert;
    uint32_t n_grp;        // groups of 4 indel draw sites (= one Philox block; RNG contract v2), all lists together
    float surv_per_pair;   // expected number of sites whose draw passes, per pair (host statistic: picks the scan mode)
    uint32_t ins_stride;   // insert-size search, first level: ins_coarse[k] = insert_thr[min((k + 1) * ins_stride, n_insert) - 1]
    uint32_t ins_coarse[31];
    uint32_t grp_begin[5]; // first group of: insertion sites fwd, deletion sites fwd, insertion sites rev, deletion sites rev; end
    uint32_t pad[3];
    int64_t frag_lo;
    uint64_t bin_thr[2][4];
    uint8_t bin_valid[2][4];
    uint32_t err_thr[ISS_QROW];
    // offsets (bytes from blob start)
    uint64_t frag_thr_off, insert_thr_off, qual_exact_off, subst_n_off, subst_thr_off, subst_alt_off, ins_base_off,
        del_thr_off, del_nz_off, basic_thr_off, basic_val_off, ent_off;
    // 16-bit quality rows (kernel phase 4): per orientation [L + 1 positions][q16_ps bytes: q16_nb rows of ISS_Q16_ROW_BYTES]
    uint64_t q16_off[2];
    uint32_t q16_nb[2];       // populated quality bins of the orientation (rows per position)
    uint32_t q16_ps[2];       // bytes per position: q16_nb * ISS_Q16_ROW_BYTES rounded up to 16 (bulk copies are 16-byte granular)
    uint8_t q16_cbin[2][4];   // bin -> row index among the populated bins
    // short rows: where no row of the model has more than 8 distinct 16-bit thresholds (binned qualities: NovaSeq, NextSeq,
    // HiSeq), a row is also stored as its <= 8 distinct non-zero values (pad 0xFFFF) + 8 bytes "entries of the full row below value c"
    uint64_t q8_off[2];
    uint32_t q8_ps[2];        // bytes per position of the short rows: q16_nb * ISS_Q8_ROW_BYTES rounded up to 16
    uint32_t q8_ok;           // the model has them
    uint32_t q8_pad;
};

// pointer view built on each device from the blob base
struct DevModel {
    const ModelHeader *h;
    const uint64_t *frag_thr;
    const uint32_t *insert_thr;
    const uint32_t *qual_exact; // [2][4][L][ISS_QROW] sorted 32-bit thresholds (tie fallback of the 16-bit search)
    const uint8_t *q16[2];      // 16-bit rows of orientation o, position-major (one bulk copy per 2 positions)
    const uint8_t *q8[2];       // the short rows, same order
    const uint8_t *subst_n;     // [2][L][4]
    const uint64_t *subst_thr;  // [2][L][4][4]
    const uint8_t *subst_alt;   // [2][L][4][4]
    const uint8_t *ins_base;    // [2][L][4] inserted base per lane (the thresholds live in `ent`)
    const uint32_t *del_thr;    // [2][L][4]
    const uint8_t *del_nz;      // [2][L]
    const uint4 *ent;           // [n_grp][2]: {reject thresholds of the group's 4 sites: the insertion threshold, or the
                                //             largest of a deletion site's 4}, {site: pos | valid << 16 | insertion lane << 17}
    const uint32_t *basic_thr;
    const uint8_t *basic_val;
};

struct DevGenome {
    const DevRecord *recs;
    const char *names;
    const uint32_t *packed;  // 2-bit words; 16 bases per word, base m of a word at bits [2m, 2m+2)
    const uint8_t *ascii;
    uint32_t n_records;
};

// One segment = a run of consecutive pairs of one work item inside a batch.
struct DevSegment {
    uint64_t pair_start;  // batch-local index of the segment's first pair
    uint64_t n;           // pairs in the segment
    uint64_t i0;          // pair index i of the first pair
    uint64_t item_base;   // P = item_base + i
    uint64_t key;         // Philox key (uniform per launch: the planner cuts batches at key changes)
    uint64_t out_base;    // byte offset of pair i in the batch stream = out_base + cf(i)  (mod 2^64)
    uint32_t rec;
    uint32_t cpu;
    uint32_t fixed_len;   // record bytes without the digits of i: 2L + 6 + name_len + 2 + digits(cpu) + 2
    uint32_t cpu_digits;
};

// cumulative FASTQ bytes of pairs 0..i-1 of a segment: cf(i) = i*(fixed_len+1) + sum_k max(0, i - 10^k)
__host__ __device__ __forceinline__ uint32_t iss_digits(uint64_t v) {
    uint32_t d = 1;
    uint64_t p = 10;
#pragma unroll 1
    for (int k = 1; k < 20; ++k) {
        if (v >= p) d++; else break;
        p *= 10;  // 10^19 < 2^64; the loop ends before 10^20
    }
    return d;
}
__host__ __device__ __forceinline__ uint64_t iss_cf(uint64_t i, uint32_t fixed_len) {
    uint64_t s = i * (uint64_t)(fixed_len + 1u);
    uint64_t p = 10;
#pragma unroll 1
    for (int k = 1; k < 20; ++k) {
        if (i > p) s += i - p; else break;
        p *= 10;
    }
    return s;
}

// Shared-memory layout of one tile, as byte offsets from the dynamic shared memory base (computed on the host so
// that they sit in the constant bank).  Offsets, not pointers: every access stays provably in the shared window.
struct Lay {
    uint32_t img[2];   // FASTQ stream images (R1, R2)
    uint32_t sP;       // u64[T]  pair ordinal
    uint32_t sOff;     // u64[T]  byte offset of the pair's record in the batch stream
    uint32_t sRecLen;  // u64[T]  record length
    uint32_t sWabs;    // u64[2T] absolute index (base or byte) of the window start in the genome array
    uint32_t sWin;     // u64[2T] window start inside the record (fs / rs)
    uint32_t sSeg;     // u32[T]
    uint32_t sAtt;     // u32[T]
    uint32_t sQrow;    // u32[2T] quality rows of the read: byte offset of its bin's row inside a position | bin << 16
    uint32_t sRel;     // u32[T]  low 32 bits of (stream offset of the record + header length): the sequence line sits at
                       //         image offset  (a0 - (uint32)g0) + sRel[pair]  in both stream images
    uint32_t sEnc;     // u32[T]
    uint32_t sSlow;    // u32[2T]
    uint32_t sFlag;    // u32[2T] read already queued for the sequential indel replay
    uint32_t sErr;     // u32[err_cap]
    uint32_t sErrThr;  // u32[64]
    uint32_t sTie;     // u32[tie_cap] positions whose 16-bit quality search met a tie (read << 16 | position)
    uint32_t sSurv;    // u32[ISS_SURV_CAP] indel sites whose draw passed, to be checked against the base under them
                       //         (read << 24 | site bits << 16 | position)
    uint32_t sCtl;     // u32[8]: 0 nErr, 1 nSlow, 2 gather-mbarrier parity, 3 bytes of the tile's last record, 4 nTie, 6 nSurv;
                       //         +28: (unused); the mbarrier of the TMA genome gather has its own 8 bytes behind (sGbar)
    uint32_t sGbar;    // u64 mbarrier of the TMA genome gather
    uint32_t sBar;     // u64[n_streams][2 stages][2]: {full, empty} mbarriers of the quality-row pipeline
    uint32_t sQst;     // [n_streams][2 stages][q_stage_cap] staged 16-bit quality rows (2 positions per stage)
    uint32_t total;    // bytes of dynamic shared memory
};

// n_streams = 2 * S position streams per tile (orientation x part), each with two stages of q_stage_cap bytes (0: no pipeline)
__host__ __device__ inline Lay iss_make_layout(uint32_t T, uint32_t img_cap, uint32_t err_cap, uint32_t tie_cap, uint32_t n_streams,
                                               uint32_t q_stage_cap) {
    const uint32_t q_stages = q_stage_cap ? 2u : 0u;
    Lay l;
    uint32_t q = 0;
    l.img[0] = q; q += img_cap;
    l.img[1] = q; q += img_cap;
    l.sP = q; q += 8u * T;
    l.sOff = q; q += 8u * T;
    l.sRecLen = q; q += 8u * T;
    l.sWabs = q; q += 16u * T;
    l.sWin = q; q += 16u * T;
    l.sSeg = q; q += 4u * T;
    l.sAtt = q; q += 4u * T;
    l.sQrow = q; q += 8u * T;
    l.sRel = q; q += 4u * T;
    l.sEnc = q; q += 4u * T;
    l.sSlow = q; q += 8u * T;
    l.sFlag = q; q += 8u * T;
    l.sErr = q; q += 4u * err_cap;
    l.sErrThr = q; q += 4u * ISS_QROW;
    l.sTie = q; q += 4u * tie_cap;
    l.sSurv = q; q += 4u * ISS_SURV_CAP;
    l.sCtl = q; q += 32u;
    l.sGbar = q; q += 16u;
    l.sBar = q; q += 16u * n_streams * q_stages;
    l.sQst = q; q += n_streams * q_stages * q_stage_cap;
    l.total = q;
    return l;
}

#define ISS_QUAL_KDE_MODE 0u
#define ISS_QUAL_CONST_MODE 1u
#define ISS_QUAL_TABLE_MODE 2u

struct GenParams {
    DevModel m;
    DevGenome g;
    // model scalars (copied from the ModelHeader so that they sit in the constant bank)
    uint32_t L, qual_mode, n_insert, frag_mode, n_frag, const_phred, n_basic;
    int32_t fixed_insert;
    int64_t frag_lo;
    uint32_t n_grp;
    uint32_t grp_begin[5];
    uint64_t bin_thr[2][4];
    uint8_t bin_valid[2][4];
    uint32_t err_thr[ISS_QROW];
    const DevSegment *segs;
    uint32_t nseg;
    uint32_t tile_pairs;      // T: pairs per tile (32 or 64); blockDim.x == 2 * S * T
    uint32_t tile_shift;      // log2(T)
    uint32_t split_shift;     // log2(S), S = threads per read (2, or 4 for long reads)
    uint64_t total_pairs;
    uint64_t key;             // Philox key of every pair of this launch (a batch never spans two workers)
    uint32_t rk[20];          // the ten Philox round keys (k0 + r*W0, k1 + r*W1): constant-bank operands of the LOP3s
    Lay lay;
    uint8_t *out1;
    uint8_t *out2;
    // quality-row pipeline (phase 4): every position stream (orientation x part) of a tile walks the same rows for every
    // tile; its producer lane stages them one block ahead with cp.async.bulk + mbarriers (two stages per stream)
    uint32_t q_stage_cap;     // bytes reserved per stage (>= q_blocks * q_stage_bytes[o], multiple of 16)
    uint32_t q_blocks;        // Philox blocks per stage: 1, or 2 (the <512, 2, 2> build)
    uint32_t q_short;         // the staged rows are the short ones (kK = 8 builds); q_pos_stride / q_stage_bytes / q_row_off describe them
    const uint8_t *q_rows[2]; // the rows the pipeline stages: m.q16 or m.q8
    uint32_t q_stage_bytes[2];  // 2 * q16_ps[o]: the rows of one block of orientation o
    uint32_t q_pos_stride[2];   // q16_ps[o]
    uint32_t q_row_off[2][4];   // bin -> byte offset of its row inside a position (q16_cbin * ISS_Q16_ROW_BYTES)
    uint32_t img_cap;         // bytes of one stream image in shared memory (multiple of 16)
    uint32_t err_cap;         // entries of the substitution list
    uint32_t tie_cap;         // entries of the tie queue (exact pass)
    uint32_t err_ovf_cap;     // entries per CTA of the substitution list's overflow in global memory
    uint32_t *err_ovf;        // [gridDim.x][err_ovf_cap]
    uint32_t ins_stride;      // two-level insert-size search (copied from the ModelHeader)
    uint32_t ins_coarse[31];
    uint32_t surv_direct;     // 1: the model's indel tables are dense (expected survivors per tile beyond the queue): scan once, after the gather
    uint32_t amplicon;
    uint32_t gc_bias;
    uint32_t chunk_magic;     // ceil(2^32 / nchunk) for idx / nchunk (unused when nchunk == 1: the value would be 2^32)
    uint32_t nchunk;          // ceil(L / 16)
    uint8_t *scratch;         // slow-path scratch: gridDim.x * ISS_SLOW_SLOTS * scratch_stride bytes
    uint32_t scratch_stride;  // per replayed read: L original bases + 4L stack + L draw masks (+ alignment)
    uint32_t scratch_mask_off;  // offset of the draw masks inside a slot's scratch (multiple of 4)
    uint32_t gc_thr;          // accept iff w < gc_thr  (ceil(0.90 * 2^32))
    unsigned long long *counters;  // [0] slow reads [1] substitution events [2] error flags
    uint32_t store_mutations;
    uint32_t tma_gather;           // 1: genome windows are staged into shared memory with cp.async.bulk (needs 15 + stage_bytes <= L)
    uint32_t stage_bytes;          // 16 * ceil((L + 125) / 64)
    uint32_t pad0;
    uint4 *events;                 // mutation records of this batch (--store_mutations)
    unsigned long long *event_count;
    unsigned long long event_cap;
};

#define ISS_DEVERR_RANGE 1u
#define ISS_DEVERR_QUALBIN 2u
#define ISS_DEVERR_CHAR 4u
#define ISS_DEVERR_EVENTS 8u
