/*
 * iss_oracle.c -- CPU restatement of InSilicoSeq's per-read-pair generation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity checker for the CUDA path and the
 * "port" CPU baseline of bench.py; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product never links or calls it.
 *
 * Parity pin: this file is validated byte-for-byte against the UNMODIFIED reference
 * (/root/reference/iss) driven by the same Philox slots (oracle/patched_reference.py) --
 * see tests/test_oracle_golden.py and the fixtures under tests/golden/ -- and through it
 * against the reference's own known-answer behaviour (list insert/pop semantics of
 * test_introduce_indels, test_adjust_seq_length_extend, test_introduce_indels_rev).
 *
 * It deliberately keeps the reference's arithmetic: uniforms are doubles u = w * 2^-32 and
 * every decision is the reference's floating-point comparison against the model's f64
 * tables.  (The CUDA path lowers the same comparisons to integer thresholds; agreement of
 * the two is therefore a real check of that lowering.)
 *
 * Reference lines restated (all relative to /root/reference):
 *   simulate_read            iss/generator.py:98-192
 *   random_insert_size       iss/error_models/kde.py:88-98
 *   introduce_indels         iss/error_models/__init__.py:158-228
 *   adjust_seq_length        iss/error_models/__init__.py:114-156
 *   gen_phred_scores         iss/error_models/kde.py:52-86
 *   mut_sequence             iss/error_models/__init__.py:69-112
 *   rev_comp                 iss/util.py:48-92
 *   FASTQ text               iss/generator.py:64-65 (Bio.SeqIO "fastq-sanger")
 *   write_mutations          iss/generator.py:598-620
 * RNG: Philox4x32-10, slot schedule "RNG contract v2" (oracle/philox.py).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ISS_ORACLE_MAX_ALT 4

typedef struct {
    int32_t read_length;      /* L */
    int32_t nq;               /* entries per quality CDF row (41 in every model iss writes) */
    int32_t n_insert;         /* len(i_size_cdf) */
    int32_t frag_mode;        /* 1: --fragment-length/-sd given (generator.py:121-123) */
    int64_t frag_lo;          /* fragment table: value = frag_lo + #{k: frag_thr[k] <= r64} */
    int32_t n_frag;
    int32_t store_mutations;
    const uint64_t *frag_thr; /* [n_frag] */
    const double *insert_cdf; /* [n_insert] */
    const double *bin_cdf;    /* [2][4] cumsum(norm_mean)/last, as numpy's choice builds it */
    const int32_t *bin_rows;  /* [2][4] number of CDF rows in quality[bin] (0 = empty list) */
    const double *qual_cdf;   /* [2][4][L][nq] */
    const double *phred_prob; /* [nq+1] phred_to_prob(q), computed by the reference function */
    const int32_t *subst_n;   /* [2][L][4] number of alternatives (3 kde/basic, 4 perfect) */
    const double *subst_cdf;  /* [2][L][4][4] choice cdf */
    const uint8_t *subst_alt; /* [2][L][4][4] ASCII */
    const double *ins_p;      /* [2][L][4] lanes in the model dict's key order */
    const uint8_t *ins_base;  /* [2][L][4] ASCII base of each lane */
    const double *del_p;      /* [2][L][4] indexed by base code A,C,G,T = 0..3 */
    const int32_t *ins_rank;  /* [2][L][4] RNG contract v2: slot rank of the site among the entries with p > 0, or -1 */
    const int32_t *del_rank;  /* [2][L]    slot rank of the position among the positions with any p > 0, or -1 */
    /* basic/perfect models: fixed insert size and host-tabulated phred draws */
    int32_t fixed_insert;     /* >=0: random_insert_size() returns this constant (basic.py:55, perfect.py:44) */
    int32_t qual_mode;        /* 0 kde CDF rows; 1 constant phred (perfect.py:35); 2 table by 32-bit draw (basic) */
    int32_t const_phred;
    int32_t n_basic;          /* qual_mode 2: phred = #{k: basic_thr[k] <= w} over sorted thresholds */
    const uint32_t *basic_thr;
    const uint8_t *basic_val; /* [n_basic+1] phred value of each interval */
} iss_oracle_model;

/* ---------------- Philox4x32-10 ---------------- */
static inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                 uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void iss_oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

typedef struct {
    uint32_t k0, k1;
    uint64_t P;       /* pair ordinal within worker */
    uint32_t attempt; /* gc_bias re-simulation counter */
    /* one-block cache per stream id (0..8) */
    uint32_t cached_blk[9];
    uint8_t cached_ok[9];
    uint32_t cache[9][4];
} draw_ctx;

static void ctx_init(draw_ctx *d, uint64_t key, uint64_t P, uint32_t attempt) {
    d->k0 = (uint32_t)key;
    d->k1 = (uint32_t)(key >> 32);
    d->P = P;
    d->attempt = attempt;
    memset(d->cached_ok, 0, sizeof d->cached_ok);
}

static inline uint32_t draw_word(draw_ctx *d, uint32_t stream, uint32_t blk, uint32_t w) {
    if (!d->cached_ok[stream] || d->cached_blk[stream] != blk) {
        philox4x32_10(blk, stream | (d->attempt << 8), (uint32_t)d->P, (uint32_t)(d->P >> 32), d->k0, d->k1,
                      d->cache[stream]);
        d->cached_blk[stream] = blk;
        d->cached_ok[stream] = 1;
    }
    return d->cache[stream][w];
}
static inline double draw_u(draw_ctx *d, uint32_t stream, uint32_t blk, uint32_t w) {
    return (double)draw_word(d, stream, blk, w) * (1.0 / 4294967296.0);
}
static inline uint64_t draw_r64(draw_ctx *d, uint32_t stream, uint32_t blk, uint32_t w) {
    uint64_t lo = draw_word(d, stream, blk, w), hi = draw_word(d, stream, blk, w + 1);
    return lo | (hi << 32);
}
static inline uint64_t randbelow(uint64_t r64, uint64_t n) { return (uint64_t)(((__uint128_t)r64 * n) >> 64); }

enum { S_PAIR = 0, K_QUAL = 0, K_SUBC = 1, K_DEL = 2, K_INS = 3 };
static inline uint32_t rstream(int o, int k) { return (uint32_t)(1 + 4 * o + k); }

/* numpy.searchsorted(a, v, side='left') on a nondecreasing array: #{j : a[j] < v} */
static int searchsorted_left(const double *a, int n, double v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = lo + ((hi - lo) >> 1);
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}
/* side='right': #{j : a[j] <= v} */
static int searchsorted_right(const double *a, int n, double v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = lo + ((hi - lo) >> 1);
        if (a[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

static int is_ambiguous_upper(int c) { /* "RYWSMKHBVDN" (error_models/__init__.py:94,190) */
    switch (c) {
    case 'R': case 'Y': case 'W': case 'S': case 'M': case 'K': case 'H': case 'B': case 'V': case 'D': case 'N':
        return 1;
    default: return 0;
    }
}
static inline int up(int c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }
static int base_code(int upper) { /* A,C,G,T -> 0..3 ; -1 otherwise (reference: KeyError) */
    switch (upper) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return -1; }
}
/* iss/util.py:57-88 complement dictionary; returns 0 for a character outside it (reference: KeyError) */
static int complement(int c) {
    switch (c) {
    case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
    case 'y': return 'r'; case 'r': return 'y'; case 'w': return 'w'; case 's': return 's';
    case 'k': return 'm'; case 'm': return 'k'; case 'n': return 'n'; case 'b': return 'v';
    case 'v': return 'b'; case 'd': return 'h'; case 'h': return 'd';
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    case 'Y': return 'R'; case 'R': return 'Y'; case 'W': return 'W'; case 'S': return 'S';
    case 'K': return 'M'; case 'M': return 'K'; case 'N': return 'N'; case 'B': return 'V';
    case 'V': return 'B'; case 'D': return 'H'; case 'H': return 'D';
    default: return 0;
    }
}

/* error codes */
enum {
    ISS_ORACLE_OK = 0,
    ISS_ORACLE_E_SHORT = 1,    /* AssertionError: read_length >= len(record) (generator.py:130) */
    ISS_ORACLE_E_CHAR = 2,     /* KeyError in the reference (base outside its dictionaries) */
    ISS_ORACLE_E_QUALBIN = 3,  /* chosen quality bin has no CDF rows (reference fails in SeqRecord) */
    ISS_ORACLE_E_SPACE = 4,
    ISS_ORACLE_E_RANGE = 5,    /* negative window start (fragment length < read length) */
    ISS_ORACLE_E_ARG = 6
};

typedef struct { /* growable VCF text sink (write_mutations, generator.py:598-620) */
    char *buf; uint64_t cap; uint64_t len; int overflow;
} vcf_sink;

static void vcf_emit(vcf_sink *v, const char *id, int idlen, int position, const char *ref, int reflen,
                     const char *alt, int altlen, int quality /* -1 => "." */) {
    if (!v || !v->buf) return;
    char line[512];
    int n = 0;
    memcpy(line + n, id, (size_t)idlen); n += idlen;
    n += sprintf(line + n, "\t%d\t.\t", position + 1);
    memcpy(line + n, ref, (size_t)reflen); n += reflen;
    line[n++] = '\t';
    memcpy(line + n, alt, (size_t)altlen); n += altlen;
    line[n++] = '\t';
    if (quality < 0) line[n++] = '.'; else n += sprintf(line + n, "%d", quality);
    line[n++] = '\t'; line[n++] = '\t'; line[n++] = '\n';
    if (v->len + (uint64_t)n > v->cap) { v->overflow = 1; return; }
    memcpy(v->buf + v->len, line, (size_t)n);
    v->len += (uint64_t)n;
}

/* introduce_indels + adjust_seq_length, literal list semantics.  m: buffer of capacity >= 5L+8, n = L on entry.
 * rec: full record bytes, reclen; bounds = (read_start, read_end). */
static int introduce_indels(const iss_oracle_model *M, draw_ctx *d, int o, uint8_t *m, const uint8_t *rec,
                            int64_t reclen, int64_t read_start, int64_t read_end, vcf_sink *vcf, const char *id,
                            int idlen) {
    const int L = M->read_length;
    int n = L;
    int position = 0;
    const double *ins_p = M->ins_p + (size_t)o * L * 4;
    const uint8_t *ins_b = M->ins_base + (size_t)o * L * 4;
    const double *del_p = M->del_p + (size_t)o * L * 4;
    for (int nucl = 0; nucl < L - 1; ++nucl) {
        if (nucl >= n) continue; /* IndexError -> continue, position not advanced (:223-224) */
        int cu = up(m[nucl]);
        if (is_ambiguous_upper(cu)) { position++; continue; }
        for (int lane = 0; lane < 4; ++lane) {
            double p = ins_p[position * 4 + lane];
            if (!(p > 0.0)) continue; /* u < p is false for p <= 0 or NaN: the site has no slot */
            const uint32_t r = (uint32_t)M->ins_rank[((size_t)o * L + position) * 4 + lane];
            double u = draw_u(d, rstream(o, K_INS), r >> 2, r & 3);
            if (u < p) {
                /* mutable_seq.insert(position + 1, base) */
                int at = position + 1;
                if (at > n) at = n;
                memmove(m + at + 1, m + at, (size_t)(n - at));
                m[at] = ins_b[position * 4 + lane];
                n++;
                if (M->store_mutations) {
                    char alt[2] = {(char)m[position], (char)ins_b[position * 4 + lane]};
                    vcf_emit(vcf, id, idlen, position, (const char *)&m[position], 1, alt, 2, -1);
                }
            }
        }
        int code = base_code(up(m[nucl]));
        if (code < 0) return ISS_ORACLE_E_CHAR;
        double pd = del_p[position * 4 + code];
        if (pd > 0.0) {
            const uint32_t r = (uint32_t)M->del_rank[(size_t)o * L + position];
            double u = draw_u(d, rstream(o, K_DEL), r >> 2, r & 3);
            if (u < pd) {
                /* mutable_seq.pop(position) */
                memmove(m + position, m + position + 1, (size_t)(n - position - 1));
                n--;
                if (M->store_mutations) {
                    /* "ref" is read AFTER the pop (:216); IndexError there aborts the iteration */
                    if (position >= n) continue; /* IndexError -> except: continue (position not advanced) */
                    vcf_emit(vcf, id, idlen, position, (const char *)&m[position], 1, ".", 1, -1);
                }
            }
        }
        position++;
    }
    /* adjust_seq_length (:114-156) */
    if (n > L) n = L;
    else if (n < L) {
        int to_add = L - n;
        if (o == 0) {
            for (int i = 0; i < to_add; ++i) {
                int64_t g = read_end + i;
                m[n++] = (g >= reclen) ? 'A' : rec[g];
            }
        } else {
            for (int i = 0; i < to_add; ++i) {
                int64_t g = read_start - 1 - i;
                if (g < 0) m[n++] = 'A';
                else {
                    int c = complement(rec[g]);
                    if (!c) return ISS_ORACLE_E_CHAR;
                    m[n++] = (uint8_t)c;
                }
            }
        }
    }
    return ISS_ORACLE_OK;
}

/* gen_phred_scores (kde.py:52-86) or the basic/perfect variants */
static int gen_phred(const iss_oracle_model *M, draw_ctx *d, int o, uint8_t *q) {
    const int L = M->read_length;
    if (M->qual_mode == 1) { for (int p = 0; p < L; ++p) q[p] = (uint8_t)M->const_phred; return ISS_ORACLE_OK; }
    if (M->qual_mode == 2) {
        for (int p = 0; p < L; ++p) {
            uint32_t w = draw_word(d, rstream(o, K_QUAL), (uint32_t)p >> 1, 2u * ((uint32_t)p & 1));
            int lo = 0, hi = M->n_basic;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (M->basic_thr[mid] <= w) lo = mid + 1; else hi = mid; }
            q[p] = M->basic_val[lo];
        }
        return ISS_ORACLE_OK;
    }
    double ub = draw_u(d, S_PAIR, 0, 1 + (uint32_t)o);
    int bin = searchsorted_right(M->bin_cdf + o * 4, 4, ub); /* np.random.choice(range(4), p=norm_mean) */
    if (bin == 4) bin = 3;                                   /* kde.py:77-78 */
    int rows = M->bin_rows[o * 4 + bin];
    if (rows < L) return ISS_ORACLE_E_QUALBIN;
    const double *cdf = M->qual_cdf + (((size_t)o * 4 + bin) * L) * M->nq;
    for (int p = 0; p < L; ++p) {
        double u = draw_u(d, rstream(o, K_QUAL), (uint32_t)p >> 1, 2u * ((uint32_t)p & 1));
        q[p] = (uint8_t)searchsorted_left(cdf + (size_t)p * M->nq, M->nq, u);
    }
    return ISS_ORACLE_OK;
}

/* mut_sequence (error_models/__init__.py:69-112) */
static int mut_sequence(const iss_oracle_model *M, draw_ctx *d, int o, uint8_t *m, const uint8_t *orig,
                        const uint8_t *q, vcf_sink *vcf, const char *id, int idlen) {
    const int L = M->read_length;
    for (int p = 0; p < L; ++p) {
        double u = draw_u(d, rstream(o, K_QUAL), (uint32_t)p >> 1, 2u * ((uint32_t)p & 1) + 1);
        int cu = up(m[p]);
        if (u > M->phred_prob[q[p]] && !is_ambiguous_upper(cu)) {
            int code = base_code(cu);
            if (code < 0) return ISS_ORACLE_E_CHAR;
            size_t e = ((size_t)o * L + p) * 4 + code;
            int na = M->subst_n[e];
            double uc = draw_u(d, rstream(o, K_SUBC), (uint32_t)p >> 2, (uint32_t)p & 3);
            int idx = searchsorted_right(M->subst_cdf + e * ISS_ORACLE_MAX_ALT, na, uc);
            if (idx >= na) idx = na - 1; /* unreachable: last cdf entry is 1.0 > u */
            uint8_t alt = M->subst_alt[e * ISS_ORACLE_MAX_ALT + idx];
            if (M->store_mutations && alt != orig[p])
                vcf_emit(vcf, id, idlen, p, (const char *)&m[p], 1, (const char *)&alt, 1, q[p]);
            m[p] = alt;
        }
    }
    return ISS_ORACLE_OK;
}

static int u64_to_dec(uint64_t v, char *out) {
    char tmp[24]; int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    for (int i = 0; i < n; ++i) out[i] = tmp[n - 1 - i];
    return n;
}

/* Number of FASTQ bytes of one read of pair i: "@{hdr}_{i}_{cpu}/1\n{seq}\n+\n{qual}\n" */
static uint64_t record_bytes(int hdrlen, uint64_t i, uint64_t cpu, int L) {
    char t[24];
    return (uint64_t)(1 + hdrlen + 1 + u64_to_dec(i, t) + 1 + u64_to_dec(cpu, t) + 2 + 1 + L + 1 + 2 + L + 1);
}

uint64_t iss_oracle_fastq_bytes(int32_t hdrlen, uint64_t cpu, int32_t L, uint64_t i0, uint64_t i1) {
    uint64_t s = 0;
    for (uint64_t i = i0; i < i1; ++i) s += record_bytes(hdrlen, i, cpu, L);
    return s;
}

/*
 * One pair.  Returns ISS_ORACLE_OK and writes the two FASTQ records, or an error code.
 * gc (optional): receives 1 when the gc_bias filter of reads_generator (generator.py:82-92) keeps the pair.
 */
static int simulate_pair(const iss_oracle_model *M, const uint8_t *rec, int64_t reclen, const char *hdr, int hdrlen,
                         uint64_t key, uint64_t P, uint32_t attempt, uint64_t i, uint64_t cpu, int amplicon,
                         uint8_t *out1, uint8_t *out2, uint64_t *nbytes, vcf_sink *vcf, uint8_t *work /* 2*(5L+8)+4L */) {
    const int L = M->read_length;
    draw_ctx d;
    ctx_init(&d, key, P, attempt);
    if (!(L < reclen)) return ISS_ORACLE_E_SHORT;

    int64_t insert_size, fragment_length;
    if (M->frag_mode) {
        uint64_t r64 = draw_r64(&d, S_PAIR, 2, 2);
        int lo = 0, hi = M->n_frag;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (M->frag_thr[mid] <= r64) lo = mid + 1; else hi = mid; }
        fragment_length = M->frag_lo + lo;
        insert_size = fragment_length - 2 * (int64_t)L;
    } else {
        if (M->fixed_insert >= 0) insert_size = M->fixed_insert;
        else insert_size = searchsorted_left(M->insert_cdf, M->n_insert, draw_u(&d, S_PAIR, 0, 0));
        fragment_length = insert_size + 2 * (int64_t)L;
    }
    int64_t fs;
    if (amplicon) fs = 0;
    else {
        int64_t width = reclen - fragment_length;
        if (width > 0) fs = (int64_t)randbelow(draw_r64(&d, S_PAIR, 1, 0), (uint64_t)width);
        else fs = (int64_t)randbelow(draw_r64(&d, S_PAIR, 1, 2), (uint64_t)(reclen - L)); /* ValueError fallback :142-144 */
    }
    int64_t fe = fs + L;
    int64_t rs, re;
    if (amplicon) { rs = reclen - L; re = rs + L; }
    else { rs = fe + insert_size; re = rs + L; }
    if (re > reclen) { /* :172-176 */
        re = L + (int64_t)randbelow(draw_r64(&d, S_PAIR, 2, 0), (uint64_t)(reclen - L));
        rs = re - L;
    }
    if (rs < 0 || fe > reclen) return ISS_ORACLE_E_RANGE;

    char id[512];
    if (hdrlen > 400) return ISS_ORACLE_E_ARG;
    memcpy(id, hdr, (size_t)hdrlen);
    int idlen = hdrlen;
    id[idlen++] = '_'; idlen += u64_to_dec(i, id + idlen);
    id[idlen++] = '_'; idlen += u64_to_dec(cpu, id + idlen);
    id[idlen++] = '/'; id[idlen++] = '1';

    uint8_t *m = work;                 /* mutable sequence, capacity 5L+8 */
    uint8_t *orig = work + (5 * L + 8);/* "original" annotation */
    uint8_t *q = orig + L;

    for (int o = 0; o < 2; ++o) {
        int64_t s = o ? rs : fs, e = o ? re : fe;
        if (o == 0) memcpy(m, rec + s, (size_t)L);
        else {
            for (int k = 0; k < L; ++k) {
                int c = complement(rec[e - 1 - k]);
                if (!c) return ISS_ORACLE_E_CHAR;
                m[k] = (uint8_t)c;
            }
        }
        memcpy(orig, m, (size_t)L);
        id[idlen - 1] = (char)('1' + o);
        int rc = introduce_indels(M, &d, o, m, rec, reclen, s, e, vcf, id, idlen);
        if (rc) return rc;
        rc = gen_phred(M, &d, o, q);
        if (rc) return rc;
        rc = mut_sequence(M, &d, o, m, orig, q, vcf, id, idlen);
        if (rc) return rc;
        uint8_t *out = o ? out2 : out1;
        uint64_t n = 0;
        out[n++] = '@'; memcpy(out + n, id, (size_t)idlen); n += (uint64_t)idlen; out[n++] = '\n';
        memcpy(out + n, m, (size_t)L); n += (uint64_t)L;
        out[n++] = '\n'; out[n++] = '+'; out[n++] = '\n';
        for (int k = 0; k < L; ++k) out[n++] = (uint8_t)(q[k] + 33);
        out[n++] = '\n';
        *nbytes = n;
    }
    return ISS_ORACLE_OK;
}

/*
 * Generate pairs i0..i1-1 of one work item (reads_generator + simulate_reads, generator.py:21-95).
 *   pair_base : ordinal of pair i=0 of this item inside its worker (RNG contract v2)
 *   out1/out2 : FASTQ sinks of capacity cap bytes each; (len1, len2) receive bytes written
 *   vcf       : optional VCF sink (store_mutations), capacity vcf_cap
 * gc_bias follows generator.py:82-92 literally: gc_fraction is 0..1 so `40 < gc < 60` never holds and the
 * pair is kept iff the PAIR.blk0.w3 draw < 0.90; a rejected pair is re-simulated with attempt+1.
 * Returns 0, or the reference's AssertionError condition (ISS_ORACLE_E_SHORT: caller logs + skips the record).
 */
int iss_oracle_generate(const iss_oracle_model *M, const uint8_t *rec, int64_t reclen, const char *hdr,
                        int32_t hdrlen, uint64_t key, uint64_t pair_base, uint64_t cpu, int32_t amplicon,
                        int32_t gc_bias, uint64_t i0, uint64_t i1, uint8_t *out1, uint8_t *out2, uint64_t cap,
                        uint64_t *len1, uint64_t *len2, char *vcf_buf, uint64_t vcf_cap, uint64_t *vcf_len) {
    const int L = M->read_length;
    uint8_t *work = (uint8_t *)malloc((size_t)(5 * L + 8) + 2 * (size_t)L + 16);
    if (!work) return ISS_ORACLE_E_SPACE;
    vcf_sink vs = {vcf_buf, vcf_cap, 0, 0};
    uint64_t n = 0;
    int rc = ISS_ORACLE_OK;
    for (uint64_t i = i0; i < i1; ++i) {
        uint64_t need = record_bytes(hdrlen, i, cpu, L);
        if (n + need > cap) { rc = ISS_ORACLE_E_SPACE; break; }
        uint32_t attempt = 0;
        for (;;) {
            uint64_t nb = 0;
            uint64_t vcf_mark = vs.len;
            rc = simulate_pair(M, rec, reclen, hdr, hdrlen, key, pair_base + i, attempt, i, cpu, amplicon, out1 + n,
                               out2 + n, &nb, M->store_mutations ? &vs : NULL, work);
            if (rc) break;
            if (gc_bias) {
                uint32_t ctr_out[4];
                philox4x32_10(0, S_PAIR | (attempt << 8), (uint32_t)(pair_base + i), (uint32_t)((pair_base + i) >> 32),
                              (uint32_t)key, (uint32_t)(key >> 32), ctr_out);
                double u = (double)ctr_out[3] * (1.0 / 4294967296.0);
                if (!(u < 0.90)) { attempt++; vs.len = vcf_mark; continue; } /* rejected: mutations never written */
            }
            n += nb;
            break;
        }
        if (rc) break;
    }
    free(work);
    *len1 = n;
    *len2 = n;
    if (vcf_len) *vcf_len = vs.len;
    if (!rc && vs.overflow) rc = ISS_ORACLE_E_SPACE;
    return rc;
}

int iss_oracle_abi_version(void) { return 1; }
