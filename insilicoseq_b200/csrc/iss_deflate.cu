// iss_deflate.cu -- gzip on the GPU for `--compress` (iss/util.py:255-268 gzips the finished files on one CPU core).
//
// The FASTQ text of a batch already sits in HBM; compressing it there is the only way the end-to-end rate can exceed
// what PCIe can carry as plain text.  Format: a gzip file is a sequence of members (RFC 1952); every 64 KB chunk of
// a stream becomes one member holding ONE dynamic-Huffman DEFLATE block (RFC 1951): literals, plus one LZ77 match per
// FASTQ record -- the name line "@name_i_cpu/1" against the previous record's (distance = that record's length; the
// lines differ in their last digits only), found by pattern ("\n@" not preceded by "\n+") and verified byte by byte,
// so any input stays valid, and never reaching back across a member boundary:
//
//   10 B gzip header | BFINAL=1 BTYPE=10 HLIT=29 HDIST=29 HCLEN=15 | 19 x 3 bit code-length-code lengths (0..15 -> 4 bits,
//   16/17/18 unused) | 286 literal/length + 30 distance code lengths, 4 bits each | tokens | EOB | pad | CRC32 | ISIZE
//
// so the block header has a fixed size (1338 bits) and every token's bit position follows from a prefix sum.
// Per call (= one FASTQ stream of one batch): byte histogram of a <= 4 MB sample + the match statistics of the first
// 16 KB -> ONE literal/length code and ONE distance code, lengths limited to 15 bits (two-queue merge + zlib's overflow
// repair, always complete codes; every symbol keeps a code) -> canonical codes + block header.  Then one CTA per chunk:
// name-line matches planned per 128-byte slice -> bit count per slice -> prefix sum -> bit packing
// into shared memory -> CRC-32 from per-slice raw CRCs (slicing-by-4) advanced by x^(8 * bytes_after) (GF(2)
// multiply, zlib's crc32_combine arithmetic) -> member written at its final offset, found by a decoupled look-back over
// the member sizes (no staging copy).  Incompressible chunks fall back to stored blocks.
// Validated by decompression (tests): gunzip(output) == input, for FASTQ and for arbitrary bytes.
#include <cuda_runtime.h>
#include <stdint.h>

#include "iss_deflate.h"

namespace {

constexpr uint32_t kChunk = ISS_GZ_CHUNK;
constexpr uint32_t kOutWords = 12288;  // 48 KB of packed bits per chunk in shared memory; beyond that: stored block
constexpr uint32_t kThreads = 512;     // threads of iss_deflate_kernel
constexpr uint32_t kSlice = kChunk / kThreads;  // bytes per thread: 128
constexpr uint32_t kPiece = 16384;     // granule of the histogram sample
constexpr uint32_t kLitLen = 286, kDist = 30;  // symbols of the two alphabets (HLIT = 29, HDIST = 29)
constexpr uint32_t kHeaderBits = 17 + 19 * 3 + (kLitLen + kDist) * 4;  // 1338
constexpr uint32_t kHdrWords = (kHeaderBits + 31) / 32;                // 42
constexpr uint32_t kMaxTok = 512;      // name-line matches per member kept (a 64 KB chunk of 150 bp FASTQ holds ~200 records)
constexpr uint32_t kSampleLz = 16384;  // bytes of the stream whose matches give the length / distance statistics

// RFC 1951 3.2.5: length codes 257..285 and distance codes 0..29
__constant__ uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073,
                                       4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

__device__ __forceinline__ uint32_t len_code(uint32_t len) {  // 3..258 -> 0..28
    uint32_t k = 28;
    while (len < kLenBase[k]) --k;
    return k;
}
__device__ __forceinline__ uint32_t dist_code(uint32_t dist) {  // 1..32768 -> 0..29 (zlib's d_code)
    if (dist <= 4u) return dist - 1u;
    const uint32_t d = dist - 1u, hb = 31u - (uint32_t)__clz((int)d);
    return 2u * hb + ((d >> (hb - 1u)) & 1u);
}
// "@..." at x starts a FASTQ record: line start, and the line before is not the "+" separator (a quality line may begin with '@')
__device__ __forceinline__ bool is_record_start(const uint8_t *__restrict__ b, uint32_t x) {
    return x >= 3u && b[x] == '@' && b[x - 1u] == '\n' && !(b[x - 2u] == '+' && b[x - 3u] == '\n');
}
// common prefix of the name lines at s and q (q < s), newline excluded, as an LZ77 match length (0: no match)
__device__ __forceinline__ uint32_t name_match(const uint8_t *__restrict__ b, uint32_t n, uint32_t s, uint32_t q) {
    if (s - q > 32768u) return 0u;
    uint32_t m = 0;
    while (s + m + 4u <= n && m + 4u <= 256u) {  // four bytes of each line per step: eight independent loads
        const uint32_t a0 = b[s + m], a1 = b[s + m + 1u], a2 = b[s + m + 2u], a3 = b[s + m + 3u];
        const uint32_t c0 = b[q + m], c1 = b[q + m + 1u], c2 = b[q + m + 2u], c3 = b[q + m + 3u];
        if (a0 != c0 || a0 == '\n') break;
        if (a1 != c1 || a1 == '\n') { m += 1u; goto done; }
        if (a2 != c2 || a2 == '\n') { m += 2u; goto done; }
        if (a3 != c3 || a3 == '\n') { m += 3u; goto done; }
        m += 4u;
    }
    while (s + m < n && m < 258u && b[s + m] == b[q + m] && b[s + m] != '\n') ++m;
done:
    return m >= 4u ? m : 0u;  // 4, not DEFLATE's 3: a match then always reaches the end of the 32-bit word it starts in (see the packing loop)
}
// 0x80 in every byte of x that is zero (exact for any x: no borrow between bytes)
__device__ __forceinline__ uint32_t zero_bytes(uint32_t x) { return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u; }
// 4 bits of a word's byte flags (0x80 per byte, as zero_bytes returns them): bit j = byte j
__device__ __forceinline__ uint32_t byte_flags4(uint32_t f) { return (((f >> 7) * 0x01020408u) >> 24) & 0xFu; }
// the bits of m that lie in runs of at least 4 consecutive ones (m: 64 positions)
__device__ __forceinline__ uint64_t runs4(uint64_t m) {
    const uint64_t a = m & (m >> 1), b = a & (a >> 2);  // b: four ones start here
    const uint64_t c = b | (b << 1);
    return c | (c << 2);
}
// bits [a, b) of a 128-bit mask held in four words (a <= b <= 128)
__device__ __forceinline__ void set_range128(uint32_t m[4], uint32_t a, uint32_t b) {
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) {
        const uint32_t lo = a > 32u * j ? a - 32u * j : 0u, hi = b > 32u * j ? b - 32u * j : 0u;  // clamped below; above: >= 32 means all
        const uint32_t mh = hi >= 32u ? 0xFFFFFFFFu : ((1u << hi) - 1u), ml = lo >= 32u ? 0xFFFFFFFFu : ((1u << lo) - 1u);
        m[j] |= mh & ~ml;
    }
}
constexpr uint32_t kPoly = 0xedb88320u;

__device__ __forceinline__ uint32_t multmodp(uint32_t a, uint32_t b) {  // a * b mod P over GF(2), reflected (zlib)
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ kPoly : b >> 1;
    }
    return p;
}

struct GzTables {
    uint32_t crc[4][256];  // slicing-by-4 CRC-32 tables: crc[k][i] = CRC of byte i followed by k zero bytes
    uint32_t x2n[32];      // x^(2^n) mod P
    uint32_t pow_slice[kChunk / kSlice];  // x^(8 * kSlice * j) mod P
    uint32_t init_full;    // 0xFFFFFFFF * x^(8 * kChunk) mod P
};
__device__ GzTables g_gz;

__device__ uint32_t x2nmodp(uint32_t n, uint32_t k) {  // x^(n * 2^k) mod P
    uint32_t p = 1u << 31;
    while (n) {
        if (n & 1u) p = multmodp(g_gz.x2n[k & 31u], p);
        n >>= 1;
        k++;
    }
    return p;
}

// The code of one iss_gz_compress call, built once from a sample of the stream and shared by all its members
// (FASTQ of one batch is statistically uniform; any complete prefix code gives a valid member).
struct GzCode {
    uint32_t cl[288];   // bit-reversed canonical code | length << 16: literals 0..255, 256 = end of block, 257..285 = match lengths
    uint32_t dl[32];    // the same for the 30 distance codes
    uint32_t hdr[44];   // the kHeaderBits bits of the block header (lengths included)
    uint32_t rmin;      // shortest distance-1 run that becomes a token: 4, or 0 = none (the sample says they would not pay)
};
constexpr uint32_t kWorkHist = 0;     // word offsets inside the caller's work area (ISS_GZ_WORK_WORDS u32)
constexpr uint32_t kWorkTicket = 260;  // next chunk to hand out
constexpr uint32_t kWorkZeroed = 264;  // words cleared at the start of a call
constexpr uint32_t kWorkCode = 264;
static_assert(kWorkCode + sizeof(GzCode) / 4 <= ISS_GZ_WORK_WORDS, "work area too small");

struct TableSmem {
    uint32_t out[44];
    uint32_t freq[288];
    uint32_t keys[288];
    uint32_t w[576];      // node weights during the merge, then node depths
    uint32_t bl[17];      // leaves per code length
    uint32_t first[17];   // first canonical code per length
    uint32_t total, overflow;
    uint16_t parent[576];
    uint32_t cl[288];
    uint8_t lens[288];
    uint32_t tokf[64];    // sample statistics: [0..28] match-length codes, [32..61] distance codes
    int32_t start[256];   // sample planning: record start inside each 64-byte slice of the first 16 KB (-1: none)
    uint32_t sub[256];    // sample statistics: bytes of each value that the sample's runs cover (they are not literals)
    int32_t gain;         // bits the sample's runs would save as tokens, at ~ -log2 p bits per literal and 10 per token
    uint32_t wbits;       // sum over the byte values of count * ~ -log2 p
    uint8_t lg[256];      // ~ -log2 p of each byte value, 1..15
};

__device__ __forceinline__ void put_bits(uint32_t *out, uint32_t pos, uint32_t v, uint32_t nbits) {
    const uint32_t wi = pos >> 5, sh = pos & 31u;
    atomicOr(&out[wi], v << sh);
    if (sh + nbits > 32u) atomicOr(&out[wi + 1], v >> (32u - sh));
}

// thread 0: Huffman tree over the sorted leaves w[0..n) by the two-queue merge (leaves [a, n) and internal nodes [b, nxt)
// are both sorted; on a tie the leaf goes first).  Both queue heads live in registers.
__device__ void merge_tree(TableSmem &s, uint32_t n) {
    const uint32_t kInf = 0xFFFFFFFFu;
    uint32_t a = 0, b = n, nxt = n;
    uint32_t wa = s.w[0], wb = kInf;
    for (uint32_t it = 0; it + 1 < n; ++it) {
        uint32_t sum = 0;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (wa <= wb) {
                s.parent[a] = (uint16_t)nxt;
                sum += wa;
                ++a;
                wa = a < n ? s.w[a] : kInf;
            } else {
                s.parent[b] = (uint16_t)nxt;
                sum += wb;
                ++b;
                wb = b < nxt ? s.w[b] : kInf;
            }
        }
        s.w[nxt] = sum;
        if (b == nxt) wb = sum;
        ++nxt;
    }
}

// Placement of the members without a second pass (decoupled look-back, Merrill & Garland 2016): status[c] holds the
// member size of chunk c tagged "aggregate" as soon as it is known (before the packing), later the running total up to
// and including c tagged "inclusive".  Warp 0 of a CTA sums its predecessors' aggregates back to the nearest inclusive
// total.  Chunks are handed out by an atomic ticket, so every predecessor is resident or finished: the wait ends.
// A wait longer than kSpinLimit cycles marks the stream as failed instead of hanging the GPU.
#define kAgg (1ull << 62)
#define kInc (2ull << 62)
#define kFail (1ull << 61)
#define kValue ((1ull << 61) - 1ull)
#define kSpinLimit 4000000000ll

// Decoupled look-back over the member sizes, in two steps so that the wait never sits in front of the packing: the size is
// published as soon as it is known (`publish`), the exclusive offset is resolved by warp 0 when its own packing is done --
// by then the predecessors' sizes are there.
__device__ __forceinline__ void look_back_publish(volatile unsigned long long *status, uint32_t c, uint32_t msize) {
    status[c] = (c == 0 ? kInc : kAgg) | msize;
}
__device__ uint64_t look_back_resolve(volatile unsigned long long *status, uint32_t c, uint32_t msize, uint32_t lane) {
    if (c == 0) return 0;
    uint64_t excl = 0;
    int64_t j = (int64_t)c - 1;
    const long long t0 = clock64();
    for (;;) {
        const int64_t idx = j - (int64_t)lane;
        const unsigned long long v = idx >= 0 ? status[idx] : kInc;  // before chunk 0: an inclusive total of 0
        const uint32_t flag = (uint32_t)(v >> 62);
        const uint32_t inc = __ballot_sync(0xFFFFFFFFu, flag == 2u), inv = __ballot_sync(0xFFFFFFFFu, flag == 0u);
        const uint32_t upto = inc ? (uint32_t)(__ffs((int)inc) - 1) : 31u;  // lanes 0..upto = the predecessors still needed
        const uint32_t need = upto == 31u ? 0xFFFFFFFFu : ((2u << upto) - 1u);
        if (inv & need) {
            if (clock64() - t0 > kSpinLimit) { excl = kFail; break; }
            continue;
        }
        uint64_t val = lane <= upto ? (v & (kValue | kFail)) : 0;
        for (uint32_t d = 16; d; d >>= 1) val += __shfl_xor_sync(0xFFFFFFFFu, val, d);
        excl += val;
        if (inc) break;
        j -= 32;
    }
    if (excl >= kFail) excl = kFail;
    if (lane == 0) status[c] = kInc | (excl == kFail ? kFail : excl + msize);
    return excl;
}

}  // namespace

// ---- 1. byte histogram of a sample: every piece_stride-th 16 KB piece of the stream (at most 256 pieces = 4 MB) ----------
__global__ void __launch_bounds__(256) iss_gz_hist_kernel(const uint8_t *__restrict__ in, uint64_t n_in, uint32_t piece_stride,
                                                          uint32_t *__restrict__ work) {
    __shared__ uint32_t h[8 * 257];
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid; i < 8 * 257; i += 256) h[i] = 0;
    __syncthreads();
    const uint64_t n_pieces = (n_in + kPiece - 1) / kPiece;
    uint32_t *hw = h + warp * 257;
    for (uint64_t pc = (uint64_t)blockIdx.x * piece_stride; pc < n_pieces; pc += (uint64_t)gridDim.x * piece_stride) {
        const uint64_t base = pc * kPiece;
        const uint32_t n = (uint32_t)((n_in - base < kPiece) ? (n_in - base) : kPiece);
        const uint8_t *src = in + base;  // 16-byte aligned: the stream base is, and kPiece is a multiple of 16
        for (uint32_t i = tid * 16u; i < n; i += 256u * 16u) {
            if (i + 16u <= n) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + i));
                const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (uint32_t k = 0; k < 16; ++k) atomicAdd(&hw[(wds[k >> 2] >> (8u * (k & 3u))) & 0xFFu], 1u);
            } else {
                for (uint32_t k = i; k < n; ++k) atomicAdd(&hw[src[k]], 1u);
            }
        }
    }
    __syncthreads();
    uint32_t f = 0;
    for (uint32_t w = 0; w < 8; ++w) f += h[w * 257 + tid];
    if (f) atomicAdd(&work[kWorkHist + tid], f);
}

// ---- 2. one CTA: the two Huffman codes of the call (every symbol keeps a code: frequency = sample count + 1) ------------
// Lengths are limited to 15 bits the way zlib's gen_bitlen does it: depth capped at 15 while descending, `overflow` =
// nodes deeper than that, then the repair loop on the per-length counts and lengths re-dealt to the leaves in frequency
// order (the rarest get the longest codes); always a complete code.
// In: s.freq[0..n).  Out: s.lens[0..n), s.cl[0..n) (bit-reversed code | length << 16).  256 threads, n <= 288.
__device__ void build_code(TableSmem &s, uint32_t n, uint32_t tid) {
    const uint32_t root = 2 * n - 2;
    if (tid == 0) s.overflow = 0;
    if (tid < 17) s.bl[tid] = 0;
    __syncthreads();
    // rank sort by (frequency, symbol); frequencies stay below 2^23 (sample <= 4 MB)
    for (uint32_t sym = tid; sym < n; sym += 256) {
        const uint32_t key = (s.freq[sym] << 9) | sym;
        uint32_t rank = 0;
        for (uint32_t o = 0; o < n; ++o) rank += (((s.freq[o] << 9) | o) < key) ? 1u : 0u;
        s.keys[rank] = key;
    }
    __syncthreads();
    for (uint32_t i = tid; i < n; i += 256) s.w[i] = s.keys[i] >> 9;
    __syncthreads();
    if (tid == 0) merge_tree(s, n);
    __syncthreads();
    // node depths by walking to the root (w[] is free now)
    for (uint32_t i = tid; i < root; i += 256) {
        uint32_t d = 0;
        for (uint32_t j = i; j != root; j = s.parent[j]) ++d;
        if (d > 15u) { d = 15u; atomicAdd(&s.overflow, 1u); }
        if (i < n) { s.w[i] = d; atomicAdd(&s.bl[d], 1u); }
    }
    __syncthreads();
    const bool repaired = s.overflow != 0;
    if (tid == 0 && repaired) {
        int32_t overflow = (int32_t)s.overflow;
        do {
            uint32_t bits = 14;
            while (s.bl[bits] == 0) bits--;
            s.bl[bits]--;
            s.bl[bits + 1] += 2;
            s.bl[15]--;
            overflow -= 2;
        } while (overflow > 0);
    }
    __syncthreads();
    for (uint32_t i = tid; i < n; i += 256) {
        uint32_t len = s.w[i];
        if (repaired) {  // leaves are sorted by increasing frequency: the first bl[15] get 15 bits, the next bl[14] 14, ...
            uint32_t cum = 0;
            for (uint32_t bits = 15; bits >= 1; --bits) {
                cum += s.bl[bits];
                if (i < cum) { len = bits; break; }
            }
        }
        s.lens[s.keys[i] & 511u] = (uint8_t)len;
    }
    if (tid == 0) {  // canonical codes: first code of each length
        uint32_t code = 0, prev = 0;
        for (uint32_t b = 1; b <= 15; ++b) {
            code = (code + prev) << 1;
            s.first[b] = code;
            prev = s.bl[b];
        }
    }
    __syncthreads();
    for (uint32_t sym = tid; sym < n; sym += 256) {  // stored bit-reversed: DEFLATE packs Huffman codes MSB first
        const uint32_t l = s.lens[sym];
        uint32_t idx = 0;
        for (uint32_t o = 0; o < sym; ++o) idx += (s.lens[o] == l) ? 1u : 0u;
        s.cl[sym] = (__brev(s.first[l] + idx) >> (32u - l)) | (l << 16);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) iss_gz_table_kernel(const uint8_t *__restrict__ in, uint64_t n_in, uint32_t *__restrict__ work) {
    __shared__ TableSmem s;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) s.total = 0;
    if (tid < 44) s.out[tid] = 0;
    if (tid < 64) s.tokf[tid] = 0;
    s.sub[tid] = 0;
    if (tid == 0) { s.gain = 0; s.wbits = 0; }
    __syncthreads();
    const uint32_t f0 = work[kWorkHist + tid];
    atomicAdd(&s.total, f0);
    s.freq[tid] = f0 + 1u;
    // match statistics from the first 16 KB of the stream: the record starts of each 64-byte slice, each name line
    // against the one before it -- the same plan the member kernel makes
    const uint32_t ns = (uint32_t)(n_in < kSampleLz ? n_in : kSampleLz);
    {
        int32_t st = -1;
        for (uint32_t x = tid * 64u; x < tid * 64u + 64u && x < ns; ++x)
            if (is_record_start(in, x)) st = (int32_t)x;
        s.start[tid] = st;
    }
    __syncthreads();
    {
        int32_t prev = -1;
        for (int32_t t = (int32_t)tid - 1; t >= 0 && prev < 0; --t) prev = s.start[t];
        // a slice holds at most a few records of very short reads: the statistics take the last start of each slice only
        if (s.start[tid] >= 0 && prev >= 0) {
            const uint32_t m = name_match(in, ns, (uint32_t)s.start[tid], (uint32_t)prev);
            if (m) {
                atomicAdd(&s.tokf[len_code(m)], 1u);
                atomicAdd(&s.tokf[32u + dist_code((uint32_t)(s.start[tid] - prev))], 1u);
            }
        }
    }
    {   // the distance-1 runs of the same window (the member kernel's rule, name lines not taken out: statistics only).
        // Do they pay?  A literal of value c costs about -log2 p(c) bits, a run token about 10: broad quality distributions
        // (MiSeq) have many short runs of symbols that cost 1-2 bits anyway -- tokens would save ~ 1 % of the bits and
        // make the packing loop diverge at every other word -- while binned qualities (NovaSeq, NextSeq, HiSeq X) are
        // made of long runs and shrink by a fifth.  Runs become tokens when the estimate exceeds 1/64 of the bits.
        uint64_t m = 0;
        for (uint32_t j = 0; j < 64u; ++j) {
            const uint32_t x = tid * 64u + j;
            if (x >= 1u && x < ns && in[x] == in[x - 1u]) m |= 1ull << j;
        }
        __syncthreads();  // s.total complete
        {
            const uint32_t q = s.total / (f0 ? f0 : 1u);
            const uint32_t lg = q > 1u ? 31u - (uint32_t)__clz((int)q) : 1u;
            s.lg[tid] = (uint8_t)(lg > 15u ? 15u : lg);
            atomicAdd(&s.wbits, f0 * s.lg[tid]);
        }
        __syncthreads();
        const uint64_t e = runs4(m);
        {
            uint64_t st = e & ~(e << 1);
            int32_t g = 0;
            while (st) {
                const uint32_t x = (uint32_t)__ffsll((long long)st) - 1u;
                st &= st - 1ull;
                const uint64_t nz = ~(e >> x);
                const uint32_t r = nz ? (uint32_t)__ffsll((long long)nz) - 1u : 64u;
                g += (int32_t)(r * s.lg[in[tid * 64u + x]]) - 10;
            }
            if (g) atomicAdd(&s.gain, g);
        }
        __syncthreads();
        // gain / (bits of the window) > 1/64, the window's bits being ns * wbits / total
        const bool run_tokens = s.gain > 0 && (uint64_t)s.gain * 64ull * s.total > (uint64_t)ns * s.wbits;
        if (tid == 0) reinterpret_cast<GzCode *>(work + kWorkCode)->rmin = run_tokens ? 4u : 0u;
        uint64_t st = run_tokens ? e & ~(e << 1) : 0ull;
        while (st) {
            const uint32_t x = (uint32_t)__ffsll((long long)st) - 1u;
            st &= st - 1ull;
            const uint64_t nz = ~(e >> x);
            const uint32_t r = nz ? (uint32_t)__ffsll((long long)nz) - 1u : 64u;
            atomicAdd(&s.tokf[len_code(r)], 1u);
            atomicAdd(&s.tokf[32u], 1u);  // distance 1 = distance code 0
            atomicAdd(&s.sub[in[tid * 64u + x]], r);
        }
    }
    __syncthreads();
    const uint32_t scale = ns ? (s.total / ns > 1u ? s.total / ns : 1u) : 1u;  // sampled bytes per byte of the 16 KB window
    if (tid == 0) s.freq[256] = (s.total >> 16) ? (s.total >> 16) : 1u;  // end of block: one per 64 KB member
    if (tid < 29) s.freq[257 + tid] = s.tokf[tid] * scale + 1u;
    {
        const uint32_t covered = s.sub[tid] * scale, f = s.freq[tid];
        s.freq[tid] = f > covered + 1u ? f - covered : 1u;
    }
    __syncthreads();
    build_code(s, kLitLen, tid);
    GzCode *code = reinterpret_cast<GzCode *>(work + kWorkCode);
    for (uint32_t sym = tid; sym < 288; sym += 256) code->cl[sym] = sym < kLitLen ? s.cl[sym] : 0u;
    // the 286 literal/length lengths, each as the 4-bit code of its value (flat code-length code, MSB first)
    for (uint32_t sym = tid; sym < kLitLen; sym += 256) put_bits(s.out, 74u + 4u * sym, __brev((uint32_t)s.lens[sym]) >> 28, 4);
    __syncthreads();
    if (tid < kDist) s.freq[tid] = s.tokf[32u + tid] * scale + 1u;
    __syncthreads();
    build_code(s, kDist, tid);
    if (tid < 32) code->dl[tid] = tid < kDist ? s.cl[tid] : 0u;
    if (tid < kDist) put_bits(s.out, 74u + 4u * (kLitLen + tid), __brev((uint32_t)s.lens[tid]) >> 28, 4);
    if (tid == 0) {
        // constant part of the block header: BFINAL=1, BTYPE=2, HLIT=29, HDIST=29, HCLEN=15, then the 19 code-length-code
        // lengths in the order 16,17,18,0,8,7,...: three zeros and sixteen fours
        uint32_t pos = 0;
        put_bits(s.out, pos, 1u, 1); pos += 1;
        put_bits(s.out, pos, 2u, 2); pos += 2;
        put_bits(s.out, pos, 29u, 5); pos += 5;
        put_bits(s.out, pos, 29u, 5); pos += 5;
        put_bits(s.out, pos, 15u, 4); pos += 4;
        pos += 9;   // lengths of 16, 17, 18 = 0
        for (int k = 0; k < 16; ++k) { put_bits(s.out, pos, 4u, 3); pos += 3; }
    }
    __syncthreads();
    if (tid < 44) code->hdr[tid] = s.out[tid];
}

// ---- 3. the members: one CTA per 64 KB chunk, 512 threads x 128-byte slices ----------------------------------------------
struct Smem {
    uint32_t out[kOutWords];
    uint32_t crcT[4][256];
    uint32_t cl[288];
    uint32_t dl[32];
    uint32_t ticket;         // the chunk being compressed
    uint32_t misc[64];       // 3 crc accumulator, 4 x^(8*(n % kSlice)), 5/6 member offset, 7 matches, 8.. warp sums
    uint32_t tok_pos[kMaxTok];   // the member's matches in stream order: start inside the chunk,
    uint32_t tok_ld[kMaxTok];    //   length | distance << 16 (distance 32768 stored as 0)
    int32_t last_start[kThreads];  // last record start inside each slice (-1: none)
    uint4 cov[kThreads];     // per slice, one bit per byte: first "equals the byte in front of it", then "covered by a match"
    uint16_t tok_off[kThreads + 1];  // matches that start before each slice
    uint8_t len8[288];
    uint32_t ltok[256];      // match length - 3 -> the length part of its token: code | extra bits << code length, bit count << 24
};

extern __shared__ __align__(16) uint8_t iss_gz_smem[];

namespace {
// bits of one match: length code + extra, distance code + extra
__device__ __forceinline__ uint32_t match_bits(const Smem &s, uint32_t len, uint32_t dist) {
    const uint32_t dc = dist_code(dist);
    return (s.ltok[len - 3u] >> 24) + (s.dl[dc] >> 16) + kDistExtra[dc];
}
__device__ __forceinline__ void flush_acc(Smem &s, uint64_t &acc, uint32_t &accbits, uint32_t &wi) {
    if (accbits >= 32u) {
        atomicOr(&s.out[wi++], (uint32_t)acc);
        acc >>= 32;
        accbits -= 32u;
    }
}
}  // namespace

__global__ void __launch_bounds__(kThreads, 3) iss_deflate_kernel(const uint8_t *__restrict__ in, uint64_t n_in, uint8_t *__restrict__ out,
                                                                  unsigned long long *status, uint32_t *__restrict__ work,
                                                                  uint32_t n_chunks) {
    Smem &s = *reinterpret_cast<Smem *>(iss_gz_smem);
    const GzCode *code = reinterpret_cast<const GzCode *>(work + kWorkCode);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    for (uint32_t i = tid; i < 1024; i += kThreads) s.crcT[i >> 8][i & 255u] = g_gz.crc[i >> 8][i & 255u];
    for (uint32_t i = tid; i < 288; i += kThreads) {
        const uint32_t e = code->cl[i];
        s.cl[i] = e;
        s.len8[i] = (uint8_t)(e >> 16);
    }
    if (tid < 32) s.dl[tid] = code->dl[tid];
    if (tid < 256) {  // one shared-memory word per match length instead of two constant-bank loads with a per-lane index
        const uint32_t len = tid + 3u, lc = len_code(len), e = code->cl[257u + lc];
        s.ltok[tid] = (e & 0xFFFFu) | ((len - kLenBase[lc]) << (e >> 16)) | (((e >> 16) + kLenExtra[lc]) << 24);
    }
    const bool run_tokens = code->rmin != 0u;
    for (;;) {
        __syncthreads();  // previous chunk fully written out; tables loaded
        // A ticket is taken when its chunk is started, never ahead: the look-back of the chunks behind it waits for this
        // chunk's size, and a ticket held back for the length of another chunk makes every one of them wait that long
        // (measured: prefetching the ticket cost a factor of five in the pipeline).
        if (tid == 0) s.ticket = atomicAdd(&work[kWorkTicket], 1u);
        __syncthreads();
        const uint32_t c = s.ticket;
        if (c >= n_chunks) break;
        const uint64_t base = (uint64_t)c * kChunk;
        const uint32_t n = (uint32_t)((n_in - base < kChunk) ? (n_in - base) : kChunk);
        const uint8_t *src = in + base;
        for (uint32_t i = tid; i < kOutWords; i += kThreads) s.out[i] = i < kHdrWords ? code->hdr[i] : 0u;
        if (tid < 64) s.misc[tid] = 0;
        // 1. my slice (kSlice contiguous bytes, read 16 at a time): literal bit count + raw CRC (slicing-by-4) + record starts
        const uint32_t lo = tid * kSlice, hi = (lo + kSlice < n) ? lo + kSlice : n;
        const uint4 *src4 = reinterpret_cast<const uint4 *>(src + lo);  // chunk base and slices are 16-byte aligned
        uint32_t bits = 0, crc = 0;
        int32_t st0 = -1, st1 = -1;  // the last two record starts of the slice (st0 < st1)
        // the four bytes in front of the slice (slices behind the end of the chunk read nothing; the chunk's first slice has none)
        uint32_t prev_word = (lo && lo < hi) ? __ldg(reinterpret_cast<const uint32_t *>(src + lo - 4u)) : 0u;
        for (uint32_t i = lo; i < hi; i += 32u) {  // one 32-byte sector per step: both halves are asked for before either is used
            const uint4 va = __ldg(src4 + ((i - lo) >> 4)), vb = i + 16u < hi ? __ldg(src4 + ((i - lo) >> 4) + 1) : make_uint4(0u, 0u, 0u, 0u);
            const uint32_t wds[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
            uint32_t same32 = 0;
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k) {
                const uint32_t wd = wds[k];
                if (i + 4u * k + 4u <= hi) {
                    bits += (uint32_t)s.len8[wd & 0xFFu] + s.len8[(wd >> 8) & 0xFFu] + s.len8[(wd >> 16) & 0xFFu] + s.len8[wd >> 24];
                    const uint32_t x = crc ^ wd;
                    crc = s.crcT[3][x & 0xFFu] ^ s.crcT[2][(x >> 8) & 0xFFu] ^ s.crcT[1][(x >> 16) & 0xFFu] ^ s.crcT[0][x >> 24];
                } else {
                    for (uint32_t j = 0; j < 4u && i + 4u * k + j < hi; ++j) {  // the stream's last, ragged word
                        const uint32_t b = (wd >> (8u * j)) & 0xFFu;
                        bits += s.len8[b];
                        crc = s.crcT[0][(crc ^ b) & 0xFFu] ^ (crc >> 8);
                    }
                }
                // record starts, from registers and without a branch for all but the starts themselves: byte j of b1 / b2 /
                // b3 is the byte 1 / 2 / 3 places in front of byte j of wd; a start is '@' behind '\n' but not behind "\n+\n"
                const uint32_t b1 = __funnelshift_l(prev_word, wd, 8), b2 = __funnelshift_l(prev_word, wd, 16), b3 = __funnelshift_l(prev_word, wd, 24);
                if (run_tokens) same32 |= byte_flags4(zero_bytes(wd ^ b1)) << (4u * k);  // runs: byte == the byte in front of it
                uint32_t starts = zero_bytes(wd ^ 0x40404040u) & zero_bytes(b1 ^ 0x0A0A0A0Au) & ~(zero_bytes(b2 ^ 0x2B2B2B2Bu) & zero_bytes(b3 ^ 0x0A0A0A0Au));
                while (starts) {  // once per record
                    const uint32_t x = i + 4u * k + ((uint32_t)(__ffs((int)starts) - 1) >> 3);
                    starts &= starts - 1u;
                    if (x >= 3u && x < hi) { st0 = st1; st1 = (int32_t)x; }
                }
                prev_word = wd;
            }
            reinterpret_cast<uint32_t *>(&s.cov[tid])[(i - lo) >> 5] = same32;
        }
        for (uint32_t j = hi > lo ? (hi - lo + 31u) >> 5 : 0u; j < 4u; ++j) reinterpret_cast<uint32_t *>(&s.cov[tid])[j] = 0u;  // no data there
        s.last_start[tid] = st1;
        __syncthreads();  // out[] initialised, misc zeroed, record starts visible
        // 2. name-line matches that start in my slice: each record start against the one before it (inside this member)
        uint32_t mlen[2] = {0, 0}, mpos[2] = {0, 0}, mdist[2] = {0, 0}, nm = 0;
        if (st1 >= 0) {
            int32_t before = -1;  // the last record start in an earlier slice of this member
            for (int32_t t = (int32_t)tid - 1; t >= 0 && before < 0; --t) before = s.last_start[t];
            const int32_t first = st0 >= 0 ? st0 : st1;
            if (before >= 0) {
                const uint32_t m = name_match(src, n, (uint32_t)first, (uint32_t)before);
                if (m) { mpos[nm] = (uint32_t)first; mlen[nm] = m; mdist[nm] = (uint32_t)(first - before); ++nm; }
            }
            if (st0 >= 0) {  // two starts in this slice (reads of a few dozen bases): the second against the first
                const uint32_t m = name_match(src, n, (uint32_t)st1, (uint32_t)st0);
                if (m) { mpos[nm] = (uint32_t)st1; mlen[nm] = m; mdist[nm] = (uint32_t)(st1 - st0); ++nm; }
            }
        }
        // exclusive scan of the match counts -> the member's matches in stream order
        uint32_t cincl = nm;
        for (uint32_t d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, cincl, d);
            if (lane >= d) cincl += v;
        }
        if (lane == 31) s.misc[32 + warp] = cincl;
        __syncthreads();
        uint32_t cwarp = 0;
        for (uint32_t w = 0; w < warp; ++w) cwarp += s.misc[32 + w];
        const uint32_t my_first = cwarp + cincl - nm;  // matches starting before my slice
        for (uint32_t j = 0; j < nm; ++j)
            if (my_first + j < kMaxTok) { s.tok_pos[my_first + j] = mpos[j]; s.tok_ld[my_first + j] = mlen[j] | (mdist[j] << 16); }
        s.tok_off[tid] = (uint16_t)(my_first < kMaxTok ? my_first : kMaxTok);
        if (tid == kThreads - 1) {
            const uint32_t tot = my_first + nm;
            s.tok_off[kThreads] = (uint16_t)(tot < kMaxTok ? tot : kMaxTok);
        }
        __syncthreads();
        // 3. my slice's plan: the bytes the name-line matches cover, then distance-1 runs (>= 4 bytes equal to the byte in
        //    front of them: quality lines of binned instruments are mostly such runs) in what is left.  Runs are cut at the
        //    slice's end and in its middle, never touch a name-line match, and a match is never followed by another one
        //    without a literal in between -- so the cover mask alone tells where tokens start and how long a run is.
        const uint32_t t_begin = s.tok_off[tid], t_end = s.tok_off[tid + 1];
        uint32_t cov_end = lo;  // bytes [lo, cov_end) belong to a match that started in an earlier slice
        if (t_begin > 0) {
            const uint32_t e = s.tok_pos[t_begin - 1] + (s.tok_ld[t_begin - 1] & 0xFFFFu);
            if (e > lo) cov_end = e < hi ? e : hi;
        }
        uint32_t nc[4] = {0u, 0u, 0u, 0u};  // name-line coverage
        set_range128(nc, 0u, cov_end - lo);
        for (uint32_t x = lo; x < cov_end; ++x) bits -= s.len8[src[x]];
        for (uint32_t t = t_begin; t < t_end; ++t) {
            const uint32_t ps = s.tok_pos[t], ld = s.tok_ld[t], len = ld & 0xFFFFu, dist = (ld >> 16) ? (ld >> 16) : 32768u;
            const uint32_t e = ps + len < hi ? ps + len : hi;
            for (uint32_t x = ps; x < e; ++x) bits -= s.len8[src[x]];
            bits += match_bits(s, len, dist);
            set_range128(nc, ps - lo, e - lo);
        }
        {
            const uint4 sm = s.cov[tid];
            uint32_t vm[4] = {0u, 0u, 0u, 0u};
            set_range128(vm, lo ? 0u : 1u, hi > lo ? hi - lo : 0u);  // the member's first byte has nothing in front of it
            // candidates: not covered, and not next to a covered byte on either side
            const uint32_t r0 = sm.x & vm[0] & ~(nc[0] | (nc[0] << 1) | (nc[0] >> 1) | (nc[1] << 31));
            const uint32_t r1 = sm.y & vm[1] & ~(nc[1] | (nc[1] << 1) | (nc[0] >> 31) | (nc[1] >> 1) | (nc[2] << 31));
            const uint32_t r2 = sm.z & vm[2] & ~(nc[2] | (nc[2] << 1) | (nc[1] >> 31) | (nc[2] >> 1) | (nc[3] << 31));
            const uint32_t r3 = sm.w & vm[3] & ~(nc[3] | (nc[3] << 1) | (nc[2] >> 31) | (nc[3] >> 1));
            const uint64_t e0 = run_tokens ? runs4((uint64_t)r0 | ((uint64_t)r1 << 32)) : 0ull;
            uint64_t m1 = (uint64_t)r2 | ((uint64_t)r3 << 32);
            if (e0 >> 63) m1 &= ~1ull;  // the halves are planned apart: no run of the second one right behind a run of the first
            const uint64_t e1 = run_tokens ? runs4(m1) : 0ull;
#pragma unroll 1
            for (uint32_t h = 0; h < 2 && (e0 | e1); ++h) {
                const uint64_t e = h ? e1 : e0;
                uint64_t st = e & ~(e << 1);  // run starts
                while (st) {
                    const uint32_t x = (uint32_t)__ffsll((long long)st) - 1u;
                    st &= st - 1ull;
                    const uint64_t nz = ~(e >> x);
                    const uint32_t r = nz ? (uint32_t)__ffsll((long long)nz) - 1u : 64u;  // (e >> x has zeros above bit 63 - x)
                    bits -= r * s.len8[src[lo + 64u * h + x - 1u]];
                    bits += match_bits(s, r, 1u);
                }
            }
            s.cov[tid] = make_uint4(nc[0] | (uint32_t)e0, nc[1] | (uint32_t)(e0 >> 32), nc[2] | (uint32_t)e1, nc[3] | (uint32_t)(e1 >> 32));
        }
        // block-wide exclusive scan of `bits`
        uint32_t incl = bits;
        for (uint32_t d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) s.misc[8 + warp] = incl;
        if (tid == 0) s.misc[4] = x2nmodp(n % kSlice, 3);
        __syncthreads();
        uint32_t warp_off = 0, total = 0;
        for (uint32_t w = 0; w < kThreads / 32u; ++w) {
            const uint32_t v = s.misc[8 + w];
            if (w < warp) warp_off += v;
            total += v;
        }
        uint32_t pos = kHeaderBits + warp_off + incl - bits;
        const uint32_t eob = s.cl[256];
        const uint32_t end_bits = kHeaderBits + total + (eob >> 16);
        const bool too_big = end_bits > kOutWords * 32u - 64u;
        const uint32_t n1 = n > 65535u ? 65535u : n, n2 = n - n1;  // stored blocks hold at most 65535 bytes each
        const uint32_t nbytes = too_big ? 5u + n1 + (n2 ? 5u + n2 : 0u) : (end_bits + 7u) >> 3;
        if (tid == 0) look_back_publish(status, c, 18u + nbytes);  // the size is known: the members behind this one can be placed
        // 4. pack my tokens
        if (!too_big) {
            uint64_t acc = 0;
            uint32_t accbits = pos & 31u, wi = pos >> 5;
            uint32_t t = t_begin;
            uint32_t next_tok = t < t_end ? s.tok_pos[t] : 0xFFFFFFFFu;
            uint4 cv = s.cov[tid];              // shift register: the low 16 bits are the group's cover bits
            uint32_t carry = cov_end > lo ? 1u : 0u;  // the byte in front of the group is covered
            for (uint32_t i = lo; i < hi; i += 16u) {
                const uint4 v = __ldg(src4 + ((i - lo) >> 4));
                const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
                const uint32_t c16 = cv.x & 0xFFFFu;
                const uint32_t st16 = c16 & ~((c16 << 1) | carry);  // a token starts where the cover starts
                carry = c16 >> 15;
                const uint32_t lit = ~c16 & (hi - i >= 16u ? 0xFFFFu : ((1u << (hi - i)) - 1u));
                // Four bytes at a time: the codes of two literals are joined in 32 bits (<= 15 bits each) before they enter
                // the 64-bit accumulator, and a byte that a match covers simply contributes a code of length 0.  A match
                // is at least 4 bytes long, so it covers the rest of the word it starts in: its token goes out behind the
                // word's literals -- the only divergent step.
#pragma unroll
                for (uint32_t kw = 0; kw < 4; ++kw) {
                    const uint32_t wd = wds[kw], l4 = lit >> (4u * kw);
                    uint32_t e0 = s.cl[wd & 0xFFu], e1 = s.cl[(wd >> 8) & 0xFFu], e2 = s.cl[(wd >> 16) & 0xFFu], e3 = s.cl[wd >> 24];
                    e0 = (l4 & 1u) ? e0 : 0u;
                    e1 = (l4 & 2u) ? e1 : 0u;
                    e2 = (l4 & 4u) ? e2 : 0u;
                    e3 = (l4 & 8u) ? e3 : 0u;
                    const uint32_t n0 = e0 >> 16, n2 = e2 >> 16;
                    const uint32_t cA = (e0 & 0xFFFFu) | ((e1 & 0xFFFFu) << n0), cB = (e2 & 0xFFFFu) | ((e3 & 0xFFFFu) << n2);
                    acc |= (uint64_t)cA << accbits;
                    accbits += n0 + (e1 >> 16);
                    flush_acc(s, acc, accbits, wi);
                    acc |= (uint64_t)cB << accbits;
                    accbits += n2 + (e3 >> 16);
                    flush_acc(s, acc, accbits, wi);
                    const uint32_t s4 = (st16 >> (4u * kw)) & 0xFu;
                    if (s4) {
                        const uint32_t x = 4u * kw + (uint32_t)__ffs((int)s4) - 1u;  // the token's first byte, inside the group
                        uint32_t tlen, tdist = 1u;
                        if (i + x == next_tok) {  // a name-line match
                            const uint32_t ld = s.tok_ld[t];
                            tlen = ld & 0xFFFFu;
                            tdist = (ld >> 16) ? (ld >> 16) : 32768u;
                            ++t;
                            next_tok = t < t_end ? s.tok_pos[t] : 0xFFFFFFFFu;
                        } else {  // a run: its length is the length of the cover (<= 64 by construction)
                            const uint32_t nz = ~__funnelshift_r(cv.x, cv.y, x);
                            if (nz) tlen = (uint32_t)__ffs((int)nz) - 1u;
                            else {
                                const uint32_t nz2 = ~__funnelshift_r(cv.y, cv.z, x);
                                tlen = nz2 ? 31u + (uint32_t)__ffs((int)nz2) : 64u;
                            }
                        }
                        const uint32_t lt = s.ltok[tlen - 3u];
                        acc |= (uint64_t)(lt & 0xFFFFFFu) << accbits;
                        accbits += lt >> 24;
                        flush_acc(s, acc, accbits, wi);
                        if (tdist == 1u) {  // distance code 0, no extra bits
                            const uint32_t de = s.dl[0];
                            acc |= (uint64_t)(de & 0xFFFFu) << accbits;
                            accbits += de >> 16;
                        } else {
                            const uint32_t dc = dist_code(tdist), de = s.dl[dc];
                            acc |= (uint64_t)((de & 0xFFFFu) | ((tdist - kDistBase[dc]) << (de >> 16))) << accbits;
                            accbits += (de >> 16) + kDistExtra[dc];
                        }
                        flush_acc(s, acc, accbits, wi);
                    }
                }
                cv.x = __funnelshift_r(cv.x, cv.y, 16);
                cv.y = __funnelshift_r(cv.y, cv.z, 16);
                cv.z = __funnelshift_r(cv.z, cv.w, 16);
                cv.w >>= 16;
            }
            if (accbits) atomicOr(&s.out[wi], (uint32_t)acc);
            if (tid == 0) put_bits(s.out, kHeaderBits + total, eob & 0xFFFFu, eob >> 16);
        }
        if (warp == 0) {  // my place in the output
            const uint64_t excl = look_back_resolve(status, c, 18u + nbytes, lane);
            if (lane == 0) { s.misc[5] = (uint32_t)excl; s.misc[6] = (uint32_t)(excl >> 32); }
        }
        // 5. CRC-32 of the chunk: raw slice CRCs advanced over the bytes that follow them, XOR-reduced
        uint32_t part = 0;
        if (hi > lo) {
            const uint32_t after = n - hi;  // = (after / kSlice) * kSlice + (n % kSlice) for every slice but the last
            uint32_t xp = g_gz.pow_slice[after / kSlice];
            if (after % kSlice) xp = multmodp(xp, s.misc[4]);
            part = multmodp(xp, crc);
        }
        for (uint32_t d = 16; d; d >>= 1) part ^= __shfl_xor_sync(0xFFFFFFFFu, part, d);
        if (lane == 0) atomicXor(&s.misc[3], part);
        __syncthreads();
        // 6. the member at its final place: header, deflate bytes, CRC32, ISIZE
        const uint64_t off = (uint64_t)s.misc[5] | ((uint64_t)s.misc[6] << 32);
        if (off >= kFail) {
            if (tid == 0 && c + 1u == n_chunks) status[n_chunks] = ~0ull;
            continue;
        }
        uint8_t *dst = out + off;
        if (!too_big) {
            // aligned 32-bit stores: output word k = bytes [head + 4k, +4) of the packed bits (funnel shift of two smem words)
            const uint8_t *ob = reinterpret_cast<const uint8_t *>(s.out);
            uint8_t *d10 = dst + 10;
            const uint32_t head0 = (4u - (uint32_t)(reinterpret_cast<uintptr_t>(d10) & 3u)) & 3u;
            const uint32_t head = head0 < nbytes ? head0 : nbytes;
            const uint32_t nwords = (nbytes - head) >> 2;
            if (tid < head) d10[tid] = ob[tid];
            uint32_t *dw = reinterpret_cast<uint32_t *>(d10 + head);
            const uint32_t sh = 8u * head;
            for (uint32_t k = tid; k < nwords; k += kThreads) dw[k] = __funnelshift_r(s.out[k], s.out[k + 1u], sh);
            const uint32_t done = head + 4u * nwords;
            if (tid < nbytes - done) d10[done + tid] = ob[done + tid];
        } else {
            if (tid == 0) {
                dst[10] = n2 ? 0 : 1;
                dst[11] = (uint8_t)n1; dst[12] = (uint8_t)(n1 >> 8); dst[13] = (uint8_t)~n1; dst[14] = (uint8_t)(~n1 >> 8);
                if (n2) {
                    uint8_t *q = dst + 15 + n1;
                    q[0] = 1; q[1] = (uint8_t)n2; q[2] = (uint8_t)(n2 >> 8); q[3] = (uint8_t)~n2; q[4] = (uint8_t)(~n2 >> 8);
                }
            }
            for (uint32_t i = tid; i < n1; i += kThreads) dst[15 + i] = src[i];
            for (uint32_t i = tid; i < n2; i += kThreads) dst[20 + n1 + i] = src[n1 + i];
        }
        if (tid == 0) {
            // CRC register start value 0xFFFFFFFF carried over the n bytes: a constant for every full chunk
            const uint32_t crc32 = s.misc[3] ^ (n == kChunk ? g_gz.init_full : multmodp(x2nmodp(n, 3), 0xFFFFFFFFu)) ^ 0xFFFFFFFFu;
            const uint8_t hdr[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 0xff};
            for (int k = 0; k < 10; ++k) dst[k] = hdr[k];
            uint8_t *t = dst + 10 + nbytes;
            for (int k = 0; k < 4; ++k) { t[k] = (uint8_t)(crc32 >> (8 * k)); t[4 + k] = (uint8_t)(n >> (8 * k)); }
            if (c + 1u == n_chunks) status[n_chunks] = off + 18u + nbytes;  // total bytes of the stream
        }
    }
}

// ---- host side -----------------------------------------------------------------------------------------
static uint32_t h_multmodp(uint32_t a, uint32_t b) {
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ kPoly : b >> 1;
    }
    return p;
}

cudaError_t iss_gz_init_device(void) {  // once per device
    static GzTables t;
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ kPoly : c >> 1;
        t.crc[0][i] = c;
    }
    for (int k = 1; k < 4; ++k)
        for (uint32_t i = 0; i < 256; ++i) t.crc[k][i] = (t.crc[k - 1][i] >> 8) ^ t.crc[0][t.crc[k - 1][i] & 0xFFu];
    uint32_t p = 1u << 30;  // x^1
    t.x2n[0] = p;
    for (int n = 1; n < 32; ++n) t.x2n[n] = p = h_multmodp(p, p);
    uint32_t xs = 1u << 31;  // x^(8 * kSlice) by repeated multiplication with x^8
    for (uint32_t j = 0; j < kSlice; ++j) xs = h_multmodp(xs, t.x2n[3]);
    t.pow_slice[0] = 1u << 31;  // x^0
    for (uint32_t j = 1; j < kChunk / kSlice; ++j) t.pow_slice[j] = h_multmodp(t.pow_slice[j - 1], xs);
    {
        uint32_t xc = 1u << 31;  // x^(8 * kChunk) = (x^(8 * kSlice))^(kChunk / kSlice)
        for (uint32_t j = 0; j < kChunk / kSlice; ++j) xc = h_multmodp(xc, xs);
        t.init_full = h_multmodp(xc, 0xFFFFFFFFu);
    }
    cudaError_t e = cudaMemcpyToSymbol(g_gz, &t, sizeof t);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(iss_deflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
}

uint32_t iss_gz_chunks(uint64_t n_in) { return (uint32_t)((n_in + kChunk - 1) / kChunk); }

cudaError_t iss_gz_compress(const uint8_t *d_in, uint64_t n_in, uint64_t *d_status, uint8_t *d_out, uint32_t *d_work, int sm_count,
                            int reuse_code, cudaStream_t stream) {
    const uint32_t nc = iss_gz_chunks(n_in);
    cudaError_t e = cudaMemsetAsync(d_status, 0, (nc + 1ull) * sizeof(uint64_t), stream);  // nc = 0: total = 0
    if (e != cudaSuccess || nc == 0) return e;
    e = cudaMemsetAsync(d_work, 0, kWorkZeroed * sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    const uint64_t n_pieces = (n_in + kPiece - 1) / kPiece;
    const uint32_t piece_stride = (uint32_t)((n_pieces + 255) / 256);
    const uint32_t n_sampled = (uint32_t)((n_pieces + piece_stride - 1) / piece_stride);
    if (!reuse_code) {  // else: the code a previous call left in d_work (the words behind kWorkZeroed survive the memset above)
        iss_gz_hist_kernel<<<n_sampled < (uint32_t)sm_count ? n_sampled : (uint32_t)sm_count, 256, 0, stream>>>(d_in, n_in, piece_stride, d_work);
        iss_gz_table_kernel<<<1, 256, 0, stream>>>(d_in, n_in, d_work);
    }
    const uint32_t grid = nc < (uint32_t)(sm_count * 3) ? nc : (uint32_t)(sm_count * 3);
    iss_deflate_kernel<<<grid, kThreads, sizeof(Smem), stream>>>(d_in, n_in, d_out, reinterpret_cast<unsigned long long *>(d_status), d_work, nc);
    return cudaGetLastError();
}
