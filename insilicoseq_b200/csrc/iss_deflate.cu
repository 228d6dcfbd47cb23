// iss_deflate.cu -- gzip on the GPU for `--compress` (iss/util.py:255-268 gzips the finished files on one CPU core).
//
// The FASTQ text of a batch already sits in HBM; compressing it there is the only way the end-to-end rate can exceed
// what PCIe can carry as plain text.  Format: a gzip file is a sequence of members (RFC 1952); every 64 KB chunk of
// a stream becomes one member holding ONE dynamic-Huffman DEFLATE block (RFC 1951) of literals only (no LZ77):
//
//   10 B gzip header | BFINAL=1 BTYPE=10 HLIT=0 HDIST=0 HCLEN=15 | 19 x 3 bit code-length-code lengths (0..15 -> 4 bits,
//   16/17/18 unused) | 257 literal/EOB lengths + 1 distance length, 4 bits each | the literals | EOB | pad | CRC32 | ISIZE
//
// so the block header has a fixed size (1106 bits) and every symbol's bit position follows from a prefix sum.
// Per call (= one FASTQ stream of one batch): byte histogram of a <= 4 MB sample -> ONE Huffman code, lengths limited to
// 15 bits (two-queue merge + zlib's overflow repair, always a complete code; every byte value keeps a code) ->
// canonical codes + block header.  Then one CTA per chunk: bit count per 128-byte slice -> prefix sum -> bit packing
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
constexpr uint32_t kHeaderBits = 17 + 19 * 3 + 258 * 4;  // 1106
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
    uint32_t cl[260];   // bit-reversed canonical code | length << 16, symbols 0..255 and 256 = end of block
    uint32_t hdr[36];   // the kHeaderBits bits of the block header (lengths included)
};
constexpr uint32_t kWorkHist = 0;     // word offsets inside the caller's work area (ISS_GZ_WORK_WORDS u32)
constexpr uint32_t kWorkTicket = 260;  // next chunk to hand out
constexpr uint32_t kWorkZeroed = 264;  // words cleared at the start of a call
constexpr uint32_t kWorkCode = 264;
static_assert(kWorkCode + sizeof(GzCode) / 4 <= ISS_GZ_WORK_WORDS, "work area too small");

struct TableSmem {
    uint32_t out[36];
    uint32_t freq[260];
    uint32_t keys[260];
    uint32_t w[516];      // node weights during the merge, then node depths
    uint32_t bl[17];      // leaves per code length
    uint32_t first[17];   // first canonical code per length
    uint32_t total, overflow;
    uint16_t parent[516];
    uint32_t cl[260];
    uint8_t lens[260];
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

__device__ uint64_t look_back(volatile unsigned long long *status, uint32_t c, uint32_t msize, uint32_t lane) {
    if (c == 0) {
        if (lane == 0) status[0] = kInc | msize;
        return 0;
    }
    if (lane == 0) status[c] = kAgg | msize;
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

// ---- 2. one CTA: the Huffman code of the call (every byte value keeps a code: frequency = sample count + 1) ------------
// Lengths are limited to 15 bits the way zlib's gen_bitlen does it: depth capped at 15 while descending, `overflow` =
// nodes deeper than that, then the repair loop on the per-length counts and lengths re-dealt to the leaves in frequency
// order (the rarest get the longest codes); always a complete code.
__global__ void __launch_bounds__(256) iss_gz_table_kernel(uint32_t *__restrict__ work) {
    __shared__ TableSmem s;
    const uint32_t tid = threadIdx.x;
    constexpr uint32_t n = 257, root = 2 * n - 2;
    if (tid == 0) s.total = s.overflow = 0;
    if (tid < 36) s.out[tid] = 0;
    if (tid < 17) s.bl[tid] = 0;
    __syncthreads();
    const uint32_t f0 = work[kWorkHist + tid];
    atomicAdd(&s.total, f0);
    s.freq[tid] = f0 + 1u;
    __syncthreads();
    if (tid == 0) s.freq[256] = (s.total >> 16) ? (s.total >> 16) : 1u;  // end of block: one per 64 KB member
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
    if (tid == 0) {
        merge_tree(s, n);
        // constant part of the block header: BFINAL=1, BTYPE=2, HLIT=0, HDIST=0, HCLEN=15, then the 19 code-length-code
        // lengths in the order 16,17,18,0,8,7,...: three zeros and sixteen fours
        uint32_t pos = 0;
        put_bits(s.out, pos, 1u, 1); pos += 1;
        put_bits(s.out, pos, 2u, 2); pos += 2;
        pos += 10;  // HLIT, HDIST = 0
        put_bits(s.out, pos, 15u, 4); pos += 4;
        pos += 9;   // lengths of 16, 17, 18 = 0
        for (int k = 0; k < 16; ++k) { put_bits(s.out, pos, 4u, 3); pos += 3; }
    }
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
        // the 257 literal/EOB lengths (+ a zero distance length), each as the 4-bit code of its value (flat code, MSB first)
        put_bits(s.out, 74u + 4u * sym, __brev(l) >> 28, 4);
    }
    __syncthreads();
    GzCode *code = reinterpret_cast<GzCode *>(work + kWorkCode);
    for (uint32_t sym = tid; sym < 260; sym += 256) code->cl[sym] = sym < n ? s.cl[sym] : 0u;
    if (tid < 36) code->hdr[tid] = s.out[tid];
}

// ---- 3. the members: one CTA per 64 KB chunk, 512 threads x 128-byte slices ----------------------------------------------
struct Smem {
    uint32_t out[kOutWords];
    uint32_t crcT[4][256];
    uint32_t cl[260];
    uint32_t misc[64];  // 3 crc accumulator, 4 x^(8*(n % kSlice)), 8.. warp sums
    uint8_t len8[260];
};

extern __shared__ __align__(16) uint8_t iss_gz_smem[];

__global__ void __launch_bounds__(kThreads, 3) iss_deflate_kernel(const uint8_t *__restrict__ in, uint64_t n_in, uint8_t *__restrict__ out,
                                                                  unsigned long long *status, uint32_t *__restrict__ work,
                                                                  uint32_t n_chunks) {
    Smem &s = *reinterpret_cast<Smem *>(iss_gz_smem);
    const GzCode *code = reinterpret_cast<const GzCode *>(work + kWorkCode);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    for (uint32_t i = tid; i < 1024; i += kThreads) s.crcT[i >> 8][i & 255u] = g_gz.crc[i >> 8][i & 255u];
    for (uint32_t i = tid; i < 260; i += kThreads) {
        const uint32_t e = code->cl[i];
        s.cl[i] = e;
        s.len8[i] = (uint8_t)(e >> 16);
    }
    for (;;) {
        __syncthreads();  // previous chunk fully written out; tables loaded
        if (tid == 0) s.misc[0] = atomicAdd(&work[kWorkTicket], 1u);
        __syncthreads();
        const uint32_t c = s.misc[0];
        if (c >= n_chunks) break;
        const uint64_t base = (uint64_t)c * kChunk;
        const uint32_t n = (uint32_t)((n_in - base < kChunk) ? (n_in - base) : kChunk);
        const uint8_t *src = in + base;
        for (uint32_t i = tid; i < kOutWords; i += kThreads) s.out[i] = i < 36u ? code->hdr[i] : 0u;
        if (tid >= 1 && tid < 64) s.misc[tid] = 0;
        // 1. my slice (kSlice contiguous bytes, read 16 at a time): bit count + raw CRC (slicing-by-4)
        const uint32_t lo = tid * kSlice, hi = (lo + kSlice < n) ? lo + kSlice : n;
        const uint4 *src4 = reinterpret_cast<const uint4 *>(src + lo);  // chunk base and slices are 16-byte aligned
        uint32_t bits = 0, crc = 0;
        for (uint32_t i = lo; i < hi; i += 16u) {
            const uint4 v = __ldg(src4 + ((i - lo) >> 4));
            const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) {
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
            }
        }
        // block-wide exclusive scan of `bits`
        uint32_t incl = bits;
        for (uint32_t d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += v;
        }
        __syncthreads();  // out[] initialised, misc zeroed
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
        if (warp == 0) {  // the size is known: publish it and find my place while the other warps pack
            const uint64_t excl = look_back(status, c, 18u + nbytes, lane);
            if (lane == 0) { s.misc[5] = (uint32_t)excl; s.misc[6] = (uint32_t)(excl >> 32); }
        }
        // 2. pack my literals
        if (!too_big) {
            uint64_t acc = 0;
            uint32_t accbits = pos & 31u, wi = pos >> 5;
            for (uint32_t i = lo; i < hi; i += 16u) {
                const uint4 v = __ldg(src4 + ((i - lo) >> 4));
                const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (uint32_t k = 0; k < 16; ++k) {
                    if (i + k < hi) {
                        const uint32_t e = s.cl[(wds[k >> 2] >> (8u * (k & 3u))) & 0xFFu];
                        acc |= (uint64_t)(e & 0xFFFFu) << accbits;
                        accbits += e >> 16;
                        if (accbits >= 32u) {
                            atomicOr(&s.out[wi++], (uint32_t)acc);
                            acc >>= 32;
                            accbits -= 32u;
                        }
                    }
                }
            }
            if (accbits) atomicOr(&s.out[wi], (uint32_t)acc);
            if (tid == 0) put_bits(s.out, kHeaderBits + total, eob & 0xFFFFu, eob >> 16);
        }
        // 3. CRC-32 of the chunk: raw slice CRCs advanced over the bytes that follow them, XOR-reduced
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
        // 4. the member at its final place: header, deflate bytes, CRC32, ISIZE
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
            const uint32_t crc32 = s.misc[3] ^ multmodp(x2nmodp(n, 3), 0xFFFFFFFFu) ^ 0xFFFFFFFFu;
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
    cudaError_t e = cudaMemcpyToSymbol(g_gz, &t, sizeof t);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(iss_deflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
}

uint32_t iss_gz_chunks(uint64_t n_in) { return (uint32_t)((n_in + kChunk - 1) / kChunk); }

cudaError_t iss_gz_compress(const uint8_t *d_in, uint64_t n_in, uint64_t *d_status, uint8_t *d_out, uint32_t *d_work, int sm_count,
                            cudaStream_t stream) {
    const uint32_t nc = iss_gz_chunks(n_in);
    cudaError_t e = cudaMemsetAsync(d_status, 0, (nc + 1ull) * sizeof(uint64_t), stream);  // nc = 0: total = 0
    if (e != cudaSuccess || nc == 0) return e;
    e = cudaMemsetAsync(d_work, 0, kWorkZeroed * sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    const uint64_t n_pieces = (n_in + kPiece - 1) / kPiece;
    const uint32_t piece_stride = (uint32_t)((n_pieces + 255) / 256);
    const uint32_t n_sampled = (uint32_t)((n_pieces + piece_stride - 1) / piece_stride);
    iss_gz_hist_kernel<<<n_sampled < (uint32_t)sm_count ? n_sampled : (uint32_t)sm_count, 256, 0, stream>>>(d_in, n_in, piece_stride, d_work);
    iss_gz_table_kernel<<<1, 256, 0, stream>>>(d_work);
    const uint32_t grid = nc < (uint32_t)(sm_count * 3) ? nc : (uint32_t)(sm_count * 3);
    iss_deflate_kernel<<<grid, kThreads, sizeof(Smem), stream>>>(d_in, n_in, d_out, reinterpret_cast<unsigned long long *>(d_status), d_work, nc);
    return cudaGetLastError();
}
