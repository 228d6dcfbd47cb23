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
// One CTA per chunk: histogram -> Huffman code lengths limited to 15 bits (two-queue merge + zlib's overflow repair,
// always a complete code) -> canonical codes -> header -> bit packing into shared memory -> CRC-32 from per-slice raw
// CRCs advanced by x^(8 * bytes_after) (GF(2) multiply, zlib's crc32_combine arithmetic) -> member into a staging slot.
// A scan + pack pass makes the members contiguous.  Incompressible chunks fall back to stored blocks.
// Validated by decompression (tests): gunzip(output) == input, for FASTQ and for arbitrary bytes.
#include <cuda_runtime.h>
#include <stdint.h>

#include "iss_deflate.h"

namespace {

constexpr uint32_t kChunk = ISS_GZ_CHUNK;
constexpr uint32_t kOutWords = 12288;  // 48 KB of packed bits per chunk in shared memory; beyond that: stored block
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
    uint32_t crc[256];     // byte-wise CRC-32 table
    uint32_t x2n[32];      // x^(2^n) mod P
    uint32_t pow256[256];  // x^(8 * 256 * j) mod P
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

struct Smem {
    uint32_t out[kOutWords];
    uint32_t hist[8 * 257];
    uint32_t freq[260];
    uint32_t keys[260];
    uint32_t w[516];
    uint32_t crcT[256];
    uint32_t misc[32];  // 0 n_sym, 1 total payload bits, 2 overflow flag, 3 crc accumulator, 4 uniform x^(8*(n%256)), 8.. warp sums
    uint16_t parent[516];
    uint16_t codes[260];  // bit-reversed canonical codes
    uint32_t cl[260];     // code | length << 16: one lookup per literal in the packing loop
    uint8_t lens[260];
};

__device__ __forceinline__ void put_bits(uint32_t *out, uint32_t pos, uint32_t v, uint32_t nbits) {
    const uint32_t wi = pos >> 5, sh = pos & 31u;
    atomicOr(&out[wi], v << sh);
    if (sh + nbits > 32u) atomicOr(&out[wi + 1], v >> (32u - sh));
}

// thread 0: code lengths for the sorted leaves keys[0..n) (freq << 9 | symbol), at most 15 bits, complete code
__device__ void build_lengths(Smem &s, uint32_t n) {
    for (uint32_t i = 0; i < 257; ++i) s.lens[i] = 0;
    if (n == 1) {  // only EOB (empty chunk): a complete code needs two codes of length 1
        s.lens[s.keys[0] & 511u] = 1;
        s.lens[(s.keys[0] & 511u) == 0 ? 1 : 0] = 1;
        return;
    }
    for (uint32_t i = 0; i < n; ++i) s.w[i] = s.keys[i] >> 9;
    uint32_t a = 0, b = n, nxt = n;
    for (uint32_t it = 0; it + 1 < n; ++it) {
        uint32_t x, y;
        if (a < n && (b >= nxt || s.w[a] <= s.w[b])) x = a++; else x = b++;
        if (a < n && (b >= nxt || s.w[a] <= s.w[b])) y = a++; else y = b++;
        s.w[nxt] = s.w[x] + s.w[y];
        s.parent[x] = (uint16_t)nxt;
        s.parent[y] = (uint16_t)nxt;
        nxt++;
    }
    // depths root-first, capped at 15 while descending (zlib gen_bitlen); s.w is reused for the depths
    uint32_t bl[17];
    for (int i = 0; i < 17; ++i) bl[i] = 0;
    uint32_t overflow = 0;
    const uint32_t root = 2 * n - 2;
    s.w[root] = 0;
    for (int i = (int)root - 1; i >= 0; --i) {
        uint32_t d = s.w[s.parent[i]] + 1;
        if (d > 15) { d = 15; overflow++; }
        s.w[i] = d;
        if ((uint32_t)i < n) bl[d]++;
    }
    if (overflow) {
        do {
            uint32_t bits = 14;
            while (bl[bits] == 0) bits--;
            bl[bits]--;
            bl[bits + 1] += 2;
            bl[15]--;
            overflow -= 2;  // may wrap below zero by one: treated as signed like zlib
        } while ((int32_t)overflow > 0);
        uint32_t i = 0;  // leaves are sorted by increasing frequency: the rarest get the longest codes
        for (uint32_t bits = 15; bits >= 1; --bits)
            for (uint32_t k = 0; k < bl[bits]; ++k) s.w[i++] = bits;
    }
    for (uint32_t i = 0; i < n; ++i) s.lens[s.keys[i] & 511u] = (uint8_t)s.w[i];
}

__device__ void assign_codes(Smem &s) {  // thread 0: canonical codes, stored bit-reversed (DEFLATE packs codes MSB first)
    uint32_t bl[17], nxt[17];
    for (int i = 0; i < 17; ++i) bl[i] = 0;
    for (uint32_t sym = 0; sym < 257; ++sym) bl[s.lens[sym]]++;
    bl[0] = 0;
    uint32_t code = 0;
    for (uint32_t b = 1; b <= 15; ++b) {
        code = (code + bl[b - 1]) << 1;
        nxt[b] = code;
    }
    for (uint32_t sym = 0; sym < 257; ++sym) {
        const uint32_t l = s.lens[sym];
        s.codes[sym] = l ? (uint16_t)(__brev(nxt[l]++) >> (32u - l)) : 0;
        s.cl[sym] = (uint32_t)s.codes[sym] | (l << 16);
    }
}

}  // namespace

extern __shared__ __align__(16) uint8_t iss_gz_smem[];

__global__ void __launch_bounds__(256) iss_deflate_kernel(const uint8_t *__restrict__ in, uint64_t n_in, uint8_t *__restrict__ stage,
                                                          uint32_t *__restrict__ sizes, uint32_t n_chunks) {
    Smem &s = *reinterpret_cast<Smem *>(iss_gz_smem);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    s.crcT[tid] = g_gz.crc[tid];
    for (uint32_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const uint64_t base = (uint64_t)c * kChunk;
        const uint32_t n = (uint32_t)((n_in - base < kChunk) ? (n_in - base) : kChunk);
        const uint8_t *src = in + base;
        uint8_t *dst = stage + (uint64_t)c * ISS_GZ_STRIDE;
        __syncthreads();  // previous chunk fully written out
        for (uint32_t i = tid; i < kOutWords; i += 256) s.out[i] = 0;
        for (uint32_t i = tid; i < 8 * 257; i += 256) s.hist[i] = 0;
        if (tid < 32) s.misc[tid] = 0;
        __syncthreads();
        // 1. histogram (per-warp copies)
        for (uint32_t i = tid; i < n; i += 256) atomicAdd(&s.hist[warp * 257 + src[i]], 1u);
        __syncthreads();
        {
            uint32_t f = 0;
            for (uint32_t w = 0; w < 8; ++w) f += s.hist[w * 257 + tid];
            s.freq[tid] = f;
            if (tid == 0) s.freq[256] = 1;  // end of block
        }
        __syncthreads();
        // 2. rank sort of the present symbols by (frequency, symbol)
        for (uint32_t sym = tid; sym < 257; sym += 256) {
            const uint32_t f = s.freq[sym];
            if (f) {
                const uint32_t key = (f << 9) | sym;
                uint32_t rank = 0;
                for (uint32_t o = 0; o < 257; ++o) {
                    const uint32_t fo = s.freq[o];
                    rank += (fo && ((fo << 9) | o) < key) ? 1u : 0u;
                }
                s.keys[rank] = key;
                atomicAdd(&s.misc[0], 1u);
            }
        }
        __syncthreads();
        // 3. code lengths + canonical codes (serial; hidden by the other CTAs of the SM)
        if (tid == 0) {
            build_lengths(s, s.misc[0]);
            assign_codes(s);
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
        // 4. the 258 code lengths, each as the 4-bit code of its value (a flat 4-bit code: code == value, MSB first)
        for (uint32_t sym = tid; sym < 257; sym += 256) put_bits(s.out, 74u + 4u * sym, __brev((uint32_t)s.lens[sym]) >> 28, 4);
        // 5. my slice (256 contiguous bytes, read 16 at a time): bit count + raw CRC
        const uint32_t lo = tid * 256u, hi = (lo + 256u < n) ? lo + 256u : n;
        uint32_t bits = 0, crc = 0;
        {
            const uint4 *src4 = reinterpret_cast<const uint4 *>(src + lo);  // chunk base and slices are 16-byte aligned
            for (uint32_t i = lo; i < hi; i += 16u) {
                const uint4 v = __ldg(src4 + ((i - lo) >> 4));
                const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (uint32_t k = 0; k < 16; ++k) {
                    if (i + k < hi) {
                        const uint32_t b = (wds[k >> 2] >> (8u * (k & 3u))) & 0xFFu;
                        bits += s.cl[b] >> 16;
                        crc = s.crcT[(crc ^ b) & 0xFFu] ^ (crc >> 8);
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
        if (lane == 31) s.misc[8 + warp] = incl;
        if (tid == 0) s.misc[4] = x2nmodp(n & 255u, 3);
        __syncthreads();
        uint32_t warp_off = 0, total = 0;
        for (uint32_t w = 0; w < 8; ++w) {
            const uint32_t v = s.misc[8 + w];
            if (w < warp) warp_off += v;
            total += v;
        }
        uint32_t pos = kHeaderBits + warp_off + incl - bits;
        const uint32_t end_bits = kHeaderBits + total + s.lens[256];
        const bool too_big = end_bits > kOutWords * 32u - 64u;
        // 6. pack my literals
        if (!too_big) {
            uint64_t acc = 0;
            uint32_t accbits = pos & 31u, wi = pos >> 5;
            const uint4 *src4 = reinterpret_cast<const uint4 *>(src + lo);
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
            if (tid == 0) put_bits(s.out, kHeaderBits + total, s.codes[256], s.lens[256]);
        }
        // 7. CRC-32 of the chunk: raw slice CRCs advanced over the bytes that follow them, XOR-reduced
        uint32_t part = 0;
        if (hi > lo) {
            const uint32_t after = n - hi;  // = (after >> 8) * 256 + (n & 255) for every slice but the last
            uint32_t xp = g_gz.pow256[after >> 8];
            if (after & 255u) xp = multmodp(xp, s.misc[4]);
            part = multmodp(xp, crc);
        }
        for (uint32_t d = 16; d; d >>= 1) part ^= __shfl_xor_sync(0xFFFFFFFFu, part, d);
        if (lane == 0) atomicXor(&s.misc[3], part);
        __syncthreads();
        // 8. the member: header, deflate bytes, CRC32, ISIZE
        uint32_t nbytes;
        if (!too_big) {
            nbytes = (end_bits + 7u) >> 3;
            const uint8_t *ob = reinterpret_cast<const uint8_t *>(s.out);
            for (uint32_t i = tid; i < nbytes; i += 256) dst[10 + i] = ob[i];
        } else {  // stored blocks: LEN <= 65535 each
            const uint32_t n1 = n > 65535u ? 65535u : n, n2 = n - n1;
            if (tid == 0) {
                dst[10] = n2 ? 0 : 1;
                dst[11] = (uint8_t)n1; dst[12] = (uint8_t)(n1 >> 8); dst[13] = (uint8_t)~n1; dst[14] = (uint8_t)(~n1 >> 8);
                if (n2) {
                    uint8_t *q = dst + 15 + n1;
                    q[0] = 1; q[1] = (uint8_t)n2; q[2] = (uint8_t)(n2 >> 8); q[3] = (uint8_t)~n2; q[4] = (uint8_t)(~n2 >> 8);
                }
            }
            for (uint32_t i = tid; i < n1; i += 256) dst[15 + i] = src[i];
            for (uint32_t i = tid; i < n2; i += 256) dst[20 + n1 + i] = src[n1 + i];
            nbytes = 5 + n1 + (n2 ? 5 + n2 : 0);
        }
        if (tid == 0) {
            const uint32_t crc32 = s.misc[3] ^ multmodp(x2nmodp(n, 3), 0xFFFFFFFFu) ^ 0xFFFFFFFFu;
            const uint8_t hdr[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 0xff};
            for (int k = 0; k < 10; ++k) dst[k] = hdr[k];
            uint8_t *t = dst + 10 + nbytes;
            for (int k = 0; k < 4; ++k) { t[k] = (uint8_t)(crc32 >> (8 * k)); t[4 + k] = (uint8_t)(n >> (8 * k)); }
            sizes[c] = 18 + nbytes;
        }
    }
}

// exclusive scan of the member sizes (one CTA) -> offsets[n_chunks + 1]
__global__ void __launch_bounds__(1024) iss_gz_scan_kernel(const uint32_t *__restrict__ sizes, uint64_t *__restrict__ offsets, uint32_t n) {
    __shared__ uint64_t warp_sum[32];
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint64_t v = i < n ? sizes[i] : 0;
        uint64_t incl = v;
        for (uint32_t d = 1; d < 32; d <<= 1) {
            const uint64_t u = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if ((threadIdx.x & 31u) >= d) incl += u;
        }
        if ((threadIdx.x & 31u) == 31u) warp_sum[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint64_t off = carry;
        for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) off += warp_sum[w];
        if (i < n) offsets[i] = off + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = off + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[n] = carry;
}

__global__ void __launch_bounds__(256) iss_gz_pack_kernel(const uint8_t *__restrict__ stage, const uint32_t *__restrict__ sizes,
                                                          const uint64_t *__restrict__ offsets, uint8_t *__restrict__ out, uint32_t n_chunks) {
    for (uint32_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const uint8_t *src = stage + (uint64_t)c * ISS_GZ_STRIDE;
        uint8_t *dst = out + offsets[c];
        const uint32_t n = sizes[c];
        for (uint32_t i = threadIdx.x; i < n; i += 256) dst[i] = src[i];
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
        t.crc[i] = c;
    }
    uint32_t p = 1u << 30;  // x^1
    t.x2n[0] = p;
    for (int n = 1; n < 32; ++n) t.x2n[n] = p = h_multmodp(p, p);
    const uint32_t x2048 = t.x2n[11];  // x^(2^11) = x^(8 * 256)
    t.pow256[0] = 1u << 31;            // x^0
    for (int j = 1; j < 256; ++j) t.pow256[j] = h_multmodp(t.pow256[j - 1], x2048);
    cudaError_t e = cudaMemcpyToSymbol(g_gz, &t, sizeof t);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(iss_deflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
}

uint32_t iss_gz_chunks(uint64_t n_in) { return (uint32_t)((n_in + kChunk - 1) / kChunk); }

cudaError_t iss_gz_compress(const uint8_t *d_in, uint64_t n_in, uint8_t *d_stage, uint32_t *d_sizes, uint64_t *d_offsets, uint8_t *d_out,
                            int sm_count, cudaStream_t stream) {
    const uint32_t nc = iss_gz_chunks(n_in);
    if (nc == 0) return cudaMemsetAsync(d_offsets, 0, sizeof(uint64_t), stream);
    const uint32_t grid = nc < (uint32_t)(sm_count * 3) ? nc : (uint32_t)(sm_count * 3);
    iss_deflate_kernel<<<grid, 256, sizeof(Smem), stream>>>(d_in, n_in, d_stage, d_sizes, nc);
    iss_gz_scan_kernel<<<1, 1024, 0, stream>>>(d_sizes, d_offsets, nc);
    iss_gz_pack_kernel<<<grid, 256, 0, stream>>>(d_stage, d_sizes, d_offsets, d_out, nc);
    return cudaGetLastError();
}
