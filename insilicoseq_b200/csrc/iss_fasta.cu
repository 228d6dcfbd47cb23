// iss_fasta.cu -- FASTA text -> the records' bases, on the device.
//
// The reference reads a genome file with Bio.SeqIO.parse and takes str(record.seq) per record (iss/generator.py:270,
// 468-490): every byte goes through Python objects.  Here the host only finds the title lines (iss_fasta_index: a memchr
// scan for '>' at line starts); the text itself is copied to HBM as it is, and five small kernels turn it into the staging
// layout that iss_genome.cu classifies and packs: title lines are overwritten with '\n', the bytes that are not white
// space are counted per 4 KB block, a prefix sum gives every block's first rank, the ranks of the record boundaries give
// the record lengths, and a scatter writes byte x of record r to stage_off[r] + (rank(x) - rank(body start of r)).
// All of it is byte work at HBM speed: 3 Gbp of text are read three times and written once.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/iss_b200.h"
#include "iss_fasta.h"

namespace {

constexpr uint32_t kBlock = 4096;  // bytes of text per CTA step: 256 threads x 16 bytes

__device__ __forceinline__ uint32_t is_white(uint32_t b) { return (b == 0x20u || (b - 9u) < 5u) ? 1u : 0u; }

// white-space flags of 16 bytes: bit j = byte j is white space
__device__ __forceinline__ uint32_t white16(const uint4 v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t m = 0;
#pragma unroll
    for (uint32_t k = 0; k < 4; ++k)
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) m |= is_white((w[k] >> (8u * j)) & 0xFFu) << (4u * k + j);
    return m;
}

__global__ void __launch_bounds__(256) fasta_blank_kernel(uint8_t *__restrict__ text, const uint64_t *__restrict__ gt, const uint64_t *__restrict__ body,
                                                          uint32_t n_recs) {
    for (uint32_t r = blockIdx.x; r < n_recs; r += gridDim.x) {
        const uint64_t a = r ? gt[r] : 0ull, b = body[r];  // record 0 also clears whatever stands in front of its title
        for (uint64_t x = a + threadIdx.x; x < b; x += blockDim.x) text[x] = '\n';
    }
}

__global__ void __launch_bounds__(256) fasta_count_kernel(const uint8_t *__restrict__ text, uint64_t n_blocks, uint32_t *__restrict__ cnt) {
    __shared__ uint32_t wsum[8];
    for (uint64_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(text + b * kBlock) + threadIdx.x);
        uint32_t c = 16u - (uint32_t)__popc(white16(v));
        for (uint32_t d = 16; d; d >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
        if ((threadIdx.x & 31u) == 0) wsum[threadIdx.x >> 5] = c;
        __syncthreads();
        if (threadIdx.x == 0) cnt[b] = wsum[0] + wsum[1] + wsum[2] + wsum[3] + wsum[4] + wsum[5] + wsum[6] + wsum[7];
        __syncthreads();
    }
}

// one CTA: thread t sums its contiguous share of the counts, the shares are scanned, the prefix is written back
__global__ void __launch_bounds__(1024) fasta_scan_kernel(const uint32_t *__restrict__ cnt, uint64_t n_blocks, uint64_t *__restrict__ pre) {
    __shared__ uint64_t part[1024];
    const uint64_t per = (n_blocks + 1023) / 1024, a = (uint64_t)threadIdx.x * per, b = a + per < n_blocks ? a + per : n_blocks;
    uint64_t s = 0;
    for (uint64_t k = a; k < b; ++k) s += cnt[k];
    part[threadIdx.x] = s;
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {  // Hillis-Steele, inclusive
        const uint64_t v = threadIdx.x >= d ? part[threadIdx.x - d] : 0ull;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    uint64_t run = part[threadIdx.x] - s;  // exclusive
    for (uint64_t k = a; k < b; ++k) { pre[k] = run; run += cnt[k]; }
    if (threadIdx.x == 1023) pre[n_blocks] = part[1023];
}

// one warp per position: the block's prefix + the non-white bytes between the block's start and the position
__global__ void __launch_bounds__(256) fasta_rank_kernel(const uint8_t *__restrict__ text, const uint64_t *__restrict__ pre, const uint64_t *__restrict__ pos,
                                                         uint32_t n_pos, uint64_t *__restrict__ rank) {
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t k = blockIdx.x * 8u + (threadIdx.x >> 5); k < n_pos; k += gridDim.x * 8u) {
        const uint64_t p = pos[k], b = p / kBlock, x0 = b * kBlock;
        uint32_t c = 0;
        for (uint64_t x = x0 + lane; x < p; x += 32u) c += 1u - is_white(text[x]);
        for (uint32_t d = 16; d; d >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
        if (lane == 0) rank[k] = pre[b] + c;
    }
}

__global__ void __launch_bounds__(256) fasta_scatter_kernel(const uint8_t *__restrict__ text, uint64_t n_blocks, const uint64_t *__restrict__ pre,
                                                            const uint64_t *__restrict__ body, const uint64_t *__restrict__ end,
                                                            const uint64_t *__restrict__ rank0, const uint64_t *__restrict__ stage_off, uint32_t n_recs,
                                                            uint8_t *__restrict__ stage) {
    __shared__ uint32_t wsum[8];
    __shared__ uint32_t first_rec;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint64_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        const uint64_t x0 = b * kBlock + 16ull * threadIdx.x;
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(text + b * kBlock) + threadIdx.x);
        const uint32_t white = white16(v), mine = 16u - (uint32_t)__popc(white);
        uint32_t incl = mine;
        for (uint32_t d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        if (threadIdx.x == 0) {  // the last record whose body starts at or before this block (bodies are in file order)
            uint32_t lo = 0, hi = n_recs;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (body[mid] <= b * kBlock) lo = mid + 1; else hi = mid; }
            first_rec = lo ? lo - 1 : 0;
        }
        __syncthreads();
        uint32_t before = 0;
        for (uint32_t w = 0; w < warp; ++w) before += wsum[w];
        uint64_t rank = pre[b] + before + incl - mine;
        if (mine) {
            uint32_t r = first_rec;
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (uint32_t j = 0; j < 16; ++j) {
                if ((white >> j) & 1u) continue;
                const uint64_t x = x0 + j;
                while (r + 1 < n_recs && x >= body[r + 1]) ++r;  // bytes between a body's end and the next body are white (blanked titles)
                if (x >= body[r] && x < end[r]) stage[stage_off[r] + (rank - rank0[r])] = (uint8_t)(w[j >> 2] >> (8u * (j & 3u)));
                ++rank;
            }
        }
        __syncthreads();
    }
}

uint32_t grid_for(uint64_t n, uint32_t cap) { return (uint32_t)(n < cap ? (n ? n : 1) : cap); }

}  // namespace

cudaError_t iss_fasta_blank(uint8_t *d_text, const uint64_t *d_gt, const uint64_t *d_body, uint32_t n_recs, cudaStream_t stream) {
    if (!n_recs) return cudaSuccess;
    fasta_blank_kernel<<<grid_for(n_recs, 65535), 256, 0, stream>>>(d_text, d_gt, d_body, n_recs);
    return cudaGetLastError();
}
cudaError_t iss_fasta_count(const uint8_t *d_text, uint64_t n_blocks, uint32_t *d_cnt, cudaStream_t stream) {
    fasta_count_kernel<<<grid_for(n_blocks, 148 * 16), 256, 0, stream>>>(d_text, n_blocks, d_cnt);
    return cudaGetLastError();
}
cudaError_t iss_fasta_scan(const uint32_t *d_cnt, uint64_t n_blocks, uint64_t *d_pre, cudaStream_t stream) {
    fasta_scan_kernel<<<1, 1024, 0, stream>>>(d_cnt, n_blocks, d_pre);
    return cudaGetLastError();
}
cudaError_t iss_fasta_ranks(const uint8_t *d_text, const uint64_t *d_pre, const uint64_t *d_pos, uint32_t n_pos, uint64_t *d_rank, cudaStream_t stream) {
    if (!n_pos) return cudaSuccess;
    fasta_rank_kernel<<<grid_for((n_pos + 7) / 8, 65535), 256, 0, stream>>>(d_text, d_pre, d_pos, n_pos, d_rank);
    return cudaGetLastError();
}
cudaError_t iss_fasta_scatter(const uint8_t *d_text, uint64_t n_blocks, const uint64_t *d_pre, const uint64_t *d_body, const uint64_t *d_end,
                              const uint64_t *d_rank0, const uint64_t *d_stage_off, uint32_t n_recs, uint8_t *d_stage, cudaStream_t stream) {
    if (!n_recs) return cudaSuccess;
    fasta_scatter_kernel<<<grid_for(n_blocks, 148 * 16), 256, 0, stream>>>(d_text, n_blocks, d_pre, d_body, d_end, d_rank0, d_stage_off, n_recs, d_stage);
    return cudaGetLastError();
}

// ---- host: where the title lines are ----------------------------------------------------------------
static inline bool host_white(unsigned char b) { return b == 0x20 || (unsigned char)(b - 9u) < 5u; }

extern "C" int iss_fasta_index(const char *text, uint64_t n, iss_fasta_rec **out, uint32_t *n_out) {
    if (!out || !n_out || (!text && n)) return ISS_E_ARG;
    *out = nullptr;
    *n_out = 0;
    size_t cap = 1024, cnt = 0;
    iss_fasta_rec *recs = (iss_fasta_rec *)malloc(cap * sizeof(iss_fasta_rec));
    if (!recs) return ISS_E_ARG;
    // first title: '>' at the very start, or the first one behind a line end (text in front of it is ignored)
    uint64_t pos = n;
    for (uint64_t x = 0; x < n;) {
        const char *g = (const char *)memchr(text + x, '>', n - x);
        if (!g) break;
        const uint64_t at = (uint64_t)(g - text);
        if (at == 0 || text[at - 1] == '\n') { pos = at; break; }
        x = at + 1;
    }
    while (pos < n) {
        uint64_t end = n;  // the next title line's '>'
        for (uint64_t x = pos + 1; x < n;) {
            const char *g = (const char *)memchr(text + x, '>', n - x);
            if (!g) break;
            const uint64_t at = (uint64_t)(g - text);
            if (text[at - 1] == '\n') { end = at; break; }
            x = at + 1;
        }
        const char *nl = (const char *)memchr(text + pos, '\n', end - pos);
        const uint64_t eol = nl ? (uint64_t)(nl - text) : end;
        iss_fasta_rec f;
        memset(&f, 0, sizeof f);
        f.gt_pos = pos;
        f.title_off = pos + 1;
        uint64_t te = eol;
        while (te > f.title_off && host_white((unsigned char)text[te - 1])) --te;  // rstrip
        f.title_len = te - f.title_off;
        uint64_t a = f.title_off;
        while (a < te && host_white((unsigned char)text[a])) ++a;  // record.id: the first white-space delimited token
        uint64_t b = a;
        while (b < te && !host_white((unsigned char)text[b])) ++b;
        f.id_off = a;
        f.id_len = (uint32_t)(b - a);
        f.body_off = eol < end ? eol + 1 : end;
        f.body_len = end - f.body_off;
        f.seq_len = 0;
        if (cnt == cap) {
            cap *= 2;
            iss_fasta_rec *bigger = (iss_fasta_rec *)realloc(recs, cap * sizeof(iss_fasta_rec));
            if (!bigger) { free(recs); return ISS_E_ARG; }
            recs = bigger;
        }
        recs[cnt++] = f;
        pos = end;
    }
    *out = recs;
    *n_out = (uint32_t)cnt;
    return ISS_OK;
}

extern "C" void iss_fasta_free(iss_fasta_rec *recs) { free(recs); }
