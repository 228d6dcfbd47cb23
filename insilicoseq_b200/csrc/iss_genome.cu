// iss_genome.cu -- records (ASCII, as Bio.SeqIO.parse yields them: iss/generator.py:270) -> the device genome blob.
//
// The reference keeps every record as a Python str inside a SeqRecord and pickles it to each worker; here the bytes go to
// HBM once and are classified and packed THERE: a record of pure upper-case ACGT becomes 2-bit words (16 bases per u32,
// code (ascii >> 1) & 3), anything else (lower case, IUPAC codes) stays as bytes; a character outside util.rev_comp's
// dictionary (iss/util.py:57-88) is reported with its record and position.  Host work is one pass of memcpy into the
// driver's staging; 4.5 Gbp used to take 3.4 s of host packing.
#include <cuda_runtime.h>
#include <stdint.h>

#include "iss_device.cuh"
#include "iss_genome.h"

namespace {

// util.rev_comp's dictionary (iss/util.py:57-88): ACGT + YRWSKMNBVDH in either case
__device__ __forceinline__ bool in_alphabet(uint32_t ch) {
    const uint32_t u = (ch & 0xDFu) - 65u;
    // A B C D G H K M N R S T V W Y
    const uint32_t mask = (1u << 0) | (1u << 1) | (1u << 2) | (1u << 3) | (1u << 6) | (1u << 7) | (1u << 10) | (1u << 12) |
                          (1u << 13) | (1u << 17) | (1u << 18) | (1u << 19) | (1u << 21) | (1u << 22) | (1u << 24);
    return u < 26u && ((mask >> u) & 1u) && (ch >= 'A');
}

__device__ __forceinline__ uint32_t find_record(const uint64_t *__restrict__ unit_begin, uint32_t n, uint64_t unit) {
    uint32_t lo = 0, hi = n;  // largest r with unit_begin[r] <= unit
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (unit_begin[mid] <= unit) lo = mid; else hi = mid;
    }
    return lo;
}

}  // namespace

// One thread per 16 staged bytes of one record.  flags[r] |= 1 when a byte is not one of "ACGT"; bad[r] = lowest
// position of a byte outside the reference's alphabet.
__global__ void __launch_bounds__(256) iss_genome_classify_kernel(const uint8_t *__restrict__ stage, const IssStagedRecord *__restrict__ recs,
                                                                  const uint64_t *__restrict__ unit_begin, uint32_t n, uint64_t n_units,
                                                                  uint32_t *__restrict__ flags, unsigned long long *__restrict__ bad) {
    for (uint64_t unit = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; unit < n_units; unit += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = find_record(unit_begin, n, unit);
        const uint64_t k0 = (unit - unit_begin[r]) * 16u, len = recs[r].len;
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(stage + recs[r].stage_off + k0));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t impure = 0;
        unsigned long long first_bad = ~0ull;
#pragma unroll
        for (uint32_t m = 0; m < 16; ++m) {
            if (k0 + m >= len) break;
            const uint32_t ch = (w[m >> 2] >> (8u * (m & 3u))) & 0xFFu;
            if (!(ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T')) {
                impure = 1;
                if (!in_alphabet(ch) && first_bad == ~0ull) first_bad = k0 + m;
            }
        }
        if (impure) atomicOr(&flags[r], 1u);
        if (first_bad != ~0ull) atomicMin(&bad[r], first_bad);
    }
}

// One thread per 16 bases: a 2-bit word of a pure record, or 16 bytes of a byte record, into the blob.
__global__ void __launch_bounds__(256) iss_genome_pack_kernel(const uint8_t *__restrict__ stage, const IssStagedRecord *__restrict__ recs,
                                                              const uint64_t *__restrict__ unit_begin, uint32_t n, uint64_t n_units,
                                                              uint32_t *__restrict__ packed, uint8_t *__restrict__ ascii) {
    for (uint64_t unit = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; unit < n_units; unit += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = find_record(unit_begin, n, unit);
        const IssStagedRecord rec = recs[r];
        const uint64_t u = unit - unit_begin[r], k0 = u * 16u;
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(stage + rec.stage_off + k0));
        if (rec.enc == ISS_ENC_2BIT) {
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            uint32_t out = 0;
#pragma unroll
            for (uint32_t m = 0; m < 16; ++m) {
                if (k0 + m >= rec.len) break;
                out |= (((w[m >> 2] >> (8u * (m & 3u))) >> 1) & 3u) << (2u * m);
            }
            packed[(rec.data_off >> 4) + u] = out;  // data_off is a multiple of 64 bases
        } else {
            uint8_t *dst = ascii + rec.data_off + k0;  // data_off is a multiple of 16 bytes
            if (k0 + 16u <= rec.len) {
                *reinterpret_cast<uint4 *>(dst) = v;
            } else {
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                for (uint32_t m = 0; k0 + m < rec.len; ++m) dst[m] = (uint8_t)(w[m >> 2] >> (8u * (m & 3u)));
            }
        }
    }
}

cudaError_t iss_genome_classify(const uint8_t *d_stage, const IssStagedRecord *d_recs, const uint64_t *d_unit_begin, uint32_t n, uint64_t n_units,
                                uint32_t *d_flags, unsigned long long *d_bad, int sm_count, cudaStream_t stream) {
    if (!n_units) return cudaSuccess;
    const uint64_t want = (n_units + 255) / 256;
    const uint32_t grid = (uint32_t)(want < (uint64_t)sm_count * 16 ? want : (uint64_t)sm_count * 16);
    iss_genome_classify_kernel<<<grid, 256, 0, stream>>>(d_stage, d_recs, d_unit_begin, n, n_units, d_flags, d_bad);
    return cudaGetLastError();
}

cudaError_t iss_genome_pack(const uint8_t *d_stage, const IssStagedRecord *d_recs, const uint64_t *d_unit_begin, uint32_t n, uint64_t n_units,
                            uint32_t *d_packed, uint8_t *d_ascii, int sm_count, cudaStream_t stream) {
    if (!n_units) return cudaSuccess;
    const uint64_t want = (n_units + 255) / 256;
    const uint32_t grid = (uint32_t)(want < (uint64_t)sm_count * 16 ? want : (uint64_t)sm_count * 16);
    iss_genome_pack_kernel<<<grid, 256, 0, stream>>>(d_stage, d_recs, d_unit_begin, n, n_units, d_packed, d_ascii);
    return cudaGetLastError();
}
