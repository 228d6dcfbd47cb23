// iss_genome.h -- device-side classification and 2-bit packing of the genome records (iss_genome.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct IssStagedRecord {
    uint64_t stage_off;  // byte offset of the record's ASCII in the staging buffer (multiple of 16; readable 16 bytes past its end)
    uint64_t len;
    uint64_t data_off;   // pack: DevRecord::data_off
    uint32_t enc;        // pack: ISS_ENC_2BIT / ISS_ENC_ASCII
    uint32_t pad;
};

// unit_begin[r] = units (of 16 bases) of the records before r; n_units = their total.
cudaError_t iss_genome_classify(const uint8_t *d_stage, const IssStagedRecord *d_recs, const uint64_t *d_unit_begin, uint32_t n, uint64_t n_units,
                                uint32_t *d_flags, unsigned long long *d_bad, int sm_count, cudaStream_t stream);
cudaError_t iss_genome_pack(const uint8_t *d_stage, const IssStagedRecord *d_recs, const uint64_t *d_unit_begin, uint32_t n, uint64_t n_units,
                            uint32_t *d_packed, uint8_t *d_ascii, int sm_count, cudaStream_t stream);
