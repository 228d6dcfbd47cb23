// iss_deflate.h -- GPU gzip (implemented in iss_deflate.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define ISS_GZ_CHUNK 65536u   // input bytes per gzip member
#define ISS_GZ_STRIDE 65664u  // staging slot per member: worst case = stored blocks (10 + 2 * 5 + 65536 + 8), 128 B aligned

cudaError_t iss_gz_init_device(void);
uint32_t iss_gz_chunks(uint64_t n_in);
// d_stage: chunks * ISS_GZ_STRIDE bytes; d_sizes: chunks u32; d_offsets: chunks + 1 u64 (offsets[chunks] = total bytes);
// d_out: the contiguous gzip stream (capacity >= chunks * ISS_GZ_STRIDE is always enough)
cudaError_t iss_gz_compress(const uint8_t *d_in, uint64_t n_in, uint8_t *d_stage, uint32_t *d_sizes, uint64_t *d_offsets, uint8_t *d_out,
                            int sm_count, cudaStream_t stream);
