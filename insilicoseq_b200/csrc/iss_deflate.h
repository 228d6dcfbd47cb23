// iss_deflate.h -- GPU gzip (implemented in iss_deflate.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define ISS_GZ_CHUNK 65536u   // input bytes per gzip member
#define ISS_GZ_STRIDE 65664u  // worst-case bytes of one member: stored blocks (10 + 2 * 5 + 65536 + 8), rounded up to 128
#define ISS_GZ_WORK_WORDS 1024u  // u32 words of per-call scratch (sample histogram, chunk ticket, the call's Huffman code)
#define ISS_GZ_FAILED(total) ((total) >> 61)  // the look-back of a member timed out: the stream is invalid

cudaError_t iss_gz_init_device(void);
uint32_t iss_gz_chunks(uint64_t n_in);
// d_in: 16-byte aligned, readable up to the next multiple of 16 bytes.  d_status: chunks + 1 u64 of scratch; after the
// call d_status[chunks] = total bytes of the gzip stream (check ISS_GZ_FAILED).  d_out: the contiguous gzip stream
// (capacity chunks * ISS_GZ_STRIDE is always enough).  d_work: ISS_GZ_WORK_WORDS u32.
// reuse_code != 0: skip the sampling and the code construction and compress with the code that an earlier call with the same
// d_work built (streams of one run are statistically alike; any complete code is valid for any input).
cudaError_t iss_gz_compress(const uint8_t *d_in, uint64_t n_in, uint64_t *d_status, uint8_t *d_out, uint32_t *d_work, int sm_count,
                            int reuse_code, cudaStream_t stream);
