// iss_fasta.h -- FASTA text to per-record ASCII on the device (iss_fasta.cu): title lines blanked, white space squeezed out
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// d_text: the text padded with '\n' to whole 4096-byte blocks.  d_gt[r] = position of record r's '>', d_body[r] = first
// byte behind its title line: [d_gt[r], d_body[r]) becomes '\n', and so does everything in front of the first title.
cudaError_t iss_fasta_blank(uint8_t *d_text, const uint64_t *d_gt, const uint64_t *d_body, uint32_t n_recs, cudaStream_t stream);
// d_cnt[b] = bytes of block b that are not white space (space, \t \n \v \f \r -- what bytes.split() and Bio strip)
cudaError_t iss_fasta_count(const uint8_t *d_text, uint64_t n_blocks, uint32_t *d_cnt, cudaStream_t stream);
// d_pre[b] = sum of d_cnt[0..b)  (d_pre has n_blocks + 1 entries)
cudaError_t iss_fasta_scan(const uint32_t *d_cnt, uint64_t n_blocks, uint64_t *d_pre, cudaStream_t stream);
// d_rank[k] = non-white bytes in front of position d_pos[k]
cudaError_t iss_fasta_ranks(const uint8_t *d_text, const uint64_t *d_pre, const uint64_t *d_pos, uint32_t n_pos, uint64_t *d_rank, cudaStream_t stream);
// every non-white byte of a record body [d_body[r], d_end[r]) -> d_stage[d_stage_off[r] + (its rank - d_rank0[r])]
cudaError_t iss_fasta_scatter(const uint8_t *d_text, uint64_t n_blocks, const uint64_t *d_pre, const uint64_t *d_body, const uint64_t *d_end,
                              const uint64_t *d_rank0, const uint64_t *d_stage_off, uint32_t n_recs, uint8_t *d_stage, cudaStream_t stream);
