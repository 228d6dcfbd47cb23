// iss_kernels.h -- launch helpers implemented in iss_kernels.cu
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "iss_device.cuh"

size_t iss_tile_smem_bytes(const Lay &l);
cudaError_t iss_kernel_configure(size_t max_smem);
cudaError_t iss_kernel_carveout(int percent);
cudaError_t iss_launch_generate(const GenParams &p, uint32_t grid, size_t smem, cudaStream_t stream);
cudaError_t iss_occupancy(uint32_t T, uint32_t split_shift, uint32_t q_blocks, uint32_t q_short, size_t smem, int *blocks_per_sm);
bool iss_has_short_row_build(uint32_t T, uint32_t split_shift);  // the geometry has a build for 8-entry rows
cudaError_t iss_launch_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t *dev_out, cudaStream_t stream);
