// iss_hostmem.h -- pinned host memory for the output ring, placed next to the GPU's PCIe root.
#pragma once
#include <stddef.h>
#include <stdint.h>

// How a ring slot's host buffer is obtained (ISS_RING_ALLOC=pinned|numa|wc, default chosen by the runtime):
//   ISS_HOSTMEM_PINNED  cudaMallocHost: pinned, cacheable; the pages land on the NUMA node of the calling thread
//   ISS_HOSTMEM_NUMA    mmap + mbind(node of the GPU) + first touch + cudaHostRegister: placement does not depend on
//                       which CPU the caller happens to run on (launchers such as torchrun do not bind ranks)
//   ISS_HOSTMEM_WC      cudaHostAlloc(cudaHostAllocWriteCombined): no snooping on device writes, but CPU reads are
//                       uncached -- only for consumers that never read the bytes on the CPU
enum { ISS_HOSTMEM_PINNED = 0, ISS_HOSTMEM_NUMA = 1, ISS_HOSTMEM_WC = 2 };

struct IssHostBuf {
    uint8_t *ptr = nullptr;
    size_t bytes = 0;
    int mode = ISS_HOSTMEM_PINNED;
    int node = -1;  // NUMA node the pages were bound to (-1: not bound / unknown)
};

int iss_gpu_numa_node(int device);  // node of the device's PCIe function, or -1
// cudaError_t as int (0 = success).  device: the current CUDA device the buffer will be used with.
int iss_host_alloc(IssHostBuf *b, size_t bytes, int mode, int device);
void iss_host_free(IssHostBuf *b);
// NUMA node that currently backs the page at p (move_pages query), or -1
int iss_page_node(const void *p);
