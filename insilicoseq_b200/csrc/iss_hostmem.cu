// iss_hostmem.cu -- see iss_hostmem.h.  Raw syscalls (mbind, move_pages): libnuma is not assumed to be installed.
#include "iss_hostmem.h"

#include <cuda_runtime.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#ifndef MPOL_BIND
#define MPOL_BIND 2
#endif
#ifndef MPOL_MF_MOVE
#define MPOL_MF_MOVE 2
#endif

int iss_gpu_numa_node(int device) {
    char bdf[32] = {0};
    if (cudaDeviceGetPCIBusId(bdf, sizeof bdf, device) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char *q = bdf; *q; ++q)
        if (*q >= 'A' && *q <= 'Z') *q = (char)(*q - 'A' + 'a');
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bdf);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

namespace {
// CPUs of a NUMA node that this thread may run on (empty set -> false)
bool node_cpus(int node, cpu_set_t *out) {
    char path[96];
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    FILE *f = fopen(path, "r");
    if (!f) return false;
    char buf[4096] = {0};
    const size_t n = fread(buf, 1, sizeof buf - 1, f);
    fclose(f);
    buf[n] = 0;
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return false;
    CPU_ZERO(out);
    int any = 0;
    for (char *q = buf; *q;) {
        char *end;
        long a = strtol(q, &end, 10), b = a;
        if (end == q) break;
        if (*end == '-') { q = end + 1; b = strtol(q, &end, 10); }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c)
            if (CPU_ISSET(c, &allowed)) { CPU_SET(c, out); any = 1; }
        q = (*end == ',') ? end + 1 : end;
        if (*end != ',' ) break;
    }
    return any != 0;
}
}  // namespace

int iss_page_node(const void *p) {
#ifdef SYS_move_pages
    void *pages[1] = {(void *)((uintptr_t)p & ~(uintptr_t)4095)};
    int status[1] = {-1};
    if (syscall(SYS_move_pages, 0, 1ul, pages, (const int *)nullptr, status, 0) != 0) return -1;
    return status[0];
#else
    (void)p;
    return -1;
#endif
}

int iss_host_alloc(IssHostBuf *b, size_t bytes, int mode, int device) {
    b->ptr = nullptr; b->bytes = bytes; b->mode = mode; b->node = -1;
    if (bytes == 0) return 0;
    if (mode == ISS_HOSTMEM_WC) {
        void *p = nullptr;
        cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocWriteCombined);
        if (e != cudaSuccess) return (int)e;
        b->ptr = (uint8_t *)p;
        return 0;
    }
    if (mode == ISS_HOSTMEM_NUMA) {
        const int node = iss_gpu_numa_node(device);
        const size_t len = (bytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
        void *p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p != MAP_FAILED) {
            madvise(p, len, MADV_HUGEPAGE);
            bool bound = false;
#ifdef SYS_mbind
            if (node >= 0 && node < 64) {
                unsigned long mask = 1ul << node;
                bound = syscall(SYS_mbind, p, len, MPOL_BIND, &mask, 65ul, 0u) == 0;
            }
#endif
            // first touch: with the policy set the pages come from `node`; where mbind is not permitted (seccomp in
            // containers) the thread moves onto the node's CPUs for the touch, which places the pages the same way
            cpu_set_t old, target;
            bool moved = false;
            if (!bound && node >= 0 && node_cpus(node, &target) && sched_getaffinity(0, sizeof old, &old) == 0)
                moved = sched_setaffinity(0, sizeof target, &target) == 0;
            for (size_t off = 0; off < len; off += 4096) ((volatile uint8_t *)p)[off] = 0;
            if (moved) sched_setaffinity(0, sizeof old, &old);
            const int actual = iss_page_node(p);  // where the first page really is (-1 when move_pages is not permitted either)
            cudaError_t e = cudaHostRegister(p, len, cudaHostRegisterPortable);
            if (e == cudaSuccess) {
                b->ptr = (uint8_t *)p;
                b->bytes = len;
                b->node = actual >= 0 ? actual : ((bound || moved) ? node : -1);
                return 0;
            }
            cudaGetLastError();
            munmap(p, len);
        }
        b->mode = mode = ISS_HOSTMEM_PINNED;  // containers without mbind / registration rights: plain pinned memory
    }
    void *p = nullptr;
    cudaError_t e = cudaMallocHost(&p, bytes);
    if (e != cudaSuccess) return (int)e;
    b->ptr = (uint8_t *)p;
    return 0;
}

void iss_host_free(IssHostBuf *b) {
    if (!b->ptr) return;
    if (b->mode == ISS_HOSTMEM_NUMA) {
        cudaHostUnregister(b->ptr);
        munmap(b->ptr, b->bytes);
    } else {
        cudaFreeHost(b->ptr);
    }
    b->ptr = nullptr;
}
