// iss_io.cu -- files-written end of the path, on top of the public C ABI (iss_ctx_submit / next / release).
//
// The reference's worker writes its FASTQ with SeqIO.write into {prefix}_R1.fastq / _R2.fastq / .vcf
// (iss/generator.py:64-66, 227-238) and the parent then re-copies every byte into the final files
// (iss/app.py:123-133, iss/util.py:216-238).  Here the consumer side of the pinned ring is native: a small pool of
// threads pwrite()s each finished batch straight from the ring slot to its place in the output file -- for plain FASTQ
// that place is known in closed form (iss_fastq_bytes), so the bytes can go to their FINAL offsets and no second copy
// is needed -- while the GPU fills the other slots.  No Python object ever touches the read bytes.
//
// Also here: iss_probe_d2h, the measurement behind the choice of ring memory (profiles/r2_d2h_probe.*).
#include <cuda_runtime.h>
#include <errno.h>
#include <nvtx3/nvToolsExt.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/statvfs.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/iss_b200.h"
#include "iss_hostmem.h"

extern "C" void iss_set_last_error(const char *msg);  // iss_runtime.cu

namespace {

struct WriteTask {
    int fd;             // < 0: memcpy into `dst` (a mapping of the output file)
    const uint8_t *ptr;
    uint64_t len;
    int64_t off;        // fd >= 0: file offset for pwrite; < 0: append (write(2) on an O_APPEND descriptor, one thread at a time)
    uint8_t *dst;
};

class WriterPool {
public:
    explicit WriterPool(unsigned n) {
        for (unsigned k = 0; k < n; ++k) threads_.emplace_back([this] { run(); });
    }
    ~WriterPool() {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    void push(const WriteTask &t) {
        { std::lock_guard<std::mutex> g(m_); q_.push_back(t); ++pending_; }
        cv_.notify_one();
    }
    int wait() {  // all pushed tasks done; returns the first errno seen (0 = none)
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [this] { return pending_ == 0; });
        return err_;
    }

private:
    void run() {
        for (;;) {
            WriteTask t;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [this] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                t = q_.front();
                q_.pop_front();
            }
            int e = 0;
            uint64_t done = 0;
            if (t.fd < 0) {  // a piece of a mapped file: page faults of different pieces are served in parallel
                memcpy(t.dst, t.ptr, t.len);
                done = t.len;
            }
            while (done < t.len) {
                const ssize_t r = t.off < 0 ? write(t.fd, t.ptr + done, t.len - done)
                                            : pwrite(t.fd, t.ptr + done, t.len - done, (off_t)(t.off + (int64_t)done));
                if (r < 0) { if (errno == EINTR) continue; e = errno; break; }
                done += (uint64_t)r;
            }
            {
                std::lock_guard<std::mutex> g(m_);
                if (e && !err_) err_ = e;
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    std::vector<std::thread> threads_;
    std::deque<WriteTask> q_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    uint64_t pending_ = 0;
    int err_ = 0;
    bool stop_ = false;
};

int fail_io(const std::string &msg) {
    iss_set_last_error(msg.c_str());
    return ISS_E_IO;
}

// One output stream.  Buffered write(2)/pwrite(2) to ONE file are serialised by the inode lock whatever the number of
// threads (measured: 2.7 GB/s into tmpfs with 4 or with 16 threads), so regular files are written through a shared mapping
// of the batch's byte range instead: ftruncate to the new end, mmap, parallel memcpy of 4 MB pieces, munmap.  Anything
// that is not a regular file (/dev/null, a pipe) and ISS_WRITER=pwrite take the write(2) path.
struct OutFile {
    int fd = -1;
    bool mapped = false;
    uint64_t size = 0;  // current file size (mapped mode)
    uint64_t pos = 0;   // where the next batch goes
};

bool room_for(int fd, uint64_t bytes) {  // a mapping of a full file system ends in SIGBUS, not in an error code
    struct statvfs v;
    if (fstatvfs(fd, &v) != 0) return true;
    return (uint64_t)v.f_bavail * v.f_frsize > bytes + (64ull << 20);
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

int iss_ctx_write_files(iss_ctx *ctx, const iss_work_item *items, uint64_t n_items, const iss_gen_opts *opts,
                        const iss_file_sink *sink, iss_write_stats *out) {
    if (!ctx || !sink) { iss_set_last_error("null argument"); return ISS_E_ARG; }
    iss_write_stats st{};
    const double t0 = now_ms();
    int fd[3] = {-1, -1, -1};
    const char *paths[3] = {sink->path_r1, sink->path_r2, sink->path_vcf};
    const char *wmode = getenv("ISS_WRITER");
    const bool want_map = !(wmode && !strcmp(wmode, "pwrite"));
    OutFile of[2];
    for (int k = 0; k < 3; ++k) {
        if (!paths[k]) continue;
        int flags = O_CREAT | (sink->truncate ? O_TRUNC : 0);
        struct stat sb;
        const bool regular = k < 2 && want_map && (stat(paths[k], &sb) != 0 || S_ISREG(sb.st_mode));
        flags |= regular ? O_RDWR : O_WRONLY;  // a shared mapping needs read access
        if (!regular && (sink->append || k == 2)) flags |= O_APPEND;
        fd[k] = open(paths[k], flags, 0666);
        if (fd[k] < 0) {
            const std::string msg = std::string("cannot open ") + paths[k] + ": " + strerror(errno);
            for (int j = 0; j < k; ++j) if (fd[j] >= 0) close(fd[j]);
            return fail_io(msg);
        }
        if (k < 2) {
            of[k].fd = fd[k];
            of[k].mapped = regular && fstat(fd[k], &sb) == 0 && S_ISREG(sb.st_mode);
            of[k].size = of[k].mapped ? (uint64_t)sb.st_size : 0;
            of[k].pos = sink->append ? of[k].size : (k == 0 ? sink->offset_r1 : sink->offset_r2);
            if (!of[k].mapped && (sink->append)) { /* O_APPEND was not set for a would-be mapping that is not regular after all */
                fcntl(fd[k], F_SETFL, fcntl(fd[k], F_GETFL) | O_APPEND);
            }
        }
    }
    auto close_all = [&] { for (int k = 0; k < 3; ++k) if (fd[k] >= 0) close(fd[k]); };
    int rc = iss_ctx_submit(ctx, items, n_items, opts);
    if (rc != ISS_OK) { close_all(); return rc; }
    const unsigned nthreads = sink->writer_threads ? sink->writer_threads : 8u;
    const uint64_t piece = 4ull << 20;
    const uint64_t page = (uint64_t)sysconf(_SC_PAGESIZE);
    int io_err = 0;
    {
        WriterPool pool(nthreads);
        iss_batch b;
        int have = 0;
        for (;;) {
            rc = iss_ctx_next(ctx, &b, &have);
            if (rc != ISS_OK || !have) break;
            const double w0 = now_ms();
            nvtxRangePushA("iss:write_batch");
            const uint8_t *src[2] = {b.r1, b.r2};
            const uint64_t len[2] = {b.nbytes, b.nbytes2};
            struct Map { uint8_t *base; uint64_t bytes; } maps[2] = {{nullptr, 0}, {nullptr, 0}};
            for (int k = 0; k < 2; ++k) {
                if (fd[k] < 0 || !src[k] || !len[k]) continue;
                OutFile &f = of[k];
                if (f.mapped) {
                    const uint64_t end = f.pos + len[k];
                    if (end > f.size) {  // grow the file -- never shrink it: another worker may have extended it meanwhile
                        struct stat sb;
                        if (fstat(f.fd, &sb) == 0 && (uint64_t)sb.st_size > f.size) f.size = (uint64_t)sb.st_size;
                        if (end > f.size) {
                            if (!room_for(f.fd, end - f.size) || ftruncate(f.fd, (off_t)end) != 0) { if (!io_err) io_err = errno ? errno : ENOSPC; continue; }
                            f.size = end;
                        }
                    }
                    const uint64_t m0 = f.pos & ~(page - 1), mlen = end - m0;
                    void *mp = mmap(nullptr, mlen, PROT_READ | PROT_WRITE, MAP_SHARED, f.fd, (off_t)m0);
                    if (mp == MAP_FAILED) {  // e.g. a file system without shared mappings: fall back to pwrite for good
                        f.mapped = false;
                    } else {
                        maps[k] = Map{(uint8_t *)mp, mlen};
                        uint8_t *dst = (uint8_t *)mp + (f.pos - m0);
                        for (uint64_t at = 0; at < len[k]; at += piece)
                            pool.push(WriteTask{-1, src[k] + at, std::min(piece, len[k] - at), 0, dst + at});
                    }
                }
                if (!f.mapped) {
                    if (sink->append) {
                        pool.push(WriteTask{fd[k], src[k], len[k], -1, nullptr});  // one task per stream: appends stay ordered
                    } else {
                        for (uint64_t at = 0; at < len[k]; at += piece)
                            pool.push(WriteTask{fd[k], src[k] + at, std::min(piece, len[k] - at), (int64_t)(f.pos + at), nullptr});
                    }
                }
                f.pos += len[k];
            }
            if (fd[2] >= 0 && b.vcf && b.vcf_bytes) {
                pool.push(WriteTask{fd[2], b.vcf, b.vcf_bytes, -1, nullptr});
                st.bytes_vcf += b.vcf_bytes;
            }
            const int e = pool.wait();  // the slot may be refilled once it is released
            for (int k = 0; k < 2; ++k)
                if (maps[k].base) munmap(maps[k].base, maps[k].bytes);
            nvtxRangePop();
            if (e && !io_err) io_err = e;
            st.write_ms += now_ms() - w0;
            st.bytes_r1 += b.nbytes;
            st.bytes_r2 += b.nbytes2;
            st.n_pairs += b.n_pairs;
            st.n_batches += 1;
            rc = iss_ctx_release(ctx, &b);
            if (rc != ISS_OK) break;
            if (io_err) { iss_ctx_abort(ctx); break; }
        }
    }
    close_all();
    st.wall_ms = now_ms() - t0;
    if (out) *out = st;
    if (rc != ISS_OK) return rc;
    if (io_err) return fail_io(std::string("write failed: ") + strerror(io_err));
    return ISS_OK;
}

// Sustained device -> host copy rate into a ring of `ring_bytes` obtained the way the runtime obtains its slots.
int iss_probe_d2h(int device, int alloc_mode, uint64_t ring_bytes, uint64_t chunk_bytes, double seconds, double *gbps, int *node) {
    if (!gbps || chunk_bytes == 0 || ring_bytes < chunk_bytes) { iss_set_last_error("bad probe arguments"); return ISS_E_ARG; }
    if (cudaSetDevice(device) != cudaSuccess) { iss_set_last_error("cudaSetDevice failed"); return ISS_E_CUDA; }
    IssHostBuf hb;
    uint8_t *d = nullptr;
    cudaStream_t s = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = ISS_OK;
    do {
        if (iss_host_alloc(&hb, ring_bytes, alloc_mode, device) != 0) { rc = ISS_E_CUDA; break; }
        if (cudaMalloc(&d, chunk_bytes) != cudaSuccess || cudaMemset(d, 0x41, chunk_bytes) != cudaSuccess) { rc = ISS_E_CUDA; break; }
        cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        const uint64_t nslot = ring_bytes / chunk_bytes;
        uint64_t k = 0, copied = 0;
        for (int w = 0; w < 4; ++w) cudaMemcpyAsync(hb.ptr + (k++ % nslot) * chunk_bytes, d, chunk_bytes, cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
        const double t_end = now_ms() + seconds * 1e3;
        cudaEventRecord(e0, s);
        const int per_sync = (int)std::max<uint64_t>(8, (64ull << 20) / chunk_bytes);
        while (now_ms() < t_end) {
            for (int w = 0; w < per_sync; ++w) {
                cudaMemcpyAsync(hb.ptr + (k++ % nslot) * chunk_bytes, d, chunk_bytes, cudaMemcpyDeviceToHost, s);
                copied += chunk_bytes;
            }
            cudaStreamSynchronize(s);
        }
        cudaEventRecord(e1, s);
        if (cudaEventSynchronize(e1) != cudaSuccess) { rc = ISS_E_CUDA; break; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        *gbps = ms > 0 ? (double)copied / (ms * 1e-3) / 1e9 : 0.0;
        if (node) *node = hb.node >= 0 ? hb.node : iss_page_node(hb.ptr);
    } while (0);
    if (rc != ISS_OK) iss_set_last_error(cudaGetErrorString(cudaGetLastError()));
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (s) cudaStreamDestroy(s);
    if (d) cudaFree(d);
    iss_host_free(&hb);
    return rc;
}

}  // extern "C"
