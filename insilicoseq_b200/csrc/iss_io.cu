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
#include <string.h>
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
    int fd;
    const uint8_t *ptr;
    uint64_t len;
    int64_t off;  // < 0: append (write(2) on an O_APPEND descriptor; only ever issued by one thread at a time)
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
    for (int k = 0; k < 3; ++k) {
        if (!paths[k]) continue;
        int flags = O_WRONLY | O_CREAT | (sink->truncate ? O_TRUNC : 0);
        if (sink->append || k == 2) flags |= O_APPEND;
        fd[k] = open(paths[k], flags, 0666);
        if (fd[k] < 0) {
            const std::string msg = std::string("cannot open ") + paths[k] + ": " + strerror(errno);
            for (int j = 0; j < k; ++j) if (fd[j] >= 0) close(fd[j]);
            return fail_io(msg);
        }
    }
    auto close_all = [&] { for (int k = 0; k < 3; ++k) if (fd[k] >= 0) close(fd[k]); };
    int rc = iss_ctx_submit(ctx, items, n_items, opts);
    if (rc != ISS_OK) { close_all(); return rc; }
    const unsigned nthreads = sink->writer_threads ? sink->writer_threads : 8u;
    const uint64_t piece = 4ull << 20;
    int io_err = 0;
    {
        WriterPool pool(nthreads);
        uint64_t off[2] = {sink->offset_r1, sink->offset_r2};
        iss_batch b;
        int have = 0;
        for (;;) {
            rc = iss_ctx_next(ctx, &b, &have);
            if (rc != ISS_OK || !have) break;
            const double w0 = now_ms();
            nvtxRangePushA("iss:write_batch");
            const uint8_t *src[2] = {b.r1, b.r2};
            const uint64_t len[2] = {b.nbytes, b.nbytes2};
            for (int k = 0; k < 2; ++k) {
                if (fd[k] < 0 || !src[k]) continue;
                if (sink->append) {
                    pool.push(WriteTask{fd[k], src[k], len[k], -1});  // one task per stream: appends stay ordered
                } else {
                    for (uint64_t at = 0; at < len[k]; at += piece)
                        pool.push(WriteTask{fd[k], src[k] + at, std::min(piece, len[k] - at), (int64_t)(off[k] + at)});
                }
                off[k] += len[k];
            }
            if (fd[2] >= 0 && b.vcf && b.vcf_bytes) {
                pool.push(WriteTask{fd[2], b.vcf, b.vcf_bytes, -1});
                st.bytes_vcf += b.vcf_bytes;
            }
            const int e = pool.wait();  // the slot may be refilled once it is released
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
