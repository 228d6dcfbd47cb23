// iss_runtime.cu -- host runtime behind the C ABI (include/iss_b200.h).
//
//   * genome blob   : FASTA records -> 2-bit packed words (pure ACGT) or bytes (anything else), one HBM allocation
//   * model blob    : integer-threshold tables + the list of positions with non-zero indel thresholds
//   * batch planner : work items -> batches of <= batch_pairs pairs -> segments with closed-form output offsets
//   * ring pipeline : n_slots x {device FASTQ buffers, pinned host buffers, stream, events}; kernel of batch k+1
//                     overlaps the D2H copy of batch k; the consumer drives it through iss_ctx_next/release
//   * NCCL          : one broadcast of the two blobs at init (rank mode for one-process-per-GPU, group mode for
//                     one process driving several GPUs); no communication afterwards
#include <cuda_runtime.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>  // header-only: ranges show up in Nsight Systems / ncu --nvtx timelines, cost ~nothing otherwise
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <functional>
#include <vector>

#include "../../include/iss_b200.h"
#include "iss_device.cuh"
#include "iss_kernels.h"
#include "iss_deflate.h"
#include "iss_fasta.h"
#include "iss_genome.h"
#include "iss_hostmem.h"

namespace {

struct NvtxRange {  // RAII range: "iss:plan+launch", "iss:wait_batch", "iss:vcf"
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

#define CU(call)                                                                                           \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess)                                                                            \
            return fail(ISS_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));                  \
    } while (0)
#define NC(call)                                                                                           \
    do {                                                                                                   \
        ncclResult_t r__ = (call);                                                                         \
        if (r__ != ncclSuccess) return fail(ISS_E_CUDA, std::string(#call) + ": " + ncclGetErrorString(r__)); \
    } while (0)

constexpr uint64_t kGenomeMagic = 0x31474e4547535349ull;  // "ISSGENG1"
constexpr uint64_t kModelMagic = 0x314c444f4d535349ull;   // "ISSMODL1"
constexpr size_t kMaxSmem = 232448;                        // 227 KB opt-in per CTA on sm_100
constexpr uint32_t kErrOvfCap = 4096;                      // overflow of the substitution list per CTA, in global memory (16 KB)
constexpr uint32_t kTieCap = 64;                           // exact-pass queue per tile (~ 11 ties expected per 64-pair tile at L = 150)

inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

struct Slot {
    uint8_t *d_out[2] = {nullptr, nullptr};
    uint8_t *h_out[2] = {nullptr, nullptr};
    IssHostBuf h_buf[2];
    DevSegment *d_segs = nullptr;
    DevSegment *h_segs = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr, ev_done = nullptr;
    cudaEvent_t ev_gz = nullptr, ev_end = nullptr;  // after the compressor; ISS_TRACE only: after the queued copies
    bool gz_code_ready[2] = {false, false};  // the slot's work areas hold the Huffman codes of this run's streams
    cudaStream_t zstream = nullptr;  // the slot's compressor kernels: a stream of the highest priority (see launch_batch)
    int state = 0;  // 0 free, 1 in flight, 2 handed to the consumer
    uint64_t nbytes = 0, n_pairs = 0, first_item = 0, batch_index = 0;
    // --store_mutations: device event list -> host text
    uint4 *d_events = nullptr;
    unsigned long long *d_evcount = nullptr;
    unsigned long long *h_evcount = nullptr;
    uint64_t event_cap = 0;
    uint32_t nseg = 0;
    // opts.compress: staging slots, member sizes/offsets, packed gzip streams, totals read back
    uint8_t *d_gz_out[2] = {nullptr, nullptr};
    uint32_t *d_gz_work[2] = {nullptr, nullptr};
    uint64_t *d_gz_status[2] = {nullptr, nullptr};
    uint64_t *h_gz_total = nullptr;  // pinned [2]
    uint64_t gz_pre[2] = {0, 0};     // compressed bytes whose D2H copy was queued before the total was known
    uint64_t gz_chunks_cap = 0;
    uint32_t gz_chunks = 0;
    uint64_t nbytes2 = 0;
    uint8_t *d_scratch = nullptr;  // slow-path scratch of the kernel running in this slot (kernels of different slots overlap)
    uint64_t scratch_bytes = 0;    // (the substitution list's overflow stretches sit behind the replay scratch)
    std::vector<uint4> h_events;
    std::string vcf;
};

struct HostRecord {
    uint64_t len;
    uint32_t name_len;
};

}  // namespace

struct iss_ctx {
    int device = 0;
    int sm_count = 0;
    // blobs
    uint8_t *d_genome = nullptr;
    uint64_t genome_bytes = 0;
    uint8_t *d_model = nullptr;
    uint64_t model_bytes = 0;
    GenomeHeader gh{};
    ModelHeader mh{};
    DevGenome gview{};
    DevModel mview{};
    std::vector<HostRecord> recs;
    std::vector<std::string> rec_names;
    bool have_genome = false, have_model = false;
    // generation state
    std::vector<iss_work_item> items;
    iss_gen_opts opts{};
    uint64_t cur_item = 0, cur_off = 0;  // planner cursor
    uint64_t batch_pairs = 0, slot_cap = 0, max_rec = 0;
    uint32_t n_slots = 0, max_segs = 0;
    uint32_t tile_pairs = 64, img_cap = 0, err_cap = 512, grid = 0, split_shift = 1;
    uint32_t q_stage_cap = 0;  // quality-row pipeline: bytes per stage
    uint32_t q_blocks = 1;     //   and Philox blocks (2 positions each) per stage
    uint32_t q_short = 0;      //   the staged rows are the model's short (8-entry) ones
    uint64_t scratch_replay_bytes = 0;  // replay part of a slot's scratch; the substitution overflow lists follow
    Lay lay{};
    size_t smem = 0;
    std::vector<Slot> slots;
    uint64_t next_launch = 0, next_return = 0;  // batch counters
    cudaEvent_t gen_done[2] = {nullptr, nullptr};  // end of the generate kernels of the last two launches (events of their slots)
    cudaEvent_t gz_done[2] = {nullptr, nullptr};   // the same for their compressors
    cudaEvent_t copy_done = nullptr;               // end of the previous launch's device-to-host copies
    bool trace = getenv("ISS_TRACE") != nullptr;
    int ring_alloc = ISS_HOSTMEM_NUMA;  // ISS_RING_ALLOC=pinned|numa|wc
    double gz_ratio[2] = {0.4, 0.4};            // compressed / plain bytes per stream: a first guess, then the last returned batch
    bool active = false, exhausted = false;
    unsigned long long *d_counters = nullptr;
    // benchmarking marks
    cudaStream_t tstream = nullptr;
    cudaEvent_t marks[ISS_N_MARKS] = {nullptr, nullptr, nullptr, nullptr};
    // NCCL
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    iss_stats stats{};
};

namespace {

int set_device(iss_ctx *c) {
    CU(cudaSetDevice(c->device));
    return ISS_OK;
}

void free_slots(iss_ctx *c) {
    c->gen_done[0] = c->gen_done[1] = c->gz_done[0] = c->gz_done[1] = c->copy_done = nullptr;  // they belong to the slots
    for (auto &s : c->slots) {
        if (s.d_out[0]) cudaFree(s.d_out[0]);
        if (s.d_out[1]) cudaFree(s.d_out[1]);
        iss_host_free(&s.h_buf[0]);
        iss_host_free(&s.h_buf[1]);
        s.h_out[0] = s.h_out[1] = nullptr;
        if (s.d_segs) cudaFree(s.d_segs);
        if (s.h_segs) cudaFreeHost(s.h_segs);
        for (int k = 0; k < 2; ++k) {
            if (s.d_gz_out[k]) cudaFree(s.d_gz_out[k]);
            if (s.d_gz_work[k]) cudaFree(s.d_gz_work[k]);
            if (s.d_gz_status[k]) cudaFree(s.d_gz_status[k]);
        }
        if (s.h_gz_total) cudaFreeHost(s.h_gz_total);
        if (s.d_scratch) cudaFree(s.d_scratch);
        if (s.d_events) cudaFree(s.d_events);
        if (s.d_evcount) cudaFree(s.d_evcount);
        if (s.h_evcount) cudaFreeHost(s.h_evcount);
        if (s.stream) cudaStreamDestroy(s.stream);
        if (s.zstream) cudaStreamDestroy(s.zstream);
        if (s.ev_k0) cudaEventDestroy(s.ev_k0);
        if (s.ev_k1) cudaEventDestroy(s.ev_k1);
        if (s.ev_done) cudaEventDestroy(s.ev_done);
        if (s.ev_gz) cudaEventDestroy(s.ev_gz);
        if (s.ev_end) cudaEventDestroy(s.ev_end);
    }
    c->slots.clear();
}

int rebuild_views(iss_ctx *c) {
    if (c->d_genome) {
        CU(cudaMemcpy(&c->gh, c->d_genome, sizeof(GenomeHeader), cudaMemcpyDeviceToHost));
        if (c->gh.magic != kGenomeMagic) return fail(ISS_E_STATE, "genome blob: bad magic");
        c->gview.recs = reinterpret_cast<const DevRecord *>(c->d_genome + c->gh.recs_off);
        c->gview.names = reinterpret_cast<const char *>(c->d_genome + c->gh.names_off);
        c->gview.packed = reinterpret_cast<const uint32_t *>(c->d_genome + c->gh.packed_off);
        c->gview.ascii = c->d_genome + c->gh.ascii_off;
        c->gview.n_records = c->gh.n_records;
        std::vector<DevRecord> dr(c->gh.n_records);
        if (c->gh.n_records)
            CU(cudaMemcpy(dr.data(), c->d_genome + c->gh.recs_off, sizeof(DevRecord) * dr.size(), cudaMemcpyDeviceToHost));
        c->recs.resize(dr.size());
        std::vector<char> names(c->gh.names_bytes + 1);
        if (c->gh.names_bytes)
            CU(cudaMemcpy(names.data(), c->d_genome + c->gh.names_off, c->gh.names_bytes, cudaMemcpyDeviceToHost));
        c->rec_names.resize(dr.size());
        for (size_t i = 0; i < dr.size(); ++i) {
            c->recs[i] = HostRecord{dr[i].len, dr[i].name_len};
            c->rec_names[i].assign(names.data() + dr[i].name_off, dr[i].name_len);
        }
        c->have_genome = true;
        c->stats.genome_bytes_device = c->genome_bytes;
    }
    if (c->d_model) {
        CU(cudaMemcpy(&c->mh, c->d_model, sizeof(ModelHeader), cudaMemcpyDeviceToHost));
        if (c->mh.magic != kModelMagic) return fail(ISS_E_STATE, "model blob: bad magic");
        const uint8_t *b = c->d_model;
        const ModelHeader &h = c->mh;
        c->mview.h = reinterpret_cast<const ModelHeader *>(b);
        c->mview.frag_thr = reinterpret_cast<const uint64_t *>(b + h.frag_thr_off);
        c->mview.insert_thr = reinterpret_cast<const uint32_t *>(b + h.insert_thr_off);
        c->mview.qual_exact = reinterpret_cast<const uint32_t *>(b + h.qual_exact_off);
        c->mview.q16[0] = b + h.q16_off[0];
        c->mview.q16[1] = b + h.q16_off[1];
        c->mview.q8[0] = h.q8_ok ? b + h.q8_off[0] : nullptr;
        c->mview.q8[1] = h.q8_ok ? b + h.q8_off[1] : nullptr;
        c->mview.subst_n = b + h.subst_n_off;
        c->mview.subst_thr = reinterpret_cast<const uint64_t *>(b + h.subst_thr_off);
        c->mview.subst_alt = b + h.subst_alt_off;
        c->mview.ins_base = b + h.ins_base_off;
        c->mview.del_thr = reinterpret_cast<const uint32_t *>(b + h.del_thr_off);
        c->mview.del_nz = b + h.del_nz_off;
        c->mview.ent = reinterpret_cast<const uint4 *>(b + h.ent_off);
        c->mview.basic_thr = reinterpret_cast<const uint32_t *>(b + h.basic_thr_off);
        c->mview.basic_val = b + h.basic_val_off;
        c->have_model = true;
        c->stats.model_bytes_device = c->model_bytes;
    }
    return ISS_OK;
}

uint32_t record_fixed_len(uint32_t L, uint32_t name_len, uint32_t cpu) {
    return 2u * L + 6u + name_len + 2u + iss_digits(cpu) + 2u;
}

// Plan and launch the next batch into slot `si`.  Returns ISS_OK; sets c->exhausted when no work is left.
int launch_batch(iss_ctx *c, uint32_t si) {
    NvtxRange nvtx("iss:plan+launch");
    Slot &s = c->slots[si];
    const uint32_t L = c->mh.L;
    uint64_t pairs = 0, bytes = 0;
    uint32_t nseg = 0;
    const uint64_t first_item = c->cur_item;
    uint64_t batch_key = 0;
    while (c->cur_item < c->items.size() && pairs < c->batch_pairs && nseg < c->max_segs) {
        const iss_work_item &it = c->items[c->cur_item];
        const uint64_t left = it.n_pairs - c->cur_off;
        if (left == 0) { c->cur_item++; c->cur_off = 0; continue; }
        if (nseg == 0) batch_key = it.key;
        else if (it.key != batch_key) break;  // one Philox key per launch: the key is a kernel parameter
        const uint64_t take = std::min(left, c->batch_pairs - pairs);
        const HostRecord &hr = c->recs[it.record];
        DevSegment &sg = s.h_segs[nseg];
        sg.pair_start = pairs;
        sg.n = take;
        sg.i0 = it.i0 + c->cur_off;
        sg.item_base = it.item_base;
        sg.key = it.key;
        sg.rec = it.record;
        sg.cpu = it.cpu_number;
        sg.cpu_digits = iss_digits(it.cpu_number);
        sg.fixed_len = record_fixed_len(L, hr.name_len, it.cpu_number);
        const uint64_t cf0 = iss_cf(sg.i0, sg.fixed_len);
        sg.out_base = bytes - cf0;  // modulo 2^64 on purpose
        bytes += iss_cf(sg.i0 + take, sg.fixed_len) - cf0;
        pairs += take;
        nseg++;
        c->cur_off += take;
        if (c->cur_off == it.n_pairs) { c->cur_item++; c->cur_off = 0; }
    }
    if (pairs == 0) { c->exhausted = true; return ISS_OK; }
    if (bytes > c->slot_cap) return fail(ISS_E_STATE, "internal: batch exceeds slot capacity");

    CU(cudaMemcpyAsync(s.d_segs, s.h_segs, sizeof(DevSegment) * nseg, cudaMemcpyHostToDevice, s.stream));
    GenParams p{};
    p.m = c->mview;
    p.g = c->gview;
    const ModelHeader &h = c->mh;
    p.L = h.L; p.qual_mode = h.qual_mode; p.n_insert = h.n_insert; p.frag_mode = h.frag_mode; p.n_frag = h.n_frag;
    p.const_phred = h.const_phred; p.n_basic = h.n_basic; p.fixed_insert = h.fixed_insert; p.frag_lo = h.frag_lo;
    p.n_grp = h.n_grp;
    memcpy(p.grp_begin, h.grp_begin, sizeof p.grp_begin);
    memcpy(p.bin_thr, h.bin_thr, sizeof p.bin_thr);
    memcpy(p.bin_valid, h.bin_valid, sizeof p.bin_valid);
    memcpy(p.err_thr, h.err_thr, sizeof p.err_thr);
    p.segs = s.d_segs;
    p.nseg = nseg;
    p.tile_pairs = c->tile_pairs;
    p.tile_shift = c->tile_pairs == 128 ? 7u : (c->tile_pairs == 64 ? 6u : 5u);
    p.split_shift = c->split_shift;
    p.total_pairs = pairs;
    p.key = batch_key;
    philox_round_keys(batch_key, p.rk);
    p.lay = c->lay;
    p.q_stage_cap = c->q_stage_cap;
    p.q_blocks = c->q_blocks;
    p.q_short = c->q_short;
    for (uint32_t o = 0; o < 2; ++o) {
        const uint32_t ps = c->q_short ? h.q8_ps[o] : h.q16_ps[o];
        p.q_pos_stride[o] = ps;
        p.q_stage_bytes[o] = 2u * ps;
        p.q_rows[o] = c->q_short ? c->mview.q8[o] : c->mview.q16[o];
        for (uint32_t b = 0; b < 4; ++b) p.q_row_off[o][b] = h.q16_cbin[o][b] * (c->q_short ? ISS_Q8_ROW_BYTES : ISS_Q16_ROW_BYTES);
    }
    p.out1 = s.d_out[0];
    p.out2 = s.d_out[1];
    p.img_cap = c->img_cap;
    p.err_cap = c->err_cap;
    p.tie_cap = kTieCap;
    p.err_ovf_cap = kErrOvfCap;
    p.err_ovf = reinterpret_cast<uint32_t *>(s.d_scratch + c->scratch_replay_bytes);
    p.ins_stride = h.ins_stride;
    memcpy(p.ins_coarse, h.ins_coarse, sizeof p.ins_coarse);
    p.surv_direct = (double)h.surv_per_pair * c->tile_pairs > ISS_SURV_CAP / 4 ? 1u : 0u;
    p.amplicon = c->opts.sequence_type == ISS_SEQ_AMPLICON;
    p.gc_bias = c->opts.gc_bias;
    p.nchunk = (L + 15u) / 16u;
    p.chunk_magic = (uint32_t)((0x100000000ull + p.nchunk - 1) / p.nchunk);
    p.scratch = s.d_scratch;
    p.scratch_mask_off = (uint32_t)align_up(5ull * L + 8, 16);
    p.scratch_stride = p.scratch_mask_off + (uint32_t)align_up(L, 16);
    p.gc_thr = 3865470567u;  // ceil(0.90 * 2^32): u < 0.90  <=>  w < gc_thr   (generator.py:88)
    p.counters = c->d_counters;
    p.store_mutations = c->mh.store_mutations;
    p.stage_bytes = 16u * ((L + 125u + 63u) / 64u);
    p.tma_gather = (15u + p.stage_bytes <= L) && c->opts.tma_gather;
    p.events = s.d_events;
    p.event_count = s.d_evcount;
    p.event_cap = s.event_cap;
    s.nseg = nseg;
    if (p.store_mutations) CU(cudaMemsetAsync(s.d_evcount, 0, sizeof(unsigned long long), s.stream));
    const uint64_t ntiles = (pairs + c->tile_pairs - 1) / c->tile_pairs;
    const uint32_t grid = (uint32_t)std::min<uint64_t>(ntiles, c->grid);
    // At most two generate kernels in flight: this one starts when the one launched two batches ago has ended.  Two still
    // hide each other's tails; but when the host hands back several slots at once (and when the ring is primed) their
    // kernels would otherwise all share the GPU, every one of them several times slower, and the first compressed batch
    // -- all the copy engine is waiting for -- would be late by that factor.
    cudaEvent_t &prev2 = c->gen_done[c->next_launch & 1u];
    if (prev2) CU(cudaStreamWaitEvent(s.stream, prev2, 0));
    // While the ring is primed in compressing mode the rule is stricter -- this kernel waits for the COMPRESSOR of the batch
    // two launches ago: the persistent generate CTAs hold all the shared memory of their SMs, so the first batch's
    // compressor (whose output the idle copy engine is waiting for) would get SMs only as each of the five kernels queued
    // behind it ends (first copy at 2.4 ms in the trace instead of 1.5).  In steady state the looser rule keeps more overlap.
    if (c->opts.compress && c->next_launch < c->n_slots && c->gz_done[c->next_launch & 1u])
        CU(cudaStreamWaitEvent(s.stream, c->gz_done[c->next_launch & 1u], 0));
    CU(cudaEventRecord(s.ev_k0, s.stream));
    CU(iss_launch_generate(p, grid, c->smem, s.stream));
    CU(cudaEventRecord(s.ev_k1, s.stream));
    prev2 = s.ev_k1;
    if (c->opts.compress) {
        s.gz_chunks = iss_gz_chunks(bytes);
        // Batches are compressed in launch order, two at a time (as the generate kernels above): the host takes batches back
        // in order, so a batch that the GPU finishes late -- it shared the machine with the five launched behind it --
        // holds up the copy engine and then frees all the slots at once.  And a slot builds the Huffman codes of its two
        // streams once per run, from its first batch: the batches of a run are statistically alike (same model, same
        // names), any complete code is valid for any input, and the two single-CTA table kernels (72 us each) would
        // otherwise sit in every batch's chain.
        // The compressor runs in a stream of the highest priority: its CTAs take the SMs that a generate kernel of a
        // LATER batch frees before that kernel's own next CTAs do -- what the copy engine is waiting for goes first.
        cudaEvent_t &prevz = c->gz_done[c->next_launch & 1u];
        CU(cudaStreamWaitEvent(s.zstream, s.ev_k1, 0));
        if (prevz) CU(cudaStreamWaitEvent(s.zstream, prevz, 0));
        uint32_t built = 0;
        for (int k = 0; k < 2; ++k) {
            CU(iss_gz_compress(s.d_out[k], bytes, s.d_gz_status[k], s.d_gz_out[k], s.d_gz_work[k], c->sm_count, s.gz_code_ready[k] ? 1 : 0, s.zstream));
            built += s.gz_code_ready[k] ? 0u : 2u;
            s.gz_code_ready[k] = bytes > 0;  // (an empty stream builds nothing)
        }
        // The event of the chain is recorded behind the kernels and IN FRONT of every copy of this batch: a device-to-host
        // copy, however small, queues in the one copy engine behind the previous batches' 27 MB copies -- with the 8-byte
        // totals between the two compressors (and in front of the event) the second compressor and the batch two launches
        // later each waited for a copy that had nothing to do with them (a bubble after every second batch in the trace).
        CU(cudaEventRecord(s.ev_gz, s.zstream));
        prevz = s.ev_gz;
        CU(cudaStreamWaitEvent(s.stream, s.ev_gz, 0));
        if (c->copy_done) CU(cudaStreamWaitEvent(s.stream, c->copy_done, 0));  // copies in launch order (see below)
        for (int k = 0; k < 2; ++k)
            CU(cudaMemcpyAsync(&s.h_gz_total[k], s.d_gz_status[k] + s.gz_chunks, sizeof(uint64_t), cudaMemcpyDeviceToHost, s.stream));
        // The host learns the compressed size only when the batch is done; queuing the D2H copy after that costs a host
        // round trip per batch, and a late small copy would also wait behind the other slots' copies in the copy
        // engine's queue.  Batches of one run compress alike (same model, same read length), so the copy is queued now
        // for 1 % more than the previous batch's ratio predicts; iss_ctx_next fetches a remainder only if that fell short.
        for (int k = 0; k < 2; ++k) {
            s.gz_pre[k] = 0;
            if (!c->opts.device_only && c->gz_ratio[k] > 0.0) {
                const uint64_t pre = std::min<uint64_t>(c->slot_cap, (uint64_t)(1.01 * c->gz_ratio[k] * (double)bytes) + 65536u);
                if (pre) CU(cudaMemcpyAsync(s.h_out[k], s.d_gz_out[k], pre, cudaMemcpyDeviceToHost, s.stream));
                s.gz_pre[k] = pre;
                c->stats.d2h_bytes += pre;
            }
        }
        c->stats.kernel_launches += 2 + built;  // member kernels, and histogram + code table where a code was built
    } else if (!c->opts.device_only) {
        // The host takes the batches back in launch order, but the copy engine serves the streams in its own order: a
        // later batch's copy that goes first makes the host wait for the earlier one and release the slots in a burst.
        if (c->copy_done) CU(cudaStreamWaitEvent(s.stream, c->copy_done, 0));
        CU(cudaMemcpyAsync(s.h_out[0], s.d_out[0], bytes, cudaMemcpyDeviceToHost, s.stream));
        CU(cudaMemcpyAsync(s.h_out[1], s.d_out[1], bytes, cudaMemcpyDeviceToHost, s.stream));
        c->stats.d2h_bytes += 2 * bytes;
    }
    if (c->mh.store_mutations)  // (behind the kernels, like every copy of the batch: see the compressor's event above)
        CU(cudaMemcpyAsync(s.h_evcount, s.d_evcount, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
    CU(cudaEventRecord(s.ev_done, s.stream));
    c->copy_done = s.ev_done;
    if (c->trace) CU(cudaEventRecord(s.ev_end, s.stream));
    s.state = 1;
    s.nbytes = bytes;
    s.n_pairs = pairs;
    s.first_item = first_item;
    s.batch_index = c->next_launch++;
    c->stats.h2d_bytes += sizeof(DevSegment) * nseg;
    c->stats.kernel_launches += 1;
    c->stats.pairs += pairs;
    c->stats.fastq_bytes += 2 * bytes;
    return ISS_OK;
}


// write_mutations (iss/generator.py:598-620): "{id}\t{position+1}\t.\t{ref}\t{alt}\t{quality}\t\t\n", in the order the
// reference appends them: forward indels, forward substitutions, reverse indels, reverse substitutions, pair by pair.
int format_vcf(iss_ctx *c, Slot &s) {
    NvtxRange nvtx("iss:vcf");
    const unsigned long long n = *s.h_evcount;
    s.vcf.clear();
    if (n == 0) return ISS_OK;
    if (n > s.event_cap) return fail(ISS_E_DEVICE, "mutation event buffer overflow");
    s.h_events.resize(n);
    CU(cudaMemcpy(s.h_events.data(), s.d_events, sizeof(uint4) * n, cudaMemcpyDeviceToHost));
    c->stats.d2h_bytes += sizeof(uint4) * n;
    std::sort(s.h_events.begin(), s.h_events.end(), [](const uint4 &a, const uint4 &b) {
        return (((uint64_t)a.y << 32) | a.x) < (((uint64_t)b.y << 32) | b.x);
    });
    s.vcf.reserve(n * 48);
    uint32_t seg = 0;
    char num[24];
    for (const uint4 &e : s.h_events) {
        const uint64_t key = ((uint64_t)e.y << 32) | e.x;
        const uint64_t gp = key >> 20;
        const uint32_t o = (key >> 19) & 1u, pos = (key >> 4) & 0x3FFFu;
        while (seg + 1 < s.nseg && s.h_segs[seg + 1].pair_start <= gp) ++seg;
        const DevSegment &sg = s.h_segs[seg];
        const uint64_t i = sg.i0 + (gp - sg.pair_start);
        s.vcf += c->rec_names[sg.rec];
        s.vcf += '_';
        s.vcf.append(num, snprintf(num, sizeof num, "%llu", (unsigned long long)i));
        s.vcf += '_';
        s.vcf.append(num, snprintf(num, sizeof num, "%u", sg.cpu));
        s.vcf += o ? "/2\t" : "/1\t";
        s.vcf.append(num, snprintf(num, sizeof num, "%u", pos + 1));
        s.vcf += "\t.\t";
        s.vcf += (char)(e.z & 0xFF);
        s.vcf += '\t';
        s.vcf += (char)((e.z >> 8) & 0xFF);
        if ((e.z >> 16) & 0xFF) s.vcf += (char)((e.z >> 16) & 0xFF);
        s.vcf += '\t';
        const uint32_t q = e.z >> 24;
        if (q == 255u) s.vcf += '.';
        else s.vcf.append(num, snprintf(num, sizeof num, "%u", q));
        s.vcf += "\t\t\n";
    }
    return ISS_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

int iss_abi_version(void) { return ISS_ABI_VERSION; }
const char *iss_last_error(void) { return g_err.c_str(); }
void iss_set_last_error(const char *msg) { g_err = msg ? msg : ""; }  // for the other translation units of the library

int iss_device_count(int *count) {
    CU(cudaGetDeviceCount(count));
    return ISS_OK;
}

int iss_device_pci_bus_id(int device, char *buf, int len) {
    if (!buf || len < 16) return fail(ISS_E_ARG, "buffer too small");
    CU(cudaDeviceGetPCIBusId(buf, len, device));
    return ISS_OK;
}

int iss_ctx_create(int device, iss_ctx **out) {
    if (!out) return fail(ISS_E_ARG, "out is null");
    int n = 0;
    CU(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(ISS_E_ARG, "no such CUDA device");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(ISS_E_CUDA, "this library is built for sm_100a (Blackwell) only");
    iss_ctx *c = new iss_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->stats.ring_numa_node = -1;
    if (const char *e = getenv("ISS_RING_ALLOC"))
        c->ring_alloc = !strcmp(e, "wc") ? ISS_HOSTMEM_WC : (!strcmp(e, "pinned") ? ISS_HOSTMEM_PINNED : ISS_HOSTMEM_NUMA);
    const int rc = [&]() -> int {
        CU(iss_kernel_configure(kMaxSmem));
        CU(iss_gz_init_device());
        CU(cudaMalloc(&c->d_counters, 4 * sizeof(unsigned long long)));
        CU(cudaMemset(c->d_counters, 0, 4 * sizeof(unsigned long long)));
        return ISS_OK;
    }();
    if (rc != ISS_OK) {
        iss_ctx_destroy(c);
        return rc;
    }
    *out = c;
    return ISS_OK;
}

void iss_ctx_destroy(iss_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    free_slots(c);
    if (c->d_genome) cudaFree(c->d_genome);
    if (c->d_model) cudaFree(c->d_model);
    if (c->d_counters) cudaFree(c->d_counters);
    if (c->comm) ncclCommDestroy(c->comm);
    for (auto &m : c->marks) if (m) cudaEventDestroy(m);
    if (c->tstream) cudaStreamDestroy(c->tstream);
    delete c;
}

// What a genome front end provides: the records' lengths and names, a way to put record r's ASCII at
// d_stage + st[r].stage_off (queued on stream 0), and the character at a position (for the error message).
struct GenomeSource {
    uint32_t n = 0;
    std::vector<uint64_t> len;
    std::vector<const char *> name;
    std::vector<uint32_t> name_len;
    std::function<int(uint8_t *, const std::vector<IssStagedRecord> &)> fill;
    std::function<unsigned(uint32_t, uint64_t)> char_at;
    uint64_t h2d_bytes = 0;
};

// Staging buffer in HBM (16-byte aligned record starts) -> classification -> blob layout -> 2-bit / byte packing, all on
// the device; the host builds the header, the record table and the names.
static int set_genome_common(iss_ctx *c, const GenomeSource &src) {
    const uint32_t n = src.n;
    std::vector<IssStagedRecord> st(n);
    std::vector<uint64_t> unit_begin(n + 1, 0);
    uint64_t stage_bytes = 0, names_bytes = 0;
    for (uint32_t r = 0; r < n; ++r) {
        st[r].stage_off = stage_bytes;
        st[r].len = src.len[r];
        st[r].data_off = 0; st[r].enc = 0; st[r].pad = 0;
        stage_bytes = align_up(stage_bytes + src.len[r], 16);
        unit_begin[r + 1] = unit_begin[r] + (src.len[r] + 15) / 16;
        names_bytes += src.name_len[r];
    }
    const uint64_t n_units = unit_begin[n];
    uint8_t *d_stage = nullptr;
    IssStagedRecord *d_recs = nullptr;
    uint64_t *d_units = nullptr;
    uint32_t *d_flags = nullptr;
    unsigned long long *d_bad = nullptr;
    uint8_t *d_blob = nullptr;
    auto cleanup = [&] {
        if (d_stage) cudaFree(d_stage);
        if (d_recs) cudaFree(d_recs);
        if (d_units) cudaFree(d_units);
        if (d_flags) cudaFree(d_flags);
        if (d_bad) cudaFree(d_bad);
    };
    uint64_t total = 0;
    const int rc = [&]() -> int {
        CU(cudaMalloc(&d_stage, stage_bytes + 32));
        CU(cudaMemsetAsync(d_stage, 0, stage_bytes + 32, 0));  // the pads between records and behind the last one read as 0
        if (int frc = src.fill(d_stage, st)) return frc;
        CU(cudaMalloc(&d_recs, sizeof(IssStagedRecord) * std::max<uint32_t>(n, 1)));
        CU(cudaMalloc(&d_units, sizeof(uint64_t) * (n + 1)));
        CU(cudaMalloc(&d_flags, sizeof(uint32_t) * std::max<uint32_t>(n, 1)));
        CU(cudaMalloc(&d_bad, sizeof(unsigned long long) * std::max<uint32_t>(n, 1)));
        CU(cudaMemsetAsync(d_flags, 0, sizeof(uint32_t) * std::max<uint32_t>(n, 1), 0));
        CU(cudaMemsetAsync(d_bad, 0xFF, sizeof(unsigned long long) * std::max<uint32_t>(n, 1), 0));
        if (n) CU(cudaMemcpyAsync(d_recs, st.data(), sizeof(IssStagedRecord) * n, cudaMemcpyHostToDevice, 0));
        CU(cudaMemcpyAsync(d_units, unit_begin.data(), sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, 0));
        // classify on the device
        CU(iss_genome_classify(d_stage, d_recs, d_units, n, n_units, d_flags, d_bad, c->sm_count, 0));
        std::vector<uint32_t> flags(n);
        std::vector<unsigned long long> bad(n);
        if (n) {
            CU(cudaMemcpy(flags.data(), d_flags, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
            CU(cudaMemcpy(bad.data(), d_bad, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost));
        }
        for (uint32_t r = 0; r < n; ++r)
            if (bad[r] != ~0ull) {
                char buf[160];
                snprintf(buf, sizeof buf, "record %u: character 0x%02x at position %llu is outside the reference's alphabet", r,
                         src.char_at(r, bad[r]), bad[r]);
                return fail(ISS_E_CHAR, buf);
            }
        // layout of the blob
        uint64_t packed_bases = 64, ascii_bytes = 64, name_off = 0;
        std::vector<DevRecord> dr(n);
        for (uint32_t r = 0; r < n; ++r) {
            const bool pure = !(flags[r] & 1u);
            dr[r].len = src.len[r];
            dr[r].enc = pure ? ISS_ENC_2BIT : ISS_ENC_ASCII;
            dr[r].name_off = (uint32_t)name_off;
            dr[r].name_len = src.name_len[r];
            dr[r].pad = 0;
            name_off += src.name_len[r];
            if (pure) { dr[r].data_off = packed_bases; packed_bases = align_up(packed_bases + src.len[r], 64); }
            else { dr[r].data_off = ascii_bytes; ascii_bytes = align_up(ascii_bytes + src.len[r], 16); }
            st[r].data_off = dr[r].data_off;
            st[r].enc = dr[r].enc;
        }
        packed_bases += 128;  // tail pad: windows may be read up to 31 bases past a record
        ascii_bytes += 64;
        GenomeHeader gh{};
        gh.magic = kGenomeMagic;
        gh.n_records = n;
        uint64_t off = align_up(sizeof(GenomeHeader), 16);
        gh.recs_off = off; off = align_up(off + sizeof(DevRecord) * (uint64_t)n, 16);
        gh.names_off = off; gh.names_bytes = names_bytes; off = align_up(off + names_bytes, 64);
        gh.packed_off = off; gh.packed_bytes = packed_bases / 4; off = align_up(off + gh.packed_bytes, 64);
        gh.ascii_off = off; gh.ascii_bytes = ascii_bytes; off = align_up(off + ascii_bytes, 64);
        gh.total_bytes = off;
        total = off;
        std::vector<uint8_t> head(gh.packed_off, 0);  // header + record table + names: the only part built on the host
        memcpy(head.data(), &gh, sizeof gh);
        if (n) memcpy(head.data() + gh.recs_off, dr.data(), sizeof(DevRecord) * (size_t)n);
        for (uint32_t r = 0; r < n; ++r) memcpy(head.data() + gh.names_off + dr[r].name_off, src.name[r], src.name_len[r]);
        CU(cudaMalloc(&d_blob, off));
        CU(cudaMemsetAsync(d_blob, 0, off, 0));
        CU(cudaMemcpyAsync(d_blob, head.data(), head.size(), cudaMemcpyHostToDevice, 0));
        if (n) CU(cudaMemcpyAsync(d_recs, st.data(), sizeof(IssStagedRecord) * n, cudaMemcpyHostToDevice, 0));
        // pack on the device
        CU(iss_genome_pack(d_stage, d_recs, d_units, n, n_units, reinterpret_cast<uint32_t *>(d_blob + gh.packed_off), d_blob + gh.ascii_off,
                           c->sm_count, 0));
        CU(cudaStreamSynchronize(0));
        return ISS_OK;
    }();
    cleanup();
    if (rc != ISS_OK) {
        if (d_blob) cudaFree(d_blob);
        return rc;
    }
    if (c->d_genome) cudaFree(c->d_genome);
    c->d_genome = d_blob;
    c->genome_bytes = total;
    c->stats.h2d_bytes += src.h2d_bytes;
    return rebuild_views(c);
}

int iss_ctx_set_genome(iss_ctx *c, const iss_record *records, uint32_t n) {
    if (!c || (!records && n)) return fail(ISS_E_ARG, "null argument");
    if (int rc = set_device(c)) return rc;
    if (c->active) return fail(ISS_E_STATE, "generation in progress");
    NvtxRange nvtx("iss:set_genome");
    GenomeSource src;
    src.n = n;
    src.len.resize(n); src.name.resize(n); src.name_len.resize(n);
    for (uint32_t r = 0; r < n; ++r) {
        if (!records[r].seq && records[r].len) return fail(ISS_E_ARG, "record with a null sequence");
        src.len[r] = records[r].len;
        src.name[r] = records[r].name;
        src.name_len[r] = records[r].name_len;
        src.h2d_bytes += records[r].len;
    }
    src.fill = [&](uint8_t *d_stage, const std::vector<IssStagedRecord> &st) -> int {
        for (uint32_t r = 0; r < n; ++r)
            if (records[r].len) CU(cudaMemcpyAsync(d_stage + st[r].stage_off, records[r].seq, records[r].len, cudaMemcpyHostToDevice, 0));
        return ISS_OK;
    };
    src.char_at = [&](uint32_t r, uint64_t pos) { return (unsigned)(uint8_t)records[r].seq[pos]; };
    return set_genome_common(c, src);
}

// FASTA text -> genome, white space and title lines taken out on the device (iss_fasta.cu).  The host only finds the title
// lines (iss_fasta_index); what the reference does per record in Python -- SeqIO.parse, str(record.seq) -- does not happen.
int iss_ctx_set_genome_fasta(iss_ctx *c, const char *text, uint64_t n_text, iss_fasta_rec *recs, uint32_t n) {
    if (!c || (!text && n_text) || (!recs && n)) return fail(ISS_E_ARG, "null argument");
    if (int rc = set_device(c)) return rc;
    if (c->active) return fail(ISS_E_STATE, "generation in progress");
    NvtxRange nvtx("iss:set_genome_fasta");
    for (uint32_t r = 0; r < n; ++r) {
        const iss_fasta_rec &f = recs[r];
        const bool ok = f.gt_pos < f.body_off && f.body_off <= n_text && f.body_len <= n_text - f.body_off && f.id_off >= f.gt_pos &&
                        f.id_off <= f.body_off && f.id_len <= f.body_off - f.id_off && (r == 0 || f.gt_pos >= recs[r - 1].body_off + recs[r - 1].body_len);
        if (!ok) return fail(ISS_E_ARG, "FASTA index does not describe the text (records must be in file order)");
    }
    const uint64_t padded = align_up(n_text + 1, 4096);  // whole 4 KB blocks, '\n' behind the text
    const uint64_t nblk = padded / 4096;
    uint8_t *d_text = nullptr;
    uint32_t *d_cnt = nullptr;
    uint64_t *d_pre = nullptr, *d_pos = nullptr, *d_rank = nullptr, *d_tab = nullptr;
    auto cleanup = [&] {
        if (d_text) cudaFree(d_text);
        if (d_cnt) cudaFree(d_cnt);
        if (d_pre) cudaFree(d_pre);
        if (d_pos) cudaFree(d_pos);
        if (d_rank) cudaFree(d_rank);
        if (d_tab) cudaFree(d_tab);
    };
    GenomeSource src;
    src.n = n;
    src.len.resize(n); src.name.resize(n); src.name_len.resize(n);
    src.h2d_bytes = n_text;
    std::vector<uint64_t> rank0(n);
    int rc = [&]() -> int {
        CU(cudaMalloc(&d_text, padded));
        CU(cudaMemsetAsync(d_text + (n_text & ~(uint64_t)15), '\n', padded - (n_text & ~(uint64_t)15), 0));
        if (n_text) CU(cudaMemcpyAsync(d_text, text, n_text, cudaMemcpyHostToDevice, 0));
        CU(cudaMalloc(&d_cnt, sizeof(uint32_t) * nblk));
        CU(cudaMalloc(&d_pre, sizeof(uint64_t) * (nblk + 1)));
        CU(cudaMalloc(&d_pos, sizeof(uint64_t) * 3 * std::max<uint32_t>(n, 1)));
        CU(cudaMalloc(&d_rank, sizeof(uint64_t) * 2 * std::max<uint32_t>(n, 1)));
        // positions: [0, n) '>' of each title, [n, 2n) body starts, [2n, 3n) body ends
        std::vector<uint64_t> pos(3 * (size_t)std::max<uint32_t>(n, 1), 0);
        for (uint32_t r = 0; r < n; ++r) { pos[r] = recs[r].gt_pos; pos[n + r] = recs[r].body_off; pos[2 * (size_t)n + r] = recs[r].body_off + recs[r].body_len; }
        CU(cudaMemcpyAsync(d_pos, pos.data(), sizeof(uint64_t) * pos.size(), cudaMemcpyHostToDevice, 0));
        CU(iss_fasta_blank(d_text, d_pos, d_pos + n, n, 0));                     // title lines (and what stands before the first) -> '\n'
        CU(iss_fasta_count(d_text, nblk, d_cnt, 0));                            // bases per 4 KB block
        CU(iss_fasta_scan(d_cnt, nblk, d_pre, 0));                              // -> bases before each block
        CU(iss_fasta_ranks(d_text, d_pre, d_pos + n, 2 * n, d_rank, 0));        // bases before each body start / end
        std::vector<uint64_t> rank(2 * (size_t)std::max<uint32_t>(n, 1), 0);
        if (n) CU(cudaMemcpy(rank.data(), d_rank, sizeof(uint64_t) * 2 * n, cudaMemcpyDeviceToHost));
        for (uint32_t r = 0; r < n; ++r) {
            rank0[r] = rank[r];
            src.len[r] = recs[r].seq_len = rank[(size_t)n + r] - rank[r];
            src.name[r] = text + recs[r].id_off;
            src.name_len[r] = recs[r].id_len;
        }
        return ISS_OK;
    }();
    if (rc != ISS_OK) { cleanup(); return rc; }
    src.fill = [&](uint8_t *d_stage, const std::vector<IssStagedRecord> &st) -> int {
        // per record: bases before its body, and where its bases go
        std::vector<uint64_t> tab(2 * (size_t)std::max<uint32_t>(n, 1), 0);
        for (uint32_t r = 0; r < n; ++r) { tab[r] = rank0[r]; tab[(size_t)n + r] = st[r].stage_off; }
        CU(cudaMalloc(&d_tab, sizeof(uint64_t) * tab.size()));
        CU(cudaMemcpyAsync(d_tab, tab.data(), sizeof(uint64_t) * tab.size(), cudaMemcpyHostToDevice, 0));
        CU(iss_fasta_scatter(d_text, nblk, d_pre, d_pos + n, d_pos + 2 * (size_t)n, d_tab, d_tab + n, n, d_stage, 0));
        CU(cudaStreamSynchronize(0));  // `tab` leaves scope
        return ISS_OK;
    };
    src.char_at = [&](uint32_t r, uint64_t p) -> unsigned {  // walk the body on the host: only on the error path
        uint64_t seen = 0;
        for (uint64_t x = recs[r].body_off; x < recs[r].body_off + recs[r].body_len; ++x) {
            const uint8_t b = (uint8_t)text[x];
            if (b == 0x20 || (uint8_t)(b - 9u) < 5u) continue;
            if (seen++ == p) return b;
        }
        return 0u;
    };
    rc = set_genome_common(c, src);
    cleanup();
    return rc;
}

int iss_ctx_set_model(iss_ctx *c, const iss_model *m) {
    if (!c || !m) return fail(ISS_E_ARG, "null argument");
    if (int rc = set_device(c)) return rc;
    if (c->active) return fail(ISS_E_STATE, "generation in progress");
    const uint32_t L = m->read_length;
    if (L < 2 || L > 8192) return fail(ISS_E_ARG, "read_length out of range [2, 8192]");
    if (m->qual_mode > ISS_QUAL_TABLE) return fail(ISS_E_ARG, "bad qual_mode");
    ModelHeader h{};
    h.magic = kModelMagic;
    h.L = L; h.qual_mode = m->qual_mode; h.n_insert = m->n_insert; h.frag_mode = m->frag_mode; h.n_frag = m->n_frag;
    h.const_phred = m->const_phred; h.n_basic = m->n_basic; h.store_mutations = m->store_mutations;
    h.fixed_insert = m->fixed_insert; h.frag_lo = m->frag_lo;
    if (m->bin_thr) memcpy(h.bin_thr, m->bin_thr, sizeof h.bin_thr);
    if (m->bin_valid) memcpy(h.bin_valid, m->bin_valid, sizeof h.bin_valid);
    if (!m->err_thr) return fail(ISS_E_ARG, "err_thr is required");
    memcpy(h.err_thr, m->err_thr, sizeof h.err_thr);
    if (!m->subst_n || !m->subst_thr || !m->subst_alt || !m->ins_thr || !m->ins_nz || !m->ins_base || !m->del_thr || !m->del_nz)
        return fail(ISS_E_ARG, "substitution / indel tables are required");
    if (m->qual_mode == ISS_QUAL_KDE && (!m->qual_thr || !m->bin_thr || !m->bin_valid)) return fail(ISS_E_ARG, "kde tables missing");
    if (m->n_insert && !m->insert_thr) return fail(ISS_E_ARG, "insert_thr missing");
    // RNG contract v2: the indel draw sites with a non-zero threshold, ranked in loop order per (orientation, kind);
    // rank r reads word r & 3 of Philox block r >> 2; only positions < L-1 count (the indel loop runs over range(L-1)).  Lists: insertion fwd, deletion fwd, insertion rev, deletion rev,
    // each padded to whole groups of 4 sites with sites that never fire (valid = 0).  A group carries what the scan
    // needs to reject it -- word > threshold for all 4 sites -- and the rare survivor is evaluated from the full tables.
    struct Grp { uint32_t thr[4]; uint32_t meta[4]; };
    std::vector<Grp> grps;
    for (uint32_t o = 0; o < 2; ++o)
        for (uint32_t kind = 0; kind < 2; ++kind) {
            h.grp_begin[2 * o + kind] = (uint32_t)grps.size();
            uint32_t rank = 0;
            auto add = [&](uint32_t thr, uint32_t meta) {
                if ((rank & 3u) == 0) grps.push_back(Grp{});
                grps.back().thr[rank & 3u] = thr;
                grps.back().meta[rank & 3u] = meta;
                ++rank;
            };
            for (uint32_t pos = 0; pos + 1 < L; ++pos) {
                const uint32_t tb = o * L + pos;
                if (kind == 0) {
                    for (uint32_t lane = 0; lane < 4; ++lane)
                        if ((m->ins_nz[tb] >> lane) & 1u) add(m->ins_thr[tb * 4 + lane], pos | (1u << 16) | (lane << 17));
                } else if (m->del_nz[tb] & 15u) {
                    uint32_t tmax = 0;
                    for (uint32_t b = 0; b < 4; ++b)
                        if ((m->del_nz[tb] >> b) & 1u) tmax = std::max(tmax, m->del_thr[tb * 4 + b]);
                    add(tmax, pos | (1u << 16));
                }
            }
        }
    h.ins_stride = std::max<uint32_t>(1u, (m->n_insert + 31u) / 32u);
    for (uint32_t k = 0; k < 31; ++k) {
        const uint64_t last = std::min<uint64_t>((uint64_t)(k + 1) * h.ins_stride, m->n_insert);
        h.ins_coarse[k] = (uint64_t)k * h.ins_stride < m->n_insert ? m->insert_thr[last - 1] : 0xFFFFFFFFu;
    }
    h.grp_begin[4] = h.n_grp = (uint32_t)grps.size();
    {
        double e = 0.0;  // a site passes when word <= thr: probability (thr + 1) / 2^32
        for (const Grp &g : grps)
            for (int k = 0; k < 4; ++k)
                if ((g.meta[k] >> 16) & 1u) e += ((double)g.thr[k] + 1.0) / 4294967296.0;
        h.surv_per_pair = (float)e;
    }
    uint64_t off = align_up(sizeof(ModelHeader), 16);
    auto place = [&](uint64_t bytes) { uint64_t at = off; off = align_up(off + std::max<uint64_t>(bytes, 16), 16); return at; };
    // 16-bit quality rows (layout: iss_kernels.cu).  Only populated bins get rows.
    const bool kde = m->qual_mode == ISS_QUAL_KDE;
    if (kde) {
        for (uint32_t o = 0; o < 2; ++o)
            for (uint32_t b = 0; b < 4; ++b) {
                if (!m->bin_valid[o * 4 + b]) continue;
                for (uint64_t pos = 0; pos < L; ++pos) {
                    const uint32_t *src = m->qual_thr + ((uint64_t)(o * 4 + b) * L + pos) * ISS_QROW;
                    for (int j = 1; j < ISS_QROW; ++j)
                        if (src[j] < src[j - 1]) return fail(ISS_E_ARG, "qual_thr rows must be sorted");
                    if (src[ISS_QROW - 1] != 0xFFFFFFFFu) return fail(ISS_E_ARG, "qual_thr rows need at least one 0xFFFFFFFF pad entry");
                }
            }
        for (uint32_t o = 0; o < 2; ++o) {
            uint32_t nb = 0;
            for (uint32_t b = 0; b < 4; ++b) {
                h.q16_cbin[o][b] = (uint8_t)nb;
                if (m->bin_valid[o * 4 + b]) ++nb;
            }
            h.q16_nb[o] = std::max(nb, 1u);
            h.q16_ps[o] = (uint32_t)align_up((uint64_t)h.q16_nb[o] * ISS_Q16_ROW_BYTES, 16);
            h.q8_ps[o] = (uint32_t)align_up((uint64_t)h.q16_nb[o] * ISS_Q8_ROW_BYTES, 16);
        }
        // short rows: every row of the model has at most 8 distinct non-zero 16-bit values (the pad 0xFFFF is one of them)
        h.q8_ok = 1;
        for (uint32_t ob = 0; ob < 8 && h.q8_ok; ++ob) {
            if (!m->bin_valid[ob]) continue;
            for (uint64_t pos = 0; pos < L && h.q8_ok; ++pos) {
                const uint32_t *src = m->qual_thr + ((uint64_t)ob * L + pos) * ISS_QROW;
                uint32_t distinct = (src[0] >> 16) ? 1u : 0u;  // the value 0 is not stored (see q8_search2)
                for (int j = 1; j < ISS_QROW; ++j) distinct += (src[j] >> 16) != (src[j - 1] >> 16) ? 1u : 0u;
                if (distinct > 8) h.q8_ok = 0;
            }
        }
        if (getenv("ISS_QROWS") && atoi(getenv("ISS_QROWS")) == 64) h.q8_ok = 0;  // A/B switch for benchmarks and tests: full rows only
    }
    h.frag_thr_off = place(8ull * m->n_frag);
    h.insert_thr_off = place(4ull * m->n_insert);
    h.qual_exact_off = place(kde ? 2ull * 4 * L * ISS_QROW * 4 : 0);
    for (uint32_t o = 0; o < 2; ++o) h.q16_off[o] = place(kde ? (uint64_t)(L + 1) * h.q16_ps[o] : 0);
    for (uint32_t o = 0; o < 2; ++o) h.q8_off[o] = place(kde && h.q8_ok ? (uint64_t)(L + 1) * h.q8_ps[o] : 0);
    h.subst_n_off = place(2ull * L * 4);
    h.subst_thr_off = place(2ull * L * 4 * ISS_MAX_ALT * 8);
    h.subst_alt_off = place(2ull * L * 4 * ISS_MAX_ALT);
    h.ins_base_off = place(2ull * L * 4);
    h.del_thr_off = place(2ull * L * 4 * 4);
    h.del_nz_off = place(2ull * L);
    h.basic_thr_off = place(4ull * m->n_basic);
    h.basic_val_off = place(1ull * m->n_basic + 1);
    h.ent_off = place(sizeof(Grp) * grps.size());
    h.total_bytes = off;
    std::vector<uint8_t> blob(off, 0);
    memcpy(blob.data(), &h, sizeof h);
    auto put = [&](uint64_t at, const void *src, uint64_t bytes) { if (src && bytes) memcpy(blob.data() + at, src, bytes); };
    put(h.frag_thr_off, m->frag_thr, 8ull * m->n_frag);
    put(h.insert_thr_off, m->insert_thr, 4ull * m->n_insert);
    if (kde) {
        put(h.qual_exact_off, m->qual_thr, 2ull * 4 * L * ISS_QROW * 4);
        for (uint32_t o = 0; o < 2; ++o) {
            uint8_t *base = blob.data() + h.q16_off[o];
            memset(base, 0xFF, (size_t)(L + 1) * h.q16_ps[o]);  // pad entries / the pad position behind an odd L never count
            for (uint32_t b = 0; b < 4; ++b) {
                if (!m->bin_valid[o * 4 + b]) continue;
                for (uint64_t pos = 0; pos < L; ++pos) {
                    const uint32_t *src = m->qual_thr + ((uint64_t)(o * 4 + b) * L + pos) * ISS_QROW;
                    uint8_t *row = base + pos * h.q16_ps[o] + (size_t)h.q16_cbin[o][b] * ISS_Q16_ROW_BYTES;  // 4-byte aligned
                    for (uint32_t j = 0; j < ISS_QROW; ++j) {
                        const uint16_t e = (uint16_t)(src[j] >> 16);
                        memcpy(row + 2 * j, &e, 2);
                    }
                }
            }
        }
    }
    if (kde && h.q8_ok) {
        for (uint32_t o = 0; o < 2; ++o) {
            uint8_t *base = blob.data() + h.q8_off[o];
            // pad positions / rows: values 0xFFFF (nothing is below a draw's tie with the pad), counts 0
            for (uint64_t pos = 0; pos <= L; ++pos)
                for (uint32_t r = 0; r < h.q16_nb[o]; ++r) {
                    uint8_t *row = base + pos * h.q8_ps[o] + (size_t)r * ISS_Q8_ROW_BYTES;
                    memset(row, 0xFF, 16);
                    memset(row + 16, 0, 8);
                }
            for (uint32_t b = 0; b < 4; ++b) {
                if (!m->bin_valid[o * 4 + b]) continue;
                for (uint64_t pos = 0; pos < L; ++pos) {
                    const uint32_t *src = m->qual_thr + ((uint64_t)(o * 4 + b) * L + pos) * ISS_QROW;
                    uint8_t *row = base + pos * h.q8_ps[o] + (size_t)h.q16_cbin[o][b] * ISS_Q8_ROW_BYTES;
                    uint32_t k = 0;
                    for (uint32_t j = 0; j < ISS_QROW; ++j) {
                        const uint16_t e = (uint16_t)(src[j] >> 16);
                        if (e == 0 || (j && e == (uint16_t)(src[j - 1] >> 16))) continue;
                        memcpy(row + 2 * k, &e, 2);   // the k-th distinct non-zero value ...
                        row[16 + k] = (uint8_t)j;     // ... and the entries of the full row below it
                        ++k;
                    }
                }
            }
        }
    }
    put(h.subst_n_off, m->subst_n, 2ull * L * 4);
    put(h.subst_thr_off, m->subst_thr, 2ull * L * 4 * ISS_MAX_ALT * 8);
    put(h.subst_alt_off, m->subst_alt, 2ull * L * 4 * ISS_MAX_ALT);
    put(h.ins_base_off, m->ins_base, 2ull * L * 4);
    put(h.del_thr_off, m->del_thr, 2ull * L * 4 * 4);
    put(h.del_nz_off, m->del_nz, 2ull * L);
    put(h.ent_off, grps.data(), sizeof(Grp) * grps.size());
    put(h.basic_thr_off, m->basic_thr, 4ull * m->n_basic);
    put(h.basic_val_off, m->basic_val, 1ull * m->n_basic + (m->basic_val ? 1 : 0));
    if (c->d_model) { cudaFree(c->d_model); c->d_model = nullptr; }
    CU(cudaMalloc(&c->d_model, off));
    CU(cudaMemcpy(c->d_model, blob.data(), off, cudaMemcpyHostToDevice));
    c->model_bytes = off;
    c->stats.h2d_bytes += off;
    return rebuild_views(c);
}

// ---- NCCL ------------------------------------------------------------------------------------------
int iss_comm_unique_id(uint8_t id[ISS_UNIQUE_ID_BYTES]) {
    static_assert(sizeof(ncclUniqueId) <= ISS_UNIQUE_ID_BYTES, "unique id size");
    ncclUniqueId u;
    NC(ncclGetUniqueId(&u));
    memset(id, 0, ISS_UNIQUE_ID_BYTES);
    memcpy(id, &u, sizeof u);
    return ISS_OK;
}

int iss_comm_init_rank(iss_ctx *c, int nranks, int rank, const uint8_t id[ISS_UNIQUE_ID_BYTES]) {
    if (!c || !id || nranks < 1 || rank < 0 || rank >= nranks) return fail(ISS_E_ARG, "bad communicator arguments");
    if (int rc = set_device(c)) return rc;
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    NC(ncclCommInitRank(&c->comm, nranks, u, rank));
    c->nranks = nranks;
    c->rank = rank;
    return ISS_OK;
}

int iss_comm_init_all(iss_ctx **ctxs, int n) {
    if (!ctxs || n < 1) return fail(ISS_E_ARG, "bad arguments");
    std::vector<int> devs(n);
    std::vector<ncclComm_t> comms(n);
    for (int i = 0; i < n; ++i) devs[i] = ctxs[i]->device;
    NC(ncclCommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; ++i) { ctxs[i]->comm = comms[i]; ctxs[i]->nranks = n; ctxs[i]->rank = i; }
    return ISS_OK;
}

int iss_ctx_bcast_all(iss_ctx **ctxs, int n, int root) {
    if (!ctxs || n < 1 || root < 0) return fail(ISS_E_ARG, "bad arguments");
    // every check that can fail happens BEFORE the first collective: a rank that returned early would leave its peers
    // blocked inside ncclBroadcast
    for (int i = 0; i < n; ++i) {
        if (!ctxs[i] || !ctxs[i]->comm) return fail(ISS_E_STATE, "communicator not initialised");
        if (ctxs[i]->active) return fail(ISS_E_STATE, "generation in progress");
        if (root >= ctxs[i]->nranks) return fail(ISS_E_ARG, "root is not a rank of the communicator");
        if (ctxs[i]->rank == root) {
            if (!ctxs[i]->have_genome || !ctxs[i]->have_model) return fail(ISS_E_STATE, "root has no genome/model");
        }
    }
    std::vector<uint64_t *> d_sizes(n, nullptr);
    auto free_sizes = [&] {
        for (int i = 0; i < n; ++i)
            if (d_sizes[i]) { cudaSetDevice(ctxs[i]->device); cudaFree(d_sizes[i]); d_sizes[i] = nullptr; }
    };
    const int rc = [&]() -> int {
        // 1. sizes
        for (int i = 0; i < n; ++i) {
            iss_ctx *c = ctxs[i];
            CU(cudaSetDevice(c->device));
            CU(cudaMalloc(&d_sizes[i], 16));
            if (c->rank == root) {
                uint64_t sz[2] = {c->genome_bytes, c->model_bytes};
                CU(cudaMemcpy(d_sizes[i], sz, 16, cudaMemcpyHostToDevice));
            }
        }
        NC(ncclGroupStart());
        for (int i = 0; i < n; ++i) {
            CU(cudaSetDevice(ctxs[i]->device));
            NC(ncclBroadcast(d_sizes[i], d_sizes[i], 16, ncclUint8, root, ctxs[i]->comm, 0));
        }
        NC(ncclGroupEnd());
        for (int i = 0; i < n; ++i) {
            iss_ctx *c = ctxs[i];
            CU(cudaSetDevice(c->device));
            CU(cudaStreamSynchronize(0));
            uint64_t sz[2];
            CU(cudaMemcpy(sz, d_sizes[i], 16, cudaMemcpyDeviceToHost));
            if (c->rank != root) {
                if (c->d_genome) { cudaFree(c->d_genome); c->d_genome = nullptr; }
                if (c->d_model) { cudaFree(c->d_model); c->d_model = nullptr; }
                c->have_genome = c->have_model = false;
                CU(cudaMalloc(&c->d_genome, sz[0]));
                CU(cudaMalloc(&c->d_model, sz[1]));
                c->genome_bytes = sz[0];
                c->model_bytes = sz[1];
            }
        }
        // 2. blobs
        NC(ncclGroupStart());
        for (int i = 0; i < n; ++i) {
            iss_ctx *c = ctxs[i];
            CU(cudaSetDevice(c->device));
            NC(ncclBroadcast(c->d_genome, c->d_genome, c->genome_bytes, ncclUint8, root, c->comm, 0));
            NC(ncclBroadcast(c->d_model, c->d_model, c->model_bytes, ncclUint8, root, c->comm, 0));
        }
        NC(ncclGroupEnd());
        for (int i = 0; i < n; ++i) {
            iss_ctx *c = ctxs[i];
            CU(cudaSetDevice(c->device));
            CU(cudaStreamSynchronize(0));
            if (int r = rebuild_views(c)) return r;
        }
        return ISS_OK;
    }();
    free_sizes();
    return rc;
}

int iss_ctx_bcast(iss_ctx *c, int root) { return iss_ctx_bcast_all(&c, 1, root); }

// ---- generation ------------------------------------------------------------------------------------
int iss_ctx_submit(iss_ctx *c, const iss_work_item *items, uint64_t n_items, const iss_gen_opts *opts) {
    if (!c || (!items && n_items)) return fail(ISS_E_ARG, "null argument");
    if (int rc = set_device(c)) return rc;
    if (!c->have_genome || !c->have_model) return fail(ISS_E_STATE, "set_genome and set_model (or bcast) first");
    if (c->active) return fail(ISS_E_STATE, "previous submit not drained");
    const uint32_t L = c->mh.L;
    iss_gen_opts o{};
    if (opts) o = *opts;
    if (o.batch_pairs == 0) o.batch_pairs = 1u << 18;
    if (o.n_slots == 0) o.n_slots = o.compress ? 6 : 3;  // the compressed path has the longer per-batch chain (compress, then copy)
    if (o.n_slots < 2 && !o.device_only) o.n_slots = 2;
    if (o.n_slots > 16) o.n_slots = 16;
    if (const char *e = getenv("ISS_TMA_GATHER")) o.tma_gather = (e[0] == '1');  // A/B switch for benchmarks
    uint64_t max_rec = 0, total = 0;
    for (uint64_t k = 0; k < n_items; ++k) {
        const iss_work_item &it = items[k];
        if (it.record >= c->recs.size()) return fail(ISS_E_ARG, "work item references an unknown record");
        const HostRecord &hr = c->recs[it.record];
        if (!(L < hr.len)) return fail(ISS_E_SHORT, "record not longer than the read length (reference: AssertionError)");
        if (it.n_pairs == 0) continue;
        const uint64_t rec = record_fixed_len(L, hr.name_len, it.cpu_number) + iss_digits(it.i0 + it.n_pairs - 1);
        max_rec = std::max(max_rec, rec);
        total += it.n_pairs;
    }
    c->items.assign(items, items + n_items);
    c->opts = o;
    c->cur_item = 0; c->cur_off = 0;
    c->next_launch = 0; c->next_return = 0;
    c->gen_done[0] = c->gen_done[1] = c->gz_done[0] = c->gz_done[1] = c->copy_done = nullptr;
    for (auto &s : c->slots) s.gz_code_ready[0] = s.gz_code_ready[1] = false;  // the model or the genome may have changed
    c->exhausted = (total == 0);
    if (total == 0) { c->active = true; return ISS_OK; }
    const uint64_t batch_pairs = std::min<uint64_t>(o.batch_pairs, total);
    // tile geometry: T pairs per CTA so that both stream images + bookkeeping fit in shared memory
    uint32_t T = 64, q_blocks = 1;
    int tile_env = 0;
    if (const char *e = getenv("ISS_TILE")) tile_env = atoi(e);  // A/B switch for benchmarks and tests: 64 or 128 pairs per tile, one block per stage
    if (tile_env == 128) T = 128;
    uint32_t img_cap = 0;
    size_t smem = 0;
    c->err_cap = 512;
    // quality-row pipeline geometry: one stage = the 16-bit rows of q_blocks Philox blocks (2 positions each) for the
    // populated bins of an orientation
    const bool kde = c->mh.qual_mode == ISS_QUAL_KDE;
    // bytes of one block's rows: the short rows where the model has them and the geometry has a build for them
    auto q_cap1_of = [&](uint32_t T_, uint32_t split) -> uint32_t {
        if (!kde) return 0u;
        const bool sh = c->mh.q8_ok && iss_has_short_row_build(T_, split);
        return 2u * (sh ? std::max(c->mh.q8_ps[0], c->mh.q8_ps[1]) : std::max(c->mh.q16_ps[0], c->mh.q16_ps[1]));
    };
    auto short_of = [&](uint32_t T_, uint32_t split) -> uint32_t { return kde && c->mh.q8_ok && iss_has_short_row_build(T_, split) ? 1u : 0u; };
    auto layout = [&](uint32_t T_, uint32_t split, uint32_t qb) {
        return iss_make_layout(T_, (uint32_t)align_up(32 + (uint64_t)T_ * max_rec, 16), c->err_cap, kTieCap, 2u << split, qb * q_cap1_of(T_, split));
    };
    Lay lay{};
    int bps = 0;
    uint32_t split_shift = 1;
    bool wide = false;
    if (kde && tile_env == 0) {
        // short reads: 128-pair tiles (512 threads, 2 CTAs/SM) with TWO blocks per stage.  Half as many CTAs per SM hold
        // half as many stages, so a stage can be twice as long in the same shared memory; every wait, hand-back and
        // cp.async.bulk of the row pipeline then serves four positions instead of two.
        const uint32_t qb = 2u;
        const Lay w = layout(128, 1, qb);
        int b = 0;
        if (w.total <= kMaxSmem && iss_occupancy(128, 1, qb, short_of(128, 1), w.total, &b) == cudaSuccess && b >= 2) {
            wide = true; T = 128; q_blocks = qb; lay = w; smem = w.total; bps = b;
            img_cap = (uint32_t)align_up(32 + (uint64_t)T * max_rec, 16);
        }
    }
    if (!wide) {
        for (;; T >>= 1) {
            img_cap = (uint32_t)align_up(32 + (uint64_t)T * max_rec, 16);
            lay = layout(T, 1, 1);
            smem = lay.total;
            if (smem <= kMaxSmem) break;
            if (T == 32) return fail(ISS_E_ARG, "records too long for one tile (read length / name length)");
        }
        CU(iss_occupancy(T, split_shift, 1, short_of(T, split_shift), smem, &bps));
        if (bps < 1) return fail(ISS_E_CUDA, "kernel does not fit on an SM");
        if (bps <= 2 && T == 64) {  // long reads: the images leave room for <= 2 CTAs/SM -> 4 threads per read keep 32 warps/SM
            int bps4 = 0;
            const Lay lay4 = layout(T, 2, 1);
            if (lay4.total <= kMaxSmem) {
                CU(iss_occupancy(T, 2, 1, short_of(T, 2), lay4.total, &bps4));
                if (bps4 >= bps) { split_shift = 2; bps = bps4; lay = lay4; smem = lay4.total; }
            }
        }
    }
    // the substitution list takes whatever shared memory the occupancy leaves (long-read tiles run 2 CTAs/SM with room to
    // spare; a model with ~ 5 errors per read fills 512 entries): largest power of two that keeps the CTAs per SM
    for (uint32_t cap = 8192; cap > c->err_cap; cap >>= 1) {
        const uint32_t keep = c->err_cap;
        c->err_cap = cap;
        const Lay big = layout(T, split_shift, q_blocks);
        int b2 = 0;
        if (big.total <= kMaxSmem && iss_occupancy(T, split_shift, q_blocks, short_of(T, split_shift), big.total, &b2) == cudaSuccess && b2 >= bps) { lay = big; smem = big.total; break; }
        c->err_cap = keep;
    }
    c->split_shift = split_shift;
    c->lay = lay;
    c->q_stage_cap = q_blocks * q_cap1_of(T, split_shift);
    c->q_blocks = q_blocks;
    c->q_short = short_of(T, split_shift);
    if (const char *e = getenv("ISS_BPS")) {  // A/B switch for benchmarks: fewer CTAs per SM (with ISS_CARVEOUT: more L1)
        const int cap = atoi(e);
        if (cap >= 1 && cap < bps) bps = cap;
    }
    if (const char *e = getenv("ISS_CARVEOUT")) CU(iss_kernel_carveout(atoi(e)));
    const uint32_t grid = (uint32_t)(c->sm_count * bps);
    const uint64_t slot_cap = align_up(batch_pairs * max_rec + 64, 256);
    const uint32_t max_segs = 1u << 16;
    // --store_mutations: room for every event a batch can produce at all -- per read at most L substitutions, 4 (L - 1)
    // insertions and L - 1 deletions -- capped at 1 GB per slot (256 events per pair at the default batch size); a model
    // that still overflows the cap gets ISS_E_DEVICE.
    const uint64_t event_cap = c->mh.store_mutations ? std::min<uint64_t>(2ull * batch_pairs * (6ull * L), (1ull << 30) / sizeof(uint4)) + 1024 : 0;
    const bool realloc = c->slots.size() != o.n_slots || (o.compress && (c->slots.empty() || !c->slots[0].d_gz_out[0])) || (c->mh.store_mutations && (c->slots.empty() || !c->slots[0].d_events || c->slots[0].event_cap < event_cap)) || c->slot_cap < slot_cap || c->max_segs != max_segs ||
                         (bool)c->opts.device_only != (c->slots.size() && c->slots[0].h_out[0] == nullptr);
    if (realloc) {
        free_slots(c);
        c->slots.resize(o.n_slots);
        for (auto &s : c->slots) {
            CU(cudaMalloc(&s.d_out[0], slot_cap));
            CU(cudaMalloc(&s.d_out[1], slot_cap));
            if (!o.device_only) {
                for (int k = 0; k < 2; ++k) {
                    CU((cudaError_t)iss_host_alloc(&s.h_buf[k], slot_cap, c->ring_alloc, c->device));
                    s.h_out[k] = s.h_buf[k].ptr;
                }
                c->stats.ring_numa_node = s.h_buf[0].node;
                c->stats.ring_alloc_mode = (uint32_t)s.h_buf[0].mode;
            }
            if (c->mh.store_mutations) {
                s.event_cap = event_cap;
                CU(cudaMalloc(&s.d_events, sizeof(uint4) * s.event_cap));
                CU(cudaMalloc(&s.d_evcount, sizeof(unsigned long long)));
                CU(cudaMallocHost(&s.h_evcount, sizeof(unsigned long long)));
            }
            if (o.compress) {
                s.gz_chunks_cap = iss_gz_chunks(slot_cap) + 1;
                for (int k = 0; k < 2; ++k) {
                    CU(cudaMalloc(&s.d_gz_out[k], s.gz_chunks_cap * ISS_GZ_STRIDE));
                    CU(cudaMalloc(&s.d_gz_work[k], ISS_GZ_WORK_WORDS * sizeof(uint32_t)));  // the call scratch
                    CU(cudaMalloc(&s.d_gz_status[k], (s.gz_chunks_cap + 1) * sizeof(uint64_t)));
                }
                CU(cudaMallocHost(&s.h_gz_total, 2 * sizeof(uint64_t)));
            }
            CU(cudaMalloc(&s.d_segs, sizeof(DevSegment) * max_segs));
            CU(cudaMallocHost(&s.h_segs, sizeof(DevSegment) * max_segs));
            CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
            if (o.compress) {
                int least = 0, greatest = 0;
                CU(cudaDeviceGetStreamPriorityRange(&least, &greatest));
                CU(cudaStreamCreateWithPriority(&s.zstream, cudaStreamNonBlocking, greatest));
            }
            CU(cudaEventCreate(&s.ev_k0));
            CU(cudaEventCreate(&s.ev_k1));
            CU(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
            CU(cudaEventCreate(&s.ev_gz));
            CU(cudaEventCreate(&s.ev_end));
        }
        c->slot_cap = slot_cap;
        c->max_segs = max_segs;
    }
    for (auto &s : c->slots) s.state = 0;
    c->scratch_replay_bytes = align_up((uint64_t)grid * ISS_SLOW_SLOTS * (align_up(5ull * L + 8, 16) + align_up(L, 16)), 256);
    const uint64_t scratch = c->scratch_replay_bytes + (uint64_t)grid * kErrOvfCap * sizeof(uint32_t);
    for (auto &s : c->slots) {
        if (scratch > s.scratch_bytes) {
            if (s.d_scratch) cudaFree(s.d_scratch);
            CU(cudaMalloc(&s.d_scratch, scratch));
            s.scratch_bytes = scratch;
        }
    }
    c->batch_pairs = batch_pairs;
    c->max_rec = max_rec;
    c->tile_pairs = T;
    c->img_cap = img_cap;
    c->smem = smem;
    c->grid = grid;
    c->n_slots = o.n_slots;
    c->active = true;
    // prime the ring
    for (uint32_t si = 0; si < c->n_slots && !c->exhausted; ++si)
        if (int rc = launch_batch(c, si)) {
            const std::string msg = g_err;
            iss_ctx_abort(c);  // a failed submit leaves a usable context
            return fail(rc, msg);
        }
    return ISS_OK;
}

int iss_ctx_next(iss_ctx *c, iss_batch *b, int *have) {
    if (!c || !b || !have) return fail(ISS_E_ARG, "null argument");
    if (int rc = set_device(c)) return rc;
    *have = 0;
    if (!c->active) return fail(ISS_E_STATE, "no submit in progress");
    if (c->next_return == c->next_launch) {  // nothing in flight
        if (c->exhausted) {
            c->active = false;
            unsigned long long cnt[4];
            CU(cudaMemcpy(cnt, c->d_counters, sizeof cnt, cudaMemcpyDeviceToHost));
            c->stats.slow_reads = cnt[0];
            c->stats.subst_errors = cnt[1];
            if (cnt[2]) {
                CU(cudaMemset(c->d_counters + 2, 0, sizeof(unsigned long long)));
                char buf[128];
                snprintf(buf, sizeof buf, "device error flags 0x%llx (1 window out of range, 2 empty quality bin, 4 bad character, 8 mutation buffer overflow)", cnt[2]);
                return fail(ISS_E_DEVICE, buf);
            }
            return ISS_OK;
        }
        return fail(ISS_E_STATE, "all slots are held by the consumer: release a batch first");
    }
    const uint32_t si = (uint32_t)(c->next_return % c->n_slots);
    Slot &s = c->slots[si];
    if (s.state != 1 || s.batch_index != c->next_return) return fail(ISS_E_STATE, "internal: ring out of order");
    {
        NvtxRange nvtx("iss:wait_batch");
        CU(cudaEventSynchronize(s.ev_done));
    }
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1));
    if (c->trace && c->marks[0]) {  // ISS_TRACE=1: device timeline of the batch relative to mark 0 (ms)
        float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
        cudaEventElapsedTime(&t0, c->marks[0], s.ev_k0);
        cudaEventElapsedTime(&t1, c->marks[0], s.ev_k1);
        if (c->opts.compress) cudaEventElapsedTime(&t2, c->marks[0], s.ev_gz);
        cudaEventElapsedTime(&t3, c->marks[0], s.ev_end);
        cudaGetLastError();
        fprintf(stderr, "[iss trace] batch %llu slot %u: kernel %.3f..%.3f  gz_end %.3f  queued_copies_end %.3f\n",
                (unsigned long long)s.batch_index, si, t0, t1, t2, t3);
    }
    c->stats.kernel_ms += ms;
    s.state = 2;
    b->r1 = c->opts.device_only ? s.d_out[0] : s.h_out[0];
    b->r2 = c->opts.device_only ? s.d_out[1] : s.h_out[1];
    b->nbytes = s.nbytes;
    b->nbytes2 = s.nbytes;
    if (c->opts.compress) {  // sizes are known only now: fetch exactly the compressed bytes
        b->nbytes = s.h_gz_total[0];
        b->nbytes2 = s.h_gz_total[1];
        if (ISS_GZ_FAILED(b->nbytes) || ISS_GZ_FAILED(b->nbytes2)) return fail(ISS_E_STATE, "internal: gzip member placement timed out");
        if (c->opts.device_only) {
            b->r1 = s.d_gz_out[0];
            b->r2 = s.d_gz_out[1];
        } else {
            if (b->nbytes > c->slot_cap || b->nbytes2 > c->slot_cap) return fail(ISS_E_STATE, "internal: gzip stream larger than the slot");
            const uint64_t tot[2] = {b->nbytes, b->nbytes2};
            bool more = false;
            for (int k = 0; k < 2; ++k) {
                if (s.nbytes) c->gz_ratio[k] = (double)tot[k] / (double)s.nbytes;
                if (tot[k] > s.gz_pre[k]) {
                    CU(cudaMemcpyAsync(s.h_out[k] + s.gz_pre[k], s.d_gz_out[k] + s.gz_pre[k], tot[k] - s.gz_pre[k], cudaMemcpyDeviceToHost, s.stream));
                    c->stats.d2h_bytes += tot[k] - s.gz_pre[k];
                    more = true;
                }
            }
            if (more) CU(cudaStreamSynchronize(s.stream));
        }
    }
    b->vcf = nullptr;
    b->vcf_bytes = 0;
    if (c->mh.store_mutations) {
        if (int rc = format_vcf(c, s)) return rc;
        b->vcf = reinterpret_cast<const uint8_t *>(s.vcf.data());
        b->vcf_bytes = s.vcf.size();
    }
    b->n_pairs = s.n_pairs;
    b->first_item = s.first_item;
    b->batch_index = s.batch_index;
    b->slot = si;
    b->kernel_ms = ms;
    c->next_return++;
    *have = 1;
    return ISS_OK;
}

int iss_ctx_release(iss_ctx *c, const iss_batch *b) {
    if (!c || !b) return fail(ISS_E_ARG, "null argument");
    if (int rc = set_device(c)) return rc;
    if (b->slot >= c->slots.size() || c->slots[b->slot].state != 2) return fail(ISS_E_STATE, "batch is not held");
    c->slots[b->slot].state = 0;
    if (c->active && !c->exhausted) return launch_batch(c, b->slot);
    return ISS_OK;
}

int iss_ctx_abort(iss_ctx *c) {
    if (!c) return fail(ISS_E_ARG, "null argument");
    if (int rc = set_device(c)) return rc;
    for (auto &s : c->slots)
        if (s.stream) cudaStreamSynchronize(s.stream);
    cudaGetLastError();
    for (auto &s : c->slots) s.state = 0;
    c->items.clear();
    c->cur_item = c->cur_off = 0;
    c->next_launch = c->next_return = 0;
    c->active = false;
    c->exhausted = false;
    if (c->d_counters) CU(cudaMemset(c->d_counters + 2, 0, sizeof(unsigned long long)));
    return ISS_OK;
}

int iss_ctx_sync(iss_ctx *c) {
    if (!c) return fail(ISS_E_ARG, "null argument");
    if (int rc = set_device(c)) return rc;
    CU(cudaDeviceSynchronize());
    return ISS_OK;
}

int iss_ctx_mark(iss_ctx *c, int idx) {
    if (!c || idx < 0 || idx >= ISS_N_MARKS) return fail(ISS_E_ARG, "bad mark index");
    if (int rc = set_device(c)) return rc;
    if (!c->tstream) CU(cudaStreamCreateWithFlags(&c->tstream, cudaStreamNonBlocking));
    if (!c->marks[idx]) CU(cudaEventCreate(&c->marks[idx]));
    for (auto &s : c->slots) {  // order the mark after everything enqueued on the slot streams
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CU(cudaEventRecord(e, s.stream));
        CU(cudaStreamWaitEvent(c->tstream, e, 0));
        CU(cudaEventDestroy(e));
    }
    CU(cudaEventRecord(c->marks[idx], c->tstream));
    return ISS_OK;
}

int iss_ctx_elapsed_ms(iss_ctx *c, int a, int b, float *ms) {
    if (!c || !ms || a < 0 || b < 0 || a >= ISS_N_MARKS || b >= ISS_N_MARKS || !c->marks[a] || !c->marks[b])
        return fail(ISS_E_ARG, "bad mark index");
    if (int rc = set_device(c)) return rc;
    CU(cudaEventSynchronize(c->marks[b]));
    CU(cudaEventElapsedTime(ms, c->marks[a], c->marks[b]));
    return ISS_OK;
}

int iss_ctx_stats(iss_ctx *c, iss_stats *out) {
    if (!c || !out) return fail(ISS_E_ARG, "null argument");
    *out = c->stats;
    return ISS_OK;
}

uint64_t iss_fastq_bytes(uint32_t name_len, uint64_t cpu, uint32_t L, uint64_t i0, uint64_t i1) {
    const uint32_t fixed = record_fixed_len(L, name_len, (uint32_t)cpu);
    return iss_cf(i1, fixed) - iss_cf(i0, fixed);
}

int iss_ctx_gzip(iss_ctx *c, const uint8_t *in, uint64_t n, uint8_t *out, uint64_t cap, uint64_t *out_n) {
    if (!c || (!in && n) || !out || !out_n) return fail(ISS_E_ARG, "null argument");
    if (int rc = set_device(c)) return rc;
    const uint32_t nc = iss_gz_chunks(n);
    uint8_t *d_in = nullptr, *d_out = nullptr;
    uint32_t *d_work = nullptr;
    uint64_t *d_status = nullptr;
    uint64_t total = 0;
    int rc = [&]() -> int {
        CU(cudaMalloc(&d_in, n + 16));
        CU(cudaMalloc(&d_out, (uint64_t)(nc + 1) * ISS_GZ_STRIDE));
        CU(cudaMalloc(&d_work, ISS_GZ_WORK_WORDS * sizeof(uint32_t)));
        CU(cudaMalloc(&d_status, (nc + 2) * sizeof(uint64_t)));
        if (n) CU(cudaMemcpy(d_in, in, n, cudaMemcpyHostToDevice));
        CU(iss_gz_compress(d_in, n, d_status, d_out, d_work, c->sm_count, 0, 0));
        CU(cudaMemcpy(&total, d_status + nc, sizeof total, cudaMemcpyDeviceToHost));
        if (ISS_GZ_FAILED(total)) return fail(ISS_E_STATE, "internal: gzip member placement timed out");
        if (total > cap) return fail(ISS_E_ARG, "output buffer too small");
        if (total) CU(cudaMemcpy(out, d_out, total, cudaMemcpyDeviceToHost));
        return ISS_OK;
    }();
    *out_n = total;
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_work); cudaFree(d_status);
    return rc;
}

int iss_ctx_philox(iss_ctx *c, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    if (!c) return fail(ISS_E_ARG, "null argument");
    if (int rc = set_device(c)) return rc;
    uint32_t *d = nullptr;
    CU(cudaMalloc(&d, 16));
    CU(iss_launch_philox(ctr, key, d, 0));
    CU(cudaMemcpy(out, d, 16, cudaMemcpyDeviceToHost));
    CU(cudaFree(d));
    return ISS_OK;
}

}  // extern "C"
