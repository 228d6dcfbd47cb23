// iss_kernels.cu -- sm_100a kernels of the read-pair generation hot path.
//
// One persistent kernel (iss_generate_kernel) produces finished FASTQ text for a batch of read pairs:
//
//   tile = T pairs (2T reads), blockDim = 2*S*T: thread <-> (pair, orientation, one of S position parts; S = 2, or 4
//   for long reads whose images leave room for only 2 CTAs/SM); a warp holds 32
//   different pairs in the same role, so its lanes walk the model tables in lockstep
//   phase 1  pair setup        insert size, window positions, quality bins        generator.py:121-177, kde.py:72-97
//   phase 2  headers + gather  2-bit (or byte) genome windows -> ASCII in smem     generator.py:146-153,179-183, util.py:48-92
//   phase 3  indel scan        Bernoulli tests in the no-event world; a read with  error_models/__init__.py:187-224
//            + replay          any hit replays the reference's list semantics
//   phase 4  phred sampling    per-position CDF search + error test                kde.py:82-85, __init__.py:94
//   phase 5  substitutions     compacted list, parallel                            __init__.py:95-109
//   phase 6  copy-out          smem image -> HBM with one cp.async.bulk per stream generator.py:64-65
//
// The FASTQ image of a tile is assembled in shared memory at the same 16-byte phase as its place in
// the output stream, so the aligned middle leaves through the TMA bulk-copy engine
// (cp.async.bulk.global.shared::cta) while the CTA starts the next tile.
//
// All randomness is Philox4x32-10 addressed by (key, pair ordinal, stream, block) -- "RNG contract v2".
// No floating point anywhere: model tables are integer thresholds (include/iss_b200.h).
#include <cuda_runtime.h>
#include <stdint.h>

#include "iss_device.cuh"
#include "iss_kernels.h"

extern __shared__ __align__(16) uint8_t iss_smem[];

namespace {

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// A, C, G, T in either case (bits 0, 2, 6, 19 of the letter index)
__device__ __forceinline__ bool is_acgt(uint32_t c) {
    uint32_t u = (c & 0xDFu) - 65u;
    return u < 26u && ((0x00080045u >> u) & 1u);
}
// base index A=0 C=1 T=2 G=3 (works for both cases); complement = code ^ 2
__device__ __forceinline__ uint32_t base_code(uint32_t c) { return (c >> 1) & 3u; }
__device__ __forceinline__ uint32_t code_to_ascii(uint32_t code) { return __byte_perm(0x47544341u, 0u, code) & 0xFFu; }

// iss/util.py:57-88 complement dictionary, case preserved.  Letters outside it were rejected at load.
//                                       @   A    B    C    D   E  F   G    H   I  J   K   L   M    N   O  P  Q   R    S    T   U   V    W   X   Y
__device__ __constant__ char kComplement[32] = {0, 'T', 'V', 'G', 'H', 0, 0, 'C', 'D', 0, 0, 'M', 0, 'K', 'N', 0, 0, 0, 'Y', 'S', 'A', 0, 'B', 'W', 0, 'R', 0, 0, 0, 0, 0, 0};
__device__ __forceinline__ uint32_t complement_ascii(uint32_t c) { return (uint32_t)kComplement[c & 31u] | (c & 0x20u); }

__device__ __forceinline__ uint32_t pick(const Philox4 &w, uint32_t j) {
    return j == 0 ? w.x : (j == 1 ? w.y : (j == 2 ? w.z : w.w));
}
__device__ __forceinline__ uint32_t pick(const uint4 &w, uint32_t j) {
    return j == 0 ? w.x : (j == 1 ? w.y : (j == 2 ? w.z : w.w));
}
__device__ __forceinline__ uint64_t u64of(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); }

__device__ __forceinline__ uint64_t *u64p(uint32_t off) { return reinterpret_cast<uint64_t *>(iss_smem + off); }
__device__ __forceinline__ uint32_t *u32p(uint32_t off) { return reinterpret_cast<uint32_t *>(iss_smem + off); }
__device__ __forceinline__ uint8_t *imgp(const Lay &l, uint32_t o) { return iss_smem + (o ? l.img[1] : l.img[0]); }
// explicit shared-window accesses for the hot loop: a 32-bit address held in ONE register, where a generic pointer costs two
// and makes the compiler re-derive the window base (S2R CgaCtaId, MOV, LEA) at every use
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

// Quality rows (kde.py:82-85: q = searchsorted(cdf_pos, u) = #{j : w > thr[j]}) are searched on the top 16 bits of the
// draw: a row is the 64 sorted entries e[j] = thr[j] >> 16 as u16 (pad = 0xFFFF; e[63] is always pad), staged into shared
// memory by cp.async.bulk (phase 4).  Branch-free lower bound, the byte offset of e[count] being the accumulator:
//     for s in 32, 16, 8, 4, 2, 1:  if (w16 > e[count + s - 1]) count += s          (LDS.U16 + ISETP + predicated IADD)
// With w16 != e[j] for all j the 16-bit order is the 32-bit order.  The row is sorted, so a tie can only start at the entry
// the search stops at: e[count] == w16 (~ 41 / 65536 of the draws) defers the position to the exact pass (phase 4b), which
// evaluates it on the 32-bit row -- nothing is called from the hot loop.
// row_off: offset of the read's staged row from iss_smem
// rowA / rowB: shared-window addresses of the read's staged rows at two positions (smem_addr(iss_smem) + offset).  The
// two lower-bound searches are independent chains of 7 dependent LDS; stepping them together halves what a warp waits.
// Short rows (kK = 8): the row holds its <= 8 DISTINCT non-zero 16-bit values (pad 0xFFFF, the last distinct value is always the
// pad) and, behind them, for c = 0..7 the number of entries of the full row below the c-th value -- which is what the
// 64-entry search would have counted, for every draw that equals none of the values (that one is a tie there too).
// Three steps instead of six, and a 24-byte row instead of 132.
__device__ __forceinline__ void q8_search2(uint32_t rowA, uint32_t wA, uint32_t rowB, uint32_t wB, uint32_t &qA, uint32_t &qB,
                                           bool &tieA, bool &tieB) {
    uint32_t a = rowA, b = rowB;
#define ISS_Q8_STEP(S)                                                                                               \
    {                                                                                                                \
        uint32_t ea, eb;                                                                                             \
        asm volatile("ld.shared.u16 %0, [%1+" #S "*2-2];" : "=r"(ea) : "r"(a));                                      \
        asm volatile("ld.shared.u16 %0, [%1+" #S "*2-2];" : "=r"(eb) : "r"(b));                                      \
        asm("{\n\t.reg .pred p;\n\tsetp.gt.u32 p, %1, %2;\n\t@p add.u32 %0, %0, " #S "*2;\n\t}" : "+r"(a) : "r"(wA), "r"(ea)); \
        asm("{\n\t.reg .pred p;\n\tsetp.gt.u32 p, %1, %2;\n\t@p add.u32 %0, %0, " #S "*2;\n\t}" : "+r"(b) : "r"(wB), "r"(eb)); \
    }
    ISS_Q8_STEP(4) ISS_Q8_STEP(2) ISS_Q8_STEP(1)
#undef ISS_Q8_STEP
    uint32_t ea, eb;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(ea) : "r"(a));
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(eb) : "r"(b));
    // a - rowA = 2 c: the count byte of value c sits at rowA + 16 + c
    asm volatile("ld.shared.u8 %0, [%1+16];" : "=r"(qA) : "r"(rowA + ((a - rowA) >> 1)));
    asm volatile("ld.shared.u8 %0, [%1+16];" : "=r"(qB) : "r"(rowB + ((b - rowB) >> 1)));
    // The value 0 (qualities below the model's lowest: cumulative probability < 2^-16) is not stored: every draw with a
    // non-zero top half is above it.  A draw whose top half IS zero goes to the exact pass with nothing known about it.
    tieA = (ea == wA) | (wA == 0u);
    tieB = (eb == wB) | (wB == 0u);
    qA = wA ? qA : 0u;
    qB = wB ? qB : 0u;
}
__device__ __forceinline__ void q16_search2(uint32_t rowA, uint32_t wA, uint32_t rowB, uint32_t wB, uint32_t &qA, uint32_t &qB,
                                            bool &tieA, bool &tieB) {
    uint32_t a = rowA, b = rowB;
#define ISS_Q16_STEP(S)                                                                                              \
    {                                                                                                                \
        uint32_t ea, eb;                                                                                             \
        asm volatile("ld.shared.u16 %0, [%1+" #S "*2-2];" : "=r"(ea) : "r"(a));                                      \
        asm volatile("ld.shared.u16 %0, [%1+" #S "*2-2];" : "=r"(eb) : "r"(b));                                      \
        asm("{\n\t.reg .pred p;\n\tsetp.gt.u32 p, %1, %2;\n\t@p add.u32 %0, %0, " #S "*2;\n\t}" : "+r"(a) : "r"(wA), "r"(ea)); \
        asm("{\n\t.reg .pred p;\n\tsetp.gt.u32 p, %1, %2;\n\t@p add.u32 %0, %0, " #S "*2;\n\t}" : "+r"(b) : "r"(wB), "r"(eb)); \
    }
    ISS_Q16_STEP(32) ISS_Q16_STEP(16) ISS_Q16_STEP(8) ISS_Q16_STEP(4) ISS_Q16_STEP(2) ISS_Q16_STEP(1)
#undef ISS_Q16_STEP
    uint32_t ea, eb;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(ea) : "r"(a));
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(eb) : "r"(b));
    tieA = ea == wA;
    tieB = eb == wB;
    qA = (a - rowA) >> 1;
    qB = (b - rowB) >> 1;
}

// the same four at once on short rows
__device__ __forceinline__ void q8_search4(uint32_t rowA, uint32_t wA, uint32_t rowB, uint32_t wB, uint32_t rowC, uint32_t wC, uint32_t rowD, uint32_t wD,
                                           uint32_t &qA, uint32_t &qB, uint32_t &qC, uint32_t &qD, bool &tieA, bool &tieB, bool &tieC, bool &tieD) {
    uint32_t a = rowA, b = rowB, c = rowC, d = rowD;
#define ISS_Q8_STEP4(S)                                                                                              \
    {                                                                                                                \
        uint32_t ea, eb, ec, ed;                                                                                     \
        asm volatile("ld.shared.u16 %0, [%1+" #S "*2-2];" : "=r"(ea) : "r"(a));                                      \
        asm volatile("ld.shared.u16 %0, [%1+" #S "*2-2];" : "=r"(eb) : "r"(b));                                      \
        asm volatile("ld.shared.u16 %0, [%1+" #S "*2-2];" : "=r"(ec) : "r"(c));                                      \
        asm volatile("ld.shared.u16 %0, [%1+" #S "*2-2];" : "=r"(ed) : "r"(d));                                      \
        asm("{\n\t.reg .pred p;\n\tsetp.gt.u32 p, %1, %2;\n\t@p add.u32 %0, %0, " #S "*2;\n\t}" : "+r"(a) : "r"(wA), "r"(ea)); \
        asm("{\n\t.reg .pred p;\n\tsetp.gt.u32 p, %1, %2;\n\t@p add.u32 %0, %0, " #S "*2;\n\t}" : "+r"(b) : "r"(wB), "r"(eb)); \
        asm("{\n\t.reg .pred p;\n\tsetp.gt.u32 p, %1, %2;\n\t@p add.u32 %0, %0, " #S "*2;\n\t}" : "+r"(c) : "r"(wC), "r"(ec)); \
        asm("{\n\t.reg .pred p;\n\tsetp.gt.u32 p, %1, %2;\n\t@p add.u32 %0, %0, " #S "*2;\n\t}" : "+r"(d) : "r"(wD), "r"(ed)); \
    }
    ISS_Q8_STEP4(4) ISS_Q8_STEP4(2) ISS_Q8_STEP4(1)
#undef ISS_Q8_STEP4
    uint32_t ea, eb, ec, ed;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(ea) : "r"(a));
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(eb) : "r"(b));
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(ec) : "r"(c));
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(ed) : "r"(d));
    asm volatile("ld.shared.u8 %0, [%1+16];" : "=r"(qA) : "r"(rowA + ((a - rowA) >> 1)));
    asm volatile("ld.shared.u8 %0, [%1+16];" : "=r"(qB) : "r"(rowB + ((b - rowB) >> 1)));
    asm volatile("ld.shared.u8 %0, [%1+16];" : "=r"(qC) : "r"(rowC + ((c - rowC) >> 1)));
    asm volatile("ld.shared.u8 %0, [%1+16];" : "=r"(qD) : "r"(rowD + ((d - rowD) >> 1)));
    tieA = (ea == wA) | (wA == 0u); tieB = (eb == wB) | (wB == 0u); tieC = (ec == wC) | (wC == 0u); tieD = (ed == wD) | (wD == 0u);
    qA = wA ? qA : 0u; qB = wB ? qB : 0u; qC = wC ? qC : 0u; qD = wD ? qD : 0u;
}
// four positions at once: the two blocks of a stage (four independent chains)
__device__ __forceinline__ void q16_search4(uint32_t rowA, uint32_t wA, uint32_t rowB, uint32_t wB, uint32_t rowC, uint32_t wC, uint32_t rowD, uint32_t wD,
                                            uint32_t &qA, uint32_t &qB, uint32_t &qC, uint32_t &qD, bool &tieA, bool &tieB, bool &tieC, bool &tieD) {
    uint32_t a = rowA, b = rowB, c = rowC, d = rowD;
#define ISS_Q16_STEP4(S)                                                                                             \
    {                                                                                                                \
        uint32_t ea, eb, ec, ed;                                                                                     \
        asm volatile("ld.shared.u16 %0, [%1+" #S "*2-2];" : "=r"(ea) : "r"(a));                                      \
        asm volatile("ld.shared.u16 %0, [%1+" #S "*2-2];" : "=r"(eb) : "r"(b));                                      \
        asm volatile("ld.shared.u16 %0, [%1+" #S "*2-2];" : "=r"(ec) : "r"(c));                                      \
        asm volatile("ld.shared.u16 %0, [%1+" #S "*2-2];" : "=r"(ed) : "r"(d));                                      \
        asm("{\n\t.reg .pred p;\n\tsetp.gt.u32 p, %1, %2;\n\t@p add.u32 %0, %0, " #S "*2;\n\t}" : "+r"(a) : "r"(wA), "r"(ea)); \
        asm("{\n\t.reg .pred p;\n\tsetp.gt.u32 p, %1, %2;\n\t@p add.u32 %0, %0, " #S "*2;\n\t}" : "+r"(b) : "r"(wB), "r"(eb)); \
        asm("{\n\t.reg .pred p;\n\tsetp.gt.u32 p, %1, %2;\n\t@p add.u32 %0, %0, " #S "*2;\n\t}" : "+r"(c) : "r"(wC), "r"(ec)); \
        asm("{\n\t.reg .pred p;\n\tsetp.gt.u32 p, %1, %2;\n\t@p add.u32 %0, %0, " #S "*2;\n\t}" : "+r"(d) : "r"(wD), "r"(ed)); \
    }
    ISS_Q16_STEP4(32) ISS_Q16_STEP4(16) ISS_Q16_STEP4(8) ISS_Q16_STEP4(4) ISS_Q16_STEP4(2) ISS_Q16_STEP4(1)
#undef ISS_Q16_STEP4
    uint32_t ea, eb, ec, ed;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(ea) : "r"(a));
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(eb) : "r"(b));
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(ec) : "r"(c));
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(ed) : "r"(d));
    tieA = ea == wA; tieB = eb == wB; tieC = ec == wC; tieD = ed == wD;
    qA = (a - rowA) >> 1; qB = (b - rowB) >> 1; qC = (c - rowC) >> 1; qD = (d - rowD) >> 1;
}

// substitution list full: the entry goes to this CTA's stretch of a list in global memory (models with ~ 4 or more errors
// per read at 64 pairs per tile); only beyond that does the tile fall back to the exact pass
__device__ __noinline__ void err_overflow(const GenParams &p, uint32_t slot, uint32_t v) {
    const uint32_t k = slot - p.err_cap;
    if (k < p.err_ovf_cap) p.err_ovf[(size_t)blockIdx.x * p.err_ovf_cap + k] = v;
}
// queue (read, 16-bit count, position) for the exact pass; a full queue makes the whole tile take it
__device__ __forceinline__ void defer_position(const Lay &l, uint32_t cap, uint32_t rd, uint32_t pos, uint32_t c16) {
    const uint32_t slot = atomicAdd(&u32p(l.sCtl)[4], 1u);
    if (slot < cap) u32p(l.sTie)[slot] = (rd << 24) | (c16 << 16) | pos;
}

// The definition the fast paths must agree with: phred of (read rd, position pos) for the quality draw wq, from the
// model's full-precision tables (kde.py:82-85 on the 64 sorted 32-bit thresholds; basic.py:41-53; perfect.py:35-42).
// c16: thresholds known to be below the draw (0: none known) -- the 16-bit search has already counted the entries whose top
// half is smaller, so only the few that share the draw's top half are compared in full
__device__ __noinline__ uint32_t quality_exact(const GenParams &p, uint32_t rd, uint32_t pos, uint32_t wq, uint32_t c16) {
    if (p.qual_mode == ISS_QUAL_CONST_MODE) return p.const_phred;
    if (p.qual_mode == ISS_QUAL_TABLE_MODE) {
        uint32_t a = 0, e = p.n_basic;
        while (a < e) { const uint32_t mid = (a + e) >> 1; if (wq >= __ldg(p.m.basic_thr + mid)) a = mid + 1; else e = mid; }
        return p.m.basic_val[a];
    }
    const uint32_t o = rd >= p.tile_pairs ? 1u : 0u;
    const uint32_t bin = u32p(p.lay.sQrow)[rd] >> 16;
    const uint32_t *__restrict__ row = p.m.qual_exact + (((size_t)(o * 4u + bin) * p.L + pos) << 6);
    uint32_t c = c16;
    if (c16) {
        while (c < ISS_QROW && wq > __ldg(row + c)) ++c;  // the row is sorted: stop at the first threshold not below the draw
        return c;
    }
#pragma unroll 1
    for (uint32_t s = 32; s >= 1; s >>= 1)
        if (wq > __ldg(row + c + s - 1u)) c += s;  // row[63] == 0xFFFFFFFF: c <= 63
    return c;
}

// 8 two-bit codes (16 bits) -> 8 nibbles: the selectors of two PRMTs into "ACTG"
__device__ __forceinline__ uint32_t spread_codes(uint32_t h) {
    uint32_t x = (h | (h << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    return (x | (x << 2)) & 0x33333333u;
}
// the four bytes of v to shared-window addresses a + kOff .. a + kOff + 3 (no alignment assumed)
template <int kOff>
__device__ __forceinline__ void sts_word_bytes(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u8 [%0+%2], %1;" ::"r"(a), "r"(v), "n"(kOff) : "memory");
    asm volatile("st.shared.u8 [%0+%2], %1;" ::"r"(a), "r"(v >> 8), "n"(kOff + 1) : "memory");
    asm volatile("st.shared.u8 [%0+%2], %1;" ::"r"(a), "r"(v >> 16), "n"(kOff + 2) : "memory");
    asm volatile("st.shared.u8 [%0+%2], %1;" ::"r"(a), "r"(v >> 24), "n"(kOff + 3) : "memory");
}
template <int kOff>
__device__ __forceinline__ void sts_word_bytes_n(uint32_t a, uint32_t v, uint32_t n) {  // only bytes with index < n
#pragma unroll
    for (int t = 0; t < 4; ++t)
        if ((uint32_t)(kOff + t) < n) asm volatile("st.shared.u8 [%0], %1;" ::"r"(a + kOff + t), "r"(v >> (8 * t)) : "memory");
}

__device__ __forceinline__ uint32_t genome_ascii(const DevGenome &g, uint32_t enc, uint64_t abs_pos) {
    if (enc == ISS_ENC_2BIT) {
        uint32_t w = __ldg(g.packed + (abs_pos >> 4));
        return code_to_ascii((w >> ((uint32_t)(abs_pos & 15u) * 2u)) & 3u);
    }
    return __ldg(g.ascii + abs_pos);
}


// bytes of dynamic shared memory for a launch (host mirror in iss_kernels.h: iss_tile_smem_bytes)

// ---- one indel draw site (RNG contract v2) that passed the group's reject test: which events does its draw allow?
//      Returns bit `lane` for an insertion site, bits 4 + b for every base b whose deletion fires at a deletion site.
__device__ __noinline__ uint32_t site_bits(const GenParams &p, uint32_t so, uint32_t kind, uint32_t word, uint32_t thr, uint32_t meta) {
    if (!((meta >> 16) & 1u) || word > thr) return 0u;  // padding, or this site of the group did not pass
    if (kind == 0u) return 1u << ((meta >> 17) & 3u);
    const uint32_t tb = so * p.L + (meta & 0xFFFFu);
    const uint32_t nz = p.m.del_nz[tb];
    const uint4 t = __ldg(reinterpret_cast<const uint4 *>(p.m.del_thr) + tb);
    return ((((nz & 1u) && word <= t.x) ? 1u : 0u) | (((nz & 2u) && word <= t.y) ? 2u : 0u) |
            (((nz & 4u) && word <= t.z) ? 4u : 0u) | (((nz & 8u) && word <= t.w) ? 8u : 0u)) << 4;
}

// ---- mbarrier / TMA bulk-copy helpers (genome windows: global -> shared, completion counted on an mbarrier) -------
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(mbar), "r"(parity)
                     : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(mbar)
                 : "memory");
}
// Staged span of a read's window: the 16-byte aligned run of packed genome bytes that covers every word the gather
// touches ([wabs - 15, wabs + L + 46) in bases), placed 16-byte aligned inside the record's (not yet written) quality line.
__device__ __forceinline__ uint64_t stage_src_byte(uint64_t wabs) { return ((wabs - 16u) >> 2) & ~(uint64_t)15u; }
__device__ __forceinline__ uint32_t stage_dst_off(uint32_t img_off, uint32_t seq_off, uint32_t L) {
    return ((img_off + seq_off + L + 3u + 15u) & ~15u);  // offset from iss_smem
}

// ---- mutation events (--store_mutations, error_models/__init__.py:98-108,197-207,211-221) -------------------
// One 16-byte record per VCF line, appended in arbitrary order; the host sorts by key and formats the text.
//   key = pair (batch-local) << 20 | orientation << 19 | phase << 18 (0 indel loop, 1 substitutions) | position << 4 | order
//   order: insertion lane 0..3, deletion 4 (the order the reference appends them inside one loop iteration)
__device__ __forceinline__ void emit_event(const GenParams &p, uint64_t gpair, uint32_t o, uint32_t phase, uint32_t pos,
                                           uint32_t order, uint32_t ref, uint32_t alt0, uint32_t alt1, uint32_t qual) {
    const unsigned long long slot = atomicAdd(p.event_count, 1ull);
    if (slot >= p.event_cap) { atomicOr(p.counters + 2, (unsigned long long)ISS_DEVERR_EVENTS); return; }
    const uint64_t key = (gpair << 20) | ((uint64_t)o << 19) | ((uint64_t)phase << 18) | ((uint64_t)pos << 4) | order;
    uint4 e;
    e.x = (uint32_t)key;
    e.y = (uint32_t)(key >> 32);
    e.z = ref | (alt0 << 8) | (alt1 << 16) | (qual << 24);
    e.w = 0;
    p.events[slot] = e;
}

// ---- substitution of one (read, position) ---------------------------------------------------------
__device__ __forceinline__ void apply_substitution(const GenParams &p, const Lay &l, uint64_t gp0, uint32_t rel0, uint32_t rd, uint32_t pos) {
    const uint32_t T = p.tile_pairs, L = p.L;
    const uint32_t o = rd >= T ? 1u : 0u, pair = rd - o * T;
    uint8_t *seq = imgp(l, o) + (uint32_t)(rel0 + u32p(l.sRel)[pair]);  // the sum wraps in 32 bits by construction
    const uint32_t c = seq[pos];
    if (!is_acgt(c)) return;  // "RYWSMKHBVDN" are never substituted (__init__.py:94)
    const uint32_t code = base_code(c);
    const uint64_t P = u64p(l.sP)[pair];
    const Philox4 w = philox4x32_10_rk(pos >> 2, iss_read_stream(o, ISS_K_SUBC) | (u32p(l.sAtt)[pair] << 8), (uint32_t)P,
                                       (uint32_t)(P >> 32), p.rk);
    const uint64_t r = pick(w, pos & 3u);
    const uint32_t e = ((o * L + pos) << 2) + code;
    const uint32_t na = p.m.subst_n[e];
    const uint64_t *thr = p.m.subst_thr + (size_t)e * ISS_MAX_ALT;
    uint32_t idx = 0;
    for (uint32_t j = 0; j < na; ++j) idx += (r >= __ldg(thr + j)) ? 1u : 0u;
    if (idx >= na) idx = na - 1u;
    const uint32_t alt = p.m.subst_alt[(size_t)e * ISS_MAX_ALT + idx];
    if (p.store_mutations) {
        // recorded only when the new base differs from the read BEFORE indels ("original", generator.py:155,186)
        const uint64_t wabs = u64p(l.sWabs)[rd];
        const uint32_t enc = u32p(l.sEnc)[pair];
        const uint32_t orig = o ? complement_ascii(genome_ascii(p.g, enc, wabs + (L - 1u - pos))) : genome_ascii(p.g, enc, wabs + pos);
        if (alt != orig) emit_event(p, gp0 + pair, o, 1u, pos, 0u, c, alt, 0u, (uint32_t)seq[L + 3u + pos] - 33u);
    }
    seq[pos] = (uint8_t)alt;
}

// ---- replay of introduce_indels + adjust_seq_length for one read, by one warp -------------------------------
// Reference list semantics (error_models/__init__.py:187-226) restated as a stream:
//   seq = out[0..nucl) ++ [cur] ++ future,   future = LIFO of inserted bases on top of the unread original.
//   insertion at index nucl+1  == push on the LIFO (a later lane lands in front of an earlier one)
//   pop(nucl)                  == drop cur; the next element takes index nucl WITHOUT being examined
//   IndexError at seq[nucl]    == the stream is exhausted: nothing further is drawn
// Every draw of the read was made in parallel beforehand (phase 3c): mask[pos] holds the insertion lanes that fire at loop
// position pos (bits 0-3) and, per base, whether a deletion fires (bits 4-7).  Where the mask is zero and the LIFO is empty
// the loop just moves one original base to the output, so the warp alternates between
//   (a) finding the next position with a non-zero mask (ballot over 32 mask bytes at a time),
//   (b) copying the run of original bases in front of it, 32 bases per step, and
//   (c) lane 0 stepping the reference's loop through that position and on while the LIFO holds bases.
// Cost is proportional to the number of events, not to L: what used to be one thread walking all L positions through
// global memory (~0.3 us per step) is a few warp-wide steps per event.
__device__ __noinline__ void replay_indels_warp(const GenParams &p, const Lay &l, uint64_t gp0, uint32_t rel0, uint32_t rd, uint8_t *scratch,
                                   const uint8_t *mask, uint32_t lane) {
    const uint32_t T = p.tile_pairs, L = p.L, FULL = 0xFFFFFFFFu;
    const uint32_t o = rd >= T ? 1u : 0u, pair = rd - o * T;
    uint8_t *seq = imgp(l, o) + (uint32_t)(rel0 + u32p(l.sRel)[pair]);  // the sum wraps in 32 bits by construction
    uint8_t *orig = scratch;       // [L]  the read before indels
    uint8_t *stk = scratch + L;    // [4L] LIFO of inserted bases (lane 0 only)
    for (uint32_t i = lane; i < L; i += 32u) orig[i] = seq[i];
    // which loop positions have a non-zero mask: one bit each, word r of the bitmap in lane r (up to 1024 positions; the loads
    // are independent of each other, so this costs one memory latency instead of one per 32 positions scanned later)
    const uint32_t nwords = (L + 30u) >> 5;  // positions 0 .. L - 2
    const bool bitmap = nwords <= 32u;
    uint32_t myword = 0;
    if (bitmap) {
#pragma unroll 4
        for (uint32_t r = 0; r < nwords; ++r) {
            const uint32_t pos = 32u * r + lane;
            const uint32_t bal = __ballot_sync(FULL, pos + 1u < L && mask[pos] != 0u);
            if (lane == r) myword = bal;
        }
    }
    __syncwarp();
    uint32_t nucl = 0, k = 0, sp = 0;  // loop position == bytes written so far; next unread original base; LIFO depth
    bool broke = false;                // the reference's loop ended on an IndexError
    while (nucl + 1u < L && !broke) {
        // (a) next loop position >= nucl whose mask is non-zero (L - 1 = none); with bases on the LIFO: right here
        uint32_t e = nucl;
        if (sp == 0) {
            e = L - 1u;
            if (bitmap) {
                const uint32_t w0 = nucl >> 5;
                const uint32_t cand = lane > w0 ? myword : (lane == w0 ? (myword & (0xFFFFFFFFu << (nucl & 31u))) : 0u);
                const uint32_t bal = __ballot_sync(FULL, cand != 0u);
                if (bal) {
                    const uint32_t src = (uint32_t)__ffs((int)bal) - 1u;
                    e = 32u * src + (uint32_t)__ffs((int)__shfl_sync(FULL, cand, src)) - 1u;
                }
            } else {
                for (uint32_t base = nucl; base + 1u < L; base += 32u) {
                    const uint32_t pos = base + lane;
                    const uint32_t m = pos + 1u < L ? mask[pos] : 0u;
                    const uint32_t bal = __ballot_sync(FULL, m != 0u);
                    if (bal) { e = base + (uint32_t)__ffs((int)bal) - 1u; break; }
                }
            }
            // (b) the run in front of it comes straight from the original read
            const uint32_t want = e - nucl, have = L - k, run = want < have ? want : have;
            for (uint32_t i = lane; i < run; i += 32u) seq[nucl + i] = orig[k + i];
            nucl += run;
            k += run;
            if (run < want) { broke = true; break; }  // original exhausted before the loop ended
            if (nucl + 1u >= L) break;
        }
        // (c) the reference's loop body, position by position, while something can happen
        if (lane == 0) {
            do {
                uint32_t cur;
                if (sp) cur = stk[--sp]; else if (k < L) cur = orig[k++]; else { broke = true; break; }
                if (!is_acgt(cur)) { seq[nucl++] = (uint8_t)cur; continue; }  // ambiguous bases draw nothing (:190-192)
                const uint32_t tb = o * L + nucl;
                const uint32_t mk = mask[nucl];
                if (mk & 15u) {
                    for (uint32_t ln = 0; ln < 4; ++ln)
                        if ((mk >> ln) & 1u) {
                            const uint32_t ib = p.m.ins_base[tb * 4 + ln];
                            stk[sp++] = (uint8_t)ib;
                            if (p.store_mutations) emit_event(p, gp0 + pair, o, 0u, nucl, ln, cur, cur, ib, 255u);
                        }
                }
                if ((mk >> (4u + base_code(cur))) & 1u) {
                    uint32_t x;
                    if (sp) x = stk[--sp]; else if (k < L) x = orig[k++]; else { broke = true; break; }  // "ref" is read after the pop: IndexError -> no record
                    if (p.store_mutations) emit_event(p, gp0 + pair, o, 0u, nucl, 4u, x, (uint32_t)'.', 0u, 255u);
                    seq[nucl++] = (uint8_t)x;
                } else {
                    seq[nucl++] = (uint8_t)cur;
                }
            } while (nucl + 1u < L && (sp != 0u || mask[nucl] != 0u));
        }
        nucl = __shfl_sync(FULL, nucl, 0);
        k = __shfl_sync(FULL, k, 0);
        sp = __shfl_sync(FULL, sp, 0);
        broke = __shfl_sync(FULL, broke ? 1u : 0u, 0) != 0u;
    }
    // the part of the list behind the loop position, truncated to L (:136-139): LIFO first, then original bases
    uint32_t outlen = nucl;
    if (lane == 0)
        while (outlen < L && sp) seq[outlen++] = stk[--sp];
    outlen = __shfl_sync(FULL, outlen, 0);
    {
        const uint32_t want = L - outlen, have = L - k, run = want < have ? want : have;
        for (uint32_t i = lane; i < run; i += 32u) seq[outlen + i] = orig[k + i];
        outlen += run;
    }
    if (outlen < L) {  // extension from the reference sequence (:141-155)
        const uint32_t enc = u32p(l.sEnc)[pair];
        const uint64_t win = u64p(l.sWin)[rd], reclen = u64p(l.sRecLen)[pair];
        const uint64_t rec_abs = u64p(l.sWabs)[rd] - win;  // absolute index of the record's first base
        for (uint32_t e = lane; outlen + e < L; e += 32u) {
            uint32_t ch;
            if (o == 0) {
                const uint64_t g = win + L + e;  // read_end + i
                ch = (g >= reclen) ? (uint32_t)'A' : genome_ascii(p.g, enc, rec_abs + g);
            } else {
                if (win < (uint64_t)e + 1u) ch = (uint32_t)'A';  // read_start - 1 - i < 0
                else ch = complement_ascii(genome_ascii(p.g, enc, rec_abs + (win - 1u - e)));
            }
            seq[outlen + e] = (uint8_t)ch;
        }
    }
    __syncwarp();
}

}  // namespace

template <int kThreads, int kMinBlocks, int kQB, int kK>
__global__ void __launch_bounds__(kThreads, kMinBlocks) iss_generate_kernel(const __grid_constant__ GenParams p) {
    const uint32_t T = p.tile_pairs, tid = threadIdx.x, L = p.L;
    const Lay &l = p.lay;
    if (tid < ISS_QROW) u32p(l.sErrThr)[tid] = p.err_thr[tid];
    const uint32_t smem0 = smem_addr(iss_smem);
    const uint32_t mbar = smem0 + l.sGbar;                      // mbarrier of the TMA genome gather
    if (tid == 0) {
        u32p(l.sCtl)[2] = 0;
        if (p.tma_gather) mbar_init(mbar, 2u * T);  // every (pair, orientation) thread arrives once per tile
    }
    const uint32_t ntiles = (uint32_t)((p.total_pairs + T - 1) / T);  // a batch is far below 2^32 tiles
    // roles: q4 = 0..2S-1 -> orientation o = q4 / S, position part = q4 % S; rd = index of the read in the tile
    const uint32_t q4 = tid >> p.tile_shift, pair = tid & (T - 1u);
    const uint32_t S = 1u << p.split_shift;
    const uint32_t o = q4 >> p.split_shift, half = q4 & (S - 1u), rd = o * T + pair;

    // ---- quality-row pipeline.  Thread (o, part) walks the Philox QUAL blocks [b_begin, b_end) of its reads, and so do
    //      the other T - 1 threads of its position stream q4 -- for EVERY tile.  The rows a stream needs are therefore a
    //      fixed cyclic sequence: block b -> the rows of positions 2b, 2b + 1 for the populated bins of orientation o,
    //      contiguous in the model blob.  Two shared-memory stages per stream: blocks b_begin, b_begin + 2, ... go through
    //      stage 0 and b_begin + 1, b_begin + 3, ... through stage 1, for every tile alike (with an odd block count the
    //      last turn of stage 1 is simply skipped), so the loop is unrolled by two with static stage addresses.  The
    //      stream's first lane issues one cp.async.bulk per block, one block ahead of the consumers; `full` mbarriers carry
    //      the byte counts, `empty` ones (one arrival per warp of the stream) hand a stage back.  Phase parities: bit s of
    //      `qpar` = consumptions of stage s so far, mod 2; the producer's wait on `empty[s]` before its (c + 1)-th fill of
    //      stage s uses parity (c + 1) & 1, which passes at once on the fresh barrier (c = 0).
    const bool kde = p.qual_mode == ISS_QUAL_KDE_MODE;
    const uint32_t nblk = (L + 1u) >> 1;
    const uint32_t b_begin = (half * nblk) >> p.split_shift, b_end = ((half + 1u) * nblk) >> p.split_shift;
    const uint32_t bar0 = smem0 + l.sBar + (q4 << 5);            // full[0] +0, empty[0] +8, full[1] +16, empty[1] +24
    const uint32_t stage0 = l.sQst + q4 * (2u * p.q_stage_cap);  // offset of the stream's two stages from iss_smem
    uint32_t qpar = 0;
    if (kde) {
        if (tid == 0) {
            const uint32_t nbar = (2u << p.split_shift) * 2u;
            for (uint32_t k = 0; k < nbar; ++k) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem0 + l.sBar + 16u * k), "r"(1u) : "memory");
                asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem0 + l.sBar + 16u * k + 8u), "r"(T >> 5) : "memory");
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    // producer lane only: stage `stg` <- the rows of block `blk`, after every warp of the stream released the stage
    auto q_fill = [&](uint32_t stg, uint32_t blk, uint32_t done /* consumptions of the stage so far */) {
        // kQB blocks per stage (their rows are contiguous in the blob); the stream's last turn may hold fewer
        const uint32_t fb = bar0 + (stg << 4), bytes1 = p.q_stage_bytes[o];
        uint32_t bytes;
        if constexpr (kQB <= 2) bytes = (kQB == 1 || blk + 1u >= b_end) ? bytes1 : 2u * bytes1;
        else { const uint32_t left = b_end - blk; bytes = (left < (uint32_t)kQB ? left : (uint32_t)kQB) * bytes1; }
        mbar_wait(fb + 8u, (done + 1u) & 1u);
        mbar_arrive_expect_tx(fb, bytes);
        bulk_load(smem0 + stage0 + stg * p.q_stage_cap, p.q_rows[o] + (size_t)blk * bytes1, bytes, fb);
    };
    if (kde && pair == 0 && blockIdx.x < ntiles && b_begin < b_end) q_fill(0u, b_begin, 0u);

    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t gp0 = (uint64_t)tile * T;
        const uint32_t npairs = (uint32_t)((p.total_pairs - gp0 < (uint64_t)T) ? (p.total_pairs - gp0) : (uint64_t)T);
        uint32_t vflag = pair < npairs ? 1u : 0u;
        asm volatile("" : "+r"(vflag));  // one live register instead of the 64-bit arithmetic above re-done wherever `valid` is used
        const bool valid = vflag != 0u;

        // ---- phase 1: one thread per pair (generator.py:121-177, kde.py:72-78,97).  Nothing here touches the stream
        //      images, which the previous tile's copy-out may still be reading; its bulk stores are waited for by thread 0
        //      before the barrier that ends the phase.
        if (tid == blockDim.x - 1u) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the thread that issued them; returns at once when nothing is pending
        if (tid == 0) {
            u32p(l.sCtl)[0] = 0; u32p(l.sCtl)[1] = 0; u32p(l.sCtl)[4] = 0; u32p(l.sCtl)[6] = 0;
            u32p(l.sCtl)[2] ^= 1u;  // phase parity of the gather mbarrier for this tile (1, 0, 1, ... ; read as parity ^ 1)
        }
        if (tid < 2u * T) u32p(l.sFlag)[tid] = 0;
        if (tid < npairs) {
            const uint64_t gp = gp0 + tid;
            uint32_t lo = 0, hi = p.nseg;
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (p.segs[mid].pair_start <= gp) lo = mid; else hi = mid;
            }
            const DevSegment *sg = p.segs + lo;
            const uint64_t i = sg->i0 + (gp - sg->pair_start);
            const uint64_t P = sg->item_base + i;
            const uint32_t Plo = (uint32_t)P, Phi = (uint32_t)(P >> 32);
            const DevRecord *rec = p.g.recs + sg->rec;
            const int64_t len = (int64_t)rec->len;
            uint32_t att = 0;
            Philox4 b0;
            for (;;) {  // --gc_bias: re-simulate the pair until the accept draw passes (generator.py:82-92)
                b0 = philox4x32_10_rk(0, ISS_S_PAIR | (att << 8), Plo, Phi, p.rk);
                if (!p.gc_bias || b0.w < p.gc_thr) break;
                ++att;
            }
            const uint32_t a8 = att << 8;
            Philox4 b2;
            bool have_b2 = false;
            int64_t insert;
            if (p.frag_mode) {  // int(np.random.normal(mu, sd)) by inversion on the integer table
                b2 = philox4x32_10_rk(2, ISS_S_PAIR | a8, Plo, Phi, p.rk);
                have_b2 = true;
                const uint64_t r64 = u64of(b2.z, b2.w);
                uint32_t a = 0, b = p.n_frag;
                while (a < b) { const uint32_t mid = (a + b) >> 1; if (r64 >= p.m.frag_thr[mid]) a = mid + 1; else b = mid; }
                insert = p.frag_lo + (int64_t)a - 2 * (int64_t)L;
            } else if (p.n_insert == 0) {
                insert = p.fixed_insert;
            } else {  // np.searchsorted(i_size_cdf, u): the CDF index is the insert size (kde.py:97)
                uint32_t c = 0;  // coarse[k] = last threshold of the k-th run of ins_stride entries (constant bank)
#pragma unroll
                for (uint32_t s = 16; s >= 1; s >>= 1)
                    if (c + s <= 31u && b0.x > p.ins_coarse[c + s - 1u]) c += s;
                // inside the run: 9-way steps with 8 independent probes each (two load latencies for runs of <= 80 entries,
                // where a binary search has six or seven -- this thread's chain of loads is what the tile's other warps wait for)
                uint32_t a = c * p.ins_stride, n = a + p.ins_stride < p.n_insert ? p.ins_stride : p.n_insert - a;
                while (n > 8u) {
                    const uint32_t q = (n + 8u) / 9u;
                    uint32_t j = 0;
#pragma unroll
                    for (uint32_t k = 1; k <= 8u; ++k) {
                        const uint32_t at = k * q - 1u;
                        const uint32_t thr = at < n ? __ldg(p.m.insert_thr + a + at) : 0xFFFFFFFFu;
                        j += b0.x > thr ? 1u : 0u;
                    }
                    a += j * q;                      // thr[a - 1] < x, and x <= thr[a + q - 1] if that entry was probed
                    const uint32_t left = n - j * q, cap = j < 8u ? q - 1u : q;  // behind the last probe nothing bounds the rest
                    n = left < cap ? left : cap;
                }
                {
                    uint32_t j = 0;
#pragma unroll
                    for (uint32_t k = 0; k < 8u; ++k) {
                        const uint32_t thr = k < n ? __ldg(p.m.insert_thr + a + k) : 0xFFFFFFFFu;
                        j += b0.x > thr ? 1u : 0u;
                    }
                    a += j;
                }
                insert = (int64_t)a;
            }
            const int64_t frag = insert + 2 * (int64_t)L;
            const Philox4 b1 = philox4x32_10_rk(1, ISS_S_PAIR | a8, Plo, Phi, p.rk);
            int64_t fs;
            if (p.amplicon) fs = 0;
            else {
                const int64_t width = len - frag;
                if (width > 0) fs = (int64_t)__umul64hi(u64of(b1.x, b1.y), (uint64_t)width);
                else fs = (int64_t)__umul64hi(u64of(b1.z, b1.w), (uint64_t)(len - (int64_t)L));  // ValueError fallback
            }
            const int64_t fe = fs + (int64_t)L;
            int64_t rs, re;
            if (p.amplicon) { rs = len - (int64_t)L; re = len; }
            else { rs = fe + insert; re = rs + (int64_t)L; }
            if (re > len) {  // generator.py:172-176
                if (!have_b2) b2 = philox4x32_10_rk(2, ISS_S_PAIR | a8, Plo, Phi, p.rk);
                re = (int64_t)L + (int64_t)__umul64hi(u64of(b2.x, b2.y), (uint64_t)(len - (int64_t)L));
                rs = re - (int64_t)L;
            }
            if (rs < 0 || fe > len) { atomicOr(p.counters + 2, (unsigned long long)ISS_DEVERR_RANGE); rs = 0; if (fe > len) fs = 0; }
            uint32_t binf = 0, binr = 0;
            if (kde) {  // np.random.choice(range(4), p=norm_mean)
                for (int j = 0; j < 4; ++j) {
                    binf += ((uint64_t)b0.y >= p.bin_thr[0][j]) ? 1u : 0u;
                    binr += ((uint64_t)b0.z >= p.bin_thr[1][j]) ? 1u : 0u;
                }
                if (binf == 4) binf = 3;
                if (binr == 4) binr = 3;
                if (!p.bin_valid[0][binf] || !p.bin_valid[1][binr]) atomicOr(p.counters + 2, (unsigned long long)ISS_DEVERR_QUALBIN);
            }
            u32p(l.sQrow)[tid] = p.q_row_off[0][binf] | (binf << 16);
            u32p(l.sQrow)[T + tid] = p.q_row_off[1][binr] | (binr << 16);
            u64p(l.sP)[tid] = P;
            u32p(l.sSeg)[tid] = lo;
            u32p(l.sAtt)[tid] = att;
            u64p(l.sRecLen)[tid] = (uint64_t)len;
            u64p(l.sWin)[tid] = (uint64_t)fs;
            u64p(l.sWin)[T + tid] = (uint64_t)rs;
            u64p(l.sWabs)[tid] = rec->data_off + (uint64_t)fs;
            u64p(l.sWabs)[T + tid] = rec->data_off + (uint64_t)rs;
            u32p(l.sEnc)[tid] = rec->enc;
            const uint64_t off = sg->out_base + iss_cf(i, sg->fixed_len);
            const uint32_t nd = iss_digits(i);
            u64p(l.sOff)[tid] = off;
            u32p(l.sRel)[tid] = (uint32_t)off + rec->name_len + nd + sg->cpu_digits + 6u;  // "@" name "_" i "_" cpu "/1\n"
            if (tid == npairs - 1) u32p(l.sCtl)[3] = sg->fixed_len + nd;  // bytes of the last record
        }
        __syncthreads();
        // image offset of a byte = rel0 + low 32 bits of its stream offset; the tile's first record starts at a0 = g0 & 15
        const uint32_t rel0 = ((uint32_t)u64p(l.sOff)[0] & 15u) - (uint32_t)u64p(l.sOff)[0];

        // ---- phase 2a: record skeleton, one thread per read:  "@{name}_{i}_{cpu}/{1|2}\n" ... "\n+\n" ... "\n"
        if (valid && half == 0) {
            const DevSegment *sg = p.segs + u32p(l.sSeg)[pair];
            uint8_t *img = imgp(l, o) + (uint32_t)(rel0 + (uint32_t)u64p(l.sOff)[pair]);  // wraps in 32 bits by construction
            const uint64_t i = u64p(l.sP)[pair] - sg->item_base;
            const DevRecord *rec = p.g.recs + sg->rec;
            uint32_t w = 0;
            img[w++] = '@';
            const char *nm = p.g.names + rec->name_off;
            for (uint32_t k = 0; k < rec->name_len; ++k) img[w++] = (uint8_t)nm[k];
            img[w++] = '_';
            {
                const uint32_t nd = iss_digits(i);
                uint64_t v = i;
                for (int d = (int)nd - 1; d >= 0; --d) { img[w + d] = (uint8_t)('0' + (uint32_t)(v % 10u)); v /= 10u; }
                w += nd;
            }
            img[w++] = '_';
            {
                uint32_t v = sg->cpu;
                const uint32_t nd = sg->cpu_digits;
                for (int d = (int)nd - 1; d >= 0; --d) { img[w + d] = (uint8_t)('0' + v % 10u); v /= 10u; }
                w += nd;
            }
            img[w++] = '/';
            img[w++] = (uint8_t)('1' + o);
            img[w++] = '\n';
            img[w + L] = '\n';
            img[w + L + 1] = '+';
            img[w + L + 2] = '\n';
            img[w + 2 * L + 3] = '\n';
            if (p.tma_gather) {
                if (rec->enc == ISS_ENC_2BIT) {  // TMA: packed window -> the record's quality line (overwritten in phase 4)
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive_expect_tx(mbar, p.stage_bytes);
                    bulk_load(smem0 + stage_dst_off(l.img[o], rel0 + u32p(l.sRel)[pair], L),
                              reinterpret_cast<const uint8_t *>(p.g.packed) + stage_src_byte(u64p(l.sWabs)[rd]), p.stage_bytes, mbar);
                } else {
                    mbar_arrive(mbar);
                }
            }
        } else if (p.tma_gather && half == 0) {
            mbar_arrive(mbar);  // lanes of a partial tile keep the arrival count constant
        }
        if (p.tma_gather) mbar_wait(mbar, u32p(l.sCtl)[2] ^ 1u);

        // ---- phase 2b: gather the windows, flattened over (read, 16-base chunk).  The sequence lines are disjoint from
        //      the skeleton bytes, so no barrier separates the two.
        {
            const uint32_t nitems = 2u * T * p.nchunk;
            for (uint32_t idx = tid; idx < nitems; idx += blockDim.x) {
                const uint32_t grd = p.nchunk == 1u ? idx : __umulhi(idx, p.chunk_magic);  // ceil(2^32 / 1) does not fit 32 bits
                const uint32_t j = idx - grd * p.nchunk;
                const uint32_t ro = grd >= T ? 1u : 0u, rp = grd - ro * T;
                if (rp >= npairs) continue;
                const uint32_t so = rel0 + u32p(l.sRel)[rp];
                uint8_t *seq = imgp(l, ro) + so;
                const uint64_t wabs = u64p(l.sWabs)[grd];
                if (u32p(l.sEnc)[rp] == ISS_ENC_2BIT) {
                    const uint64_t q = ro ? (wabs + L - 16u * j - 16u) : (wabs + 16u * j);
                    const uint32_t sh = ((uint32_t)q & 15u) * 2u;
                    uint32_t w0, w1;
                    if (p.tma_gather) {  // words of the span staged by TMA in the record's quality line
                        const uint32_t *st = u32p(stage_dst_off(l.img[ro], so, L));
                        const uint32_t wi = (uint32_t)((q >> 4) - (stage_src_byte(wabs) >> 2));
                        w0 = st[wi];
                        w1 = st[wi + 1];
                    } else {
                        const uint64_t wi = q >> 4;  // 64-bit word index: packed genomes beyond 2^36 bases stay addressable
                        w0 = __ldg(p.g.packed + wi);
                        w1 = __ldg(p.g.packed + wi + 1);
                    }
                    uint32_t w = __funnelshift_r(w0, w1, sh);
                    if (ro) {  // reverse complement of the 16 bases: complement = code ^ 2, then the 2-bit fields in reverse order
                        w = __brev(w ^ 0xAAAAAAAAu);
                        w = ((w & 0xAAAAAAAAu) >> 1) | ((w & 0x55555555u) << 1);
                    }
                    // base m of w -> position 16 j + m of the read, for both strands.  2-bit codes -> PRMT selector nibbles
                    // (8 bases per 32-bit selector pair) -> four ASCII words -> byte stores (the line has no alignment).
                    const uint32_t s0 = spread_codes(w & 0xFFFFu), s1 = spread_codes(w >> 16);
                    const uint32_t W0 = __byte_perm(0x47544341u, 0u, s0), W1 = __byte_perm(0x47544341u, 0u, s0 >> 16);
                    const uint32_t W2 = __byte_perm(0x47544341u, 0u, s1), W3 = __byte_perm(0x47544341u, 0u, s1 >> 16);
                    const uint32_t dst = smem0 + l.img[ro] + so + 16u * j;  // shared-window address of position 16 j
                    const uint32_t n = L - 16u * j;                        // bases of this chunk that belong to the read (>= 1)
                    if (n >= 16u) {
                        sts_word_bytes<0>(dst, W0); sts_word_bytes<4>(dst, W1); sts_word_bytes<8>(dst, W2); sts_word_bytes<12>(dst, W3);
                    } else {
                        sts_word_bytes_n<0>(dst, W0, n); sts_word_bytes_n<4>(dst, W1, n); sts_word_bytes_n<8>(dst, W2, n); sts_word_bytes_n<12>(dst, W3, n);
                    }
                } else {
#pragma unroll 4
                    for (uint32_t m = 0; m < 16; ++m) {
                        const uint32_t k = 16u * j + m;
                        if (k < L) {
                            if (ro) seq[k] = (uint8_t)complement_ascii(__ldg(p.g.ascii + wabs + (L - 1u - k)));
                            else seq[k] = __ldg(p.g.ascii + wabs + k);
                        }
                    }
                }
            }
        }
        if (p.tma_gather) __syncthreads();  // block-uniform: the staged windows live in the quality lines that phase 4 writes

        // ---- phase 3: indel scan in the no-event world (error_models/__init__.py:187-224).  The pair's combined list
        //      of non-zero positions (forward then reverse) is cut in equal parts, one per thread of the pair.  Whether a
        //      site that passes its draw really fires depends on the base under it, which another thread may still be
        //      gathering: such survivors (rare) are queued and resolved behind the barrier that ends this phase.
        const uint64_t P = valid ? u64p(l.sP)[pair] : 0;
        const uint32_t Plo = (uint32_t)P, Phi = (uint32_t)(P >> 32);
        const uint32_t att8 = valid ? (u32p(l.sAtt)[pair] << 8) : 0;
        const uint32_t seq_off = rel0 + (valid ? u32p(l.sRel)[pair] : 0);
        uint8_t *seq = imgp(l, o) + seq_off;
        const uint32_t g_begin = (q4 * p.n_grp) >> (p.split_shift + 1u), g_end = ((q4 + 1u) * p.n_grp) >> (p.split_shift + 1u);
        if (p.surv_direct) {  // dense indel tables: the queue would overflow in every tile -- scan once, behind the barrier
            if (tid == 0 && p.n_grp) u32p(l.sCtl)[6] = ISS_SURV_CAP + 1u;
        } else if (valid && p.n_grp) {
            // one loop per list (insertion fwd, deletion fwd, insertion rev, deletion rev): the stream id and the rank of a
            // group inside its list are loop invariants, and a group's meta words are fetched only for a survivor
#pragma unroll 1
            for (uint32_t li = 0; li < 4u; ++li) {
                const uint32_t lb = p.grp_begin[li], le = p.grp_begin[li + 1u];
                const uint32_t ga = g_begin > lb ? g_begin : lb, gb = g_end < le ? g_end : le;
                const uint32_t so = li >> 1, kind = li & 1u;
                const uint32_t stream = iss_read_stream(so, kind ? ISS_K_DEL : ISS_K_INS) | att8;
                for (uint32_t g = ga; g < gb; ++g) {
                    const uint4 t = __ldg(p.m.ent + 2u * g);
                    const Philox4 w = philox4x32_10_rk(g - lb, stream, Plo, Phi, p.rk);
                    if ((w.x <= t.x) | (w.y <= t.y) | (w.z <= t.z) | (w.w <= t.w)) {  // rare: some site may fire
                        const uint4 mt = __ldg(p.m.ent + 2u * g + 1u);
#pragma unroll 1
                        for (uint32_t r = 0; r < 4; ++r) {
                            const uint32_t meta = pick(mt, r);
                            const uint32_t bits = site_bits(p, so, kind, pick(w, r), pick(t, r), meta);
                            if (!bits) continue;
                            const uint32_t slot = atomicAdd(&u32p(l.sCtl)[6], 1u);
                            if (slot < ISS_SURV_CAP) u32p(l.sSurv)[slot] = ((so * T + pair) << 24) | (bits << 16) | (meta & 0xFFFFu);
                        }
                    }
                }
            }
        }

        // ---- phase 4: phred scores + substitution tests.  Thread (o, part) walks its part of the read's Philox QUAL
        //      blocks; the 32 lanes of a warp are 32 pairs at the same position, reading <= 4 staged quality rows.
        {
            uint8_t *qual = seq + L + 3;
            const uint32_t sQ = iss_read_stream(o, ISS_K_QUAL) | att8;
            if (kde) {
                const uint32_t row0 = smem0 + stage0 + (valid ? (u32p(l.sQrow)[rd] & 0xFFFFu) : 0u);  // shared-window address of the read's row in stage 0, first position
                const uint32_t qual_s = smem0 + (o ? l.img[1] : l.img[0]) + seq_off + L + 3u;  // the quality line, and the
                const uint32_t thr_s = smem0 + l.sErrThr;                                       //   error thresholds by phred
                const bool producer = pair == 0;
                const bool more_tiles = tile + gridDim.x < ntiles;
                // one block: wait for its rows, draw + search two positions, hand the stage back
                // one block: draw + search two positions against the rows at r0
                auto q_block = [&](uint32_t b, uint32_t r0) {
                    const uint32_t pos = 2u * b;
                    const Philox4 w = philox4x32_10_rk(b, sQ, Plo, Phi, p.rk);
                    bool tie0, tie1;
                    uint32_t q0, q1;  // the row behind an odd L is all pad: q1 = 0 there, and not used
                    if (kK == 8) q8_search2(r0, w.x >> 16, r0 + p.q_pos_stride[o], w.z >> 16, q0, q1, tie0, tie1);
                    else q16_search2(r0, w.x >> 16, r0 + p.q_pos_stride[o], w.z >> 16, q0, q1, tie0, tie1);
                    if (tie0) {
                        defer_position(l, p.tie_cap, rd, pos, q0);
                    } else {
                        sts_u8(qual_s + pos, q0 + 33u);
                        if (w.y > lds_u32(thr_s + 4u * q0)) {
                            const uint32_t slot = atomicAdd(&u32p(l.sCtl)[0], 1u);
                            if (slot < p.err_cap) u32p(l.sErr)[slot] = (rd << 16) | pos; else err_overflow(p, slot, (rd << 16) | pos);
                        }
                    }
                    if (pos + 1u < L) {
                        if (tie1) {
                            defer_position(l, p.tie_cap, rd, pos + 1u, q1);
                        } else {
                            sts_u8(qual_s + pos + 1u, q1 + 33u);
                            if (w.w > lds_u32(thr_s + 4u * q1)) {
                                const uint32_t slot = atomicAdd(&u32p(l.sCtl)[0], 1u);
                                if (slot < p.err_cap) u32p(l.sErr)[slot] = (rd << 16) | (pos + 1u); else err_overflow(p, slot, (rd << 16) | (pos + 1u));
                            }
                        }
                    }
                };
                // two full blocks of a stage at once: four searches stepped together
                auto q_emit = [&](uint32_t pos, uint32_t q, bool tie, uint32_t werr) {
                    if (tie) {
                        defer_position(l, p.tie_cap, rd, pos, q);
                    } else {
                        sts_u8(qual_s + pos, q + 33u);
                        if (werr > lds_u32(thr_s + 4u * q)) {
                            const uint32_t slot = atomicAdd(&u32p(l.sCtl)[0], 1u);
                            if (slot < p.err_cap) u32p(l.sErr)[slot] = (rd << 16) | pos; else err_overflow(p, slot, (rd << 16) | pos);
                        }
                    }
                };
                auto q_block2 = [&](uint32_t b, uint32_t r0) {
                    const uint32_t pos = 2u * b;
                    const Philox4 w0 = philox4x32_10_rk(b, sQ, Plo, Phi, p.rk);
                    const Philox4 w1 = philox4x32_10_rk(b + 1u, sQ, Plo, Phi, p.rk);
                    const uint32_t r1 = r0 + p.q_stage_bytes[o], ps = p.q_pos_stride[o];
                    bool t0, t1, t2, t3;
                    uint32_t q0, q1, q2, q3;
                    if (kK == 8) q8_search4(r0, w0.x >> 16, r0 + ps, w0.z >> 16, r1, w1.x >> 16, r1 + ps, w1.z >> 16, q0, q1, q2, q3, t0, t1, t2, t3);
                    else q16_search4(r0, w0.x >> 16, r0 + ps, w0.z >> 16, r1, w1.x >> 16, r1 + ps, w1.z >> 16, q0, q1, q2, q3, t0, t1, t2, t3);
                    q_emit(pos, q0, t0, w0.y);
                    q_emit(pos + 1u, q1, t1, w0.w);
                    q_emit(pos + 2u, q2, t2, w1.y);
                    if (pos + 3u < L) q_emit(pos + 3u, q3, t3, w1.w);
                };
                // one turn of a stage: wait for its rows, its kQB blocks, hand the stage back
                auto q_stage = [&](uint32_t b, uint32_t stg) {
                    mbar_wait(bar0 + (stg << 4), (qpar >> stg) & 1u);
                    qpar ^= 1u << stg;
                    if (valid) {
                        const uint32_t r0 = row0 + stg * p.q_stage_cap;
                        if constexpr (kQB == 2) {
                            if (b + 1u < b_end) q_block2(b, r0); else q_block(b, r0);
                        } else if constexpr (kQB == 1) {
                            q_block(b, r0);
                        } else if constexpr (kQB > 2) {
#pragma unroll
                            for (uint32_t j = 1; j < (uint32_t)kQB; ++j)
                                if (b + j < b_end) q_block(b + j, r0 + j * p.q_stage_bytes[o]);
                        }
                    }
                    __syncwarp();
                    if ((tid & 31u) == 0) mbar_arrive(bar0 + (stg << 4) + 8u);  // this warp is done with the stage
                };
                for (uint32_t b = b_begin; b < b_end; b += 2u * kQB) {
                    const bool second = b + kQB < b_end;
                    if (producer && second) q_fill(1u, b + kQB, (qpar >> 1) & 1u);
                    q_stage(b, 0u);
                    if (producer) {  // stage 0 for the next turn: the next blocks, or the first ones of the next tile
                        if (b + 2u * kQB < b_end) q_fill(0u, b + 2u * kQB, qpar & 1u);
                        else if (more_tiles) q_fill(0u, b_begin, qpar & 1u);
                    }
                    if (second) q_stage(b + kQB, 1u);
                }
            } else if (valid) {
                for (uint32_t b = b_begin; b < b_end; ++b) {
                    const Philox4 w = philox4x32_10_rk(b, sQ, Plo, Phi, p.rk);
#pragma unroll
                    for (uint32_t h = 0; h < 2; ++h) {
                        const uint32_t pos = 2u * b + h;
                        if (pos >= L) break;
                        const uint32_t wq = h ? w.z : w.x, wt = h ? w.w : w.y;
                        uint32_t q;
                        if (p.qual_mode == ISS_QUAL_CONST_MODE) q = p.const_phred;
                        else {
                            uint32_t a = 0, e = p.n_basic;
                            while (a < e) { const uint32_t mid = (a + e) >> 1; if (wq >= __ldg(p.m.basic_thr + mid)) a = mid + 1; else e = mid; }
                            q = p.m.basic_val[a];
                        }
                        qual[pos] = (uint8_t)(q + 33u);
                        if (wt > u32p(l.sErrThr)[q]) {
                            const uint32_t slot = atomicAdd(&u32p(l.sCtl)[0], 1u);
                            if (slot < p.err_cap) u32p(l.sErr)[slot] = (rd << 16) | pos; else err_overflow(p, slot, (rd << 16) | pos);
                        }
                    }
                }
            }
        }
        __syncthreads();

        // ---- phase 3b: which reads have an indel event?  (block-uniform; no survivor at all in most tiles of the bundled
        //      models).  A survivor fires if the base under it is unambiguous and -- for a deletion -- is one of the bases
        //      whose threshold its draw passed.  An overflowing queue (dense, fitted or adversarial tables) falls back to
        //      every thread re-walking its groups with the gathered bases in place.
        {
            const uint32_t nsurv = u32p(l.sCtl)[6];
            if (nsurv) {
                if (nsurv <= ISS_SURV_CAP) {
                    for (uint32_t e = tid; e < nsurv; e += blockDim.x) {
                        const uint32_t v = u32p(l.sSurv)[e];
                        const uint32_t srd = v >> 24, bits = (v >> 16) & 0xFFu, so = srd >= T ? 1u : 0u, sp_ = srd - so * T;
                        const uint32_t c = (imgp(l, so) + (uint32_t)(rel0 + u32p(l.sRel)[sp_]))[v & 0xFFFFu];
                        if (!is_acgt(c)) continue;  // ambiguous bases draw nothing (:190-192)
                        if (((bits & 15u) || ((bits >> (4u + base_code(c))) & 1u)) && atomicExch(&u32p(l.sFlag)[srd], 1u) == 0u)
                            u32p(l.sSlow)[atomicAdd(&u32p(l.sCtl)[1], 1u)] = srd;
                    }
                } else if (valid) {
                    uint32_t fired = 0;  // bit so set when read orientation so has an event
                    for (uint32_t g = g_begin; g < g_end; ++g) {
                        const uint32_t li = (g >= p.grp_begin[1] ? 1u : 0u) + (g >= p.grp_begin[2] ? 1u : 0u) + (g >= p.grp_begin[3] ? 1u : 0u);
                        const uint32_t so = li >> 1, kind = li & 1u;
                        const uint4 t = __ldg(p.m.ent + 2u * g), mt = __ldg(p.m.ent + 2u * g + 1u);
                        const Philox4 w = philox4x32_10_rk(g - p.grp_begin[li], iss_read_stream(so, kind ? ISS_K_DEL : ISS_K_INS) | att8, Plo, Phi, p.rk);
                        if ((w.x <= t.x) | (w.y <= t.y) | (w.z <= t.z) | (w.w <= t.w)) {
                            const uint8_t *sq = imgp(l, so) + seq_off;
#pragma unroll 1
                            for (uint32_t r = 0; r < 4; ++r) {
                                const uint32_t meta = pick(mt, r);
                                const uint32_t bits = site_bits(p, so, kind, pick(w, r), pick(t, r), meta);
                                if (!bits) continue;
                                const uint32_t c = sq[meta & 0xFFFFu];
                                if (!is_acgt(c)) continue;
                                if ((bits & 15u) || ((bits >> (4u + base_code(c))) & 1u)) fired |= 1u << so;
                            }
                        }
                    }
                    for (uint32_t so = 0; so < 2; ++so)
                        if (((fired >> so) & 1u) && atomicExch(&u32p(l.sFlag)[so * T + pair], 1u) == 0u)
                            u32p(l.sSlow)[atomicAdd(&u32p(l.sCtl)[1], 1u)] = so * T + pair;
                }
                __syncthreads();
                const uint32_t nslow = u32p(l.sCtl)[1];
                if (nslow) {
                    // phase 3c: all indel draws of the queued reads, in parallel.  A draw is addressed by (read, position, lane)
                    // and does not depend on the evolving sequence -- only WHICH deletion threshold applies depends on the
                    // current base -- so the mask holds the insertion lanes that fire and, per base, whether a deletion
                    // fires.  The masks live in the slot's scratch (global memory, behind the replay's own buffers).
                    const uint32_t n_o0 = p.grp_begin[2] - p.grp_begin[0], n_o1 = p.grp_begin[4] - p.grp_begin[2];
                    const uint32_t ngrp = (n_o0 > n_o1) ? n_o0 : n_o1;  // groups of 4 sites (one Philox block each) per read
                    const uint32_t mask_words = (L + 3u) >> 2;
                    for (uint32_t base = 0; base < nslow; base += ISS_SLOW_SLOTS) {
                        const uint32_t nround = (nslow - base < ISS_SLOW_SLOTS) ? (nslow - base) : ISS_SLOW_SLOTS;
                        uint8_t *cta_scratch = p.scratch + (size_t)blockIdx.x * ISS_SLOW_SLOTS * p.scratch_stride;
                        for (uint32_t i = tid; i < nround * mask_words; i += blockDim.x) {
                            const uint32_t sidx = i / mask_words;
                            reinterpret_cast<uint32_t *>(cta_scratch + (size_t)sidx * p.scratch_stride + p.scratch_mask_off)[i - sidx * mask_words] = 0;
                        }
                        __syncthreads();
                        for (uint32_t idx = tid; idx < nround * ngrp; idx += blockDim.x) {
                            const uint32_t sidx = idx / ngrp, gq = idx - sidx * ngrp;
                            const uint32_t srd = u32p(l.sSlow)[base + sidx];
                            const uint32_t so = srd >= T ? 1u : 0u, sp_ = srd - so * T;
                            const uint32_t g = p.grp_begin[2u * so] + gq;
                            if (g >= p.grp_begin[2u * so + 2u]) continue;
                            const uint32_t kind = g >= p.grp_begin[2u * so + 1u] ? 1u : 0u;
                            const uint64_t sP = u64p(l.sP)[sp_];
                            const uint4 t = __ldg(p.m.ent + 2u * g), mt = __ldg(p.m.ent + 2u * g + 1u);
                            const Philox4 w = philox4x32_10_rk(g - p.grp_begin[2u * so + kind], iss_read_stream(so, kind ? ISS_K_DEL : ISS_K_INS) | (u32p(l.sAtt)[sp_] << 8),
                                                               (uint32_t)sP, (uint32_t)(sP >> 32), p.rk);
                            uint32_t *masks32 = reinterpret_cast<uint32_t *>(cta_scratch + (size_t)sidx * p.scratch_stride + p.scratch_mask_off);
#pragma unroll 1
                            for (uint32_t r = 0; r < 4; ++r) {
                                const uint32_t meta = pick(mt, r);
                                const uint32_t bits = site_bits(p, so, kind, pick(w, r), pick(t, r), meta);
                                if (bits) {
                                    const uint32_t bi = meta & 0xFFFFu;
                                    atomicOr(&masks32[bi >> 2], bits << (8u * (bi & 3u)));
                                }
                            }
                        }
                        __syncthreads();
                        for (uint32_t sidx = tid >> 5; sidx < nround; sidx += blockDim.x >> 5) {  // one warp per read
                            uint8_t *slot = cta_scratch + (size_t)sidx * p.scratch_stride;
                            replay_indels_warp(p, l, gp0, rel0, u32p(l.sSlow)[base + sidx], slot, slot + p.scratch_mask_off, tid & 31u);
                        }
                        __syncthreads();
                    }
                    if (tid == 0) atomicAdd(p.counters + 0, (unsigned long long)nslow);
                }
            }
        }

        // ---- phase 4b: the exact pass.  Positions whose 16-bit search met a tie are evaluated on the 32-bit row: the draw is
        //      recomputed (a Philox word is a pure function of its address), the phred written, the error test made.  If the
        //      queue or the substitution list overflowed (models far outside the bundled ones: ~ every base an error), every
        //      thread walks its positions again on the exact tables and applies its substitutions directly.
        const uint32_t ntie = u32p(l.sCtl)[4];
        const uint32_t nerr_fast = u32p(l.sCtl)[0];
        const bool redo_all = ntie > p.tie_cap || nerr_fast > p.err_cap + p.err_ovf_cap;  // block-uniform
        if (ntie && !redo_all) {
            for (uint32_t e = blockDim.x - 1u - tid; e < ntie; e += blockDim.x) {  // from the CTA's last threads: beside the substitutions, which start at thread 0
                const uint32_t v = u32p(l.sTie)[e];
                const uint32_t trd = v >> 24, pos = v & 0xFFFFu;
                const uint32_t to = trd >= T ? 1u : 0u, tp = trd - to * T;
                const uint64_t tP = u64p(l.sP)[tp];
                const Philox4 w = philox4x32_10_rk(pos >> 1, iss_read_stream(to, ISS_K_QUAL) | (u32p(l.sAtt)[tp] << 8), (uint32_t)tP,
                                                   (uint32_t)(tP >> 32), p.rk);
                const uint32_t q = quality_exact(p, trd, pos, (pos & 1u) ? w.z : w.x, (v >> 16) & 0xFFu);
                (imgp(l, to) + (uint32_t)(rel0 + u32p(l.sRel)[tp]) + L + 3u)[pos] = (uint8_t)(q + 33u);
                if (((pos & 1u) ? w.w : w.y) > u32p(l.sErrThr)[q]) {  // a handful per tile: applied in place, no second list
                    apply_substitution(p, l, gp0, rel0, trd, pos);
                    atomicAdd(p.counters + 1, 1ull);
                }
            }
        }

        // ---- phase 5: substitutions, compacted over the whole tile (error_models/__init__.py:95-109)
        if (!redo_all) {
            const uint32_t nerr = nerr_fast;
            const uint32_t *ovf = p.err_ovf + (size_t)blockIdx.x * p.err_ovf_cap - p.err_cap;  // entries err_cap .. nerr - 1
            for (uint32_t e = tid; e < nerr; e += blockDim.x) {
                const uint32_t v = e < p.err_cap ? u32p(l.sErr)[e] : ovf[e];
                apply_substitution(p, l, gp0, rel0, v >> 16, v & 0xFFFFu);
            }
            if (tid == 0 && nerr) atomicAdd(p.counters + 1, (unsigned long long)nerr);
        } else if (valid) {
            uint8_t *qual = seq + L + 3;
            const uint32_t sQ = iss_read_stream(o, ISS_K_QUAL) | att8;
            uint32_t napplied = 0;
#pragma unroll 1
            for (uint32_t b = b_begin; b < b_end; ++b) {
                const Philox4 w = philox4x32_10_rk(b, sQ, Plo, Phi, p.rk);
#pragma unroll 1
                for (uint32_t h = 0; h < 2; ++h) {
                    const uint32_t pos = 2u * b + h;
                    if (pos >= L) break;
                    const uint32_t q = quality_exact(p, rd, pos, h ? w.z : w.x, 0u);
                    qual[pos] = (uint8_t)(q + 33u);
                    if ((h ? w.w : w.y) > u32p(l.sErrThr)[q]) { apply_substitution(p, l, gp0, rel0, rd, pos); ++napplied; }
                }
            }
            if (napplied) atomicAdd(p.counters + 1, (unsigned long long)napplied);
        }
        // The copy-out geometry is read BEFORE the barrier (it dates from phase 1): behind it no thread reads the per-pair
        // arrays any more, so the next tile's phase 1 -- which no barrier separates from this copy-out -- may overwrite them.
        const uint64_t g0 = u64p(l.sOff)[0];
        const uint32_t tile_bytes = (uint32_t)(u64p(l.sOff)[npairs - 1] - g0) + u32p(l.sCtl)[3];
        // generic-proxy writes to shared memory must be visible to the async (TMA) proxy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();

        // ---- phase 6: copy-out.  [g0, g0 + tile_bytes) of each stream; the 16 B aligned middle goes through the
        //      bulk-copy engine, the ragged head/tail (< 16 B each) through plain byte stores.
        {
            const uint32_t a0 = (uint32_t)(g0 & 15u);
            uint32_t head = (16u - a0) & 15u;
            if (head > tile_bytes) head = tile_bytes;
            const uint32_t mid = (tile_bytes - head) & ~15u;
            const uint32_t tail = tile_bytes - head - mid;
            const uint32_t edge = head + tail;
            for (uint32_t idx = tid; idx < 2u * edge; idx += blockDim.x) {
                const uint32_t st = idx >= edge ? 1u : 0u;
                const uint32_t e = idx - st * edge;
                const uint32_t byte = e < head ? e : (head + mid + (e - head));
                (st ? p.out2 : p.out1)[g0 + byte] = imgp(l, st)[a0 + byte];
            }
#ifndef ISS_NO_BULK_STORE
            if (tid == blockDim.x - 1u && mid) {  // a thread without a pair in phase 1: its wait for these stores runs beside that phase
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p.out1 + g0 + head),
                             "r"(smem_addr(imgp(l, 0) + a0 + head)), "r"(mid)
                             : "memory");
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p.out2 + g0 + head),
                             "r"(smem_addr(imgp(l, 1) + a0 + head)), "r"(mid)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
#else
            for (uint32_t idx = tid; idx < 2u * (mid >> 4); idx += blockDim.x) {
                const uint32_t st = idx >= (mid >> 4) ? 1u : 0u;
                const uint32_t e = idx - st * (mid >> 4);
                const uint4 v = *reinterpret_cast<const uint4 *>(imgp(l, st) + a0 + head + 16u * e);
                *reinterpret_cast<uint4 *>((st ? p.out2 : p.out1) + g0 + head + 16u * e) = v;
            }
#endif
        }
        // the next tile's phase-1 barrier orders these reads before any new write to the images
    }
    if (tid == blockDim.x - 1u) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void iss_philox_kernel(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out) {
    const Philox4 w = philox4x32_10(c0, c1, c2, c3, k0, k1);
    out[0] = w.x; out[1] = w.y; out[2] = w.z; out[3] = w.w;
}

// ---- host-visible launch helpers --------------------------------------------------------------------
size_t iss_tile_smem_bytes(const Lay &l) { return l.total; }

// Builds of the one kernel template <threads, CTAs/SM, Philox blocks per quality-row stage, entries per staged row>:
// 256 x 4 (64-pair tiles), 512 x 2 with one block per stage (long reads, each read split over 4 threads) and 512 x 2 with two
// blocks per stage (128-pair tiles of short reads); the 512-thread ones also for short (8-entry) rows.  (Four blocks per
// stage, which the short rows have room for, measured 6 % slower than two: 7.89e8 against 8.41e8 pairs/s on configs[2] --
// the loop body is unrolled per block and stage, and the kernel is 120 KB of code as it is.)
template <typename F>
static cudaError_t for_each_build(F f) {
    cudaError_t e = f(iss_generate_kernel<256, 4, 1, 64>);
    if (e != cudaSuccess) return e;
    e = f(iss_generate_kernel<512, 2, 1, 64>);
    if (e != cudaSuccess) return e;
    e = f(iss_generate_kernel<512, 2, 2, 64>);
    if (e != cudaSuccess) return e;
    e = f(iss_generate_kernel<512, 2, 1, 8>);
    if (e != cudaSuccess) return e;
    return f(iss_generate_kernel<512, 2, 2, 8>);
}

cudaError_t iss_kernel_configure(size_t max_smem) {  // once per device
    return for_each_build([&](auto k) { return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem); });
}

cudaError_t iss_kernel_carveout(int percent) {  // shared-memory share of the SM's unified L1/shared array
    return for_each_build([&](auto k) { return cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, percent); });
}

bool iss_has_short_row_build(uint32_t T, uint32_t split_shift) { return (2u << split_shift) * T > 256u; }

cudaError_t iss_launch_generate(const GenParams &p, uint32_t grid, size_t smem, cudaStream_t stream) {
    const uint32_t threads = (2u << p.split_shift) * p.tile_pairs;
    if (threads > 256 && p.q_blocks == 2) {
        if (p.q_short) iss_generate_kernel<512, 2, 2, 8><<<grid, threads, smem, stream>>>(p);
        else iss_generate_kernel<512, 2, 2, 64><<<grid, threads, smem, stream>>>(p);
    } else if (threads > 256) {
        if (p.q_short) iss_generate_kernel<512, 2, 1, 8><<<grid, threads, smem, stream>>>(p);
        else iss_generate_kernel<512, 2, 1, 64><<<grid, threads, smem, stream>>>(p);
    } else {
        iss_generate_kernel<256, 4, 1, 64><<<grid, threads, smem, stream>>>(p);
    }
    return cudaGetLastError();
}

cudaError_t iss_occupancy(uint32_t T, uint32_t split_shift, uint32_t q_blocks, uint32_t q_short, size_t smem, int *blocks_per_sm) {
    const int threads = (int)((2u << split_shift) * T);
    if (threads > 256 && q_blocks == 2)
        return q_short ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, iss_generate_kernel<512, 2, 2, 8>, threads, smem)
                       : cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, iss_generate_kernel<512, 2, 2, 64>, threads, smem);
    if (threads > 256)
        return q_short ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, iss_generate_kernel<512, 2, 1, 8>, threads, smem)
                       : cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, iss_generate_kernel<512, 2, 1, 64>, threads, smem);
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, iss_generate_kernel<256, 4, 1, 64>, threads, smem);
}

cudaError_t iss_launch_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t *dev_out, cudaStream_t stream) {
    iss_philox_kernel<<<1, 1, 0, stream>>>(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], dev_out);
    return cudaGetLastError();
}
