// Functional model of the tcgen05 / mbarrier / bulk-copy PTX subset used by the fp16-piece tensor-core kernels
// (fastdiff_b200/csrc/fd_kernels_f16.cuh), for the CPU fibre emulator.  TEST INFRASTRUCTURE, never shipped.
//   * shared-memory "addresses" are byte offsets into the CTA's dynamic shared memory (1024-byte aligned like the real window);
//   * an mbarrier keeps its state in its own 64-bit word {pending arrivals, arrival count, outstanding tx bytes, phase};
//   * cp.async.bulk global->shared lands its bytes at issue or, on request, when its mbarrier is first polled (two legal schedules, see
//     bulk_g2s); shared->global lands when the issuing thread waits for it; tcgen05.mma executes synchronously at issue (so
//     tcgen05.commit is a plain arrive); polling loops yield to the other fibres of the CTA;
//   * tcgen05.mma kind::f16, cta_group::1, K = 16, both operands K-major SWIZZLE_128B from shared-memory descriptors: the swizzle
//     XORs ABSOLUTE address bits [7,10) into [4,7) (descriptor base_offset 0, as measured on B200), D = 128 lanes x N fp32 columns;
//   * TMEM = 128 lanes x 512 fp32 columns per CTA.
#pragma once
#include <vector>
#include <stdio.h>

namespace fd {

struct EmuMbar { int16_t pending; int16_t count; int32_t tx; uint32_t phase; };   // overlays the kernel's uint64_t (+ 4 bytes of it)
static_assert(sizeof(EmuMbar) <= 12, "mbarrier state");
// the kernels lay their mbarriers out as consecutive uint64_t: keep the model inside 8 bytes
struct EmuMbar8 { int16_t pending; int16_t count; int32_t tx_phase; };            // tx in the upper 31 bits (signed), phase in bit 0
static_assert(sizeof(EmuMbar8) == 8, "mbarrier state");

inline thread_local float emu_tmem_all[2][128][512];   // one TMEM per CTA of a pair
#define emu_tmem (emu_tmem_all[emu::t_cta_rank])

inline uint32_t smem_u32(const void* p) { return (uint32_t)((const unsigned char*)p - emu::t_dyn_smem); }
inline unsigned char* emu_smem_ptr(uint32_t a) { return emu::t_dyn_smem + a; }

inline void emu_mbar_settle(EmuMbar8* b) {
    if (b->pending == 0 && (b->tx_phase >> 1) == 0) { b->tx_phase ^= 1; b->pending = b->count; }
}
inline void mbar_init(uint64_t* bar, uint32_t count) { EmuMbar8* b = (EmuMbar8*)bar; b->pending = b->count = (int16_t)count; b->tx_phase = 0; }
inline void mbar_init_fence() {}
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {   // mbarrier.arrive.expect_tx
    EmuMbar8* b = (EmuMbar8*)bar;
    b->tx_phase += (int32_t)(bytes << 1);
    --b->pending;
    emu_mbar_settle(b);
}
inline void mbar_arrive(uint64_t* bar) { EmuMbar8* b = (EmuMbar8*)bar; --b->pending; emu_mbar_settle(b); }
inline void emu_mbar_complete_tx(uint64_t* bar, uint32_t bytes) { EmuMbar8* b = (EmuMbar8*)bar; b->tx_phase -= (int32_t)(bytes << 1); emu_mbar_settle(b); }
// global -> shared bulk copies.  Two schedules, both legal on the hardware, chosen by bit 0 of cudaemu_set_bulk_late():
//   early (default): the bytes land at issue -- the adversarial case for a target that something still READS (write-after-read);
//   late: the bytes land when a thread first polls the copy's mbarrier -- the adversarial case for a target that is read, or written
//         by ordinary stores, BEFORE the barrier was waited for.
// The emulated parity tests run the pipelined / ring-fed kernels under both.
struct EmuBulkLoad { void* dst; const void* src; uint32_t bytes; uint64_t* bar; unsigned serial; };
inline std::vector<EmuBulkLoad>& emu_load_pending() { static thread_local std::vector<EmuBulkLoad> v; return v; }
inline void emu_deliver_loads(uint64_t* bar) {
    auto& v = emu_load_pending();
    size_t keep = 0;
    for (size_t i = 0; i < v.size(); ++i) {
        if (v[i].serial != emu::t_cta_serial) continue;                       // left over from an earlier CTA of this OS thread
        if (v[i].bar == bar) { memcpy(v[i].dst, v[i].src, v[i].bytes); emu_mbar_complete_tx(bar, v[i].bytes); __atomic_fetch_add(&emu::g_late_ops[0], 1, __ATOMIC_RELAXED); }
        else v[keep++] = v[i];
    }
    v.resize(keep);
}
inline bool emu_mma_queue_empty();
inline void emu_run_mmas_until(uint64_t* bar);
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
    const EmuMbar8* b = (const EmuMbar8*)bar;
    for (uint32_t it = 0; it < (1u << 22); ++it) {
        // late schedules: hold the asynchronous work back until this thread has polled for two full rounds of the fibre scheduler, i.e.
        // until every other fibre of the CTA has run as far as it can without it (blocked at a barrier or polling as well)
        if (it >= 2) {
            if ((emu::g_bulk_late & 1) && !emu_load_pending().empty()) emu_deliver_loads(bar);
            if ((emu::g_bulk_late & 2) && !emu_mma_queue_empty()) emu_run_mmas_until(bar);
        }
        if ((uint32_t)(b->tx_phase & 1) != (parity & 1)) return;
        emu::yield();
    }
    fprintf(stderr, "tcemu: mbarrier wait timed out (protocol bug)\n");
    abort();
}
inline void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    if (emu::g_bulk_late & 1) { emu_load_pending().push_back({smem_dst, gsrc, bytes, bar, emu::t_cta_serial}); return; }
    memcpy(smem_dst, gsrc, bytes);
    emu_mbar_complete_tx(bar, bytes);
}
inline void bulk_prefetch_l2(const void*, uint32_t) {}
inline void bulk_g2s_once(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) { bulk_g2s(smem_dst, gsrc, bytes, bar); }   // (L2 policy hint: no functional effect)
// shared -> global bulk copy (bulk async-groups).  The model performs the copy as LATE as the program allows -- when the issuing thread
// executes wait_group(.read) -- so a staging buffer that is overwritten before its copies were waited for delivers the wrong bytes and
// the parity tests see it (the adversarial schedule for source-reuse hazards).  A kernel must end with bulk_wait_all() in every issuing
// thread (the real one must, too).
struct EmuBulkStore { void* dst; const void* src; uint32_t bytes; unsigned who; };
inline std::vector<EmuBulkStore>& emu_bulk_pending() { static thread_local std::vector<EmuBulkStore> v; return v; }   // per OS thread = per CTA (pair) in flight
inline unsigned emu_bulk_me() { return emu::t_linear_tid + 65536u * emu::t_cta_rank; }
inline void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) { emu_bulk_pending().push_back({gdst, smem_src, bytes, emu_bulk_me()}); }
inline void emu_bulk_flush() {
    auto& v = emu_bulk_pending();
    const unsigned me = emu_bulk_me();
    size_t keep = 0;
    for (size_t i = 0; i < v.size(); ++i) {
        if (v[i].who == me) memcpy(v[i].dst, v[i].src, v[i].bytes);
        else v[keep++] = v[i];
    }
    v.resize(keep);
}
inline void bulk_commit() {}
inline void bulk_wait_read0() { emu_bulk_flush(); }
inline void bulk_wait_all() { emu_bulk_flush(); }
inline bool elect_one() { return (emu::t_linear_tid & 31) == 0; }
inline void tc_fence_before() {}
inline void tc_fence_after() {}
inline void fence_async_smem() {}
// tcgen05.mma is asynchronous as well.  Early schedule (default): it executes at issue and tcgen05.commit is a plain arrive.  Late schedule
// (bit 1 of cudaemu_set_bulk_late(), 1-CTA kind::f16 MMAs): MMAs and commits queue up in program order and run when a thread first polls a
// committed mbarrier -- an operand tile refilled, or an accumulator read, before the commit was waited for then shows in the results.
struct EmuMmaOp { int kind; uint32_t d; uint64_t a, b; uint32_t idesc, acc; uint64_t* bar; unsigned serial, who; };   // kind 0 = mma, 1 = commit
inline std::vector<EmuMmaOp>& emu_mma_pending() { static thread_local std::vector<EmuMmaOp> v; return v; }
inline void emu_mma_f16_now(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate);
inline bool emu_mma_queue_empty() { return emu_mma_pending().empty(); }
inline void emu_run_mmas_until(uint64_t* bar) {       // run the issuing thread's queue in order up to (and including) its first commit on `bar`
    auto& v = emu_mma_pending();
    size_t keep = 0;
    for (size_t i = 0; i < v.size(); ++i)              // entries left over from an earlier CTA of this OS thread
        if (v[i].serial == emu::t_cta_serial) v[keep++] = v[i];
    v.resize(keep);
    size_t upto = 0;
    for (size_t i = 0; i < v.size(); ++i)
        if (v[i].kind == 1 && v[i].bar == bar) { upto = i + 1; break; }
    if (!upto) return;
    const unsigned who = v[upto - 1].who;              // only that thread's MMAs are ordered before this commit: the other issuers' stay queued
    keep = 0;
    for (size_t i = 0; i < v.size(); ++i) {
        if (i < upto && v[i].who == who) {
            if (v[i].kind == 0) { emu_mma_f16_now(v[i].d, v[i].a, v[i].b, v[i].idesc, v[i].acc); __atomic_fetch_add(&emu::g_late_ops[1], 1, __ATOMIC_RELAXED); }
            else mbar_arrive(v[i].bar);
        } else {
            v[keep++] = v[i];
        }
    }
    v.resize(keep);
}
inline void tc_commit(uint64_t* bar) {
    if (emu::g_bulk_late & 2) { emu_mma_pending().push_back({1, 0, 0, 0, 0, 0, bar, emu::t_cta_serial, emu_bulk_me()}); return; }
    mbar_arrive(bar);
}
inline void tmem_alloc(uint32_t* dst_smem, uint32_t) { *dst_smem = 0; }
inline void tmem_dealloc(uint32_t, uint32_t) {}
inline void group_sync(int id, int nthreads) { emu::barrier(id, nthreads); }
inline void tmem_ld_wait() {}

inline uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
constexpr uint32_t umma_idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
constexpr uint32_t umma_idesc_tf32(int M, int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

inline uint32_t umma_desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
inline uint64_t umma_desc_at(uint32_t lo) { return ((uint64_t)0x40004040u << 32) | lo; }
inline uint64_t umma_desc_sw64(uint32_t smem_addr) {   // K-major SWIZZLE_64B: rows of 64 B, 8-row groups 512 B apart (layout type 4)
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
// element (row, k) of a K-major operand slice: layout 2 = SWIZZLE_128B (128-byte rows, address bits [7,10) XORed into [4,7)),
// layout 4 = SWIZZLE_64B (64-byte rows, bits [7,9) XORed into [4,6)) -- both on ABSOLUTE address bits, as measured on B200 (tests/microbench/tc_probe.cu)
inline float emu_f16_at(uint32_t desc_start, uint32_t sbo, int row, int k, int layout = 2) {
    const uint32_t pitch = layout == 4 ? 64u : 128u;
    const uint32_t lin = desc_start + (uint32_t)(row >> 3) * sbo + (uint32_t)(row & 7) * pitch + (uint32_t)k * 2u;
    const uint32_t phys = layout == 4 ? (lin ^ (((lin >> 7) & 3u) << 4)) : (lin ^ (((lin >> 7) & 7u) << 4));
    uint16_t h;
    memcpy(&h, emu_smem_ptr(phys), 2);
    return f16_bits_to_float_soft(h);
}
inline void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if (emu::g_bulk_late & 2) { emu_mma_pending().push_back({0, d_tmem, a_desc, b_desc, idesc, accumulate, nullptr, emu::t_cta_serial, emu_bulk_me()}); return; }
    emu_mma_f16_now(d_tmem, a_desc, b_desc, idesc, accumulate);
}
inline void emu_mma_f16_now(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    const int N = (int)((idesc >> 17) & 0x3F) << 3, M = (int)((idesc >> 24) & 0x1F) << 4;
    const uint32_t a0 = (uint32_t)(a_desc & 0x3FFF) << 4, b0 = (uint32_t)(b_desc & 0x3FFF) << 4;
    const uint32_t sa = (uint32_t)((a_desc >> 32) & 0x3FFF) << 4, sb = (uint32_t)((b_desc >> 32) & 0x3FFF) << 4;
    const int la = (int)((a_desc >> 61) & 7), lb = (int)((b_desc >> 61) & 7);
    if (M != 128 || (la != 2 && la != 4) || (lb != 2 && lb != 4)) { fprintf(stderr, "tcemu: unsupported MMA shape/layout\n"); abort(); }
    const int lane0 = (int)(d_tmem >> 16), col0 = (int)(d_tmem & 0xFFFF);
    float bt[256][16];
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < 16; ++k) bt[n][k] = emu_f16_at(b0, sb, n, k, lb);
    for (int m = 0; m < M; ++m) {
        float a[16];
        for (int k = 0; k < 16; ++k) a[k] = emu_f16_at(a0, sa, m, k, la);
        for (int n = 0; n < N; ++n) {
            float acc = accumulate ? emu_tmem[lane0 + m][col0 + n] : 0.f;
            for (int k = 0; k < 16; ++k) acc += a[k] * bt[n][k];
            emu_tmem[lane0 + m][col0 + n] = acc;
        }
    }
}
// kind::tf32: K = 8 fp32 elements per instruction (32 bytes of a K-major row).  The pieces the kernels feed have at most 11
// significant bits (hi) or are tiny (lo), so the products are formed in fp32 without modelling the operand conversion to tf32.
inline float emu_f32_at(uint32_t desc_start, uint32_t sbo, int row, int k) {
    const uint32_t lin = desc_start + (uint32_t)(row >> 3) * sbo + (uint32_t)(row & 7) * 128u + (uint32_t)k * 4u;
    const uint32_t phys = lin ^ (((lin >> 7) & 7u) << 4);
    float f;
    memcpy(&f, emu_smem_ptr(phys), 4);
    return f;
}
inline void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    const int N = (int)((idesc >> 17) & 0x3F) << 3, M = (int)((idesc >> 24) & 0x1F) << 4;
    const uint32_t a0 = (uint32_t)(a_desc & 0x3FFF) << 4, b0 = (uint32_t)(b_desc & 0x3FFF) << 4;
    const uint32_t sa = (uint32_t)((a_desc >> 32) & 0x3FFF) << 4, sb = (uint32_t)((b_desc >> 32) & 0x3FFF) << 4;
    if (M != 128 || ((a_desc >> 61) & 7) != 2 || ((b_desc >> 61) & 7) != 2) { fprintf(stderr, "tcemu: unsupported MMA shape/layout\n"); abort(); }
    const int lane0 = (int)(d_tmem >> 16), col0 = (int)(d_tmem & 0xFFFF);
    float bt[256][8];
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < 8; ++k) bt[n][k] = emu_f32_at(b0, sb, n, k);
    for (int m = 0; m < M; ++m) {
        float a[8];
        for (int k = 0; k < 8; ++k) a[k] = emu_f32_at(a0, sa, m, k);
        for (int n = 0; n < N; ++n) {
            float acc = accumulate ? emu_tmem[lane0 + m][col0 + n] : 0.f;
            for (int k = 0; k < 8; ++k) acc += a[k] * bt[n][k];
            emu_tmem[lane0 + m][col0 + n] = acc;
        }
    }
}
inline uint16_t emu_f16_sat(float f);
// ---- CTA pairs (cta_group::2): the kernel_conv GEMM ---------------------------------------------------------------------------
// TMA tensor maps: the driver's opaque CUtensorMap is replaced by a plain description of the 2-D fp32-sized tensor and its box.
struct CUtensorMap { unsigned char opaque[128]; } __attribute__((aligned(64)));
struct EmuTmap { const unsigned char* base; uint64_t cols, rows, stride_bytes; uint32_t box_cols, box_rows; };
static_assert(sizeof(EmuTmap) <= sizeof(CUtensorMap), "tensor map model");
inline void emu_make_map_2d(CUtensorMap* m, const float* base, uint64_t cols, uint64_t rows, uint64_t row_stride_bytes, uint32_t box_cols,
                            uint32_t box_rows) {
    EmuTmap t{(const unsigned char*)base, cols, rows, row_stride_bytes, box_cols, box_rows};
    memset(m, 0, sizeof *m);
    memcpy(m, &t, sizeof t);
}
inline uint32_t cluster_ctarank() { return emu::cluster_rank(); }
inline void cluster_sync_all() { emu::cluster_barrier(); }
inline uint64_t* emu_bar_of_rank(uint64_t* bar, unsigned rank) {   // the barrier at the same shared-memory offset in CTA `rank` of the pair
    return (uint64_t*)(emu::cluster_smem(rank) + ((unsigned char*)bar - emu::t_dyn_smem));
}
// cp.async.bulk.tensor.2d ... cta_group::2: box {c0 .. c0+box_cols, c1 .. c1+box_rows} (out-of-range rows/cols read as zero) into
// this CTA's shared memory as 128-byte rows, SWIZZLE_128B (dst is 1024-byte aligned: chunk j of row r at position j ^ (r & 7));
// the bytes complete on the LEADER's copy of the barrier.
inline void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    EmuTmap t;
    memcpy(&t, map, sizeof t);
    if (t.box_cols != 32) { fprintf(stderr, "tcemu: TMA box must be 128 bytes wide\n"); abort(); }
    unsigned char* dst = (unsigned char*)smem_dst;
    for (uint32_t r = 0; r < t.box_rows; ++r) {
        const int64_t row = (int64_t)c1 + r;
        for (uint32_t j = 0; j < 8; ++j) {
            float chunk[4] = {0.f, 0.f, 0.f, 0.f};
            for (int e = 0; e < 4; ++e) {
                const int64_t col = (int64_t)c0 + j * 4 + e;
                if (row >= 0 && (uint64_t)row < t.rows && col >= 0 && (uint64_t)col < t.cols)
                    memcpy(&chunk[e], t.base + (uint64_t)row * t.stride_bytes + (uint64_t)col * 4, 4);
            }
            memcpy(dst + r * 128 + ((j ^ (r & 7)) << 4), chunk, 16);
        }
    }
    emu_mbar_complete_tx(emu_bar_of_rank(bar, 0), t.box_rows * 128u);
}
inline void tc_commit_2sm(uint64_t* bar) { mbar_arrive(emu_bar_of_rank(bar, 0)); mbar_arrive(emu_bar_of_rank(bar, 1)); }
inline void mbar_arrive_leader(uint64_t* bar) { mbar_arrive(emu_bar_of_rank(bar, 0)); }
inline void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t) { *dst_smem = 0; }
inline void tmem_dealloc_2sm(uint32_t, uint32_t) {}
inline float emu_f16_at_rank(unsigned rank, uint32_t desc_start, uint32_t sbo, int row, int k) {
    const uint32_t lin = desc_start + (uint32_t)(row >> 3) * sbo + (uint32_t)(row & 7) * 128u + (uint32_t)k * 2u;
    const uint32_t phys = lin ^ (((lin >> 7) & 7u) << 4);
    uint16_t h;
    memcpy(&h, emu::cluster_smem(rank) + phys, 2);
    return f16_bits_to_float_soft(h);
}
// tcgen05.mma.cta_group::2.kind::f16, M = 256, N = 256: CTA r of the pair supplies A rows [128 r, +128) and B rows [128 r, +128) from
// the same shared-memory offsets; CTA r's TMEM receives its 128 rows of D (all 256 columns).
inline void umma_f16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    const int N = (int)((idesc >> 17) & 0x3F) << 3, M = (int)((idesc >> 24) & 0x1F) << 4;
    if (M != 256 || N != 256) { fprintf(stderr, "tcemu: unsupported 2-CTA MMA shape\n"); abort(); }
    const uint32_t a0 = (uint32_t)(a_desc & 0x3FFF) << 4, b0 = (uint32_t)(b_desc & 0x3FFF) << 4;
    const uint32_t sa = (uint32_t)((a_desc >> 32) & 0x3FFF) << 4, sb = (uint32_t)((b_desc >> 32) & 0x3FFF) << 4;
    const int lane0 = (int)(d_tmem >> 16), col0 = (int)(d_tmem & 0xFFFF);
    static thread_local float bt[256][16];
    for (int n = 0; n < 256; ++n)
        for (int k = 0; k < 16; ++k) bt[n][k] = emu_f16_at_rank((unsigned)(n >> 7), b0, sb, n & 127, k);
    for (unsigned r = 0; r < 2; ++r)
        for (int m = 0; m < 128; ++m) {
            float a[16];
            for (int k = 0; k < 16; ++k) a[k] = emu_f16_at_rank(r, a0, sa, m, k);
            float* drow = &emu_tmem_all[r][lane0 + m][col0];
            for (int n = 0; n < 256; ++n) {
                float acc = accumulate ? drow[n] : 0.f;
                for (int k = 0; k < 16; ++k) acc += a[k] * bt[n][k];
                drow[n] = acc;
            }
        }
}
inline void umma_tf32_2sm(uint32_t, uint64_t, uint64_t, uint32_t, uint32_t) { fprintf(stderr, "tcemu: the tf32 CTA-pair GEMM is not modelled\n"); abort(); }
inline uint16_t f16_sat_bits(float f) { return emu_f16_sat(f); }

template <int N> inline void emu_tmem_ld(uint32_t taddr, uint32_t (&v)[N]) {
    const int lane = (int)(taddr >> 16) + (int)(emu::t_linear_tid & 31), col = (int)(taddr & 0xFFFF);
    for (int i = 0; i < N; ++i) memcpy(&v[i], &emu_tmem[lane][col + i], 4);
}
inline void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) { emu_tmem_ld<32>(taddr, v); }
inline void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) { emu_tmem_ld<16>(taddr, v); }
inline void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&v)[8]) { emu_tmem_ld<8>(taddr, v); }
inline void tmem_ld_32x32b_x4(uint32_t taddr, uint32_t* v) { uint32_t t[4]; emu_tmem_ld<4>(taddr, t); for (int i = 0; i < 4; ++i) v[i] = t[i]; }

// packed fp16 helpers (cvt.rn.satfinite.f16x2.f32 d, a, b: a -> upper half, b -> lower half)
inline uint16_t emu_f16_sat(float f) {
    if (f > F16_MAX) f = F16_MAX;
    if (f < -F16_MAX) f = -F16_MAX;
    return f16_bits_rn_soft(f);
}
inline uint32_t pack_f16x2_sat(float upper, float lower) { return ((uint32_t)emu_f16_sat(upper) << 16) | emu_f16_sat(lower); }
inline float2 unpack_f16x2(uint32_t u) { float2 r; r.x = f16_bits_to_float_soft((uint16_t)(u & 0xFFFF)); r.y = f16_bits_to_float_soft((uint16_t)(u >> 16)); return r; }
inline float ex2_approx(float x) { return exp2f(x); }
inline float rcp_approx(float x) { return 1.f / x; }
#define FD_OPAQUE(x) ((void)0)
#define FD_OPAQUE2(x, y) ((void)0)

}  // namespace fd

// warp primitives as the kernels use them: lane-0 broadcasts of warp-uniform values, and __syncwarp as a real barrier
template <typename T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
static inline void __syncwarp() { emu::syncwarp(); }

