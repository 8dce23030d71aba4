// tcgen05 / TMEM / TMA kernels (sm_100a) for the heavy contractions.
//
// Two families share this file.  The DEFAULT mode FD_MODE_TC_3XF16 runs kind::f16 MMAs on fp16 hi/lo pieces of power-of-two
// prescaled operands (fd_common.cuh: f16_split, S16_*): k_kc_gemm_tc2<true, 16>, k_lvc_layer_h,
// k_kp_hidden_tc.  The older kind::tf32 family (k_kc_gemm_tc, k_kc_gemm_tc2<false, 8>, k_lvc_layer_tc, and still the only
// version of k_dblock0_tc / k_upsample_tc) is described first:
//
// Arithmetic: kind::tf32 MMAs with fp32 accumulation in TMEM.  FD_MODE_TC_3XTF32 splits BOTH operands into
// explicit tf32 pieces  x = hi + lo  (hi = RN_tf32(x), lo = RN_tf32(x - hi); low 13 mantissa bits stored as zero, so
// the tensor core's fp32->tf32 conversion is exact whatever its rounding) and accumulates hi*hi + hi*lo + lo*hi:
// ~2^-21 relative per product, i.e. fp32-level.  FD_MODE_TC_TF32 issues only hi*hi.
//
// K8+K9  kernel_conv + bias_conv GEMM (modules.py:330-331), "swap-AB":
//     D[n, p] = sum_k WT[n][k] * hkpad[p*64 + k]        n < 24832 (M side, 128 per tile), p = padded frame row (N side)
//   A = weights, K-major rows of 192 floats (packer sections LBn_KCT_HI / _LO), TMA box {32 k, 128 n}
//   B = kernel-predictor hidden rows; the im2col row of frame p is the 192 contiguous floats starting at
//       hk[p*64], so k-atom a (32 floats) of row p is the plain 2-D box at (col (a&1)*32, row p + a/2)
//   D in TMEM: lane = n, column = frame  ->  the epilogue thread of lane n holds consecutive frames in
//   registers and a warp stores 32 consecutive n of one frame = one coalesced 128-byte line of
//   kern[(b,f)][n]; no smem staging, bias added per lane.
// Warp roles (320 threads): warp 0 TMA producer + TMEM allocator, warp 1 MMA issuer, warps 2-9 epilogue (two per TMEM lane
// quarter, each taking 128 of the 256 columns).
// Persistent CTAs (grid = #SM), 2-stage smem ring (96 KB/stage), 2 x 256-column TMEM accumulators.
#pragma once
// The fp16-piece kernels (k_lvc_layer_h, k_kp_hidden_tc) also compile for the CPU fibre emulator (FD_EMU: tests/cudaemu/tcemu.h
// supplies functional models of the PTX wrappers); everything else in this file is GPU-only.
#ifndef FD_EMU
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#endif
#include <string>
#include "fd_blob.h"
#include "fd_common.cuh"
#ifdef FD_EMU
#include "tcemu.h"
#endif

namespace fd {

#ifndef FD_EMU
// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug traps instead of hanging the GPU box.
// MBAR_HINT_NS > 0: try_wait carries a suspend-time hint, so a waiting warp sleeps in hardware up to that long per poll instead of
// re-issuing the poll loop every ~100 cycles (ncu, round 2: a third of k_lvc_p's issued instructions were polls).
#ifndef MBAR_HINT_NS
#define MBAR_HINT_NS 0
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
#pragma unroll 1   // (ptxas otherwise unrolls the poll 32 times at every wait site: ~70 of the 357 instructions of k_lvc_p's gate loop, I-cache misses)
    for (uint32_t it = 0; it < (1u << 22); ++it) {
#if MBAR_HINT_NS > 0
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)MBAR_HINT_NS) : "memory");
#else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
#endif
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
// One lane of a converged warp (cute::elect_one_sync): the compiler recognises the elect.sync predicate as single-thread,
// so uniform-datapath instructions (UTCHMMA, UTCBAR, UTMALDG) inside the branch are emitted straight-line.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by one thread.
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B operand tile: rows of 128 bytes, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor,
// cute/arch/mma_sm100_desc.hpp: start>>4 [0,14), LBO [16,30), SBO [32,46), version=1 [46,48), layout [61,64) = 2).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;              // LBO (unused for swizzled K-major; canonical value 1)
    d |= (uint64_t)(1024 >> 4) << 32;    // SBO: 8 rows * 128 B
    d |= (uint64_t)1 << 46;              // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;              // SWIZZLE_128B
    return d;
}
// The same descriptor in two steps for issue loops that build many of them: the low word of the descriptor of `smem_addr` once, then every
// other start address of the same tile is that word plus the byte offset / 16 (no carry out of the 14-bit field below 256 KB) -- one add.
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t umma_desc_at(uint32_t lo) { return ((uint64_t)0x40004040u << 32) | lo; }   // SBO 1024, version 1, SWIZZLE_128B
// K-major, SWIZZLE_64B operand tile: rows of 64 bytes, 8-row groups 512 B apart, 16-byte chunk c of row r at position c ^ ((r >> 1) & 3)
// (layout type 4; addressing confirmed on B200 with tests/microbench/tc_probe.cu, probe 2: start + 32 B selects the second K = 16 slice).
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;     // SBO: 8 rows * 64 B
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;              // SWIZZLE_64B
    return d;
}
// K-major, no swizzle ("interleave"): 8-row x 16-byte core matrices, LBO between K chunks, SBO between 8-row groups (microbenchmark only).
__device__ __forceinline__ uint64_t umma_desc_ns(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 (1<<4), A=B=tf32 (2<<7, 2<<10), K-major both, N>>3 at 17, M>>4 at 24.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// kind::f16 with A = B = fp16 (format code 0), D = f32; one instruction covers K = 16 (32 bytes of a K-major row).
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// --- wrappers the fp16-piece kernels use in both builds (the emulator has models of the same names) ---
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t ncols) {       // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
// two floats -> packed fp16 pair, round to nearest even, saturating at +-65504: `upper` lands in bits [16,32)
__device__ __forceinline__ uint32_t pack_f16x2_sat(float upper, float lower) {
    uint32_t d;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(upper), "f"(lower));
    return d;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {   // the same warp of both CTAs of a pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t base, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}
// one float -> fp16 bits, round to nearest even, saturating at +-65504
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_32x32b_x4(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void group_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// shared -> global bulk copy tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // sources may be overwritten
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }          // writes complete
__device__ __forceinline__ uint16_t f16_sat_bits(float f) { uint16_t h; asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(f)); return h; }
// opaque to the optimiser: keeps ptxas from hoisting ~100 descriptors out of a tile loop
#define FD_OPAQUE(x) asm volatile("" : "+r"(x))
#define FD_OPAQUE2(x, y) asm volatile("" : "+r"(x), "+r"(y))

// ---------------------------------------------------------------------------------------------------------
// kernel_conv GEMM
// ---------------------------------------------------------------------------------------------------------
#endif  // !FD_EMU
constexpr int KCG_BM = 128;                 // n rows per tile (MMA M)
constexpr int KCG_BN = 256;                 // frames per tile (MMA N)
constexpr int KCG_KATOM = 32;               // floats per 128-byte swizzle row
constexpr int KCG_NATOM = KCK / KCG_KATOM;  // 6
constexpr int KCG_STAGES = 2;
constexpr int KCG_A_BYTES = KCG_BM * 128;   // 16 KB
constexpr int KCG_B_BYTES = KCG_BN * 128;   // 32 KB
constexpr int KCG_STAGE_BYTES = 2 * KCG_A_BYTES + 2 * KCG_B_BYTES;  // A_hi | A_lo | B_hi | B_lo = 96 KB
constexpr int KCG_SMEM_BYTES = KCG_STAGES * KCG_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

struct KcgMaps {               // per LVC block: weights hi/lo, hidden hi/lo
    CUtensorMap w_hi[NBLK], w_lo[NBLK], h_hi[NBLK], h_lo[NBLK];
};

#ifndef FD_EMU
__global__ void __launch_bounds__(320, 1)
k_kc_gemm_tc(const __grid_constant__ KcgMaps maps, const float* __restrict__ bias0, const float* __restrict__ bias1,
             const float* __restrict__ bias2, float* __restrict__ kern_all, int B, int Tm, int three_pass) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space
    uint64_t* bars = (uint64_t*)(smem + KCG_STAGES * KCG_STAGE_BYTES);
    uint64_t* full_bar = bars;                    // [STAGES]  TMA -> MMA
    uint64_t* empty_bar = bars + KCG_STAGES;      // [STAGES]  MMA -> TMA
    uint64_t* tfull_bar = bars + 2 * KCG_STAGES;  // [2]       MMA -> epilogue
    uint64_t* tempty_bar = tfull_bar + 2;         // [2]       epilogue -> MMA
    uint32_t* tmem_base_s = (uint32_t*)(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int M = B * (Tm + 2) - 2;                       // padded frame rows
    const int f_tiles = (M + KCG_BN - 1) / KCG_BN;
    const int n_tiles = KCN / KCG_BM;                     // 194
    const int tiles_per_blk = n_tiles * f_tiles;
    const int total_tiles = NBLK * tiles_per_blk;

    if (threadIdx.x == 0) {
        for (int s = 0; s < KCG_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int blk = tile / tiles_per_blk, r = tile % tiles_per_blk;
                const int ft = r / n_tiles, nt = r % n_tiles;   // n fastest: consecutive CTAs share the frame tile
                const CUtensorMap* wh = &maps.w_hi[blk]; const CUtensorMap* wl = &maps.w_lo[blk];
                const CUtensorMap* hh = &maps.h_hi[blk]; const CUtensorMap* hl = &maps.h_lo[blk];
                for (int a = 0; a < KCG_NATOM; ++a) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    unsigned char* st = smem + stage * KCG_STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], three_pass ? KCG_STAGE_BYTES : (KCG_A_BYTES + KCG_B_BYTES));
                    tma_load_2d(st, wh, a * KCG_KATOM, nt * KCG_BM, &full_bar[stage]);
                    tma_load_2d(st + 2 * KCG_A_BYTES, hh, (a & 1) * KCG_KATOM, ft * KCG_BN + (a >> 1), &full_bar[stage]);
                    if (three_pass) {
                        tma_load_2d(st + KCG_A_BYTES, wl, a * KCG_KATOM, nt * KCG_BM, &full_bar[stage]);
                        tma_load_2d(st + 2 * KCG_A_BYTES + KCG_B_BYTES, hl, (a & 1) * KCG_KATOM, ft * KCG_BN + (a >> 1), &full_bar[stage]);
                    }
                    if (++stage == KCG_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_tf32(KCG_BM, KCG_BN);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);   // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * KCG_BN;
                for (int a = 0; a < KCG_NATOM; ++a) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + stage * KCG_STAGE_BYTES);
                    const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + KCG_A_BYTES);
                    const uint64_t b_hi = umma_desc_sw128(st + 2 * KCG_A_BYTES), b_lo = umma_desc_sw128(st + 2 * KCG_A_BYTES + KCG_B_BYTES);
#pragma unroll
                    for (int k = 0; k < KCG_KATOM / 8; ++k) {   // UMMA_K = 8 tf32 = 32 bytes: advance start address by 2 (16-B units)
                        const uint64_t adv = (uint64_t)(k * 2);
                        umma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc, (a | k) ? 1u : 0u);
                        if (three_pass) {
                            umma_tf32(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                            umma_tf32(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
                        }
                    }
                    tc_commit(&empty_bar[stage]);              // smem slot free once these MMAs have read it
                    if (++stage == KCG_STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(&tfull_bar[acc]);                    // accumulator complete
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 2) {
        // ================= epilogue: TMEM -> registers -> (+bias) -> global =================
        const int q = warp & 3;              // TMEM lane quarter this warp may access
        const int chalf = (warp - 2) >> 2;   // which 128 of the 256 frame columns
        uint32_t acc = 0, acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int blk = tile / tiles_per_blk, r = tile % tiles_per_blk;
            const int ft = r / n_tiles, nt = r % n_tiles;
            const int n = nt * KCG_BM + q * 32 + lane;
            const float* bias = blk == 0 ? bias0 : (blk == 1 ? bias1 : bias2);
            const float bv = bias[n];
            float* kern = kern_all + (size_t)blk * B * Tm * KCN;
            int p = ft * KCG_BN + chalf * 128;         // padded row of this warp's first column
            int center = p + 1, bb = center / (Tm + 2), fp = center % (Tm + 2);
            const bool fast = (fp >= 1) && (fp + 127 <= Tm) && (p + 127 < M);   // 128 valid frames of one item
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * KCG_BN + chalf * 128;
            if (fast) {
                float* o = kern + ((size_t)bb * Tm + (fp - 1)) * KCN + n;
#pragma unroll 1
                for (int c0 = 0; c0 < 128; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) o[(size_t)j * KCN] = __uint_as_float(v[j]) + bv;   // immediate-offset stores
                    o += (size_t)32 * KCN;
                }
            } else {
#pragma unroll 1
                for (int c0 = 0; c0 < 128; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (p < M && fp >= 1 && fp <= Tm)
                            kern[((size_t)bb * Tm + (fp - 1)) * KCN + n] = __uint_as_float(v[j]) + bv;
                        ++p;
                        if (++fp == Tm + 2) { fp = 0; ++bb; }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------
// kernel_conv GEMM, CTA-pair version (cta_group::2).  The 1-CTA kernel above is shared-memory-bandwidth bound (per 128-cycle
// N=256 MMA: 12 KB of operand reads + the TMA refill = 158 B/clk against 128 B/clk; profiles/r01_tcgen05_findings.md).  Here a
// cluster of two CTAs computes a 256 (n) x 256 (frames) tile: each CTA stages its own 128 weight rows and HALF of the frame
// rows (128), the leader issues one M=256 MMA per k-step that reads both halves, and each CTA ends up with its 128 n-rows x
// 256 frames in its own TMEM.  Per SM: 8 KB operand reads per MMA + 64 KB TMA per stage = ~106 B/clk.
//   * TMA loads of BOTH CTAs signal the LEADER's full barrier (cp.async.bulk.tensor...cta_group::2, barrier address with the
//     peer bit cleared); the leader arms it with the byte count of the pair.
//   * tcgen05.commit...multicast::cluster (mask 0b11) releases the smem slot / publishes the accumulator in both CTAs.
//   * epilogue warps of both CTAs arrive (remotely for the peer) on the leader's tmem-empty barrier.
// 3-stage ring of 64 KB per CTA, 2 x 256-column accumulators, 8 epilogue warps per CTA as in the 1-CTA kernel.
// ---------------------------------------------------------------------------------------------------------
#endif  // !FD_EMU
constexpr int KC2_STAGES = 3;
constexpr int KC2_A_BYTES = 128 * 128;       // 16 KB: 128 weight rows x one 32-float k-atom
constexpr int KC2_B_BYTES = 128 * 128;       // 16 KB: this CTA's 128 of the 256 frame rows
constexpr int KC2_STAGE_BYTES = 2 * KC2_A_BYTES + 2 * KC2_B_BYTES;   // A_hi | A_lo | B_hi | B_lo = 64 KB
constexpr int KC2_SMEM_BYTES = KC2_STAGES * KC2_STAGE_BYTES + 1024 + 256;
// RES form (k_kc_gemm_tc2<true, 16, B0P, true>, option "kc_res" = 1; parity-green on B200 and measured SLOWER than the default, see below):
// the frame tile stays resident, only the weights stream.
//   * a CTA pair walks a CONTIGUOUS range of the (block, frame tile, n tile) sequence, so the 256 frames of a tile are loaded once per ~97 items
//     and not once per item; the three k-atoms of the hidden im2col are the SAME rows shifted by one (taps p, p+1, p+2): one 130-row SWIZZLE_128B
//     tile per piece serves all three through shifted descriptor starts (absolute-address swizzle, base_offset 0 -- the LVC kernels' trick,
//     here on a cta_group::2 B operand);
//   * what streams through the ring is the weight k-atom only (A_hi | A_lo = 32 KB): 4 stages = 1.33 items of prefetch at half the bytes
//     per item; two frame-tile buffers, so the next tile's load overlaps the last items of the current one;
//   * the epilogue warps read all 64 accumulator columns before the first store and release the stage at once (KC_EARLY_RELEASE).
// Result (profiles/r02_kc_gemm_res_ab.txt): 0.92 ms per launch against 0.57 -- the operand side is fine (no `full` waits), the STORES slow down
// (2,200-6,300 cycles per 32 frames instead of ~2,800).  The likely cause: with the merged-N image the two 64-byte halves of a 128-byte line of the predicted-kernel tensor belong to items 8 n-tile pairs apart; strided
// over the pairs they are written within microseconds of each other by neighbouring pairs and leave L2 as full lines, in a contiguous range the
// same pair writes them ~40 us (~140 MB of stores) apart and the half lines go to DRAM one by one.  The early release alone, on the whole-stage
// ring, was slower too (2.60 vs 2.33 ms per call): the 3-slot ring holds exactly one item, so the wait just moves from `tempty` to `full`.
constexpr int KC3_ASTAGES = 4;
constexpr int KC3_A_STAGE = 2 * KC2_A_BYTES;          // 32 KB
constexpr int KC3_BROWS = 130;                        // 128 frames + 2 rows of im2col halo
constexpr int KC3_B_PIECE = 17 * 1024;                // 130 x 128 B rounded up to the 1024-byte swizzle period
constexpr int KC3_B_BUF = 2 * KC3_B_PIECE;            // hi | lo
constexpr int KC3_RING_BYTES = KC3_ASTAGES * KC3_A_STAGE + 2 * KC3_B_BUF;   // 200,704
constexpr int KC3_SMEM_BYTES = KC3_RING_BYTES + 1024 + 256;

#ifndef FD_EMU
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    const uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;   // Sm100MmaPeerBitMask: the leader CTA's copy of the barrier
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(mbar) : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar) {   // arrives on the barrier at this offset in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {   // arrive on CTA 0's copy of `bar`
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(0u));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");   // cutlass ClusterBarrier::arrive(cta_id) form

}

#endif  // !FD_EMU
// F16 = true (mode tc_3xf16): the operands are fp16 pieces (weights LBn_KCT_F16 prescaled per tensor, hidden rows of 64 fp16
// = 128 B prescaled by S16_HK), a k-atom is 64 values = one tap of the im2col (box at row p + a), every MMA is kind::f16 with
// K = 16, and the epilogue undoes the scales: kern = acc * inv[blk] + bias.  Half the MMAs and half the operand bytes of the
// tf32 variant for the same 22 significant bits per operand.
#ifndef KC_STORE_CS
#define KC_STORE_CS 1     // 1: streaming (evict-first) cache hint on the predicted-kernel piece stores: the 2 GB stream then does not push the operand
#endif                    //    tiles out of L2 -- GEMM prefix of a reverse step 0.776 -> 0.750 ms in the 2 s power-profile loops, same joules (profiles/r02_energy_ab.txt)
#ifndef KC_EARLY_RELEASE
#define KC_EARLY_RELEASE 1   // 1 (RES form only): an epilogue warp reads all of its 64 accumulator columns before its first store and releases the stage
                             // right away.  With the whole-stage ring this was SLOWER (2.60 vs 2.33 ms per call): the operand fetch of the next item became the wait.
#endif
// EPW = epilogue warps per CTA (8 or 16: EPW/4 per TMEM lane quarter, each taking 256/(EPW/4) of the 256 frame columns).
// B0P (experimental, option "tc_b0" = 1): block 0's predicted kernels are written as fp16 pieces as well (its weight rows then come
// in the same SWIZZLE_128B image order as blocks 1 and 2: sections LB0_KCT_F16P / LB0_KC_BP) for the tensor-core block-0 consumer.
// (Measured and dropped in round 2: the un-swapped form -- frames on the MMA's M side, so that an epilogue thread owns whole 128-byte operand
// rows and writes them with 16-byte stores, bit-identical output -- ran at 1.60 ms per launch against 0.61 ms: a warp store then touches 32
// different lines with 16 bytes each, and the L2 request rate, not the instruction count, is what bounds this store stream.)
// Optional role timeline of the kernel_conv GEMM (-DKC_TIMELINE=1, GPU build): CTA 0 stamps clock64 for its first 32 items --
// [role 0 TMA producer | 1 MMA issuer | 2 epilogue warp 2][item][8 slots]; fd_debug_read("kc_timeline"), tests/gpu_scripts/kc_timeline.py.
#if defined(KC_TIMELINE) && !defined(FD_EMU)
constexpr int KC_TL_ITEMS = 32;
__device__ unsigned long long g_kc_timeline[3 * KC_TL_ITEMS * 8];
#define KC_STAMP(role, it, slot) do { if (blockIdx.x == 0 && (it) < KC_TL_ITEMS) g_kc_timeline[((role) * KC_TL_ITEMS + (it)) * 8 + (slot)] = clock64(); } while (0)
#else
#define KC_STAMP(role, it, slot) do { } while (0)
#endif
template <bool F16, int EPW, bool B0P = false, bool RES = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EPW, 1)
k_kc_gemm_tc2(const __grid_constant__ KcgMaps maps, const float* __restrict__ bias0, const float* __restrict__ bias1,
              const float* __restrict__ bias2, float* __restrict__ kern_all, int B, int Tm, int three_pass,
              float inv0, float inv1, float inv2, int exp_mask) {
    pdl_trigger();

    constexpr int NATOM = F16 ? 3 : 6;
    constexpr int CPW = 256 / (EPW / 4);   // frame columns per epilogue warp
    FD_DYN_SMEM(unsigned char, smem_raw);
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    static_assert(!RES || (F16 && EPW == 16), "the resident-frame-tile form exists for the fp16-piece GEMM only");
    constexpr int NST = RES ? KC3_ASTAGES : KC2_STAGES;
    uint64_t* bars = (uint64_t*)(smem + (RES ? KC3_RING_BYTES : KC2_STAGES * KC2_STAGE_BYTES));
    uint64_t* full_bar = bars;                     // [STAGES]  (leader's copy is used) TMA of both CTAs -> leader MMA
    uint64_t* empty_bar = bars + NST;              // [STAGES]  leader MMA -> TMA producer of each CTA (multicast commit)
    uint64_t* tfull_bar = bars + 2 * NST;          // [2]       leader MMA -> epilogue of each CTA (multicast commit)
    uint64_t* tempty_bar = tfull_bar + 2;          // [2]       (leader's copy) epilogue warps of both CTAs -> leader MMA
    uint64_t* bfull_bar = tempty_bar + 2;          // [2] RES:  frame tile landed (leader's copy)
    uint64_t* bempty_bar = bfull_bar + 2;          // [2] RES:  leader MMA -> TMA producer of each CTA: the tile's last MMA has completed
    uint32_t* tmem_base_s = (uint32_t*)(bempty_bar + 2);

    // warp index broadcast from lane 0: the role branches below are then provably warp-uniform for ptxas (uniform registers stay usable)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int M = B * (Tm + 2) - 2;
    const int f_tiles = (M + 255) / 256;
    const int n_pairs = KCN / 256;                 // 97
    const int items_per_blk = n_pairs * f_tiles;
    const int total_items = NBLK * items_per_blk;
    const int pair_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    // items of this CTA pair: strided over the pairs (all pairs on the same frame tile at the same time), or -- RES -- one contiguous range
    const int i_lo = RES ? (int)((long long)total_items * pair_id / n_clusters) : pair_id;
    const int i_hi = RES ? (int)((long long)total_items * (pair_id + 1) / n_clusters) : total_items;
    const int i_step = RES ? 1 : n_clusters;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 2 * EPW); mbar_init(&bfull_bar[a], 1); mbar_init(&bempty_bar[a], 1); }
        mbar_init_fence();
    }
    if (warp == 0) {   // collective over the pair: the same warp of both CTAs
        tmem_alloc_2sm(tmem_base_s, 512u);
    }
    tc_fence_before();
    cluster_sync_all();     // barriers of both CTAs initialised and TMEM allocated before any remote access
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;
    pdl_wait();   // programmatic dependent launch: everything above touched constants, shared memory and TMEM only

    if (warp == 0) {
        // ================= TMA producer (each CTA loads its own A rows and its half of the frames) =================
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            [[maybe_unused]] int it_no = 0;
            [[maybe_unused]] int cur_tile = -1;
            [[maybe_unused]] uint32_t n_tiles_loaded = 0;
            for (int item = i_lo; item < i_hi; item += i_step, ++it_no) {
                const int blk = item / items_per_blk, r = item % items_per_blk;
                const int ft = r / n_pairs, nt = (r % n_pairs) * 2 + (int)rank;
                const CUtensorMap* wh = &maps.w_hi[blk]; const CUtensorMap* wl = &maps.w_lo[blk];
                KC_STAMP(0, it_no, 0);
                const CUtensorMap* hh = &maps.h_hi[blk]; const CUtensorMap* hl = &maps.h_lo[blk];
                if constexpr (RES) {
                    if (blk * f_tiles + ft != cur_tile) {   // next frame tile: 130 rows per piece into the buffer the MMA issuer has released
                        cur_tile = blk * f_tiles + ft;
                        const uint32_t bb = n_tiles_loaded & 1u, bph = (n_tiles_loaded >> 1) & 1u;
                        mbar_wait(&bempty_bar[bb], bph ^ 1);
                        unsigned char* bt = smem + KC3_ASTAGES * KC3_A_STAGE + bb * KC3_B_BUF;
                        if (rank == 0) mbar_expect_tx(&bfull_bar[bb], 2u * 2u * (uint32_t)(KC3_BROWS * 128));
                        tma_load_2d_2sm(bt, hh, 0, ft * 256 + (int)rank * 128, &bfull_bar[bb]);
                        tma_load_2d_2sm(bt + KC3_B_PIECE, hl, 0, ft * 256 + (int)rank * 128, &bfull_bar[bb]);
                        ++n_tiles_loaded;
                    }
                    for (int a = 0; a < NATOM; ++a) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        KC_STAMP(0, it_no, 1 + a);
                        unsigned char* st = smem + stage * KC3_A_STAGE;
                        if (rank == 0) mbar_expect_tx(&full_bar[stage], 2u * (uint32_t)KC3_A_STAGE);
                        tma_load_2d_2sm(st, wh, a * 32, nt * 128, &full_bar[stage]);
                        tma_load_2d_2sm(st + KC2_A_BYTES, wl, a * 32, nt * 128, &full_bar[stage]);
                        if (++stage == NST) { stage = 0; phase ^= 1; }
                    }
                } else
                for (int a = 0; a < NATOM; ++a) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    KC_STAMP(0, it_no, 1 + a);
                    unsigned char* st = smem + stage * KC2_STAGE_BYTES;
                    // timing experiment (exp_mask & 16, WRONG results): weights loaded for the first tile only -> half the operand feed
                    const bool skip_a = (exp_mask & 16) && item != pair_id;
                    if (rank == 0) mbar_expect_tx(&full_bar[stage], 2u * (skip_a ? 2u * KC2_B_BYTES : (three_pass ? KC2_STAGE_BYTES : (KC2_A_BYTES + KC2_B_BYTES))));
                    const int frow = ft * 256 + (int)rank * 128 + (F16 ? a : (a >> 1));
                    const int fcol = F16 ? 0 : (a & 1) * 32;
                    if (!skip_a) tma_load_2d_2sm(st, wh, a * 32, nt * 128, &full_bar[stage]);
                    tma_load_2d_2sm(st + 2 * KC2_A_BYTES, hh, fcol, frow, &full_bar[stage]);
                    if (three_pass) {
                        if (!skip_a) tma_load_2d_2sm(st + KC2_A_BYTES, wl, a * 32, nt * 128, &full_bar[stage]);
                        tma_load_2d_2sm(st + 2 * KC2_A_BYTES + KC2_B_BYTES, hl, fcol, frow, &full_bar[stage]);
                    }
                    if (++stage == KC2_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: leader CTA only =================
        if (rank == 0 && elect_one()) {
            constexpr uint32_t idesc = F16 ? umma_idesc_f16(256, 256) : umma_idesc_tf32(256, 256);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            [[maybe_unused]] int it_no = 0;
            [[maybe_unused]] int cur_tile = -1;
            [[maybe_unused]] uint32_t n_tiles_used = 0, bb = 0;
            auto tile_of = [&](int item) -> int { return item < i_hi ? (item / items_per_blk) * f_tiles + (item % items_per_blk) / n_pairs : -1; };
            for (int item = i_lo; item < i_hi; item += i_step, ++it_no) {
                KC_STAMP(1, it_no, 0);
                if constexpr (RES) {
                    if (tile_of(item) != cur_tile) {
                        cur_tile = tile_of(item);
                        bb = n_tiles_used & 1u;
                        mbar_wait(&bfull_bar[bb], (n_tiles_used >> 1) & 1u);
                        ++n_tiles_used;
                    }
                }
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                KC_STAMP(1, it_no, 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                if constexpr (RES) {
                    const uint32_t bt = smem_u32(smem + KC3_ASTAGES * KC3_A_STAGE) + bb * KC3_B_BUF;
                    for (int a = 0; a < NATOM; ++a) {
                        mbar_wait(&full_bar[stage], phase);
                        KC_STAMP(1, it_no, 2 + a);
                        tc_fence_after();
                        const uint32_t st = smem_u32(smem + stage * KC3_A_STAGE);
                        const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + KC2_A_BYTES);
                        // tap a of the im2col = the tile's rows shifted by a: start address + a * 128 B, base_offset 0
                        const uint64_t b_hi = umma_desc_sw128(bt + a * 128), b_lo = umma_desc_sw128(bt + KC3_B_PIECE + a * 128);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t adv = (uint64_t)(k * 2);
                            umma_f16_2sm(d_tmem, a_hi + adv, b_hi + adv, idesc, (a | k) ? 1u : 0u);
                            umma_f16_2sm(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                            umma_f16_2sm(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
                        }
                        tc_commit_2sm(&empty_bar[stage]);
                        if (++stage == NST) { stage = 0; phase ^= 1; }
                    }
                    if (tile_of(item + 1) != cur_tile) tc_commit_2sm(&bempty_bar[bb]);   // the tile's last item: its buffer goes back once these MMAs are done
                } else
                for (int a = 0; a < NATOM; ++a) {
                    mbar_wait(&full_bar[stage], phase);
                    KC_STAMP(1, it_no, 2 + a);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + stage * KC2_STAGE_BYTES);
                    const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + KC2_A_BYTES);
                    const uint64_t b_hi = umma_desc_sw128(st + 2 * KC2_A_BYTES), b_lo = umma_desc_sw128(st + 2 * KC2_A_BYTES + KC2_B_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t adv = (uint64_t)(k * 2);
                        if (F16) {
                            umma_f16_2sm(d_tmem, a_hi + adv, b_hi + adv, idesc, (a | k) ? 1u : 0u);
                            if (three_pass && !(exp_mask & 2)) {
                                umma_f16_2sm(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                                umma_f16_2sm(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
                            }
                        } else {
                            umma_tf32_2sm(d_tmem, a_hi + adv, b_hi + adv, idesc, (a | k) ? 1u : 0u);
                            if (three_pass) {
                                umma_tf32_2sm(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                                umma_tf32_2sm(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
                            }
                        }
                    }
                    tc_commit_2sm(&empty_bar[stage]);
                    if (++stage == KC2_STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit_2sm(&tfull_bar[acc]);
                KC_STAMP(1, it_no, 5);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ================= epilogue (both CTAs): own TMEM (128 n rows x 256 frames) -> (+bias) -> global =================
        const int q = warp & 3;
        const int cpart = (warp - 2) >> 2;
        uint32_t acc = 0, acc_phase = 0;
        [[maybe_unused]] int it_no = 0;
        // the bias of the NEXT item is loaded one item ahead: a dependent global load at the top of every item (~1,000 of the ~7,000 cycles an
        // item's epilogue took) sat between the accumulator hand-over and the first store (round-2 GEMM timeline)
        auto bias_of = [&](int item) -> float {
            if (item >= i_hi) return 0.f;
            const int blk = item / items_per_blk, nt = ((item % items_per_blk) % n_pairs) * 2 + (int)rank;
            const float* bias = blk == 0 ? bias0 : (blk == 1 ? bias1 : bias2);
            return bias[nt * 128 + q * 32 + lane];
        };
        float bv_next = bias_of(i_lo);
        for (int item = i_lo; item < i_hi; item += i_step, ++it_no) {
            const int blk = item / items_per_blk, r = item % items_per_blk;
            const int ft = r / n_pairs, nt = (r % n_pairs) * 2 + (int)rank;
            const int n = nt * 128 + q * 32 + lane;
            if (warp == 2 && lane == 0) KC_STAMP(2, it_no, 0);
            const float bv = bv_next;
            bv_next = bias_of(item + i_step);
            const float inv = F16 ? (blk == 0 ? inv0 : (blk == 1 ? inv1 : inv2)) : 1.f;
            float* kern = kern_all + (size_t)blk * B * Tm * KCN;
            // F16, blocks 1 and 2 (tensor-core LVC consumers): the predicted kernel w is written as fp16 pieces of w*S16_KERN straight
            // into the SWIZZLE_128B smem image of the LVC B operand -- per (layer, tap) 64 rows (o) of 128 B = [32 i hi | 32 i lo],
            // 16-byte chunk c at position c ^ (o & 7) -- in the SAME 24,576 bytes the fp32 image occupies.  Lane n holds element
            // (l, k, o, i) and the 32 lanes of a warp the 32 i of one row.  The 64 biases per layer stay fp32.
            const bool pieces = F16 && (B0P || blk >= 1) && !(exp_mask & 32);   // exp 32: timing experiment, fp32 full-line stores for every block (WRONG for the LVC consumer)
            const int rem = n % KPL;
            const bool is_w = rem < KK * LVC_OUT;          // warp-uniform: 6144 and KPL are multiples of 32
            // two 16-bit stores per value (a warp covers the 64 contiguous bytes of the hi half and of the lo half of one row: full sectors)
            int word = n, hw_hi = 0, hw_lo = 0;             // 32-bit word / 16-bit halfword indices inside the frame record
            if (pieces && is_w) {
                const int ko = rem >> 5, oo = ko & 63, ci = ((((rem >> 2) & 7) ^ (oo & 7)) << 2) + (rem & 3);
                if ((exp_mask & 64) && blk >= 1) {
                    // merged-N image (k_lvc_p, blocks 1 and 2), all SWIZZLE_128B: per layer
                    //   T01 [128 rows x 128 B]: row R = piece * 64 + o holds [tap 0: 32 i | tap 1: 32 i] of that piece, chunk c at c ^ (R & 7) -- the B
                    //       operand of ONE N = 128 MMA per activation piece and K-slice for taps 0 and 1 (hi and lo kernel products side by side);
                    //   T2  [64 rows x 128 B]: row o = [32 i hi | 32 i lo] of tap 2, chunk c at c ^ (o & 7) (three-pass form).
                    // (A 64-byte-row SWIZZLE_64B image holding all three taps merged was measured first: its B rows fetch ~2.5x slower.)
                    // A warp (the 32 i of one (k, o)) still writes two runs of 64 contiguous, 64-byte-aligned bytes.
                    const int tap = ko >> 6, sw = oo & 7;
                    if (tap < 2) {
                        const int base = 2 * (n - rem) + oo * 64 + ((((tap << 2) + (ci >> 3)) ^ sw) << 3) + (ci & 7);
                        hw_hi = base;
                        hw_lo = base + 64 * 64;
                    } else {
                        const int base = 2 * (n - rem) + 8192 + oo * 64 + (ci & 7);
                        hw_hi = base + (((ci >> 3) ^ sw) << 3);
                        hw_lo = base + (((4 + (ci >> 3)) ^ sw) << 3);
                    }
                } else {
                const int base = 2 * ((n - rem) + ko * 32) + (ci & 7);
                hw_hi = base + (((ci >> 3) ^ (oo & 7)) << 3);
                hw_lo = base + (((4 + (ci >> 3)) ^ (oo & 7)) << 3);
                }
            }
            const float inv_s = inv * S16_KERN, bv_s = bv * S16_KERN;
            // warp-uniform; broadcast from lane 0 so that ptxas KNOWS it (otherwise every store below re-materialises its uniform
            // memory descriptor with two R2UR: 40% of the epilogue's instructions)
            const bool as_pieces = __shfl_sync(0xffffffffu, (int)(pieces && is_w), 0) != 0;
            auto put_pieces = [&](uint16_t* ph, uint16_t* pl, float accv) {
                const float sv = fmaf(accv, inv_s, bv_s);
                const uint16_t h16 = f16_sat_bits(sv);
                const uint16_t l16 = f16_sat_bits(sv - f16_bits_to_float(h16));
#if KC_STORE_CS && !defined(FD_EMU)
                asm volatile("st.global.cs.b16 [%0], %1;" ::"l"(ph), "h"(h16) : "memory");
                asm volatile("st.global.cs.b16 [%0], %1;" ::"l"(pl), "h"(l16) : "memory");
#else
                *ph = h16;
                *pl = l16;
#endif
            };
            int p = ft * 256 + cpart * CPW;
            int center = p + 1, bb = center / (Tm + 2), fp = center % (Tm + 2);
            const bool fast = __shfl_sync(0xffffffffu, (int)((fp >= 1) && (fp + CPW - 1 <= Tm) && (p + CPW - 1 < M)), 0) != 0;
            mbar_wait(&tfull_bar[acc], acc_phase);
            if (warp == 2 && lane == 0) KC_STAMP(2, it_no, 1);
            tc_fence_after();
            // The accumulator stage goes back to the MMA issuer as soon as this warp's LAST tcgen05.ld has completed -- half an item's stores
            // earlier than at the end of the item: the two accumulator stages both take about the store time, so every cycle of hand-over counts.
            auto release_acc = [&]() {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(&tempty_bar[acc]);
            };
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256 + cpart * CPW;
            if (exp_mask & 1) {
                release_acc();
                // timing experiment: no epilogue work at all
            } else if (exp_mask & 4) {
                // timing experiment: TMEM loads + arithmetic, no global stores (the store is predicated on a value that never occurs)
#pragma unroll 1
                for (int c0 = 0; c0 < CPW; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + c0, v);
                    tmem_ld_wait();
                        if (c0 + 32 >= CPW) release_acc();   // the accumulator is in registers: hand it back before the stores
                    float acc2 = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc2 += fmaf(__uint_as_float(v[j]), inv_s, bv_s);
                    if (acc2 == 1.2345e-33f) kern[n] = acc2;
                }
            } else if ((exp_mask & 8) && fast) {
                // timing experiment (WRONG data layout): the same bytes and records, but written as 16-byte stores -- lane l writes one
                // 16-B chunk of the record of frame j0 + (l & 7): 1/8 of the store instructions of the real path
                float* o = kern + ((size_t)bb * Tm + (fp - 1)) * KCN + (n & ~31) + ((lane >> 3) << 2);
#pragma unroll 1
                for (int c0 = 0; c0 < CPW; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + c0, v);
                    tmem_ld_wait();
                        if (c0 + 32 >= CPW) release_acc();   // the accumulator is in registers: hand it back before the stores
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        float* r = o + (size_t)(j + (lane & 7)) * KCN;
                        *reinterpret_cast<uint4*>(r) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        *reinterpret_cast<uint4*>(r + 16) = make_uint4(v[j + 4], v[j + 5], v[j + 6], v[j + 7]);
                    }
                    o += (size_t)32 * KCN;
                }
            } else if (fast) {
                float* o = kern + ((size_t)bb * Tm + (fp - 1)) * KCN;   // record of this warp's first frame; frames are KCN words apart
                if (as_pieces) {
                    uint16_t* ph = reinterpret_cast<uint16_t*>(o) + hw_hi;
                    uint16_t* pl = reinterpret_cast<uint16_t*>(o) + hw_lo;
                    if constexpr (CPW == 64 && RES && KC_EARLY_RELEASE) {
                        // both 32-column halves go to registers first: the accumulator stage returns to the MMA issuer ~half an item's store
                        // time earlier than with load / store / load / store (the hand-over, not the store rate, sets this kernel's pace)
                        uint32_t v0[32], v1[32];
                        tmem_ld_32x32b_x32(taddr, v0);
                        tmem_ld_32x32b_x32(taddr + 32, v1);
                        tmem_ld_wait();
                        release_acc();
                        if (warp == 2 && lane == 0) KC_STAMP(2, it_no, 2);
#pragma unroll
                        for (int j = 0; j < 32; ++j) put_pieces(ph + (size_t)j * (2 * KCN), pl + (size_t)j * (2 * KCN), __uint_as_float(v0[j]));
                        if (warp == 2 && lane == 0) KC_STAMP(2, it_no, 3);
                        ph += (size_t)64 * KCN; pl += (size_t)64 * KCN;
#pragma unroll
                        for (int j = 0; j < 32; ++j) put_pieces(ph + (size_t)j * (2 * KCN), pl + (size_t)j * (2 * KCN), __uint_as_float(v1[j]));
                        if (warp == 2 && lane == 0) KC_STAMP(2, it_no, 5);
                    } else
#pragma unroll 1
                    for (int c0 = 0; c0 < CPW; c0 += 32) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(taddr + c0, v);
                        tmem_ld_wait();
                        if (c0 + 32 >= CPW) release_acc();   // the accumulator is in registers: hand it back before the stores
                        if (warp == 2 && lane == 0) KC_STAMP(2, it_no, 2 + (c0 >> 5) * 2);
#pragma unroll
                        for (int j = 0; j < 32; ++j) put_pieces(ph + (size_t)j * (2 * KCN), pl + (size_t)j * (2 * KCN), __uint_as_float(v[j]));
                        if (warp == 2 && lane == 0) KC_STAMP(2, it_no, 3 + (c0 >> 5) * 2);
                        ph += (size_t)64 * KCN; pl += (size_t)64 * KCN;
                    }
                } else {
                    o += word;
                    if constexpr (CPW == 64 && RES && KC_EARLY_RELEASE) {
                        uint32_t v0[32], v1[32];
                        tmem_ld_32x32b_x32(taddr, v0);
                        tmem_ld_32x32b_x32(taddr + 32, v1);
                        tmem_ld_wait();
                        release_acc();
#pragma unroll
                        for (int j = 0; j < 32; ++j) o[(size_t)j * KCN] = F16 ? fmaf(__uint_as_float(v0[j]), inv, bv) : __uint_as_float(v0[j]) + bv;
                        o += (size_t)32 * KCN;
#pragma unroll
                        for (int j = 0; j < 32; ++j) o[(size_t)j * KCN] = F16 ? fmaf(__uint_as_float(v1[j]), inv, bv) : __uint_as_float(v1[j]) + bv;
                    } else
#pragma unroll 1
                    for (int c0 = 0; c0 < CPW; c0 += 32) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(taddr + c0, v);
                        tmem_ld_wait();
                        if (c0 + 32 >= CPW) release_acc();   // the accumulator is in registers: hand it back before the stores
#pragma unroll
                        for (int j = 0; j < 32; ++j) o[(size_t)j * KCN] = F16 ? fmaf(__uint_as_float(v[j]), inv, bv) : __uint_as_float(v[j]) + bv;
                        o += (size_t)32 * KCN;
                    }
                }
            } else if (Tm + 2 > 34) {
                // GAP path: the 64-frame span crosses an utterance boundary (30 % of the items at T' = 861 have one such warp group) or the end of the
                // tensor.  Rows of padded index p map to frame records consecutively except for the two pad rows between utterances, so per 32-row
                // sub-chunk there is one record base, at most one gap (rows j1, j1 + 1 skipped, later rows shifted by two records), possibly a leading
                // pad row and a row limit -- all warp-uniform.  (The generic per-row walk below cost ~14,000 cycles per item against ~5,000 for the
                // straight path and set the pace of the whole kernel: round-2 GEMM timeline.)
                auto gap_chunk = [&](const int c0, const uint32_t (&v)[32]) {
                    // (broadcast from lane 0 so that ptxas KNOWS these are warp-uniform: uniform predicates and branches, no per-store R2UR)
                    const int pc = __shfl_sync(0xffffffffu, p + c0, 0), cen = pc + 1, bbc = cen / (Tm + 2), fpc = cen % (Tm + 2);
                    const int j1 = Tm + 1 - fpc, jlo = fpc == 0 ? 1 : 0, jhi = M - pc;   // end pad row; leading pad row; rows past the end
                    float* rec0 = kern + ((long long)bbc * Tm + fpc - 1) * KCN;           // record of row 0 (never dereferenced when row 0 is a pad row)
                    // rows before the gap: record row0 + j; rows after it: row0 + j - 2 -- two base pointers, constant offsets, uniform predicates
                    const int ja = j1 < jhi ? j1 : jhi;        // rows [jlo, ja) via base A; rows (j1 + 1, jhi) via base B
                    // most sub-chunks of a gap span have no gap INSIDE: all 32 rows on one base -> the straight store sequence
                    const bool all_a = jlo == 0 && j1 >= 32 && jhi >= 32, all_b = j1 + 1 < 0 && jhi >= 32;
                    if (as_pieces && (all_a || all_b)) {
                        uint16_t* ph1 = reinterpret_cast<uint16_t*>(rec0) + hw_hi - (all_b ? 4 * KCN : 0);
                        uint16_t* pl1 = reinterpret_cast<uint16_t*>(rec0) + hw_lo - (all_b ? 4 * KCN : 0);
#pragma unroll
                        for (int j = 0; j < 32; ++j) put_pieces(ph1 + (size_t)j * (2 * KCN), pl1 + (size_t)j * (2 * KCN), __uint_as_float(v[j]));
                    } else if (as_pieces) {
                        uint16_t* pha = reinterpret_cast<uint16_t*>(rec0) + hw_hi;
                        uint16_t* pla = reinterpret_cast<uint16_t*>(rec0) + hw_lo;
                        uint16_t* phb = pha - 4 * KCN;
                        uint16_t* plb = pla - 4 * KCN;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float sv = fmaf(__uint_as_float(v[j]), inv_s, bv_s);
                            const uint16_t h16 = f16_sat_bits(sv);
                            const uint16_t l16 = f16_sat_bits(sv - f16_bits_to_float(h16));
                            if (j >= jlo && j < ja) { pha[(size_t)j * (2 * KCN)] = h16; pla[(size_t)j * (2 * KCN)] = l16; }
                            else if (j > j1 + 1 && j < jhi) { phb[(size_t)j * (2 * KCN)] = h16; plb[(size_t)j * (2 * KCN)] = l16; }
                        }
                    } else {
                        float* oa = rec0 + word;
                        float* ob = oa - 2 * KCN;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float val = F16 ? fmaf(__uint_as_float(v[j]), inv, bv) : __uint_as_float(v[j]) + bv;
                            if (j >= jlo && j < ja) oa[(size_t)j * KCN] = val;
                            else if (j > j1 + 1 && j < jhi) ob[(size_t)j * KCN] = val;
                        }
                    }
                };
                if constexpr (CPW == 64 && RES && KC_EARLY_RELEASE) {
                    uint32_t v0[32], v1[32];
                    tmem_ld_32x32b_x32(taddr, v0);
                    tmem_ld_32x32b_x32(taddr + 32, v1);
                    tmem_ld_wait();
                    release_acc();
                    gap_chunk(0, v0);
                    gap_chunk(32, v1);
                } else {
#pragma unroll 1
                    for (int c0 = 0; c0 < CPW; c0 += 32) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(taddr + c0, v);
                        tmem_ld_wait();
                        if (c0 + 32 >= CPW) release_acc();   // the accumulator is in registers: hand it back before the stores
                        gap_chunk(c0, v);
                    }
                }
            } else {   // tiny utterances (several boundaries per sub-chunk): generic per-row walk
#pragma unroll 1
                for (int c0 = 0; c0 < CPW; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + c0, v);
                    tmem_ld_wait();
                        if (c0 + 32 >= CPW) release_acc();   // the accumulator is in registers: hand it back before the stores
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (p < M && fp >= 1 && fp <= Tm) {   // uniform across the warp (depends on the column only)
                            float* rec = kern + ((size_t)bb * Tm + (fp - 1)) * KCN;
                            if (as_pieces) put_pieces(reinterpret_cast<uint16_t*>(rec) + hw_hi, reinterpret_cast<uint16_t*>(rec) + hw_lo, __uint_as_float(v[j]));
                            else rec[word] = F16 ? fmaf(__uint_as_float(v[j]), inv, bv) : __uint_as_float(v[j]) + bv;
                        }
                        ++p;
                        if (++fp == Tm + 2) { fp = 0; ++bb; }
                    }
                }
            }
            if (warp == 2 && lane == 0) KC_STAMP(2, it_no, 6);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    cluster_sync_all();     // no CTA of the pair exits (or frees TMEM) while the other may still touch its smem / barriers
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, 512u);
    }
}

#ifndef FD_EMU
// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcState {
    int device = 0, sm_count = 148;
    PFN_encodeTiled encode = nullptr;
    const float* blob = nullptr;
    uint64_t sec_off[FD_S_COUNT];
    CUtensorMap w_hi[NBLK], w_lo[NBLK];
    CUtensorMap w16_hi[NBLK], w16_lo[NBLK];   // fp16 pieces (LBn_KCT_F16): rows of 96 fp32-sized elements = 192 fp16
    CUtensorMap w16p_hi, w16p_lo;             // block 0 in image row order (LB0_KCT_F16P; experimental, built on first use)
    int b0p_ready = 0;
    float scales16[64];                       // host copy of section SCALES16
    int kc_2cta = 1;       // kernel_conv GEMM on CTA pairs (cta_group::2, default); option "kc_2cta" = 0 selects the 1-CTA kernel
    int kc_exp = 0;        // timing experiments only (option "kc_exp"): 1 = epilogue does nothing, 2 = hi*hi MMAs only (WRONG results)
    int lvc_exp = 0;       // timing experiments only (option "lvc_exp"): 1 = no second conv pass, 2 / 4 = hi*hi only in the LVC / conv (WRONG results)
    int lvc_groups = 2;    // tc_3xf16, block 2: independent 8-warp groups per CTA (2 or 3; option "lvc_groups")
    int b0_attr_set = 0;   // experimental block-0 kernel: attribute set on first use
    int lvc_swizzle = 0;   // LVC operand tiles: 0 = no-swizzle panels, 1 = SWIZZLE_128B + base_offset, 2 = SWIZZLE_128B, base_offset 0
    bool ok = false;
};

static inline int tc_make_map_2d(TcState* s, CUtensorMap* m, const float* base, uint64_t cols, uint64_t rows, uint64_t row_stride_bytes,
                                 uint32_t box_cols, uint32_t box_rows, std::string& err) {
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = s->encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"; return -3; }
    return 0;
}

static inline cudaError_t tc_set_lvc_attrs();
static inline void tc_destroy(void* st) { delete (TcState*)st; }
static inline bool tc_available(void* st) { return st && ((TcState*)st)->ok; }
static inline void tc_set_lvc_swizzle(void* st, int v) { if (st) ((TcState*)st)->lvc_swizzle = v; }
static inline void tc_set_kc_2cta(void* st, int v) { if (st) ((TcState*)st)->kc_2cta = v; }
static inline void tc_set_lvc_groups(void* st, int v) { if (st) ((TcState*)st)->lvc_groups = v; }
static inline void tc_set_lvc_exp(void* st, int v) { if (st) ((TcState*)st)->lvc_exp = v; }
static inline void tc_set_kc_exp(void* st, int v) { if (st) ((TcState*)st)->kc_exp = v; }

static inline int tc_init(void** state, int device, const float* blob, const uint64_t* sec_off, std::string& err) {
    TcState* old = (TcState*)*state;
    TcState* s = new TcState();
    if (old) {   // a weight reload keeps the tuning options set through fd_set_option
        s->lvc_swizzle = old->lvc_swizzle; s->kc_2cta = old->kc_2cta; s->lvc_groups = old->lvc_groups; s->lvc_exp = old->lvc_exp; s->kc_exp = old->kc_exp;
    }
    tc_destroy(old);
    *state = s;
    s->device = device; s->blob = blob;
    for (int i = 0; i < FD_S_COUNT; ++i) s->sec_off[i] = sec_off[i];
    cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, device);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
        err = "cuTensorMapEncodeTiled entry point unavailable"; return -3;
    }
    s->encode = (PFN_encodeTiled)fn;
    for (int n = 0; n < NBLK; ++n) {
        if (tc_make_map_2d(s, &s->w_hi[n], blob + sec_off[FD_S_LB0_KCT_HI + n * FD_LB_STRIDE], KCK, KCN, KCK * 4, KCG_KATOM, KCG_BM, err)) return -3;
        if (tc_make_map_2d(s, &s->w_lo[n], blob + sec_off[FD_S_LB0_KCT_LO + n * FD_LB_STRIDE], KCK, KCN, KCK * 4, KCG_KATOM, KCG_BM, err)) return -3;
    }
    for (int n = 0; n < NBLK; ++n) {
        const float* w16 = blob + sec_off[FD_S_LB0_KCT_F16 + n];
        if (tc_make_map_2d(s, &s->w16_hi[n], w16, KCK / 2, KCN, KCK * 2, KCG_KATOM, KCG_BM, err)) return -3;
        if (tc_make_map_2d(s, &s->w16_lo[n], w16 + (size_t)KCN * (KCK / 2), KCK / 2, KCN, KCK * 2, KCG_KATOM, KCG_BM, err)) return -3;
    }
    if (cudaMemcpy(s->scales16, blob + sec_off[FD_S_SCALES16], sizeof s->scales16, cudaMemcpyDeviceToHost) != cudaSuccess) {
        err = "reading SCALES16 failed"; return -3;
    }
    if (cudaFuncSetAttribute(k_kc_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, KCG_SMEM_BYTES) != cudaSuccess ||
        tc_set_lvc_attrs() != cudaSuccess) {
        err = "cudaFuncSetAttribute(tensor-core kernels) failed"; return -3;
    }
    s->ok = true;
    return 0;
}

// hk_hi / hk_lo: (3, B, T'+2, 64) each, written by k_kp_hidden.
// kimg (mode tc_3xf16): 1 = blocks 1, 2 written in the merged-N image of k_lvc_p (exp_mask bit 64), 0 = the [hi | lo] row image of k_lvc_layer_h
static inline int tc_kc_gemm(void* state, int mode, const float* hk_hi, const float* hk_lo, float* kern, int B, int Tm, cudaStream_t st,
                             std::string& err, uint64_t* launches, int b0_pieces = 0, int kimg = 0, int res = 0, int max_clusters = 0) {
    TcState* s = (TcState*)state;
    if (!s || !s->ok) { err = "tensor-core path not initialised"; return -4; }
    if (b0_pieces && mode == FD_MODE_TC_3XF16 && !s->b0p_ready) {   // experimental path: maps + attribute on first use only
        const float* w16 = s->blob + s->sec_off[FD_S_LB0_KCT_F16P];
        if (tc_make_map_2d(s, &s->w16p_hi, w16, KCK / 2, KCN, KCK * 2, KCG_KATOM, KCG_BM, err)) return -3;
        if (tc_make_map_2d(s, &s->w16p_lo, w16 + (size_t)KCN * (KCK / 2), KCK / 2, KCN, KCK * 2, KCG_KATOM, KCG_BM, err)) return -3;
        cudaError_t ea = cudaFuncSetAttribute(k_kc_gemm_tc2<true, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, KC2_SMEM_BYTES);
        if (ea != cudaSuccess) { err = std::string("k_kc_gemm_tc2<f16, b0 pieces>: shared-memory attribute: ") + cudaGetErrorString(ea); return -3; }
        ea = cudaFuncSetAttribute(k_kc_gemm_tc2<true, 16, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, KC3_SMEM_BYTES);
        if (ea != cudaSuccess) { err = std::string("k_kc_gemm_tc2<f16, b0 pieces, resident tile>: shared-memory attribute: ") + cudaGetErrorString(ea); return -3; }
        s->b0p_ready = 1;
    }
    KcgMaps maps;
    const uint64_t rows = (uint64_t)B * (Tm + 2);
    for (int n = 0; n < NBLK; ++n) {
        maps.w_hi[n] = s->w_hi[n]; maps.w_lo[n] = s->w_lo[n];
        if (tc_make_map_2d(s, &maps.h_hi[n], hk_hi + (size_t)n * rows * HID, HID, rows, HID * 4, KCG_KATOM, KCG_BN, err)) return -3;
        if (tc_make_map_2d(s, &maps.h_lo[n], hk_lo + (size_t)n * rows * HID, HID, rows, HID * 4, KCG_KATOM, KCG_BN, err)) return -3;
    }
    const int M = B * (Tm + 2) - 2;
    if (mode == FD_MODE_TC_3XF16) {   // fp16 pieces: hk rows are 64 fp16 = 32 fp32-sized elements = one 128-byte k-atom
        const bool res_form = res != 0;
        for (int n = 0; n < NBLK; ++n) {
            maps.w_hi[n] = s->w16_hi[n]; maps.w_lo[n] = s->w16_lo[n];
            // (resident form: one 130-row box per frame tile and piece -- the three im2col taps are shifted views of it)
            if (tc_make_map_2d(s, &maps.h_hi[n], hk_hi + (size_t)n * rows * (HID / 2), HID / 2, rows, HID * 2, KCG_KATOM, res_form ? KC3_BROWS : 128, err)) return -3;
            if (tc_make_map_2d(s, &maps.h_lo[n], hk_lo + (size_t)n * rows * (HID / 2), HID / 2, rows, HID * 2, KCG_KATOM, res_form ? KC3_BROWS : 128, err)) return -3;
        }
        const int items = NBLK * (KCN / 256) * ((M + 255) / 256);
        int clusters = s->sm_count / 2; if (items < clusters) clusters = items;
        if (max_clusters > 0 && clusters > max_clusters) clusters = max_clusters;
        float inv[NBLK];
        for (int n = 0; n < NBLK; ++n) inv[n] = 1.f / (s->scales16[n] * S16_HK);
        if (b0_pieces && res_form) {
            maps.w_hi[0] = s->w16p_hi; maps.w_lo[0] = s->w16p_lo;
            fd_launch_pdl(k_kc_gemm_tc2<true, 16, true, true>, dim3(2 * clusters), dim3(64 + 32 * 16), KC3_SMEM_BYTES, st, maps, s->blob + s->sec_off[FD_S_LB0_KC_BP], s->blob + s->sec_off[FD_S_LB1_KC_B],
                                                                      s->blob + s->sec_off[FD_S_LB2_KC_B], kern, B, Tm, 1, inv[0], inv[1], inv[2], s->kc_exp | (kimg ? 64 : 0));
        } else if (b0_pieces) {
            maps.w_hi[0] = s->w16p_hi; maps.w_lo[0] = s->w16p_lo;
            fd_launch_pdl(k_kc_gemm_tc2<true, 16, true>, dim3(2 * clusters), dim3(64 + 32 * 16), KC2_SMEM_BYTES, st, maps, s->blob + s->sec_off[FD_S_LB0_KC_BP], s->blob + s->sec_off[FD_S_LB1_KC_B],
                                                                      s->blob + s->sec_off[FD_S_LB2_KC_B], kern, B, Tm, 1, inv[0], inv[1], inv[2], s->kc_exp | (kimg ? 64 : 0));
        } else if (res_form) {
            fd_launch_pdl(k_kc_gemm_tc2<true, 16, false, true>, dim3(2 * clusters), dim3(64 + 32 * 16), KC3_SMEM_BYTES, st, maps, s->blob + s->sec_off[FD_S_LB0_KC_B], s->blob + s->sec_off[FD_S_LB1_KC_B],
                                                                  s->blob + s->sec_off[FD_S_LB2_KC_B], kern, B, Tm, 1, inv[0], inv[1], inv[2], s->kc_exp | (kimg ? 64 : 0));
        } else
        fd_launch_pdl(k_kc_gemm_tc2<true, 16>, dim3(2 * clusters), dim3(64 + 32 * 16), KC2_SMEM_BYTES, st, maps, s->blob + s->sec_off[FD_S_LB0_KC_B], s->blob + s->sec_off[FD_S_LB1_KC_B],
                                                                  s->blob + s->sec_off[FD_S_LB2_KC_B], kern, B, Tm, 1, inv[0], inv[1], inv[2], s->kc_exp | (kimg ? 64 : 0));
        cudaError_t e2 = cudaGetLastError();
        if (e2 != cudaSuccess) { err = std::string("launch of k_kc_gemm_tc2<f16> failed: ") + cudaGetErrorString(e2); return -3; }
        ++*launches;
        return 0;
    }
    if (s->kc_2cta) {
        for (int n = 0; n < NBLK; ++n) {   // frame boxes of 128 rows: each CTA of a pair loads half of the 256-frame tile
            if (tc_make_map_2d(s, &maps.h_hi[n], hk_hi + (size_t)n * rows * HID, HID, rows, HID * 4, KCG_KATOM, 128, err)) return -3;
            if (tc_make_map_2d(s, &maps.h_lo[n], hk_lo + (size_t)n * rows * HID, HID, rows, HID * 4, KCG_KATOM, 128, err)) return -3;
        }
        const int items = NBLK * (KCN / 256) * ((M + 255) / 256);
        int clusters = s->sm_count / 2; if (items < clusters) clusters = items;
        k_kc_gemm_tc2<false, 8><<<2 * clusters, 320, KC2_SMEM_BYTES, st>>>(maps, s->blob + s->sec_off[FD_S_LB0_KC_B], s->blob + s->sec_off[FD_S_LB1_KC_B],
                                                                   s->blob + s->sec_off[FD_S_LB2_KC_B], kern, B, Tm, mode == 1 ? 1 : 0, 1.f, 1.f, 1.f, 0);
        cudaError_t e2 = cudaGetLastError();
        if (e2 != cudaSuccess) { err = std::string("launch of k_kc_gemm_tc2 failed: ") + cudaGetErrorString(e2); return -3; }
        ++*launches;
        return 0;
    }
    const int total = NBLK * (KCN / KCG_BM) * ((M + KCG_BN - 1) / KCG_BN);
    const int grid = total < s->sm_count ? total : s->sm_count;
    k_kc_gemm_tc<<<grid, 320, KCG_SMEM_BYTES, st>>>(maps, s->blob + s->sec_off[FD_S_LB0_KC_B], s->blob + s->sec_off[FD_S_LB1_KC_B],
                                                    s->blob + s->sec_off[FD_S_LB2_KC_B], kern, B, Tm, mode == 1 ? 1 : 0);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("launch of k_kc_gemm_tc failed: ") + cudaGetErrorString(e); return -3; }
    ++*launches;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// K11+K12+K13  one LVC layer on tensor cores (modules.py:208-217), blocks 1 (hop 64) and 2 (hop 256).
//
// Per tile of 128 time steps (one batch item).  Every operand is a K-major SWIZZLE_128B tile: row r = 128 B (32 channels)
// at r*128, its 16-byte chunk c stored at chunk position c ^ (r & 7).  A dilated-conv tap is the SAME tile read from
// `start + shift*128 B` -- measured on B200: shifts that are not multiples of 8 rows are handled correctly with descriptor
// base_offset = 0 (the swizzle XOR uses absolute smem address bits), so the im2col over time costs nothing.
//   conv :  D1[128 rows x 32]  = sum_{k<3} A[rows + (k-1)*dil][32 ci] * Wc[k][32 co x 32 ci]^T    rows t0-1 .. t0+126;
//           the LVC also needs y at t0+127, t0+128: those 2 rows are computed with FFMA while the MMAs run
//   y = lrelu(D1 + b) -> tf32 pieces -> rows of the Y tile (aliasing the A tile)
//   lvc  :  D2[128 x 64] = sum_{k<3} Y[rows + k][32] * Wl[f][k][64 o x 32 i]^T   with the per-frame predicted kernels
//           (hop 64: two frames per tile, each run over the whole M-tile into its own TMEM columns; rows pick their frame)
//   out = xs + sigmoid(D2[:, :32] + b) * tanh(D2[:, 32:] + b),  xs = x + skip re-read from global
// 3xTF32: every MMA is issued for (hi,hi), (hi,lo), (lo,hi) into one accumulator.
//
// Data movement: the raw inputs of a tile -- x rows, skip rows (block 1) or the audio window (block 2), and the predicted
// kernels, which the KC GEMM already wrote in this tile layout -- are fetched by cp.async.bulk (one elected thread, mbarrier
// complete_tx) ONE TILE AHEAD, straight into the tile's own operand regions (x rows land row-major in the rows of the A_lo
// tile, skip rows in A_hi, kernels in LW_hi), and are then transformed in place (x+skip, lrelu, Veltkamp tf32 split; the 8
// lanes of a row read, __syncwarp, write the permuted chunks).  No register staging, no LDG queue pressure, no extra smem;
// the copies for tile i+1 are issued the moment tile i's LVC MMAs have completed and overlap its gate epilogue.
//
// One persistent CTA per SM with GROUPS independent 8-warp groups, each owning a tile slot (operand tiles, TMEM columns,
// mbarriers, a named barrier) and walking its own tile sequence, so one group's transforms / epilogues overlap the other's
// MMAs.  One elected lane of a group's warp 0 issues its MMAs and bulk copies.
// ---------------------------------------------------------------------------------------------------------
#endif  // !FD_EMU
constexpr int LT_TT = 128;
constexpr int LT_A_BYTES = 24576;                   // 192 rows x 128 B (rows 0..183 used: t0-28 .. t0+155)
constexpr int LT_CW_BYTES = 3 * C * 128;            // 12288
constexpr int LT_LW_BYTES = 3 * LVC_OUT * 128;      // 24576
constexpr int LT_AU = 192;                          // audio window (SKIP_FIRST): positions t0-32 .. t0+159
template <int HOP>
__host__ __device__ constexpr int lt_nf() { return LT_TT / HOP > 0 ? LT_TT / HOP : 1; }
template <int HOP>
__host__ __device__ constexpr int lt_slot_bytes() { return 2 * LT_A_BYTES + lt_nf<HOP>() * 2 * LT_LW_BYTES; }   // A hi | A lo | LW hi[NF] | LW lo[NF]
template <int HOP>
__host__ __device__ constexpr int lt_small_bytes() { return 2 * (lt_nf<HOP>() * 256 + LT_AU * 4); }            // double-buffered lbias | audio
constexpr int LT_SHARED_BYTES = 2 * LT_CW_BYTES + (7 * C + C + C) * 4 + 128;   // conv W pieces, first_w, first_b, conv_b, barriers + tmem ptr
template <int HOP, int GROUPS>
constexpr int lt_smem_bytes() { return GROUPS * (lt_slot_bytes<HOP>() + lt_small_bytes<HOP>()) + LT_SHARED_BYTES + 1024; }

__device__ __forceinline__ uint32_t swz128(int row, int c) { return (uint32_t)(row * 128 + ((c ^ (row & 7)) << 4)); }
#ifndef FD_EMU
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// The predicted-kernel tensor is read exactly once per reverse step (2 GB at config 2): an evict-first L2 policy on those loads keeps the stream
// from displacing the activation rows the next layer re-reads and the weight images (LVC_KERN_EF, measured: profiles/r02_energy_ab.txt).
#ifndef LVC_KERN_EF
#define LVC_KERN_EF 1
#endif
__device__ __forceinline__ void bulk_g2s_once(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
#if LVC_KERN_EF
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
#else
    bulk_g2s(smem_dst, gsrc, bytes, bar);
#endif
}
#endif  // !FD_EMU
// ---- small helpers shared by the kernels that also compile for the emulator ----
template <int N> __device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[N]);
template <> __device__ __forceinline__ void tmem_ld_cols<8>(uint32_t taddr, uint32_t (&v)[8]) { tmem_ld_32x32b_x8(taddr, v); }
template <> __device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, uint32_t (&v)[16]) { tmem_ld_32x32b_x16(taddr, v); }
template <> __device__ __forceinline__ void tmem_ld_cols<32>(uint32_t taddr, uint32_t (&v)[32]) { tmem_ld_32x32b_x32(taddr, v); }
// 8 consecutive floats (32-byte aligned) to global memory.  FD_VEC256 = 1 (experiment, off: not measured yet) uses the 256-bit store of
// sm_100 (STG.E.ENL2.256): the epilogues below write rows with one lane per row, so a 16-byte store fills half a 32-byte sector per
// lane and a warp store touches 32 half-sectors; the 256-bit form writes whole sectors with half the store instructions.
#ifndef FD_VEC256
#define FD_VEC256 1   // measured on B200 (round 2): upsample 0.79 -> 0.64, LVC block 2 4.47 -> 4.31 ms per N=4 call at config 2
#endif
__device__ __forceinline__ void st_global_f8(float* p, const float4 a, const float4 b) {
#if FD_VEC256 && !defined(FD_EMU)
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
#else
    *reinterpret_cast<float4*>(p) = a;
    *reinterpret_cast<float4*>(p + 4) = b;
#endif
}
// tf32 pieces by Veltkamp splitting (3 FP ops): hi = x rounded to nearest at 11 significant bits (low 13 mantissa bits
// zero -> exactly a tf32), lo = x - hi exactly.  lo is handed to the tensor core as is (its own tf32 conversion of lo costs
// <= 2^-11 |lo| <= 2^-22 |x|).  `cvt.rna.tf32.f32` is emulated with ~8 integer instructions on sm_100a (ncu: it was the
// majority of this kernel's instruction stream), and _rn intrinsics keep the compiler from contracting c - x into an FMA.
__device__ __forceinline__ float split_hi(float x) {
    const float c = __fmul_rn(x, 8193.0f);
    return __fsub_rn(c, __fsub_rn(c, x));
}
__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
    hi = make_float4(split_hi(v.x), split_hi(v.y), split_hi(v.z), split_hi(v.w));
    lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
}
#ifndef FD_EMU
// gate non-linearities from ex2.approx/rcp.approx: abs error ~2e-7 (the precise tanhf/expf forms cost ~10x the instructions)
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + exp2f(-1.4426950408889634f * x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.f - __fdividef(2.f, 1.f + exp2f(2.8853900817779268f * x)); }

// phase timeline of CTA 0 / group 0 (clock64 at phase boundaries) -- compiled in with -DFD_LVC_TIMELINE, read through
// fd_debug_read("lvc_timeline").
__device__ unsigned long long g_lvc_timeline[128];
#ifdef FD_LVC_TIMELINE
#define LT_STAMP(i) do { if (stamp && tile_no < 12) g_lvc_timeline[tile_no * 10 + (i)] = clock64(); } while (0)
#else
#define LT_STAMP(i) do { } while (0)
#endif

struct LvcTcParams {
    const float* cw_hi; const float* cw_lo;      // [3][32 co][8 chunks ^ (co&7)][4] tf32 pieces of this layer's dilated conv
    const float* conv_b;                         // [32]
    const float* first_w; const float* first_b;  // [7][32], [32]   (SKIP_FIRST)
};

// GW = warps per group (8 or 16): the single-group variant runs 16 so that its transform / epilogue phases take half as long.
template <int HOP, bool SKIP_FIRST, int GROUPS, int GW>
__global__ void __launch_bounds__(32 * GW * GROUPS, 1)
k_lvc_layer_tc(LvcTcParams p, const float* __restrict__ x_in, const float* __restrict__ skip, const float* __restrict__ kern,
               float* __restrict__ x_out, int B, int T, int Tm, int dil, int three_pass) {
    constexpr int NF = lt_nf<HOP>();
    constexpr int GT = 32 * GW;              // threads per group
    constexpr int NPART = GW / 4;            // warps per TMEM lane quarter
    constexpr int CH = 32 / NPART;           // gate channels per warp in the LVC epilogue (16 or 8)
    // Merge the (hi,hi) and (hi,lo) passes into one MMA of twice the N (W_hi | W_lo as one operand): 88 vs 120 cycles per conv
    // k-step, 112 vs 144 per LVC k-step (profiles/r01_tcgen05_findings.md) at the price of extra TMEM reads in the epilogues.
    // Pays off for the single-group variant (nothing overlaps its MMAs); with two groups the MMAs are already hidden.
    constexpr bool MERGE = (GROUPS == 1);
    constexpr int SLOT = lt_slot_bytes<HOP>();
    constexpr int SMALL = lt_small_bytes<HOP>();
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space
    // B-operand tiles are stored per tap as [hi rows | lo rows] so that (W_hi | W_lo) is ONE operand of twice the N:
    //   conv: cw + k*8192 + {0: hi 32 rows, 4096: lo 32 rows};   lvc: lw + fi*49152 + k*16384 + {0: hi 64 rows, 8192: lo 64 rows}
    unsigned char* cw = smem + GROUPS * SLOT;
    unsigned char* small0 = cw + 2 * LT_CW_BYTES;            // [GROUPS][2 buffers][lbias NF*64 | audio LT_AU]
    float* fw_s = (float*)(small0 + GROUPS * SMALL);         // [7][32] first_audio_conv weights
    float* fb_s = fw_s + 7 * C;                              // [32]
    float* cb_s = fb_s + C;                                  // [32] dilated-conv bias
    uint64_t* bars = (uint64_t*)(cb_s + C);                  // [GROUPS][4]: conv MMAs, LVC MMAs, loads, (pad)
    uint32_t* tmem_base_s = (uint32_t*)(bars + 8);

    const int tid = threadIdx.x, g = tid / GT, gt = tid % GT, gw = gt >> 5, lane = tid & 31;
    unsigned char* slot = smem + g * SLOT;
    unsigned char* a_hi = slot;                              // A tile (hi) | raw skip rows (block 1) | later: Y tile (hi)
    unsigned char* a_lo = a_hi + LT_A_BYTES;                 // A tile (lo) | raw x rows            | later: Y tile (lo)
    unsigned char* lw = a_lo + LT_A_BYTES;                   // [NF][3 taps][hi 8 KB (raw on arrival) | lo 8 KB]
    unsigned char* small = small0 + g * SMALL;
    uint64_t* bar = bars + 4 * g;

    if (tid == 0) {
        for (int i = 0; i < 4 * GROUPS; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {   // per-layer constants: the global order of the conv weights is already the swizzled smem image
        const float4* sh = reinterpret_cast<const float4*>(p.cw_hi); const float4* sl = reinterpret_cast<const float4*>(p.cw_lo);
        for (int i = tid; i < LT_CW_BYTES / 16; i += GT * GROUPS) {   // i = tap*256 + float4 within the 4 KB tap tile
            const int k = i >> 8, w = i & 255;
            reinterpret_cast<float4*>(cw + k * 8192)[w] = sh[i];
            reinterpret_cast<float4*>(cw + k * 8192 + 4096)[w] = sl[i];
        }
        if (tid < 7 * C) fw_s[tid] = SKIP_FIRST ? p.first_w[tid] : 0.f;
        if (tid < C) { fb_s[tid] = SKIP_FIRST ? p.first_b[tid] : 0.f; cb_s[tid] = p.conv_b[tid]; }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // TMEM columns of a group (256 apart): conv [0,64) = {[0,32): A_hi W_hi + A_lo W_hi, [32,64): A_hi W_lo};
    // LVC frame fi at [64 + 128 fi, +128) = {[0,64): Y_hi W_hi + Y_lo W_hi, [64,128): Y_hi W_lo}; the epilogues add the halves.
    const uint32_t tmem_base = *tmem_base_s + g * 256;
    constexpr uint32_t idesc_conv = umma_idesc_tf32(128, 32), idesc_conv2 = umma_idesc_tf32(128, 64);
    constexpr uint32_t idesc_lvc = umma_idesc_tf32(128, 64), idesc_lvc2 = umma_idesc_tf32(128, 128);

    // this thread's fixed role in the A transform: channel chunk c4 = gt & 7 -> keep its first-conv taps in registers
    const int c4 = gt & 7;
    float fwr[7][4], fbr[4];
    if (SKIP_FIRST) {
#pragma unroll
        for (int k = 0; k < 7; ++k)
#pragma unroll
            for (int q = 0; q < 4; ++q) fwr[k][q] = fw_s[k * C + c4 * 4 + q];
#pragma unroll
        for (int q = 0; q < 4; ++q) fbr[q] = fb_s[c4 * 4 + q];
    }
    // warp-uniform copies (shfl from lane 0) so the issue paths are provably uniform
    const int gw_u = __shfl_sync(0xffffffffu, gw, 0), g_u = __shfl_sync(0xffffffffu, g, 0);
    const uint32_t slot_u = smem_u32(smem) + (uint32_t)(g_u * SLOT);
    const uint32_t cw_u = smem_u32(cw);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);

    const int ntt = (T + LT_TT - 1) / LT_TT, total = B * ntt, tstride = gridDim.x * GROUPS;
    const int r_lo = 27 - dil, r_hi = 157 + dil;   // A rows ar <-> t = t0 - 28 + ar that the 130 conv outputs touch

    // Issue the bulk copies of one tile into this group's slot / small buffer `buf` (call from ONE thread).
    auto issue_loads = [&](int tile, int buf) {
        const int b = tile / ntt, t0 = (tile % ntt) * LT_TT;
        float* lbias = (float*)(small + buf * (SMALL / 2));
        float* au = lbias + NF * 64;
        const int ar0 = max(r_lo, 28 - t0), ar1 = min(r_hi, T - t0 + 28);
        const int i0 = max(0, 32 - t0), i1 = min(LT_AU, T - t0 + 32);
        uint32_t bytes = 0;
        if (ar1 > ar0) bytes += (uint32_t)(ar1 - ar0) * 128u * (SKIP_FIRST ? 1u : 2u);
        if (SKIP_FIRST && i1 > i0) bytes += (uint32_t)(i1 - i0) * 4u;
#pragma unroll
        for (int fi = 0; fi < NF; ++fi) if (t0 / HOP + fi < Tm) bytes += LT_LW_BYTES + 256;
        mbar_expect_tx(&bar[2], bytes);
        if (ar1 > ar0) {
            const size_t off = ((size_t)b * T + (t0 - 28 + ar0)) * C;
            bulk_g2s(a_lo + ar0 * 128, x_in + off, (uint32_t)(ar1 - ar0) * 128u, &bar[2]);
            if (!SKIP_FIRST) bulk_g2s(a_hi + ar0 * 128, skip + off, (uint32_t)(ar1 - ar0) * 128u, &bar[2]);
        }
        if (SKIP_FIRST && i1 > i0) bulk_g2s(au + i0, skip + (size_t)b * T + (t0 - 32 + i0), (uint32_t)(i1 - i0) * 4u, &bar[2]);
#pragma unroll
        for (int fi = 0; fi < NF; ++fi) {
            const int f = t0 / HOP + fi;
            if (f < Tm) {
                const float* src = kern + ((size_t)b * Tm + f) * KCN;
#pragma unroll
                for (int k = 0; k < 3; ++k) bulk_g2s(lw + fi * 2 * LT_LW_BYTES + k * 16384, src + k * 2048, 8192, &bar[2]);
                bulk_g2s(lbias + fi * 64, src + KK * LVC_OUT, 256, &bar[2]);
            }
        }
    };

    // predicted LVC kernels -> tf32 pieces (hi in place, lo in the half-tile next to it; element-wise: the KC GEMM already wrote
    // them in this tile layout).  With two groups it runs in phase 1 (under the OTHER group's MMAs it would compete with their
    // operand fetch for smem bandwidth -- measured slower); the single-group variant hides it under its own conv MMAs.
    auto lw_split_tile = [&](int t0) {
#pragma unroll
        for (int fi = 0; fi < NF; ++fi) {
            if (t0 / HOP + fi < Tm) {
#pragma unroll
                for (int i = 0; i < 1536 / GT; ++i) {
                    const int e = gt + i * GT;   // float4 index inside the 24 KB of raw kernels: tap = e >> 9
                    float4* ph = reinterpret_cast<float4*>(lw + fi * 2 * LT_LW_BYTES + (e >> 9) * 16384) + (e & 511);
                    float4 hi, lo;
                    split4(*ph, hi, lo);
                    *ph = hi;
                    ph[512] = lo;   // + 8 KB
                }
            }
        }
    };

    int tile = blockIdx.x * GROUPS + g;
    if (tile < total && gw_u == 0) { if (elect_one()) issue_loads(tile, 0); __syncwarp(); }
#ifdef FD_LVC_TIMELINE
    const bool stamp = (blockIdx.x == 0 && tid == 0 && HOP == 256);
    int tile_no = -1;
#endif
    uint32_t parity = 0;
    for (; tile < total; tile += tstride, parity ^= 1) {
#ifdef FD_LVC_TIMELINE
        ++tile_no;
#endif
        const int b = tile / ntt, t0 = (tile % ntt) * LT_TT;
        const float* lbias = (const float*)(small + parity * (SMALL / 2));
        float* au_s = (float*)lbias + NF * 64;
        LT_STAMP(0);
        // ---------------- phase 1: raw tile (bulk-copied one tile ago) -> tf32 operand tiles, in place ----------------
        mbar_wait(&bar[2], parity);
        LT_STAMP(1);
        if (SKIP_FIRST) {   // audio positions outside [0,T) are zero (the first conv zero-pads)
            if (gt < LT_AU) { const int pos = t0 - 32 + gt; if (pos < 0 || pos >= T) au_s[gt] = 0.f; }
            group_sync(1 + g, GT);
        }
        if (!MERGE) lw_split_tile(t0);
        LT_STAMP(2);
#pragma unroll
        for (int i = 0; i < 1536 / GT; ++i) {   // A rows: the 8 lanes of a row read its raw chunks, then write the swizzled tf32 pieces
            const int ar = r_lo + (gt >> 3) + i * (GT / 8), t = t0 - 28 + ar;
            const bool active = ar < r_hi;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (active && t >= 0 && t < T) {
                const float4 xv = *reinterpret_cast<const float4*>(a_lo + ar * 128 + c4 * 16);
                float4 sk;
                if (SKIP_FIRST) {
                    sk = make_float4(fbr[0], fbr[1], fbr[2], fbr[3]);
#pragma unroll
                    for (int k = 0; k < 7; ++k) {
                        const float a = au_s[ar + k + 1];
                        sk.x = fmaf(fwr[k][0], a, sk.x); sk.y = fmaf(fwr[k][1], a, sk.y);
                        sk.z = fmaf(fwr[k][2], a, sk.z); sk.w = fmaf(fwr[k][3], a, sk.w);
                    }
                } else {
                    sk = *reinterpret_cast<const float4*>(a_hi + ar * 128 + c4 * 16);
                }
                v.x = lrelu(xv.x + sk.x, 0.2f); v.y = lrelu(xv.y + sk.y, 0.2f);
                v.z = lrelu(xv.z + sk.z, 0.2f); v.w = lrelu(xv.w + sk.w, 0.2f);
            }
            float4 hi, lo;
            split4(v, hi, lo);
            __syncwarp();   // every lane of the row has read its raw chunk before any lane overwrites the row
            if (active) {
                *reinterpret_cast<float4*>(a_hi + swz128(ar, c4)) = hi;
                *reinterpret_cast<float4*>(a_lo + swz128(ar, c4)) = lo;
            }
        }
        LT_STAMP(3);   // A built
        fence_async_smem();
        group_sync(1 + g, GT);
        LT_STAMP(4);
        // ---------------- phase 2: dilated conv on tensor cores (+ 2 halo rows on FFMA meanwhile) ----------------
        if (gw_u == 0) {
            tc_fence_after();
            uint32_t slot_t = slot_u, cw_t = cw_u;
            asm volatile("" : "+r"(slot_t), "+r"(cw_t));   // opaque per tile: keeps ptxas from hoisting ~100 descriptors out of the tile loop
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const uint32_t sh = (uint32_t)(27 + (k - 1) * dil) * 128u;   // row shift: start address + shift*128 B, base_offset 0
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint64_t dah = umma_desc_sw128(slot_t + sh + j * 32), dal = umma_desc_sw128(slot_t + LT_A_BYTES + sh + j * 32);
                        const uint64_t db = umma_desc_sw128(cw_t + k * 8192 + j * 32);   // rows 0-31 W_hi, rows 32-63 W_lo
                        if (three_pass && MERGE) {
                            umma_tf32(tmem_u, dah, db, idesc_conv2, (k | j) ? 1u : 0u);   // cols [0,32) += A_hi W_hi ; [32,64) += A_hi W_lo
                            umma_tf32(tmem_u, dal, db, idesc_conv, 1u);                   // cols [0,32) += A_lo W_hi
                        } else if (three_pass) {
                            umma_tf32(tmem_u, dah, db, idesc_conv, (k | j) ? 1u : 0u);
                            umma_tf32(tmem_u, dah, db + (4096 >> 4), idesc_conv, 1u);     // W_lo half-tile
                            umma_tf32(tmem_u, dal, db, idesc_conv, 1u);
                        } else {
                            umma_tf32(tmem_u, dah, db, idesc_conv, (k | j) ? 1u : 0u);
                        }
                    }
                }
                tc_commit(&bar[0]);
            }
            __syncwarp();
        }
        if (MERGE) lw_split_tile(t0);   // single group: nothing else overlaps the conv MMAs, so hide the split under them
        float halo = 0.f;   // warps 6,7 of the group: conv outputs yr = 128 (warp 6), 129 (warp 7), lane = co
        if (gw >= GW - 2) {
            const int yr = 128 + (gw - (GW - 2));
            float acc = cb_s[lane];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int ar = yr + 27 + (k - 1) * dil;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 ah4 = *reinterpret_cast<const float4*>(a_hi + swz128(ar, c));
                    const float4 al4 = *reinterpret_cast<const float4*>(a_lo + swz128(ar, c));
                    const float4 wh4 = *reinterpret_cast<const float4*>(cw + k * 8192 + swz128(lane, c));
                    const float4 wl4 = *reinterpret_cast<const float4*>(cw + k * 8192 + 4096 + swz128(lane, c));
                    acc = fmaf(ah4.x + al4.x, wh4.x + wl4.x, acc); acc = fmaf(ah4.y + al4.y, wh4.y + wl4.y, acc);
                    acc = fmaf(ah4.z + al4.z, wh4.z + wl4.z, acc); acc = fmaf(ah4.w + al4.w, wh4.w + wl4.w, acc);
                }
            }
            const int t = t0 - 1 + yr;
            halo = (t >= 0 && t < T) ? lrelu(acc, 0.2f) : 0.f;
        }
        // ---------------- phase 3: y = lrelu(conv + b) -> tf32 rows of the Y tile (over the A tile) ----------------
        LT_STAMP(5);
        mbar_wait(&bar[0], parity);
        tc_fence_after();
        LT_STAMP(6);   // conv MMAs complete
        group_sync(1 + g, GT);   // Y aliases A: the halo warps must have finished READING A before anyone writes Y
        if (gw < 4 * (GW / 8)) {   // GW/8 warps per lane quarter, each taking 32/(GW/8) of the 32 conv channels
            constexpr int P3N = GW / 8, P3C = 32 / P3N;   // columns per warp: 32 or 16
            const int q3 = gw & 3, part3 = gw >> 2;
            uint32_t v[P3C];
            const uint32_t ta3 = tmem_base + ((uint32_t)(q3 * 32) << 16) + part3 * P3C;
            tmem_ld_cols<P3C>(ta3, v);
            if (three_pass && MERGE) {
                uint32_t v2[P3C];
                tmem_ld_cols<P3C>(ta3 + 32, v2);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < P3C; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(v2[i]));
            }
            tmem_ld_wait();
            const int yr = q3 * 32 + lane, t = t0 - 1 + yr;
            const bool in = (t >= 0 && t < T);
#pragma unroll
            for (int cc = 0; cc < P3C / 4; ++cc) {
                const int c = part3 * (P3C / 4) + cc;
                float4 y;
                y.x = in ? lrelu(__uint_as_float(v[cc * 4 + 0]) + cb_s[c * 4 + 0], 0.2f) : 0.f;
                y.y = in ? lrelu(__uint_as_float(v[cc * 4 + 1]) + cb_s[c * 4 + 1], 0.2f) : 0.f;
                y.z = in ? lrelu(__uint_as_float(v[cc * 4 + 2]) + cb_s[c * 4 + 2], 0.2f) : 0.f;
                y.w = in ? lrelu(__uint_as_float(v[cc * 4 + 3]) + cb_s[c * 4 + 3], 0.2f) : 0.f;
                float4 hi, lo;
                split4(y, hi, lo);
                *reinterpret_cast<float4*>(a_hi + swz128(yr, c)) = hi;
                *reinterpret_cast<float4*>(a_lo + swz128(yr, c)) = lo;
            }
        } else if (gw >= GW - 2) {
            const int yr = 128 + (gw - (GW - 2));
            const float hi = split_hi(halo), lo = halo - hi;
            *reinterpret_cast<float*>(a_hi + swz128(yr, lane >> 2) + (lane & 3) * 4) = hi;
            *reinterpret_cast<float*>(a_lo + swz128(yr, lane >> 2) + (lane & 3) * 4) = lo;
        }
        fence_async_smem();
        tc_fence_before();
        group_sync(1 + g, GT);
        LT_STAMP(7);   // Y written
        // ---------------- phase 4: location-variable conv on tensor cores ----------------
        if (gw_u == 0) {
            tc_fence_after();
            uint32_t slot_t = slot_u;
            asm volatile("" : "+r"(slot_t));
            if (elect_one()) {
#pragma unroll
                for (int fi = 0; fi < NF; ++fi) {
                    const uint32_t d = tmem_u + 64 + fi * 128;
                    const uint32_t lwb = slot_t + 2 * LT_A_BYTES + fi * 2 * LT_LW_BYTES;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint64_t dah = umma_desc_sw128(slot_t + k * 128 + j * 32), dal = umma_desc_sw128(slot_t + LT_A_BYTES + k * 128 + j * 32);
                            const uint64_t db = umma_desc_sw128(lwb + k * 16384 + j * 32);   // rows 0-63 W_hi, rows 64-127 W_lo
                            if (three_pass && MERGE) {
                                umma_tf32(d, dah, db, idesc_lvc2, (k | j) ? 1u : 0u);   // cols [0,64) += Y_hi W_hi ; [64,128) += Y_hi W_lo
                                umma_tf32(d, dal, db, idesc_lvc, 1u);                   // cols [0,64) += Y_lo W_hi
                            } else if (three_pass) {
                                umma_tf32(d, dah, db, idesc_lvc, (k | j) ? 1u : 0u);
                                umma_tf32(d, dah, db + (8192 >> 4), idesc_lvc, 1u);     // W_lo half-tile
                                umma_tf32(d, dal, db, idesc_lvc, 1u);
                            } else {
                                umma_tf32(d, dah, db, idesc_lvc, (k | j) ? 1u : 0u);
                            }
                        }
                    }
                }
                tc_commit(&bar[1]);
            }
            __syncwarp();
        }
        // ---------------- phase 5: gate + residual -> global ----------------
        {
            const int q = gw & 3, part = gw >> 2;              // lane quarter / which CH of the 32 gate channels
            const int r = q * 32 + lane, t = t0 + r;
            const int fi = (HOP >= LT_TT) ? 0 : r / HOP;       // warp-uniform (HOP is a multiple of 32)
            const size_t row = ((size_t)b * T + (t < T ? t : 0)) * C + part * CH;
            float4 xs[CH / 4];                                 // residual base, fetched while the MMAs run
#pragma unroll
            for (int c = 0; c < CH / 4; ++c) {
                xs[c] = *reinterpret_cast<const float4*>(x_in + row + c * 4);
                if (SKIP_FIRST) {
                    const int o = part * CH + c * 4;
                    float4 sk = *reinterpret_cast<const float4*>(fb_s + o);
#pragma unroll
                    for (int k = 0; k < 7; ++k) {
                        const float a = au_s[29 + r + k];
                        const float4 w = *reinterpret_cast<const float4*>(fw_s + k * C + o);
                        sk.x = fmaf(w.x, a, sk.x); sk.y = fmaf(w.y, a, sk.y); sk.z = fmaf(w.z, a, sk.z); sk.w = fmaf(w.w, a, sk.w);
                    }
                    xs[c].x += sk.x; xs[c].y += sk.y; xs[c].z += sk.z; xs[c].w += sk.w;
                } else {
                    const float4 sk = *reinterpret_cast<const float4*>(skip + row + c * 4);
                    xs[c].x += sk.x; xs[c].y += sk.y; xs[c].z += sk.z; xs[c].w += sk.w;
                }
            }
            LT_STAMP(8);   // LVC MMAs issued, residual prefetched
            mbar_wait(&bar[1], parity);
            tc_fence_after();
            LT_STAMP(9);   // LVC MMAs complete: the slot's operand tiles are free -> fetch the next tile while this one is gated
            if (gw_u == 0 && tile + tstride < total) { if (elect_one()) issue_loads(tile + tstride, (int)(parity ^ 1)); __syncwarp(); }
            uint32_t zs[CH], zt[CH];
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + 64 + fi * 128 + part * CH;
            tmem_ld_cols<CH>(ta, zs);
            tmem_ld_cols<CH>(ta + 32, zt);
            if (three_pass && MERGE) {   // + the (hi, lo) partial sums in the upper 64 columns
                uint32_t a1[CH], a2[CH];
                tmem_ld_cols<CH>(ta + 64, a1);
                tmem_ld_cols<CH>(ta + 96, a2);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    zs[i] = __float_as_uint(__uint_as_float(zs[i]) + __uint_as_float(a1[i]));
                    zt[i] = __float_as_uint(__uint_as_float(zt[i]) + __uint_as_float(a2[i]));
                }
            }
            tmem_ld_wait();
            if (t < T) {
                const float* lb = lbias + fi * 64 + part * CH;
#pragma unroll
                for (int c = 0; c < CH / 4; ++c) {
                    float4 o4;
                    o4.x = xs[c].x + fast_sigmoid(__uint_as_float(zs[c * 4 + 0]) + lb[c * 4 + 0]) * fast_tanh(__uint_as_float(zt[c * 4 + 0]) + lb[32 + c * 4 + 0]);
                    o4.y = xs[c].y + fast_sigmoid(__uint_as_float(zs[c * 4 + 1]) + lb[c * 4 + 1]) * fast_tanh(__uint_as_float(zt[c * 4 + 1]) + lb[32 + c * 4 + 1]);
                    o4.z = xs[c].z + fast_sigmoid(__uint_as_float(zs[c * 4 + 2]) + lb[c * 4 + 2]) * fast_tanh(__uint_as_float(zt[c * 4 + 2]) + lb[32 + c * 4 + 2]);
                    o4.w = xs[c].w + fast_sigmoid(__uint_as_float(zs[c * 4 + 3]) + lb[c * 4 + 3]) * fast_tanh(__uint_as_float(zt[c * 4 + 3]) + lb[32 + c * 4 + 3]);
                    *reinterpret_cast<float4*>(x_out + row + c * 4) = o4;
                }
            }
        }
        tc_fence_before();
        group_sync(1 + g, GT);   // the group's TMEM columns are free for its next tile
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_base_s), "r"(512u) : "memory");
    }
}

#endif  // !FD_EMU
#ifndef LT_STAMP
#define LT_STAMP(i) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------------------
// K11+K12+K13, mode tc_3xf16: the same LVC layer on kind::f16 MMAs over fp16 PIECES (hi | lo) of power-of-two prescaled
// operands.  Same tile walk, tap-as-shifted-start trick, one-tile-ahead bulk prefetch and two independent 8-warp groups per CTA
// as k_lvc_layer_tc above; what changes:
//   * one operand row = 128 B = [32 channels hi | 32 channels lo] (fp16; 16-byte chunk c at position c ^ (row & 7)), so an
//     operand tile is HALF the bytes of the tf32 hi + lo pair and the three passes of a tap are the K-slices
//       (hi,hi): A + {0,32} x B + {0,32}     (hi,lo): A + {0,32} x B + {64,96}     (lo,hi): A + {64,96} x B + {0,32}
//     -> 6 MMAs of K = 16 per tap instead of 12 of K = 8;
//   * the predicted kernels arrive from the kernel_conv GEMM already as pieces in this layout: no split phase, no lo half-tile;
//   * the raw x rows are transformed in place to pieces (8 lanes per row: LDS.128 -> STS.64 hi + STS.64 lo);
//   * block 2 keeps xs = x + first_conv(audio) (fp32) of the 128 output rows in smem for the gate epilogue instead of recomputing
//     it (7 FMA per channel) and re-reading x from global;
//   * all 8 warps of a group take part in the conv epilogue (2 per TMEM lane quarter, 16 channels each);
//   * tile walk: every group owns a contiguous chunk of tiles and walks it in DESCENDING time order, so the two extra conv rows
//     (yr = 128, 129) are Y rows 0, 1 of the tile processed just before and are carried over in 256 B of smem; only chunk starts
//     and utterance ends take them from a second MMA pass over A rows +128 -- the same instruction sequence, hence the same bits:
//     every conv output is independent of the tiling and of the batch composition;
//   * block 1 adds the skip to the rows it PRODUCES (skip_out: the next layer's x += audio_down, same rounding order), so only the
//     first layer of the block loads the skip tile; block 2 recomputes first_conv(audio) on the way in (compile-time);
//   * sigmoid(a) * tanh(b) = (1 - E) / ((1 + e^-a)(1 + E)), E = e^-2b: two ex2 and ONE rcp per gate.
// Scales: A pieces hold v*S16_ACT, conv weights w*S (per tensor, SCALES16), predicted kernels w*S16_KERN; the epilogues multiply
// the accumulators by inv_c = 1/(S16_ACT*S) and inv_l = 1/(S16_ACT*S16_KERN).
// ---------------------------------------------------------------------------------------------------------
constexpr int LH_A_BYTES = 24576;                    // 192 rows x 128 B
constexpr int LH_XS_BYTES = 16384;                   // 128 rows x 128 B fp32 (SKIP_FIRST)
constexpr int LH_LW_BYTES = 24576;                   // per frame: 3 taps x 64 rows x 128 B
constexpr int LH_CW_BYTES = 3 * C * 128;             // 12288
template <int HOP, bool SKIP_FIRST>
__host__ __device__ constexpr int lh_slot_bytes() { return LH_A_BYTES + (SKIP_FIRST ? LH_XS_BYTES : LH_A_BYTES) + lt_nf<HOP>() * LH_LW_BYTES; }
constexpr int LH_SHARED_BYTES = LH_CW_BYTES + (7 * C + C + C + C) * 4 + 3 * 3 * 64 * 4 + 3 * 512 + 192;   // conv W, first_w, first_b, conv_b (x1, x16), spare, carried rows, barriers + tmem ptr
template <int HOP, bool SKIP_FIRST, int GROUPS>
constexpr int lh_smem_bytes() { return GROUPS * (lh_slot_bytes<HOP, SKIP_FIRST>() + lt_small_bytes<HOP>()) + LH_SHARED_BYTES + 1024; }

struct LvcHParams {
    const float* cw16;                           // [3 taps][32 co][128 B] SWIZZLE_128B image of this layer's dilated conv (LBn_CONV_F16)
    const float* conv_b;                         // [32]
    const float* first_w; const float* first_b;  // [7][32], [32]   (SKIP_FIRST)
};

// 4 floats -> fp16 pieces of v*S16_ACT: hi (4 halves in a uint2), lo likewise
__device__ __forceinline__ void split4_f16_pre(const float sx, const float sy, const float sz, const float sw, uint2& hi, uint2& lo);
__device__ __forceinline__ void split4_f16(const float4 v, uint2& hi, uint2& lo) {
    split4_f16_pre(v.x * S16_ACT, v.y * S16_ACT, v.z * S16_ACT, v.w * S16_ACT, hi, lo);
}
// same, for values that already carry the prescale
__device__ __forceinline__ void split4_f16_pre(const float sx, const float sy, const float sz, const float sw, uint2& hi, uint2& lo) {
    hi.x = pack_f16x2_sat(sy, sx);
    hi.y = pack_f16x2_sat(sw, sz);
    const float2 h01 = unpack_f16x2(hi.x), h23 = unpack_f16x2(hi.y);
    lo.x = pack_f16x2_sat(sy - h01.y, sx - h01.x);
    lo.y = pack_f16x2_sat(sw - h23.y, sz - h23.x);
}
// sigmoid(a) * tanh(b) with two ex2 and one rcp; b clamped from below at -15 (tanh is -1 to fp32 precision below -9.01) so E stays finite
#ifndef LH_GATE_ASM
#define LH_GATE_ASM 1
#endif
__device__ __forceinline__ float gate_st(float a, float b) {
    const float bc = fmaxf(b, -15.f);   // E = e^-2b must stay finite (E -> 0 for large b is harmless); tanh(-15) = -1 to fp32 precision
#if !LH_GATE_ASM
    const float E2 = exp2f(-2.8853900817779268f * bc), A2 = exp2f(-1.4426950408889634f * a);
    return (1.f - E2) / ((1.f + A2) * (1.f + E2));
#endif
    const float E = ex2_approx(-2.8853900817779268f * bc), A = ex2_approx(-1.4426950408889634f * a);   // A = +inf for a << 0 -> result 0
    return (1.f - E) * rcp_approx((1.f + A) * (1.f + E));
}
// leaky ReLU with slope 0.2 as max(v, 0.2 v): two instructions instead of compare + multiply + select
__device__ __forceinline__ float lrelu02(float v) { return fmaxf(v, 0.2f * v); }
// S16_ACT * lrelu_0.2(v) in the same three instructions (the prescale of the fp16 pieces folded into the two products; exact: power of two)
__device__ __forceinline__ float lrelu02_s(float v) { return fmaxf(v * S16_ACT, v * (0.2f * S16_ACT)); }

template <int HOP, bool SKIP_FIRST, int GROUPS>
__global__ void __launch_bounds__(256 * GROUPS, 1)
k_lvc_layer_h(LvcHParams p, const float* __restrict__ x_in, const float* __restrict__ skip, const float* __restrict__ kern,
              float* __restrict__ x_out, int B, int T, int Tm, int dil, float inv_c, float inv_l, int exp_mask, int skip_in_rt, int skip_out_rt) {
    // where the skip is added is a run-time choice for block 1 only; block 2 (SKIP_FIRST) always adds it on the way in (compile-time)
    const bool skip_in = SKIP_FIRST ? true : (skip_in_rt != 0), skip_out = SKIP_FIRST ? false : (skip_out_rt != 0);
    constexpr int NF = lt_nf<HOP>();
    constexpr int GT = 256;                   // threads per group (8 warps)
    constexpr int SLOT = lh_slot_bytes<HOP, SKIP_FIRST>();
    constexpr int SMALL = lt_small_bytes<HOP>();
    constexpr int S_OFF = LH_A_BYTES;                                              // raw skip rows (block 1) | xs rows (block 2)
    constexpr int LW_OFF = LH_A_BYTES + (SKIP_FIRST ? LH_XS_BYTES : LH_A_BYTES);
    FD_DYN_SMEM(unsigned char, smem_raw);
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* cw = smem + GROUPS * SLOT;                 // [3 taps][32 rows][128 B]
    unsigned char* small0 = cw + LH_CW_BYTES;                 // [GROUPS][2 buffers][lbias NF*64 | audio LT_AU]
    float* fw_s = (float*)(small0 + GROUPS * SMALL);          // [7][32]
    float* fb_s = fw_s + 7 * C;                               // [32]
    float* cb_s = fb_s + C;                                   // [32]
    float* cbs_s = cb_s + C;                                  // [32] conv bias * S16_ACT
    float* hp_s = cbs_s + C;                                  // [GROUPS][3][64] (unused since the halo rows are carried / come from the second MMA pass)
    unsigned char* carry_s = (unsigned char*)(hp_s + GROUPS * 3 * 64);   // [GROUPS][2][256 B]: Y rows 0, 1 of the previous tile (pieces)
    uint64_t* bars = (uint64_t*)(carry_s + GROUPS * 512);     // [GROUPS][4]: conv MMAs, LVC MMAs, loads, (pad)
    uint32_t* tmem_base_s = (uint32_t*)(bars + 4 * GROUPS);

    const int tid = threadIdx.x, g = tid / GT, gt = tid % GT, gw = gt >> 5, lane = tid & 31;
    unsigned char* slot = smem + g * SLOT;
    unsigned char* a_t = slot;                                // A tile (raw x rows on arrival) | later: Y tile
    unsigned char* s_t = slot + S_OFF;
    unsigned char* lw = slot + LW_OFF;
    unsigned char* small = small0 + g * SMALL;
    uint64_t* bar = bars + 4 * g;
    unsigned char* carry = carry_s + g * 512;

    if (tid == 0) {
        for (int i = 0; i < 4 * GROUPS; ++i) mbar_init(&bars[i], 1);
        mbar_init_fence();
    }
    if (tid < 32) tmem_alloc(tmem_base_s, 512u);
    {
        const float4* src = reinterpret_cast<const float4*>(p.cw16);
        for (int i = tid; i < LH_CW_BYTES / 16; i += GT * GROUPS) reinterpret_cast<float4*>(cw)[i] = src[i];
        if (tid < 7 * C) fw_s[tid] = SKIP_FIRST ? p.first_w[tid] : 0.f;
        if (tid < C) { fb_s[tid] = SKIP_FIRST ? p.first_b[tid] : 0.f; cb_s[tid] = p.conv_b[tid]; cbs_s[tid] = p.conv_b[tid] * S16_ACT; }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // TMEM columns of a group (256 apart): conv [0,32); LVC frame fi at [32 + 64 fi, +64); second conv pass (rows +128) after them
    constexpr uint32_t D2_COL = 32 + NF * 64;
    constexpr uint32_t TSTRIDE = (D2_COL + 32 <= 128) ? 128 : 256;
    static_assert(GROUPS * TSTRIDE <= 512, "TMEM columns");
    const uint32_t tmem_base = *tmem_base_s + g * TSTRIDE;
    constexpr uint32_t idesc_conv = umma_idesc_f16(128, 32), idesc_lvc = umma_idesc_f16(128, 64);

    const int c4 = gt & 7;   // this thread's channel quad in the A transform
    // its row within each 32-row block of the A transform.  LH_ROW_SPREAD = 1 (experiment, off: not measured yet) gives a warp the rows
    // r, r+4, r+8, r+12 instead of 4 consecutive ones: their swizzle XOR then differs in bit 2, so the hi (and lo) 8-byte piece stores of
    // a warp cover both halves of the 128-byte bank line -- 2 shared-memory wavefronts per STS.64 instead of 4 (ncu: 1.8 M excess per launch)
    const int prow = gt >> 3;
    const int gw_u = __shfl_sync(0xffffffffu, gw, 0), g_u = __shfl_sync(0xffffffffu, g, 0);
    const uint32_t slot_u = smem_u32(smem) + (uint32_t)(g_u * SLOT);
    const uint32_t cw_u = smem_u32(cw);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);

    const int ntt = (T + LT_TT - 1) / LT_TT, total = B * ntt;
    const int r_lo = 27 - dil, r_hi = 157 + dil;   // A rows ar <-> t = t0 - 28 + ar that the 130 conv outputs touch

    // keep_lw: the tile uses the same frame as the one just processed (hop 256: two tiles per frame, walked back to back): its
    // predicted kernels are already in place, only the 64 biases (double-buffered with the audio window) are fetched again
    auto issue_loads = [&](int tile, int buf, bool keep_lw) {
        const int b = tile / ntt, t0 = (tile % ntt) * LT_TT;
        float* lbias = (float*)(small + buf * (SMALL / 2));
        float* au = lbias + NF * 64;
        const int ar0 = max(r_lo, 28 - t0), ar1 = min(r_hi, T - t0 + 28);
        const int i0 = max(0, 32 - t0), i1 = min(LT_AU, T - t0 + 32);
        uint32_t bytes = 0;
        if (ar1 > ar0) bytes += (uint32_t)(ar1 - ar0) * 128u * ((SKIP_FIRST || !skip_in) ? 1u : 2u);
        if (SKIP_FIRST && i1 > i0) bytes += (uint32_t)(i1 - i0) * 4u;
#pragma unroll
        for (int fi = 0; fi < NF; ++fi) if (t0 / HOP + fi < Tm) bytes += (keep_lw ? 0 : LH_LW_BYTES) + 256;
        mbar_expect_tx(&bar[2], bytes);
        if (ar1 > ar0) {
            const size_t off = ((size_t)b * T + (t0 - 28 + ar0)) * C;
            bulk_g2s(a_t + ar0 * 128, x_in + off, (uint32_t)(ar1 - ar0) * 128u, &bar[2]);
            if (!SKIP_FIRST && skip_in) bulk_g2s(s_t + ar0 * 128, skip + off, (uint32_t)(ar1 - ar0) * 128u, &bar[2]);
        }
        if (SKIP_FIRST && i1 > i0) bulk_g2s(au + i0, skip + (size_t)b * T + (t0 - 32 + i0), (uint32_t)(i1 - i0) * 4u, &bar[2]);
#pragma unroll
        for (int fi = 0; fi < NF; ++fi) {
            const int f = t0 / HOP + fi;
            if (f < Tm) {
                const float* src = kern + ((size_t)b * Tm + f) * KCN;
                if (!keep_lw) bulk_g2s(lw + fi * LH_LW_BYTES, src, LH_LW_BYTES, &bar[2]);   // 3 taps x 64 rows x 128 B of pieces, ready to use
                bulk_g2s(lbias + fi * 64, src + KK * LVC_OUT, 256, &bar[2]);
            }
        }
    };

    // L2 prefetch of the tile after the next one (x rows, skip rows, predicted kernels: all come from HBM), so that the bulk copies
    // issued one tile ahead are L2 hits (call from ONE thread)
    auto prefetch_l2 = [&](int tile) {
        const int b = tile / ntt, t0 = (tile % ntt) * LT_TT;
        const int ar0 = max(r_lo, 28 - t0), ar1 = min(r_hi, T - t0 + 28);
        if (ar1 > ar0) {
            const size_t off = ((size_t)b * T + (t0 - 28 + ar0)) * C;
            bulk_prefetch_l2(x_in + off, (uint32_t)(ar1 - ar0) * 128u);
            if (!SKIP_FIRST && skip_in) bulk_prefetch_l2(skip + off, (uint32_t)(ar1 - ar0) * 128u);
        }
#pragma unroll
        for (int fi = 0; fi < NF; ++fi) {
            const int f = t0 / HOP + fi;
            if (f < Tm) bulk_prefetch_l2(kern + ((size_t)b * Tm + f) * KCN, (uint32_t)(KPL * 4));
        }
    };

    // Tile walk: every group owns one contiguous chunk of tiles and walks it in DESCENDING time order.  The two conv rows past the
    // M = 128 tile that the LVC taps of the last output rows need (yr = 128, 129) are then exactly Y rows 0, 1 of the tile processed
    // just before (t0 + 128): they are carried over in shared memory (256 B) instead of being recomputed.  Only the first tile of a
    // chunk and the last tile of an utterance have no predecessor: they get the rows from a second MMA pass over A rows +128 (same
    // instruction sequence as the carried rows -> the same bits, whatever the chunking or the batch composition).
    const int ngroups = gridDim.x * GROUPS, chunk = (total + ngroups - 1) / ngroups;
    const int tile_lo = (blockIdx.x * GROUPS + g) * chunk, tile_hi = min(total, tile_lo + chunk) - 1;
    int tile = tile_hi;
    if (tile >= tile_lo && gw_u == 0) {
        if (elect_one()) { issue_loads(tile, 0, false); if ((exp_mask & 8) && tile - 1 >= tile_lo) prefetch_l2(tile - 1); }
        __syncwarp();
    }
#ifdef FD_LVC_TIMELINE
    const bool stamp = (blockIdx.x == 0 && tid == 0 && HOP == 256);
    int tile_no = -1;
#endif
    uint32_t parity = 0;
    int b = tile >= 0 ? tile / ntt : 0, tt = tile >= 0 ? tile % ntt : 0;   // maintained incrementally (no per-tile division)
    for (; tile >= tile_lo; --tile, parity ^= 1) {
#ifdef FD_LVC_TIMELINE
        ++tile_no;
#endif
        const bool have_carry = (tile != tile_hi) && (tt != ntt - 1);   // the previous tile of this group was (b, tt + 1)
        LT_STAMP(0);
        const int t0 = tt * LT_TT;
        const float* lbias = (const float*)(small + parity * (SMALL / 2));
        float* au_s = (float*)lbias + NF * 64;
        // ---------------- phase 1: raw rows (bulk-copied one tile ago) -> fp16 pieces, in place ----------------
        mbar_wait(&bar[2], parity);
        LT_STAMP(1);
        if (SKIP_FIRST) {   // audio positions outside [0,T) are zero (the first conv zero-pads)
            if (gt < LT_AU) { const int pos = t0 - 32 + gt; if (pos < 0 || pos >= T) au_s[gt] = 0.f; }
            group_sync(1 + g, GT);
        }
        LT_STAMP(2);
        {   // pass 1: every thread pulls its 6 (row, channel quad) items into registers; one __syncwarp; pass 2: transform + store.
            // skip_in (first layer of a block): the rows are x and the skip is added here; later layers read z = x + skip, written so by
            // the previous layer's epilogue (skip_out) -- the reference's "x += audio_down" of the next layer, moved to where the row is
            // produced: 128 rows instead of the 130 + 2 dil rows this phase touches, same rounding sequence.
            float fwr[7][4], fbr[4];   // first-conv taps of this thread's channel quad: live in this phase only (reloaded per tile)
            if (SKIP_FIRST && skip_in) {
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    const float4 w4 = *reinterpret_cast<const float4*>(fw_s + k * C + c4 * 4);
                    fwr[k][0] = w4.x; fwr[k][1] = w4.y; fwr[k][2] = w4.z; fwr[k][3] = w4.w;
                }
                const float4 b4 = *reinterpret_cast<const float4*>(fb_s + c4 * 4);
                fbr[0] = b4.x; fbr[1] = b4.y; fbr[2] = b4.z; fbr[3] = b4.w;
            }
            float4 xv[1536 / GT], sv[1536 / GT];
#pragma unroll
            for (int i = 0; i < 1536 / GT; ++i) {
                const int ar = r_lo + prow + i * (GT / 8), t = t0 - 28 + ar;
                xv[i] = make_float4(0.f, 0.f, 0.f, 0.f); sv[i] = xv[i];
                if (ar < r_hi && t >= 0 && t < T) {
                    xv[i] = *reinterpret_cast<const float4*>(a_t + ar * 128 + c4 * 16);
                    if (!SKIP_FIRST && skip_in) sv[i] = *reinterpret_cast<const float4*>(s_t + ar * 128 + c4 * 16);
                }
            }
            __syncwarp();   // the 8 lanes of a row sit in one warp: every raw chunk has been read before any row is overwritten
#pragma unroll
            for (int i = 0; i < 1536 / GT; ++i) {
                const int ar = r_lo + prow + i * (GT / 8), t = t0 - 28 + ar;
                const bool active = ar < r_hi;
                float4 pre = xv[i];   // zero outside [0,T)
                if (skip_in && active && t >= 0 && t < T) {
                    float4 sk;
                    if (SKIP_FIRST) {
                        sk = make_float4(fbr[0], fbr[1], fbr[2], fbr[3]);
#pragma unroll
                        for (int k = 0; k < 7; ++k) {
                            const float a = au_s[ar + k + 1];
                            sk.x = fmaf(fwr[k][0], a, sk.x); sk.y = fmaf(fwr[k][1], a, sk.y);
                            sk.z = fmaf(fwr[k][2], a, sk.z); sk.w = fmaf(fwr[k][3], a, sk.w);
                        }
                    } else {
                        sk = sv[i];
                    }
                    pre = make_float4(xv[i].x + sk.x, xv[i].y + sk.y, xv[i].z + sk.z, xv[i].w + sk.w);
                }
                uint2 hi, lo;
                split4_f16_pre(lrelu02_s(pre.x), lrelu02_s(pre.y), lrelu02_s(pre.z), lrelu02_s(pre.w), hi, lo);
                if (active) {
                    const int sw = ar & 7;
                    *reinterpret_cast<uint2*>(a_t + ar * 128 + (((c4 >> 1) ^ sw) << 4) + (c4 & 1) * 8) = hi;
                    *reinterpret_cast<uint2*>(a_t + ar * 128 + (((4 + (c4 >> 1)) ^ sw) << 4) + (c4 & 1) * 8) = lo;
                    if (SKIP_FIRST && ar >= 28 && ar < 28 + LT_TT)   // xs of output row r = ar - 28, chunk c4 at position c4 ^ (r & 7)
                        *reinterpret_cast<float4*>(s_t + (ar - 28) * 128 + ((c4 ^ ((ar - 28) & 7)) << 4)) = pre;
                }
            }
        }
        LT_STAMP(3);
        fence_async_smem();
        group_sync(1 + g, GT);
        LT_STAMP(4);
        // ---------------- phase 2: dilated conv on tensor cores (+ the 2 extra rows on FFMA meanwhile) ----------------
        if (gw_u == 0) {
            tc_fence_after();
            uint32_t slot_t = slot_u, cw_t = cw_u;
            FD_OPAQUE2(slot_t, cw_t);   // opaque per tile: keeps ptxas from hoisting the descriptors out of the tile loop
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const uint32_t sh = (uint32_t)(27 + (k - 1) * dil) * 128u;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint64_t dah = umma_desc_sw128(slot_t + sh + j * 32), dal = umma_desc_sw128(slot_t + sh + 64 + j * 32);
                        const uint64_t dbh = umma_desc_sw128(cw_t + k * 4096 + j * 32), dbl = umma_desc_sw128(cw_t + k * 4096 + 64 + j * 32);
                        umma_f16(tmem_u, dah, dbh, idesc_conv, (k | j) ? 1u : 0u);
                        if (!(exp_mask & 4)) {
                            umma_f16(tmem_u, dah, dbl, idesc_conv, 1u);
                            umma_f16(tmem_u, dal, dbh, idesc_conv, 1u);
                        }
                    }
                }
                // second pass over A rows +128: its output rows 0 and 1 are the conv rows yr = 128, 129 that the LVC taps of the last
                // output rows need.  The other 126 rows read past the A tile (whatever bytes follow it in this slot) and are never used.
                if (!have_carry)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const uint32_t sh = (uint32_t)(128 + 27 + (k - 1) * dil) * 128u;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint64_t dah = umma_desc_sw128(slot_t + sh + j * 32), dal = umma_desc_sw128(slot_t + sh + 64 + j * 32);
                        const uint64_t dbh = umma_desc_sw128(cw_t + k * 4096 + j * 32), dbl = umma_desc_sw128(cw_t + k * 4096 + 64 + j * 32);
                        umma_f16(tmem_u + D2_COL, dah, dbh, idesc_conv, (k | j) ? 1u : 0u);
                        umma_f16(tmem_u + D2_COL, dah, dbl, idesc_conv, 1u);
                        umma_f16(tmem_u + D2_COL, dal, dbh, idesc_conv, 1u);
                    }
                }
                tc_commit(&bar[0]);
            }
            __syncwarp();
        }
        // ---------------- phase 3: y = lrelu(conv + b) -> fp16 pieces, rows of the Y tile (over the A tile) ----------------
        LT_STAMP(5);
        mbar_wait(&bar[0], parity);
        tc_fence_after();
        LT_STAMP(6);
        {
            const int q3 = gw & 3, part3 = gw >> 2;   // lane quarter / which 16 of the 32 conv channels
            // 16 accumulator columns of row yr -> S16_ACT * lrelu(acc*inv_c + b) (the pieces' prescale folded into the FFMA) -> pieces
            const float inv_cs = inv_c * S16_ACT;
            auto emit_row = [&](const uint32_t (&v)[16], int yr, unsigned char* copy_to) {
                const int t = t0 - 1 + yr;
                const bool in = (t >= 0 && t < T);
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    float y[8];
                    const int cb0 = part3 * 16 + cc * 8;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float tt2 = fmaf(__uint_as_float(v[cc * 8 + e]), inv_cs, cbs_s[cb0 + e]);
                        y[e] = in ? fmaxf(tt2, 0.2f * tt2) : 0.f;
                    }
                    uint2 h0, l0, h1, l1;
                    split4_f16_pre(y[0], y[1], y[2], y[3], h0, l0);
                    split4_f16_pre(y[4], y[5], y[6], y[7], h1, l1);
                    const int chunk = part3 * 2 + cc, sw = yr & 7;
                    *reinterpret_cast<uint4*>(a_t + yr * 128 + ((chunk ^ sw) << 4)) = make_uint4(h0.x, h0.y, h1.x, h1.y);
                    *reinterpret_cast<uint4*>(a_t + yr * 128 + (((4 + chunk) ^ sw) << 4)) = make_uint4(l0.x, l0.y, l1.x, l1.y);
                    if (copy_to) {   // rows 0, 1 again, for the next tile of this group (row yr of a 2-row image; 128 & 7 == 0, 129 & 7 == 1)
                        *reinterpret_cast<uint4*>(copy_to + yr * 128 + ((chunk ^ sw) << 4)) = make_uint4(h0.x, h0.y, h1.x, h1.y);
                        *reinterpret_cast<uint4*>(copy_to + yr * 128 + (((4 + chunk) ^ sw) << 4)) = make_uint4(l0.x, l0.y, l1.x, l1.y);
                    }
                }
            };
            uint32_t v[16];
            tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q3 * 32) << 16) + part3 * 16, v);
            tmem_ld_wait();
            emit_row(v, q3 * 32 + lane, (q3 == 0 && lane < 2) ? carry + parity * 256 : nullptr);
            if (have_carry) {
                if (gt < 16)   // rows 128, 129 = rows 0, 1 of the previous tile, already pieces in exactly this 256-byte layout
                    reinterpret_cast<uint4*>(a_t + 128 * 128)[gt] = reinterpret_cast<const uint4*>(carry + (parity ^ 1) * 256)[gt];
            } else if (q3 == 0) {   // rows 128, 129 = rows 0, 1 of the second pass (TMEM lanes 0, 1)
                tmem_ld_32x32b_x16(tmem_base + D2_COL + part3 * 16, v);
                tmem_ld_wait();
                if (lane < 2) emit_row(v, 128 + lane, nullptr);
            }
        }
        fence_async_smem();
        tc_fence_before();
        group_sync(1 + g, GT);
        LT_STAMP(7);
        // ---------------- phase 4: location-variable conv on tensor cores ----------------
        if (gw_u == 0) {
            tc_fence_after();
            uint32_t slot_t = slot_u;
            FD_OPAQUE(slot_t);
            if (elect_one()) {
#pragma unroll
                for (int fi = 0; fi < NF; ++fi) {
                    const uint32_t d = tmem_u + 32 + fi * 64;
                    const uint32_t lwb = slot_t + LW_OFF + fi * LH_LW_BYTES;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint64_t dah = umma_desc_sw128(slot_t + k * 128 + j * 32), dal = umma_desc_sw128(slot_t + k * 128 + 64 + j * 32);
                            const uint64_t dbh = umma_desc_sw128(lwb + k * 8192 + j * 32), dbl = umma_desc_sw128(lwb + k * 8192 + 64 + j * 32);
                            umma_f16(d, dah, dbh, idesc_lvc, (k | j) ? 1u : 0u);
                            if (!(exp_mask & 2)) {
                                umma_f16(d, dah, dbl, idesc_lvc, 1u);
                                umma_f16(d, dal, dbh, idesc_lvc, 1u);
                            }
                        }
                    }
                }
                tc_commit(&bar[1]);
            }
            __syncwarp();
        }
        // ---------------- phase 5: gate + residual -> global ----------------
        {
            const int q = gw & 3, part = gw >> 2;              // lane quarter / which 16 of the 32 gate channels
            const int r = q * 32 + lane, t = t0 + r;
            const int fi = (HOP >= LT_TT) ? 0 : r / HOP;       // warp-uniform (HOP is a multiple of 32)
            const size_t row = ((size_t)b * T + (t < T ? t : 0)) * C + part * 16;
            float4 xs[4], so[4];                               // residual base x + skip; skip to add to the output (skip_out)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                so[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (SKIP_FIRST) {
                    xs[c] = *reinterpret_cast<const float4*>(s_t + r * 128 + (((part * 4 + c) ^ (r & 7)) << 4));
                    if (skip_out) {   // first_conv(audio) at this row, 4 channels at a time (audio position t + k - 3 = au_s[29 + r + k])
                        const int o = part * 16 + c * 4;
                        float4 sk = *reinterpret_cast<const float4*>(fb_s + o);
#pragma unroll
                        for (int k = 0; k < 7; ++k) {
                            const float a = au_s[29 + r + k];
                            const float4 w = *reinterpret_cast<const float4*>(fw_s + k * C + o);
                            sk.x = fmaf(w.x, a, sk.x); sk.y = fmaf(w.y, a, sk.y); sk.z = fmaf(w.z, a, sk.z); sk.w = fmaf(w.w, a, sk.w);
                        }
                        so[c] = sk;
                    }
                } else {
                    xs[c] = *reinterpret_cast<const float4*>(x_in + row + c * 4);
                    if (skip_in || skip_out) {
                        const float4 sk = *reinterpret_cast<const float4*>(skip + row + c * 4);
                        if (skip_in) { xs[c].x += sk.x; xs[c].y += sk.y; xs[c].z += sk.z; xs[c].w += sk.w; }
                        if (skip_out) so[c] = sk;
                    }
                }
            }
            LT_STAMP(8);
            mbar_wait(&bar[1], parity);
            tc_fence_after();
            LT_STAMP(9);
            // LVC MMAs complete: the A/Y tile and the kernels are free -> fetch the next tile while this one is gated
            if (gw_u == 0 && tile - 1 >= tile_lo) {
                if (elect_one()) {   // tile - 1 = (b, tt - 1) shares this tile's frame iff hop 256 and tt is odd
                    issue_loads(tile - 1, (int)(parity ^ 1), HOP == 2 * LT_TT && (tt & 1) && !(exp_mask & 16));
                    if ((exp_mask & 8) && tile - 2 >= tile_lo) prefetch_l2(tile - 2);
                }
                __syncwarp();
            }
            uint32_t zs[16], zt[16];
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + 32 + fi * 64 + part * 16;
            tmem_ld_32x32b_x16(ta, zs);
            tmem_ld_32x32b_x16(ta + 32, zt);
            tmem_ld_wait();
            if (t < T) {
                const float* lb = lbias + fi * 64 + part * 16;
#if FD_VEC256
                float4 o_prev = make_float4(0.f, 0.f, 0.f, 0.f);
#endif
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float4 o4;
                    o4.x = xs[c].x + gate_st(fmaf(__uint_as_float(zs[c * 4 + 0]), inv_l, lb[c * 4 + 0]), fmaf(__uint_as_float(zt[c * 4 + 0]), inv_l, lb[32 + c * 4 + 0]));
                    o4.y = xs[c].y + gate_st(fmaf(__uint_as_float(zs[c * 4 + 1]), inv_l, lb[c * 4 + 1]), fmaf(__uint_as_float(zt[c * 4 + 1]), inv_l, lb[32 + c * 4 + 1]));
                    o4.z = xs[c].z + gate_st(fmaf(__uint_as_float(zs[c * 4 + 2]), inv_l, lb[c * 4 + 2]), fmaf(__uint_as_float(zt[c * 4 + 2]), inv_l, lb[32 + c * 4 + 2]));
                    o4.w = xs[c].w + gate_st(fmaf(__uint_as_float(zs[c * 4 + 3]), inv_l, lb[c * 4 + 3]), fmaf(__uint_as_float(zt[c * 4 + 3]), inv_l, lb[32 + c * 4 + 3]));
                    if (skip_out) {   // the next layer's "x += audio_down": (x + gate) + skip, the reference's rounding sequence
                        o4.x = __fadd_rn(o4.x, so[c].x); o4.y = __fadd_rn(o4.y, so[c].y); o4.z = __fadd_rn(o4.z, so[c].z); o4.w = __fadd_rn(o4.w, so[c].w);
                    }
#if FD_VEC256
                    if (c & 1) st_global_f8(x_out + row + (c - 1) * 4, o_prev, o4); else o_prev = o4;
#else
                    *reinterpret_cast<float4*>(x_out + row + c * 4) = o4;
#endif
                }
            }
        }
        tc_fence_before();
        group_sync(1 + g, GT);   // the group's TMEM columns and xs rows are free for its next tile
        if (--tt < 0) { tt = ntt - 1; --b; }
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) {
        tc_fence_after();
        tmem_dealloc(*tmem_base_s, 512u);
    }
}

}  // namespace fd
#include "fd_kernels_lvc_b0.cuh"   // k_lvc_layer_b0h: LVC block 0 (hop 8) in swapped-operand form
#include "fd_kernels_lvcp.cuh"     // k_lvc_p: LVC layers of blocks 1, 2 on the default path (piece-row protocol, warp-specialised pipeline)
#include "fd_kernels_up4.cuh"      // k_upsample_p4: block 2 upsampling + skip on kind::f16 pieces (default path)
namespace fd {

// ---------------------------------------------------------------------------------------------------------
// K5+K6+K7 on tensor cores (mode tc_3xf16): KernelPredictor hidden stack (modules.py:202-203, 328-329)
//   cond = mel + fc_t(e);  h0 = lrelu_.1(conv5(cond));  h = h0 + R(h0),  R = 6 x [conv3 + lrelu_.1]
// One tile = 128 frame rows (frames f0-8 .. f0+119) of one (LVC block, item); the 112 rows 8..119 are exact after the 7 layers
// (each conv eats one row -- two for the k = 5 input conv -- from either edge; rows outside [0, T') are forced to zero after
// every layer, which is the zero padding the reference's convs see).  All seven convs are kind::f16 MMAs on fp16 pieces:
//   A = activation tiles, rows = frames, 128 B = 64 channels (hi tile, lo tile; pieces of v*S16_HK); a tap = start + shift*128 B
//       (cond has 80 channels: channels 64..79 live in a third tile whose rows are [16 hi | 16 lo | unused])
//   B = weight tiles streamed from the blob (LBn_KPW_F16, 28 slots of 16 KB in consumption order) through a 6-slot smem ring by
//       one elected thread (cp.async.bulk + mbarrier), released by tcgen05.commit
//   D = 64 fp32 columns in TMEM; epilogue (16 warps: 4 per lane quarter, 16 channels each): acc*inv + bias -> lrelu -> zero outside
//       [0,T') -> pieces -> the other A tile pair.  h0 stays in registers (the thread <-> (row, channels) mapping is fixed), and
//       the last epilogue writes h = h0 + lrelu(.) as fp32 and as pieces (the kernel_conv GEMM's B operand).
// Persistent, grid = min(#tiles, #SM), 512 threads.
// ---------------------------------------------------------------------------------------------------------
constexpr int KT_VALID = 112, KT_PAD = 8;
constexpr int KT_TILE = (128 + 2 * KT_PAD) * 128;      // 18432 B
constexpr int KT_SLOT = 16384, KT_NSLOT = 6, KT_NLOAD = 28;
constexpr int KT_SMEM_BYTES = 5 * KT_TILE + KT_NSLOT * KT_SLOT + 7 * HID * 4 + COND * 4 + (2 * KT_NSLOT + 1) * 8 + 16 + 1024;

struct KpTcParams {
    const float* w16[NBLK];       // LBn_KPW_F16
    const float* in_b[NBLK];      // [64]
    const float* res_b[NBLK];     // [6][64]
    float inv[NBLK][8];           // 1 / (S16_HK * S_w) per layer (0: input conv, 1..6: residual convs)
};

__global__ void __launch_bounds__(512, 1)
k_kp_hidden_tc(const __grid_constant__ KpTcParams p, const float* __restrict__ mel, const float* __restrict__ cnoise,
               float* __restrict__ hk_all, float* __restrict__ hk_hi_all, float* __restrict__ hk_lo_all, int B, int Tm) {
    pdl_trigger();

    FD_DYN_SMEM(unsigned char, smem_raw);
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    // tiles: 0 = P_hi, 1 = P_lo (cond ch 0..63, then odd layers' output), 2 = Q_hi, 3 = Q_lo (even layers' output), 4 = cond ch 64..79
    unsigned char* ring = smem + 5 * KT_TILE;
    float* bias_s = (float*)(ring + KT_NSLOT * KT_SLOT);   // [7][64]
    float* cn_s = bias_s + 7 * HID;                         // [80] fc_t(e) of this (block, item)
    uint64_t* full_bar = (uint64_t*)(cn_s + COND);          // [6]
    uint64_t* empty_bar = full_bar + KT_NSLOT;              // [6]
    uint64_t* mma_bar = empty_bar + KT_NSLOT;               // [1]
    uint32_t* tmem_base_s = (uint32_t*)(mma_bar + 1);

    const int tid = threadIdx.x, lane = tid & 31;
    const int gw = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (tid == 0) {
        for (int i = 0; i < KT_NSLOT; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(mma_bar, 1);
        mbar_init_fence();
    }
    if (gw == 0) tmem_alloc(tmem_base_s, 64u);
    for (int i = tid; i < 5 * KT_TILE / 16; i += 512) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);   // pad rows stay 0
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;
    pdl_wait();   // programmatic dependent launch: everything above touched constants, shared memory and TMEM only
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t smem_u = smem_u32(smem);
    constexpr uint32_t idesc = umma_idesc_f16(128, 64);

    const int ntf = (Tm + KT_VALID - 1) / KT_VALID, per_blk = B * ntf, total = NBLK * per_blk;
    // ring counters, kept identical in every lane of warp 0 (the elected lane works on copies; the deterministic update is applied by
    // all lanes afterwards, so it does not matter which lane elect.sync picks next time).  Both advance KT_NLOAD per tile.
    uint32_t ld_issued = 0, ld_used = 0;
    uint32_t par_mma = 0;
    int cur_blk = -1;

    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int blk = tile / per_blk, rem = tile % per_blk, b = rem / ntf, f0 = (rem % ntf) * KT_VALID;
        const float* wsrc = FD_SEL3(p.w16, blk);
        if (blk != cur_blk) {   // biases of this block (the previous tile's last __syncthreads ordered its reads before these writes)
            if (tid < HID) bias_s[tid] = FD_SEL3(p.in_b, blk)[tid];
            if (tid < 6 * HID) bias_s[HID + tid] = FD_SEL3(p.res_b, blk)[tid];
            cur_blk = blk;
        }
        if (tid < COND) cn_s[tid] = cnoise[(blk * B + b) * COND + tid];
        __syncthreads();
        // ---- cond rows r = -2 .. 129 (frames f0-10 .. f0+121) -> pieces in tiles 0,1 (ch 0..63) and 4 (ch 64..79) ----
        for (int idx = tid; idx < COND * 132; idx += 512) {
            const int ci = idx / 132, rr = idx % 132, f = f0 - 10 + rr, row = KT_PAD - 2 + rr;
            float v = 0.f;   // the conv zero-pads cond (= mel + noise): outside [0,T') it is 0, not the noise
            if (f >= 0 && f < Tm) v = mel[((size_t)b * COND + ci) * Tm + f] + cn_s[ci];
            uint16_t h16, l16;
            f16_split(v, S16_HK, h16, l16);
            if (ci < 64) {
                const uint32_t off = (uint32_t)row * 128u + ((((uint32_t)ci >> 3) ^ ((uint32_t)row & 7u)) << 4) + ((uint32_t)ci & 7u) * 2u;
                *reinterpret_cast<uint16_t*>(smem + off) = h16;
                *reinterpret_cast<uint16_t*>(smem + KT_TILE + off) = l16;
            } else {
                const uint32_t c2 = (uint32_t)ci - 64u;   // hi chunk c2>>3 (0..1), lo chunk 2 + (c2>>3)
                unsigned char* t4 = smem + 4 * KT_TILE + (uint32_t)row * 128u;
                *reinterpret_cast<uint16_t*>(t4 + (((c2 >> 3) ^ ((uint32_t)row & 7u)) << 4) + (c2 & 7u) * 2u) = h16;
                *reinterpret_cast<uint16_t*>(t4 + (((2u + (c2 >> 3)) ^ ((uint32_t)row & 7u)) << 4) + (c2 & 7u) * 2u) = l16;
            }
        }
        fence_async_smem();
        __syncthreads();

        float h0[16];   // this thread's 16 channels of h0 (row = TMEM lane), kept for the final residual
#pragma unroll 1
        for (int layer = 0; layer < 7; ++layer, par_mma ^= 1) {
            const int src = (layer & 1) ? 2 : 0;   // layer 0 reads tiles 0,1(+4) -> writes 2,3; layer 1 reads 2,3 -> writes 0,1; ...
            if (gw == 0) {
                tc_fence_after();
                const int nslots = layer == 0 ? 10 : 3;
                const uint32_t tile_base_ld = ld_used - (ld_used % KT_NLOAD);
                const uint32_t ld_used_start = ld_used;
                if (elect_one()) {
                    const uint32_t at = smem_u + (uint32_t)(src * KT_TILE), x1t = smem_u + 4u * KT_TILE, ringu = smem_u + 5u * KT_TILE;
                    for (int sidx = 0; sidx < nslots; ++sidx) {
                        // keep the ring full: loads of this tile up to 6 ahead of the slot about to be consumed; load i goes to ring slot
                        // i % 6 once the MMAs that read its previous content have completed (empty barrier)
                        while (ld_issued < ld_used + KT_NSLOT && ld_issued < tile_base_ld + KT_NLOAD) {
                            const uint32_t rs = ld_issued % KT_NSLOT, li = ld_issued % KT_NLOAD;
                            if (ld_issued >= KT_NSLOT) mbar_wait(&empty_bar[rs], ((ld_issued / KT_NSLOT) - 1) & 1);
                            const bool half = (li < 10) && (li & 1);     // the ci 64..79 slots carry one 8 KB tile
                            const uint32_t bytes = half ? 8192u : 16384u;
                            mbar_expect_tx(&full_bar[rs], bytes);
                            bulk_g2s(ring + rs * KT_SLOT, wsrc + (size_t)li * (KT_SLOT / 4), bytes, &full_bar[rs]);
                            ++ld_issued;
                        }
                        const uint32_t rs = ld_used % KT_NSLOT;
                        mbar_wait(&full_bar[rs], (ld_used / KT_NSLOT) & 1);
                        tc_fence_after();
                        const uint32_t wt = ringu + rs * KT_SLOT;
                        const uint32_t w_lo = umma_desc_lo(wt), a_lo = umma_desc_lo(at);   // descriptors below: one add each (offsets in 16-byte units)
                        if (layer == 0 && (sidx & 1)) {   // cond channels 64..79 of tap j: one K = 16 step, pieces at byte 0 (hi) and 32 (lo)
                            const int j = sidx >> 1;
                            const uint32_t sh = (uint32_t)(KT_PAD + j - 2) * 128u;
                            const uint64_t da = umma_desc_sw128(x1t + sh), db = umma_desc_sw128(wt);
                            umma_f16(tmem_u, da, db, idesc, 1u);
                            umma_f16(tmem_u, da, db + 2, idesc, 1u);        // A_hi x W_lo (+32 B)
                            umma_f16(tmem_u, da + 2, db, idesc, 1u);        // A_lo x W_hi
                        } else {
                            const int j = layer == 0 ? (sidx >> 1) : sidx;
                            const uint32_t sh = (uint32_t)(KT_PAD + j - (layer == 0 ? 2 : 1)) * 8u;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                const uint64_t dah = umma_desc_at(a_lo + sh + ks * 2), dal = umma_desc_at(a_lo + KT_TILE / 16 + sh + ks * 2);
                                const uint64_t dbh = umma_desc_at(w_lo + ks * 2), dbl = umma_desc_at(w_lo + 512 + ks * 2);
                                umma_f16(tmem_u, dah, dbh, idesc, (sidx | ks) ? 1u : 0u);
                                umma_f16(tmem_u, dah, dbl, idesc, 1u);
                                umma_f16(tmem_u, dal, dbh, idesc, 1u);
                            }
                        }
                        tc_commit(&empty_bar[rs]);   // slot free once these MMAs have read it
                        ++ld_used;
                    }
                    tc_commit(mma_bar);
                }
                __syncwarp();
                ld_used = ld_used_start + nslots;   // what the elected lane did, applied in every lane
                ld_issued = min(ld_used + KT_NSLOT - 1, tile_base_ld + (uint32_t)KT_NLOAD);
            }
            mbar_wait(mma_bar, par_mma);
            tc_fence_after();
            {   // epilogue
                const int q = gw & 3, part = gw >> 2, row = q * 32 + lane, f = f0 - 8 + row;
                uint32_t v[16];
                tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + part * 16, v);
                tmem_ld_wait();
                const float inv = FD_SEL3(p.inv, blk)[layer];
                const float* bias = bias_s + layer * HID + part * 16;
                const bool in = (f >= 0 && f < Tm);
                float y[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) y[i] = in ? lrelu(fmaf(__uint_as_float(v[i]), inv, bias[i]), 0.1f) : 0.f;
                if (layer == 0) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) h0[i] = y[i];
                }
                if (layer < 6) {
                    unsigned char* dst = smem + (src ^ 2) * KT_TILE + (uint32_t)(row + KT_PAD) * 128u;
                    const uint32_t sw = (uint32_t)(row + KT_PAD) & 7u;
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        uint2 ha, la, hb, lb;
                        const float s = S16_HK / S16_ACT;   // split4_f16 prescales by S16_ACT
                        split4_f16(make_float4(y[cc * 8 + 0] * s, y[cc * 8 + 1] * s, y[cc * 8 + 2] * s, y[cc * 8 + 3] * s), ha, la);
                        split4_f16(make_float4(y[cc * 8 + 4] * s, y[cc * 8 + 5] * s, y[cc * 8 + 6] * s, y[cc * 8 + 7] * s), hb, lb);
                        const uint32_t pos = (((uint32_t)(part * 2 + cc)) ^ sw) << 4;
                        *reinterpret_cast<uint4*>(dst + pos) = make_uint4(ha.x, ha.y, hb.x, hb.y);
                        *reinterpret_cast<uint4*>(dst + KT_TILE + pos) = make_uint4(la.x, la.y, lb.x, lb.y);
                    }
                } else if (row >= 8 && row < 8 + KT_VALID && f < Tm) {
                    const size_t o = ((size_t)(blk * B + b) * (Tm + 2) + (size_t)(1 + f)) * HID + part * 16;
                    float hv[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) hv[i] = h0[i] + y[i];
#pragma unroll
                    for (int c = 0; c < 4; ++c) *reinterpret_cast<float4*>(hk_all + o + c * 4) = make_float4(hv[c * 4], hv[c * 4 + 1], hv[c * 4 + 2], hv[c * 4 + 3]);
                    const float s = S16_HK / S16_ACT;
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        uint2 ha, la, hb, lb;
                        split4_f16(make_float4(hv[cc * 8 + 0] * s, hv[cc * 8 + 1] * s, hv[cc * 8 + 2] * s, hv[cc * 8 + 3] * s), ha, la);
                        split4_f16(make_float4(hv[cc * 8 + 4] * s, hv[cc * 8 + 5] * s, hv[cc * 8 + 6] * s, hv[cc * 8 + 7] * s), hb, lb);
                        *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(hk_hi_all) + o + cc * 8) = make_uint4(ha.x, ha.y, hb.x, hb.y);
                        *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(hk_lo_all) + o + cc * 8) = make_uint4(la.x, la.y, lb.x, lb.y);
                    }
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (gw == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 64u);
    }
}

// ---------------------------------------------------------------------------------------------------------
// K3+K4  first_audio_conv + DiffusionDBlock 0 on tensor cores (FastDiff_model.py:89, modules.py:127-138):
//   xs[o] = first_conv(audio)[4 o]  (evaluated only at the kept positions);
//   out = conv_d4(lrelu(conv_d2(lrelu(conv_d1(lrelu(xs)))))) + W1x1 xs + b
// One tile = 128 rows (positions o0-7 .. o0+120), 114 valid outputs (halo 7 recomputed).  Same building blocks as the LVC
// kernel: SWIZZLE_128B K-major tiles, a dilated tap = start address + shift*128 B, 3xTF32 with the (hi,hi)/(hi,lo) passes
// merged into one N=64 MMA (W_hi | W_lo as one operand), the 1x1 residual accumulated into the last conv's TMEM tile, the
// audio window of the next tile fetched by cp.async.bulk one tile ahead.  One persistent CTA per SM, 16 warps.
// ---------------------------------------------------------------------------------------------------------
constexpr int DT_VALID = 114, DT_PAD = 8;
constexpr int DT_ATILE = (128 + 2 * DT_PAD) * 128;     // 18432 B per piece (hi or lo)
constexpr int DT_AU = 520;                             // audio window: positions 4*o0 - 32 .. + 519
constexpr int DT_SMEM_BYTES = 4 * DT_ATILE + 2 * 16384 + 3 * 24576 + 8192 + 2 * (DT_AU * 4 + 32) + (7 * C + C + 4 * C) * 4 + 64 + 1024;

struct DbTcParams {
    const float* cw_hi; const float* cw_lo;     // [3][3][32][8][4]
    const float* rw_hi; const float* rw_lo;     // [32][8][4]
    const float* conv_b; const float* res_b;    // [3][32], [32]
    const float* first_w; const float* first_b; // [7][32], [32]
};

__global__ void __launch_bounds__(512, 1)
k_dblock0_tc(DbTcParams p, const float* __restrict__ audio, float* __restrict__ out, int B, int L, int To, int three_pass) {
    pdl_trigger();

    FD_DYN_SMEM(unsigned char, smem_raw);
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* a_tile0 = smem;                                          // two ping-pong tiles, each hi | lo, rows 0..143 (row r at r + 8)
    unsigned char* x_hi = smem + 4 * DT_ATILE;                             // xs (pre-activation) for the 1x1 residual, 128 rows
    unsigned char* x_lo = x_hi + 16384;
    unsigned char* wc = x_lo + 16384;                                      // [3 layers][3 taps][hi 4 KB | lo 4 KB]
    unsigned char* wr = wc + 3 * 24576;                                    // residual: hi 4 KB | lo 4 KB
    float* au0 = (float*)(wr + 8192);                                      // [2][DT_AU + 8]
    float* fw_s = au0 + 2 * (DT_AU + 8);
    float* fb_s = fw_s + 7 * C;
    float* cb_s = fb_s + C;                                                // [3][32] conv biases, then [32] residual bias
    uint64_t* bar = (uint64_t*)(cb_s + 4 * C);                             // [0] MMAs, [1] audio loads
    uint32_t* tmem_base_s = (uint32_t*)(bar + 2);

    const int tid = threadIdx.x, gw = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init_fence(); }
    if (tid < 32) tmem_alloc(tmem_base_s, 64u);
    {
        const float4* sh = reinterpret_cast<const float4*>(p.cw_hi); const float4* sl = reinterpret_cast<const float4*>(p.cw_lo);
        for (int i = tid; i < 3 * 3 * 256; i += 512) {   // i = (layer*3 + tap)*256 + float4 within the 4 KB tap tile
            const int lt = i >> 8, w = i & 255;
            reinterpret_cast<float4*>(wc + lt * 8192)[w] = sh[i];
            reinterpret_cast<float4*>(wc + lt * 8192 + 4096)[w] = sl[i];
        }
        if (tid < 256) {
            reinterpret_cast<float4*>(wr)[tid] = reinterpret_cast<const float4*>(p.rw_hi)[tid];
            reinterpret_cast<float4*>(wr + 4096)[tid] = reinterpret_cast<const float4*>(p.rw_lo)[tid];
        }
        if (tid < 7 * C) fw_s[tid] = p.first_w[tid];
        if (tid < C) { fb_s[tid] = p.first_b[tid]; cb_s[3 * C + tid] = p.res_b[tid]; }
        if (tid < 3 * C) cb_s[tid] = p.conv_b[tid];
        for (int i = tid; i < 4 * DT_ATILE / 16; i += 512) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);   // pad rows stay 0
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;
    pdl_wait();   // programmatic dependent launch: everything above touched constants, shared memory and TMEM only
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const int gw_u = __shfl_sync(0xffffffffu, gw, 0);
    const uint32_t smem_u = smem_u32(smem);
    constexpr uint32_t idesc32 = umma_idesc_tf32(128, 32), idesc64 = umma_idesc_tf32(128, 64);

    const int c4 = tid & 7;
    float fwr[7][4], fbr[4];
#pragma unroll
    for (int k = 0; k < 7; ++k)
#pragma unroll
        for (int q = 0; q < 4; ++q) fwr[k][q] = fw_s[k * C + c4 * 4 + q];
#pragma unroll
    for (int q = 0; q < 4; ++q) fbr[q] = fb_s[c4 * 4 + q];

    const int ntt = (To + DT_VALID - 1) / DT_VALID, total = B * ntt;
    auto issue_audio = [&](int tile, int buf) {   // au[i] <-> audio position 4*o0 - 32 + i
        const int b = tile / ntt, o0 = (tile % ntt) * DT_VALID;
        const long p0 = 4L * o0 - 32;
        const int i0 = (int)(p0 < 0 ? -p0 : 0), i1 = (int)((L - p0) < DT_AU ? (L - p0) : DT_AU);
        const uint32_t bytes = i1 > i0 ? (uint32_t)(i1 - i0) * 4u : 0u;
        mbar_expect_tx(&bar[1], bytes);
        if (bytes) bulk_g2s(au0 + buf * (DT_AU + 8) + i0, audio + (size_t)b * L + (p0 + i0), bytes, &bar[1]);
    };
    int tile = blockIdx.x;
    if (tile < total && gw_u == 0) { if (elect_one()) issue_audio(tile, 0); __syncwarp(); }
    uint32_t par_ld = 0, par_mma = 0;
    for (; tile < total; tile += gridDim.x, par_ld ^= 1) {
        const int b = tile / ntt, o0 = (tile % ntt) * DT_VALID;
        float* au = au0 + par_ld * (DT_AU + 8);
        mbar_wait(&bar[1], par_ld);
        {   // audio positions outside [0, L) are zero (first_audio_conv zero-pads)
            const long p0 = 4L * o0 - 32;
            for (int i = tid; i < DT_AU; i += 512) if (p0 + i < 0 || p0 + i >= L) au[i] = 0.f;
        }
        __syncthreads();
        if (gw_u == 0 && tile + (int)gridDim.x < total) { if (elect_one()) issue_audio(tile + gridDim.x, (int)(par_ld ^ 1)); __syncwarp(); }
        // ---- xs = first_conv(audio) at the kept positions; A0 = lrelu(xs), X = xs (tf32 pieces) ----
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int row = (tid >> 3) + i * 64, o = o0 - 7 + row;
            float4 xs = make_float4(0.f, 0.f, 0.f, 0.f), a = xs;
            if (o >= 0 && o < To) {
                xs = make_float4(fbr[0], fbr[1], fbr[2], fbr[3]);
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    const float s1 = au[4 * row + k + 1];   // audio position 4 o + k - 3
                    xs.x = fmaf(fwr[k][0], s1, xs.x); xs.y = fmaf(fwr[k][1], s1, xs.y);
                    xs.z = fmaf(fwr[k][2], s1, xs.z); xs.w = fmaf(fwr[k][3], s1, xs.w);
                }
                a = make_float4(lrelu(xs.x, 0.2f), lrelu(xs.y, 0.2f), lrelu(xs.z, 0.2f), lrelu(xs.w, 0.2f));
            }
            float4 hi, lo;
            split4(a, hi, lo);
            *reinterpret_cast<float4*>(a_tile0 + swz128(row + DT_PAD, c4)) = hi;
            *reinterpret_cast<float4*>(a_tile0 + DT_ATILE + swz128(row + DT_PAD, c4)) = lo;
            split4(xs, hi, lo);
            *reinterpret_cast<float4*>(x_hi + swz128(row, c4)) = hi;
            *reinterpret_cast<float4*>(x_lo + swz128(row, c4)) = lo;
        }
        fence_async_smem();
        __syncthreads();
        // ---- three dilated convs, each: MMAs -> epilogue -> next tile ----
#pragma unroll 1
        for (int layer = 0; layer < 3; ++layer, par_mma ^= 1) {
            const int src = layer & 1, dil = 1 << layer;
            if (gw_u == 0) {
                tc_fence_after();
                uint32_t at = smem_u + (uint32_t)(src * 2 * DT_ATILE), wt = smem_u + 4 * DT_ATILE + 2 * 16384 + (uint32_t)(layer * 24576);
                FD_OPAQUE2(at, wt);
                if (elect_one()) {
                    const uint32_t a_lo = umma_desc_lo(at), w_lo = umma_desc_lo(wt);   // descriptors: one add each (offsets in 16-byte units)
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const uint32_t sh = (uint32_t)(DT_PAD + (k - 1) * dil) * 8u;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint64_t dah = umma_desc_at(a_lo + sh + j * 2), dal = umma_desc_at(a_lo + DT_ATILE / 16 + sh + j * 2);
                            const uint64_t db = umma_desc_at(w_lo + k * 512 + j * 2);   // rows 0-31 W_hi, 32-63 W_lo
                            if (three_pass) {
                                umma_tf32(tmem_u, dah, db, idesc64, (k | j) ? 1u : 0u);
                                umma_tf32(tmem_u, dal, db, idesc32, 1u);
                            } else {
                                umma_tf32(tmem_u, dah, db, idesc32, (k | j) ? 1u : 0u);
                            }
                        }
                    }
                    if (layer == 2) {   // + W1x1 xs accumulated into the same tile
                        const uint32_t xt = smem_u + 4 * DT_ATILE, rt = smem_u + 4 * DT_ATILE + 2 * 16384 + 3 * 24576;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint64_t dxh = umma_desc_sw128(xt + j * 32), dxl = umma_desc_sw128(xt + 16384 + j * 32);
                            const uint64_t dr = umma_desc_sw128(rt + j * 32);
                            if (three_pass) {
                                umma_tf32(tmem_u, dxh, dr, idesc64, 1u);
                                umma_tf32(tmem_u, dxl, dr, idesc32, 1u);
                            } else {
                                umma_tf32(tmem_u, dxh, dr, idesc32, 1u);
                            }
                        }
                    }
                    tc_commit(&bar[0]);
                }
                __syncwarp();
            }
            mbar_wait(&bar[0], par_mma);
            tc_fence_after();
            {   // epilogue: 4 warps per lane quarter, 8 channels each
                const int q = gw & 3, part = gw >> 2, row = q * 32 + lane, o = o0 - 7 + row;
                uint32_t v[8];
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + part * 8;
                tmem_ld_cols<8>(ta, v);
                if (three_pass) {
                    uint32_t v2[8];
                    tmem_ld_cols<8>(ta + 32, v2);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(v2[i]));
                }
                tmem_ld_wait();
                const float* bias = cb_s + layer * C + part * 8;
                if (layer < 2) {
                    const bool in = (o >= 0 && o < To);
                    unsigned char* dst = a_tile0 + (src ^ 1) * 2 * DT_ATILE;
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        float4 y;
                        y.x = in ? lrelu(__uint_as_float(v[cc * 4 + 0]) + bias[cc * 4 + 0], 0.2f) : 0.f;
                        y.y = in ? lrelu(__uint_as_float(v[cc * 4 + 1]) + bias[cc * 4 + 1], 0.2f) : 0.f;
                        y.z = in ? lrelu(__uint_as_float(v[cc * 4 + 2]) + bias[cc * 4 + 2], 0.2f) : 0.f;
                        y.w = in ? lrelu(__uint_as_float(v[cc * 4 + 3]) + bias[cc * 4 + 3], 0.2f) : 0.f;
                        float4 hi, lo;
                        split4(y, hi, lo);
                        *reinterpret_cast<float4*>(dst + swz128(row + DT_PAD, part * 2 + cc)) = hi;
                        *reinterpret_cast<float4*>(dst + DT_ATILE + swz128(row + DT_PAD, part * 2 + cc)) = lo;
                    }
                } else if (row >= 7 && row < 7 + DT_VALID && o < To) {
                    const float* rb = cb_s + 3 * C + part * 8;
                    float4 o0v, o1v;
                    o0v.x = __uint_as_float(v[0]) + bias[0] + rb[0]; o0v.y = __uint_as_float(v[1]) + bias[1] + rb[1];
                    o0v.z = __uint_as_float(v[2]) + bias[2] + rb[2]; o0v.w = __uint_as_float(v[3]) + bias[3] + rb[3];
                    o1v.x = __uint_as_float(v[4]) + bias[4] + rb[4]; o1v.y = __uint_as_float(v[5]) + bias[5] + rb[5];
                    o1v.z = __uint_as_float(v[6]) + bias[6] + rb[6]; o1v.w = __uint_as_float(v[7]) + bias[7] + rb[7];
                    float* dstg = out + ((size_t)b * To + o) * C + part * 8;
#if FD_VEC256
                    st_global_f8(dstg, o0v, o1v);
#else
                    *reinterpret_cast<float4*>(dstg) = o0v;
                    *reinterpret_cast<float4*>(dstg + 4) = o1v;
#endif
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) {
        tc_fence_after();
        tmem_dealloc(*tmem_base_s, 64u);
    }
}

// ---------------------------------------------------------------------------------------------------------
// K10  LVC-block upsample on tensor cores: lrelu_.2 -> ConvTranspose1d(32,32,k=2r,stride=r,pad=r/2)  (modules.py:205-206)
//   out[r m + ph] = b + lrelu(in[m + sh]) W[kk1] + lrelu(in[m + sh - 1]) W[kk1 + r],   sh = (ph + r/2)/r, kk1 = (ph + r/2) % r
// One tile = 128 input rows (+1 halo row either side) -> 128 r output rows.  For every output phase ph the two taps are two
// row-shifted reads of the SAME SWIZZLE_128B input tile; the r phase accumulators (64 TMEM columns each: merged tf32 passes)
// are all live at once and one epilogue interleaves them into consecutive output rows.  Input rows of the next tile arrive by
// cp.async.bulk while the current tile is written out.  16 warps; r = 4: 2 CTAs/SM (99 KB smem, 256 TMEM columns each).
// ---------------------------------------------------------------------------------------------------------
constexpr int UT_AROWS = 136;                          // input rows m0-1 .. m0+128 (+pad)
constexpr int UT_ATILE = UT_AROWS * 128;               // 17408 B per piece
constexpr int UT_PEXTRA = 6144;                        // POUT, r = 4: two audio windows (128 r + 8 floats each) + first conv taps and bias
template <int R> constexpr int ut_smem_bytes() { return 2 * UT_ATILE + 2 * R * 8192 + 256 + 64 + 1024; }

// POUT (the default path of mode tc_3xf16): the epilogue adds the block's skip (r = 8: rows of the DBlock output; r = 4:
// first_conv(audio), 7 taps) -- the reference's first "x += audio_down" -- and writes z = up + skip as the PIECE ROWS the first LVC
// layer consumes (fd_kernels_lvcp.cuh: fp16 hi | lo of 16 * lrelu(z), chunk c at c ^ (t & 7), padded row layout) instead of fp32 rows.
struct UpPOut {
    const float* skip;       // r = 8: (B, T, 32) rows; r = 4: audio (B, T)
    const float* first_w;    // [7][32]  (r = 4)
    const float* first_b;    // [32]
    float* p_out;            // padded piece rows (B items of T = r Tin rows)
    unsigned int* sat;
};

// Optional phase timeline (-DUT_TIMELINE=1, GPU build): thread 0 of CTA 0 of the r = 4 piece-row launch stamps clock64 per tile (8 slots x 24 tiles).
#if defined(UT_TIMELINE) && !defined(FD_EMU)
__device__ unsigned long long g_ut_timeline[24 * 8];
#define UT_STAMP(slot) do { if (R == 4 && POUT && blockIdx.x == 0 && tid == 0 && tl_n < 24) g_ut_timeline[tl_n * 8 + (slot)] = clock64(); } while (0)
#else
#define UT_STAMP(slot) do { } while (0)
#endif
template <int R, bool POUT = false>
__global__ void __launch_bounds__(512, (R == 4 ? 2 : 1))
k_upsample_tc(const float* __restrict__ w_hi, const float* __restrict__ w_lo, const float* __restrict__ bias,
              const float* __restrict__ in, float* __restrict__ out, int B, int Tin, int three_pass, const UpPOut po = UpPOut()) {
    pdl_trigger();
    FD_DYN_SMEM(unsigned char, smem_raw);
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* a_hi = smem;                          // tile row ar <-> input row m0 - 1 + ar
    unsigned char* a_lo = a_hi + UT_ATILE;               // raw rows land here
    unsigned char* wt = a_lo + UT_ATILE;                 // [2R taps][hi 32 rows 4 KB | lo 32 rows 4 KB]
    float* b_s = (float*)(wt + 2 * R * 8192);
    uint64_t* bar = (uint64_t*)(b_s + 64);               // [0] MMAs, [1] loads
    uint32_t* tmem_base_s = (uint32_t*)(bar + 2);
    float* au_s = (float*)(tmem_base_s + 4);             // POUT, r = 4: [2] audio positions r m0 - 4 .. r m0 + 128 r + 3 (bulk-copied one tile ahead)
    constexpr int AUW = 128 * R + 8;
    float* pfw_s = au_s + 2 * AUW;                       // [7][32]
    float* pfb_s = pfw_s + 7 * C;                        // [32]
    constexpr uint32_t NCOLS = R * 64;

    const int tid = threadIdx.x, gw = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init_fence(); }
    if (tid < 32) tmem_alloc(tmem_base_s, NCOLS);
    if (POUT && R == 4) {
        if (tid < 7 * C) pfw_s[tid] = po.first_w[tid];
        if (tid < C) pfb_s[tid] = po.first_b[tid];
    }
    float vmax = 0.f;
    for (int i = tid; i < 2 * R * 256; i += 512) {   // i = tap*256 + float4 within the 4 KB tap tile
        const int k = i >> 8, w = i & 255;
        reinterpret_cast<float4*>(wt + k * 8192)[w] = reinterpret_cast<const float4*>(w_hi)[i];
        reinterpret_cast<float4*>(wt + k * 8192 + 4096)[w] = reinterpret_cast<const float4*>(w_lo)[i];
    }
    if (tid < C) b_s[tid] = bias[tid];
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;
    pdl_wait();   // programmatic dependent launch: everything above touched constants, shared memory and TMEM only
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const int gw_u = __shfl_sync(0xffffffffu, gw, 0);
    const uint32_t smem_u = smem_u32(smem);
    constexpr uint32_t idesc32 = umma_idesc_tf32(128, 32), idesc64 = umma_idesc_tf32(128, 64);

    const int ntt = (Tin + 127) / 128, total = B * ntt;
    auto issue_rows = [&](int tile, uint32_t buf) {
        const int b = tile / ntt, m0 = (tile % ntt) * 128;
        const int ar0 = m0 == 0 ? 1 : 0, ar1 = min(130, Tin - m0 + 1);   // rows inside [0, Tin)
        const uint32_t bytes = (uint32_t)(ar1 - ar0) * 128u;
        // POUT, r = 4: the audio window of the tile's outputs rides on the same barrier (positions inside [0, Tout): whole 16-byte groups)
        const int i0 = m0 == 0 ? 4 : 0, i1 = min(AUW, R * (Tin - m0) + 4);
        const uint32_t abytes = (POUT && R == 4) ? (uint32_t)(i1 - i0) * 4u : 0u;
        mbar_expect_tx(&bar[1], bytes + abytes);
        bulk_g2s(a_lo + ar0 * 128, in + ((size_t)b * Tin + (m0 - 1 + ar0)) * C, bytes, &bar[1]);
        if (POUT && R == 4) bulk_g2s(au_s + buf * AUW + i0, po.skip + (size_t)b * Tin * R + (R * m0 - 4 + i0), abytes, &bar[1]);
    };
    int tile = blockIdx.x;
    if (tile < total && gw_u == 0) { if (elect_one()) issue_rows(tile, 0u); __syncwarp(); }
    uint32_t parity = 0;
    [[maybe_unused]] int tl_n = 0;
    for (; tile < total; tile += gridDim.x, parity ^= 1, ++tl_n) {
        const int b = tile / ntt, m0 = (tile % ntt) * 128;
        UT_STAMP(0);
        const float* au_t = au_s + parity * AUW;   // this tile's window (buffer = tile parity)
        if (POUT && R == 4) {   // positions outside the utterance are zero (the first conv zero-pads): utterance ends only; the copy never touches them
            const int Tout = Tin * R;
            if (m0 == 0 || R * m0 + AUW - 4 > Tout)
                for (int i = tid; i < AUW; i += 512) {
                    const int pos = R * m0 - 4 + i;
                    if (pos < 0 || pos >= Tout) au_s[parity * AUW + i] = 0.f;
                }
        }
        mbar_wait(&bar[1], parity);
        UT_STAMP(1);
        // ---- lrelu + tf32 split, in place (8 lanes per row) ----
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int ar = (tid >> 3) + i * 64, c4 = tid & 7, m = m0 - 1 + ar;
            const bool active = ar < 130;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (active && m >= 0 && m < Tin) {
                const float4 x = *reinterpret_cast<const float4*>(a_lo + ar * 128 + c4 * 16);
                v = make_float4(lrelu(x.x, 0.2f), lrelu(x.y, 0.2f), lrelu(x.z, 0.2f), lrelu(x.w, 0.2f));
            }
            float4 hi, lo;
            split4(v, hi, lo);
            __syncwarp();
            if (active) {
                *reinterpret_cast<float4*>(a_hi + swz128(ar, c4)) = hi;
                *reinterpret_cast<float4*>(a_lo + swz128(ar, c4)) = lo;
            }
        }
        fence_async_smem();
        __syncthreads();
        UT_STAMP(2);
        if (gw_u == 0) {
            tc_fence_after();
            uint32_t at = smem_u, wu = smem_u + 2 * UT_ATILE;
            FD_OPAQUE2(at, wu);
            if (elect_one()) {
                const uint32_t a_lo = umma_desc_lo(at), w_lo = umma_desc_lo(wu);   // descriptors: one add each (offsets in 16-byte units)
#pragma unroll
                for (int ph = 0; ph < R; ++ph) {
                    const int sh = (ph + R / 2) / R, kk1 = (ph + R / 2) % R;
                    const uint32_t d = tmem_u + ph * 64;
#pragma unroll
                    for (int tap = 0; tap < 2; ++tap) {   // tap 0: input m + sh (weights kk1); tap 1: input m + sh - 1 (weights kk1 + R)
                        const uint32_t arow = (uint32_t)(1 + sh - tap) * 8u, kk = (uint32_t)(kk1 + tap * R);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint64_t dah = umma_desc_at(a_lo + arow + j * 2), dal = umma_desc_at(a_lo + UT_ATILE / 16 + arow + j * 2);
                            const uint64_t db = umma_desc_at(w_lo + kk * 512 + j * 2);
                            if (three_pass) {
                                umma_tf32(d, dah, db, idesc64, (tap | j) ? 1u : 0u);
                                umma_tf32(d, dal, db, idesc32, 1u);
                            } else {
                                umma_tf32(d, dah, db, idesc32, (tap | j) ? 1u : 0u);
                            }
                        }
                    }
                }
                tc_commit(&bar[0]);
            }
            __syncwarp();
        }
        UT_STAMP(3);
        mbar_wait(&bar[0], parity);
        UT_STAMP(4);
        tc_fence_after();
        if (gw_u == 0 && tile + (int)gridDim.x < total) { if (elect_one()) issue_rows(tile + gridDim.x, parity ^ 1u); __syncwarp(); }   // tile is free
        {   // epilogue: thread = (input row m, 8 channels); its r outputs are consecutive rows r m + ph
            const int q = gw & 3, part = gw >> 2, m = m0 + q * 32 + lane;
            const float* bb = b_s + part * 8;
            float* dst = out + (((size_t)b * Tin + (m < Tin ? m : 0)) * R) * C + part * 8;
#pragma unroll
            for (int ph = 0; ph < R; ++ph) {
                uint32_t v[8];
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + ph * 64 + part * 8;
                tmem_ld_cols<8>(ta, v);
                if (three_pass) {
                    uint32_t v2[8];
                    tmem_ld_cols<8>(ta + 32, v2);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(v2[i]));
                }
                tmem_ld_wait();
                if (POUT) {
                    if (m < Tin) {
                        const int Tout = Tin * R, t = R * m + ph;
                        float sk[8];
                        if (R == 4) {
                            const float4 b0 = *reinterpret_cast<const float4*>(pfb_s + part * 8), b1 = *reinterpret_cast<const float4*>(pfb_s + part * 8 + 4);
                            sk[0] = b0.x; sk[1] = b0.y; sk[2] = b0.z; sk[3] = b0.w; sk[4] = b1.x; sk[5] = b1.y; sk[6] = b1.z; sk[7] = b1.w;
#pragma unroll
                            for (int k = 0; k < 7; ++k) {
                                const float x = au_t[t - R * m0 + 1 + k];   // audio position t + k - 3
                                const float4 w0 = *reinterpret_cast<const float4*>(pfw_s + k * C + part * 8), w1 = *reinterpret_cast<const float4*>(pfw_s + k * C + part * 8 + 4);
                                sk[0] = fmaf(w0.x, x, sk[0]); sk[1] = fmaf(w0.y, x, sk[1]); sk[2] = fmaf(w0.z, x, sk[2]); sk[3] = fmaf(w0.w, x, sk[3]);
                                sk[4] = fmaf(w1.x, x, sk[4]); sk[5] = fmaf(w1.y, x, sk[5]); sk[6] = fmaf(w1.z, x, sk[6]); sk[7] = fmaf(w1.w, x, sk[7]);
                            }
                        } else {
                            const float4* sp = reinterpret_cast<const float4*>(po.skip + ((size_t)b * Tout + t) * C + part * 8);
                            const float4 s0 = sp[0], s1 = sp[1];
                            sk[0] = s0.x; sk[1] = s0.y; sk[2] = s0.z; sk[3] = s0.w; sk[4] = s1.x; sk[5] = s1.y; sk[6] = s1.z; sk[7] = s1.w;
                        }
                        float z[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) z[i] = lrelu02_s(__fadd_rn(__uint_as_float(v[i]) + bb[i], sk[i]));   // up + skip, the reference's order
                        uint4 hi, lo;
                        lp_split8(z, hi, lo, vmax);
                        uint4* dst = reinterpret_cast<uint4*>(po.p_out + lp_row_of(b, Tout, t) * C);
                        const int sw = t & 7;
                        dst[part ^ sw] = hi;
                        dst[(4 + part) ^ sw] = lo;
                    }
                } else
                if (m < Tin) {
                    float4 o0 = make_float4(__uint_as_float(v[0]) + bb[0], __uint_as_float(v[1]) + bb[1], __uint_as_float(v[2]) + bb[2], __uint_as_float(v[3]) + bb[3]);
                    float4 o1 = make_float4(__uint_as_float(v[4]) + bb[4], __uint_as_float(v[5]) + bb[5], __uint_as_float(v[6]) + bb[6], __uint_as_float(v[7]) + bb[7]);
#if FD_VEC256
                    st_global_f8(dst + ph * C, o0, o1);
#else
                    *reinterpret_cast<float4*>(dst + ph * C) = o0;
                    *reinterpret_cast<float4*>(dst + ph * C + 4) = o1;
#endif
                }
            }
        }
        UT_STAMP(5);
        tc_fence_before();
        __syncthreads();
        UT_STAMP(6);
    }
    if (POUT && po.sat && vmax > F16_MAX) *po.sat = 1u;
    tc_fence_before();
    __syncthreads();
    if (tid < 32) {
        tc_fence_after();
        tmem_dealloc(*tmem_base_s, NCOLS);
    }
}

// x_in/x_out: (B,T,32); skip: (B,T,32) buffer or (block 2) the audio (B,T); kern: this layer's slice of the predicted kernels.
#ifndef FD_EMU
static inline int tc_lvc_layer(void* state, int mode, int blk, int layer, const float* x_in, const float* skip, const float* kern,
                               float* x_out, int B, int T, int Tm, int dil, cudaStream_t st, std::string& err, uint64_t* launches,
                               bool* done) {
    *done = false;
    TcState* s = (TcState*)state;
    if (!s || !s->ok) { err = "tensor-core path not initialised"; return -4; }
    if (blk == 0) return 0;   // hop 8: stays on the SIMT kernel
    if (mode == FD_MODE_TC_3XF16) {
        LvcHParams hp;
        hp.cw16 = s->blob + s->sec_off[blk == 1 ? FD_S_LB1_CONV_F16 : FD_S_LB2_CONV_F16] + (size_t)layer * (LH_CW_BYTES / 4);
        hp.conv_b = s->blob + s->sec_off[FD_S_LB0_CONV_B + blk * FD_LB_STRIDE] + layer * C;
        hp.first_w = s->blob + s->sec_off[FD_S_FIRST_W];
        hp.first_b = s->blob + s->sec_off[FD_S_FIRST_B];
        const float inv_c = 1.f / (S16_ACT * s->scales16[4 + 4 * blk + layer]), inv_l = 1.f / (S16_ACT * S16_KERN);
        const int tiles = B * ((T + LT_TT - 1) / LT_TT);
        // where the skip is added: block 1 adds it to the rows it produces (skip_out; only the first layer adds it on the way in), which
        // drops the skip tile from the loads (-7 %); block 2 recomputes first_conv(audio) on the way in for every layer (in its gate
        // epilogue the same work sits on the critical path: +7 %)
        const int skip_in = (blk == 2 || layer == 0) ? 1 : 0, skip_out = (blk == 1 && layer < LAYERS - 1) ? 1 : 0;
        const int ng = (blk == 2 && s->lvc_groups == 3) ? 3 : 2;
        const int per = (tiles + ng - 1) / ng, grid = per < s->sm_count ? per : s->sm_count;
        if (blk == 1)     k_lvc_layer_h<64, false, 2><<<grid, 512, lh_smem_bytes<64, false, 2>(), st>>>(hp, x_in, skip, kern, x_out, B, T, Tm, dil, inv_c, inv_l, s->lvc_exp, skip_in, skip_out);
        else if (ng == 3) k_lvc_layer_h<256, true, 3><<<grid, 768, lh_smem_bytes<256, true, 3>(), st>>>(hp, x_in, skip, kern, x_out, B, T, Tm, dil, inv_c, inv_l, s->lvc_exp, skip_in, skip_out);
        else              k_lvc_layer_h<256, true, 2><<<grid, 512, lh_smem_bytes<256, true, 2>(), st>>>(hp, x_in, skip, kern, x_out, B, T, Tm, dil, inv_c, inv_l, s->lvc_exp, skip_in, skip_out);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { err = std::string("launch of k_lvc_layer_h failed: ") + cudaGetErrorString(e); return -3; }
        ++*launches;
        *done = true;
        return 0;
    }
    LvcTcParams p;
    p.cw_hi = s->blob + s->sec_off[FD_S_LB0_CONVT_HI + blk * FD_LB_STRIDE] + (size_t)layer * 3 * 8 * C * 4;
    p.cw_lo = s->blob + s->sec_off[FD_S_LB0_CONVT_LO + blk * FD_LB_STRIDE] + (size_t)layer * 3 * 8 * C * 4;
    p.conv_b = s->blob + s->sec_off[FD_S_LB0_CONV_B + blk * FD_LB_STRIDE] + layer * C;
    p.first_w = s->blob + s->sec_off[FD_S_FIRST_W];
    p.first_b = s->blob + s->sec_off[FD_S_FIRST_B];
    const int total = B * ((T + LT_TT - 1) / LT_TT);
    const int tp = mode != FD_MODE_TC_TF32 ? 1 : 0;
    if (blk == 1) {
        const int grid = total < s->sm_count ? total : s->sm_count;
        k_lvc_layer_tc<64, false, 1, 16><<<grid, 512, lt_smem_bytes<64, 1>(), st>>>(p, x_in, skip, kern, x_out, B, T, Tm, dil, tp);
    } else {
        const int pairs = (total + 1) / 2, grid = pairs < s->sm_count ? pairs : s->sm_count;
        k_lvc_layer_tc<256, true, 2, 8><<<grid, 512, lt_smem_bytes<256, 2>(), st>>>(p, x_in, skip, kern, x_out, B, T, Tm, dil, tp);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("launch of k_lvc_layer_tc failed: ") + cudaGetErrorString(e); return -3; }
    ++*launches;
    *done = true;
    return 0;
}

// LVC block 0 on tensor cores (k_lvc_layer_b0h); the kernel's shared-memory attribute is set here on first use.
static inline int tc_lvc_layer_b0(void* state, int layer, const float* x_in, const float* skip, const float* kern, float* x_out,
                                  int B, int T, int Tm, int dil, cudaStream_t st, std::string& err, uint64_t* launches) {
    TcState* s = (TcState*)state;
    if (!s || !s->ok) { err = "tensor-core path not initialised"; return -4; }
    if (!s->b0_attr_set) {
        cudaError_t ea = cudaFuncSetAttribute(k_lvc_layer_b0h, cudaFuncAttributeMaxDynamicSharedMemorySize, LB0_SMEM_BYTES);
        if (ea != cudaSuccess) { err = std::string("k_lvc_layer_b0h: shared-memory attribute: ") + cudaGetErrorString(ea); return -3; }
        s->b0_attr_set = 1;
    }
    LvcHParams hp;
    hp.cw16 = s->blob + s->sec_off[FD_S_LB0_CONV_F16] + (size_t)layer * (LH_CW_BYTES / 4);
    hp.conv_b = s->blob + s->sec_off[FD_S_LB0_CONV_B] + layer * C;
    hp.first_w = nullptr; hp.first_b = nullptr;
    const float inv_c = 1.f / (S16_ACT * s->scales16[4 + layer]), inv_l = 1.f / (S16_ACT * S16_KERN);
    const int tiles = B * ((T + LT_TT - 1) / LT_TT), grid = tiles < s->sm_count ? tiles : s->sm_count;
    fd_launch_pdl(k_lvc_layer_b0h, dim3(grid), dim3(512), LB0_SMEM_BYTES, st, hp, x_in, skip, kern, x_out, B, T, Tm, dil, inv_c, inv_l, layer == 0 ? 1 : 0, layer < LAYERS - 1 ? 1 : 0);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("launch of k_lvc_layer_b0h failed: ") + cudaGetErrorString(e); return -3; }
    ++*launches;
    return 0;
}

// in (B,Tin,32) -> out (B,Tin*r,32) for LVC blocks 1 (r = 8) and 2 (r = 4)
static inline int tc_upsample(void* state, int mode, int blk, const float* in, float* out, int B, int Tin, cudaStream_t st, std::string& err,
                              uint64_t* launches) {
    TcState* s = (TcState*)state;
    if (!s || !s->ok) { err = "tensor-core path not initialised"; return -4; }
    const float* wh = s->blob + s->sec_off[blk == 1 ? FD_S_LB1_UPT_HI : FD_S_LB2_UPT_HI];
    const float* wl = s->blob + s->sec_off[blk == 1 ? FD_S_LB1_UPT_LO : FD_S_LB2_UPT_LO];
    const float* bias = s->blob + s->sec_off[FD_S_LB0_UP_B + blk * FD_LB_STRIDE];
    const int total = B * ((Tin + 127) / 128);
    if (blk == 1) {
        const int grid = total < s->sm_count ? total : s->sm_count;
        k_upsample_tc<8><<<grid, 512, ut_smem_bytes<8>(), st>>>(wh, wl, bias, in, out, B, Tin, mode != FD_MODE_TC_TF32 ? 1 : 0);
    } else {
        const int grid = total < 2 * s->sm_count ? total : 2 * s->sm_count;
        k_upsample_tc<4><<<grid, 512, ut_smem_bytes<4>(), st>>>(wh, wl, bias, in, out, B, Tin, mode != FD_MODE_TC_TF32 ? 1 : 0);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("launch of k_upsample_tc failed: ") + cudaGetErrorString(e); return -3; }
    ++*launches;
    return 0;
}

// audio (B,L) -> d0 (B, L/4, 32)
static inline int tc_dblock0(void* state, int mode, const float* audio, float* d0, int B, int L, cudaStream_t st, std::string& err,
                             uint64_t* launches) {
    TcState* s = (TcState*)state;
    if (!s || !s->ok) { err = "tensor-core path not initialised"; return -4; }
    DbTcParams p;
    p.cw_hi = s->blob + s->sec_off[FD_S_DB0_CONVT_HI]; p.cw_lo = s->blob + s->sec_off[FD_S_DB0_CONVT_LO];
    p.rw_hi = s->blob + s->sec_off[FD_S_DB0_REST_HI];  p.rw_lo = s->blob + s->sec_off[FD_S_DB0_REST_LO];
    p.conv_b = s->blob + s->sec_off[FD_S_DB0_CONV_B];  p.res_b = s->blob + s->sec_off[FD_S_DB0_RES_B];
    p.first_w = s->blob + s->sec_off[FD_S_FIRST_W];    p.first_b = s->blob + s->sec_off[FD_S_FIRST_B];
    const int To = L / 4, total = B * ((To + DT_VALID - 1) / DT_VALID);
    const int grid = total < s->sm_count ? total : s->sm_count;
    fd_launch_pdl(k_dblock0_tc, dim3(grid), dim3(512), DT_SMEM_BYTES, st, p, audio, d0, B, L, To, mode != FD_MODE_TC_TF32 ? 1 : 0);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("launch of k_dblock0_tc failed: ") + cudaGetErrorString(e); return -3; }
    ++*launches;
    return 0;
}

static inline int tc_kp_hidden(void* state, const float* mel, const float* cnoise, float* hk, float* hk_hi, float* hk_lo, int B, int Tm,
                               cudaStream_t st, std::string& err, uint64_t* launches) {
    TcState* s = (TcState*)state;
    if (!s || !s->ok) { err = "tensor-core path not initialised"; return -4; }
    KpTcParams p;
    for (int n = 0; n < NBLK; ++n) {
        p.w16[n] = s->blob + s->sec_off[FD_S_LB0_KPW_F16 + n];
        p.in_b[n] = s->blob + s->sec_off[FD_S_LB0_KPIN_B + n * FD_LB_STRIDE];
        p.res_b[n] = s->blob + s->sec_off[FD_S_LB0_KPRES_B + n * FD_LB_STRIDE];
        for (int l = 0; l < 7; ++l) p.inv[n][l] = 1.f / (S16_HK * s->scales16[16 + 8 * n + l]);
        p.inv[n][7] = 0.f;
    }
    const int total = NBLK * B * ((Tm + KT_VALID - 1) / KT_VALID);
    const int grid = total < s->sm_count ? total : s->sm_count;
    fd_launch_pdl(k_kp_hidden_tc, dim3(grid), dim3(512), KT_SMEM_BYTES, st, p, mel, cnoise, hk, hk_hi, hk_lo, B, Tm);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("launch of k_kp_hidden_tc failed: ") + cudaGetErrorString(e); return -3; }
    ++*launches;
    return 0;
}

// ---- default path of mode tc_3xf16 for LVC blocks 1, 2: piece-row protocol (fd_kernels_lvcp.cuh) ----
// up-sampling + first skip add -> piece rows of layer 0
static inline int tc_upsample_p(void* state, int blk, const float* in, const float* skip, float* p_out, unsigned int* sat, int B, int Tin,
                                cudaStream_t st, std::string& err, uint64_t* launches) {
    TcState* s = (TcState*)state;
    if (!s || !s->ok) { err = "tensor-core path not initialised"; return -4; }
    const float* wh = s->blob + s->sec_off[blk == 1 ? FD_S_LB1_UPT_HI : FD_S_LB2_UPT_HI];
    const float* wl = s->blob + s->sec_off[blk == 1 ? FD_S_LB1_UPT_LO : FD_S_LB2_UPT_LO];
    const float* bias = s->blob + s->sec_off[FD_S_LB0_UP_B + blk * FD_LB_STRIDE];
    UpPOut po;
    po.skip = skip; po.first_w = s->blob + s->sec_off[FD_S_FIRST_W]; po.first_b = s->blob + s->sec_off[FD_S_FIRST_B]; po.p_out = p_out; po.sat = sat;
    const int total = B * ((Tin + 127) / 128);
    if (blk == 1) {
        const int grid = total < s->sm_count ? total : s->sm_count;
        fd_launch_pdl(k_upsample_tc<8, true>, dim3(grid), dim3(512), ut_smem_bytes<8>() + UT_PEXTRA, st, wh, wl, bias, in, p_out, B, Tin, 1, po);
    } else {
        const int grid = total < 2 * s->sm_count ? total : 2 * s->sm_count;
        fd_launch_pdl(k_upsample_tc<4, true>, dim3(grid), dim3(512), ut_smem_bytes<4>() + UT_PEXTRA, st, wh, wl, bias, in, p_out, B, Tin, 1, po);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("launch of k_upsample_tc<POUT> failed: ") + cudaGetErrorString(e); return -3; }
    ++*launches;
    return 0;
}

// block 2 upsampling + skip on kind::f16 pieces (k_upsample_p4, fd_kernels_up4.cuh)
static inline int tc_upsample_p4(void* state, const float* in, const float* audio, float* p_out, unsigned int* sat, int B, int Tin,
                                 cudaStream_t st, std::string& err, uint64_t* launches) {
    TcState* s = (TcState*)state;
    if (!s || !s->ok) { err = "tensor-core path not initialised"; return -4; }
    Up4Params p;
    p.w16 = s->blob + s->sec_off[FD_S_LB2_UP_F16M];
    p.first16u = s->blob + s->sec_off[FD_S_FIRST_F16U];
    p.bias = s->blob + s->sec_off[FD_S_LB0_UP_B + 2 * FD_LB_STRIDE];
    p.in = in; p.audio = audio; p.p_out = p_out; p.sat = sat; p.B = B; p.Tin = Tin;
    p.inv = 1.f / (S16_ACT * s->scales16[41]);
    const int total = B * ((Tin + 127) / 128);
    const int grid = total < 2 * s->sm_count ? total : 2 * s->sm_count;
    fd_launch_pdl(k_upsample_p4, dim3(grid), dim3(512), U4_SMEM_BYTES, st, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("launch of k_upsample_p4 failed: ") + cudaGetErrorString(e); return -3; }
    ++*launches;
    return 0;
}

// one LVC layer of block 1 / 2: piece rows in -> piece rows out (p_out) or fp32 rows out (f_out, last layer of the block)
static inline int tc_lvc_p_layer(void* state, int blk, int layer, const float* p_in, const float* skip, const float* kern, float* p_out,
                                 float* f_out, unsigned int* sat, int B, int T, int Tm, int dil, cudaStream_t st, std::string& err, uint64_t* launches) {
    TcState* s = (TcState*)state;
    if (!s || !s->ok) { err = "tensor-core path not initialised"; return -4; }
    LvcPParams p;
    p.cw16 = s->blob + s->sec_off[blk == 1 ? FD_S_LB1_CONV_F16M : FD_S_LB2_CONV_F16M] + (size_t)layer * (LP_CW_BYTES / 4);
    p.conv_b = s->blob + s->sec_off[FD_S_LB0_CONV_B + blk * FD_LB_STRIDE] + layer * C;
    p.first16 = s->blob + s->sec_off[FD_S_FIRST_F16];
    p.p_in = p_in; p.skip = skip; p.kern = kern; p.p_out = p_out; p.f_out = f_out; p.sat = sat;
    p.B = B; p.T = T; p.Tm = Tm; p.dil = dil;
    p.inv_c = 1.f / (S16_ACT * s->scales16[4 + 4 * blk + layer]);
    p.inv_l = 1.f / (S16_ACT * S16_KERN);
    p.inv_sk = 1.f / (LP_S_AU * s->scales16[40]);
    const int tiles = B * ((T + LP_TT - 1) / LP_TT);
    const int grid = tiles < s->sm_count ? tiles : s->sm_count;
    if (blk == 1) fd_launch_pdl(k_lvc_p<64>, dim3(grid), dim3(LP_THREADS), lp_smem_bytes<64>(), st, p);
    else          fd_launch_pdl(k_lvc_p<256>, dim3(grid), dim3(LP_THREADS), lp_smem_bytes<256>(), st, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("launch of k_lvc_p failed: ") + cudaGetErrorString(e); return -3; }
    ++*launches;
    return 0;
}

static inline int tc_zero_pads(float* buf, int B, int T, cudaStream_t st, std::string& err, uint64_t* launches) {
    fd_launch_pdl(k_zero_pads, dim3(B + 1), dim3(256), 0, st, buf, B, T);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("launch of k_zero_pads failed: ") + cudaGetErrorString(e); return -3; }
    ++*launches;
    return 0;
}

static inline cudaError_t tc_set_lvc_attrs() {
    {
        cudaError_t ek = cudaFuncSetAttribute(k_kp_hidden_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, KT_SMEM_BYTES);
        if (ek != cudaSuccess) return ek;
    }
    cudaError_t e0 = cudaFuncSetAttribute(k_dblock0_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM_BYTES);
    if (e0 != cudaSuccess) return e0;
    e0 = cudaFuncSetAttribute(k_kc_gemm_tc2<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, KC2_SMEM_BYTES);
    if (e0 != cudaSuccess) return e0;
    e0 = cudaFuncSetAttribute(k_kc_gemm_tc2<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, KC2_SMEM_BYTES);
    if (e0 != cudaSuccess) return e0;
    e0 = cudaFuncSetAttribute(k_kc_gemm_tc2<true, 16, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, KC3_SMEM_BYTES);
    if (e0 != cudaSuccess) return e0;
    e0 = cudaFuncSetAttribute(k_upsample_tc<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ut_smem_bytes<4>());
    if (e0 != cudaSuccess) return e0;
    e0 = cudaFuncSetAttribute(k_upsample_tc<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, ut_smem_bytes<8>());
    if (e0 != cudaSuccess) return e0;
    e0 = cudaFuncSetAttribute(k_upsample_p4, cudaFuncAttributeMaxDynamicSharedMemorySize, U4_SMEM_BYTES);
    if (e0 != cudaSuccess) return e0;
    e0 = cudaFuncSetAttribute(k_upsample_tc<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ut_smem_bytes<4>() + UT_PEXTRA);
    if (e0 != cudaSuccess) return e0;
    e0 = cudaFuncSetAttribute(k_upsample_tc<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ut_smem_bytes<8>() + UT_PEXTRA);
    if (e0 != cudaSuccess) return e0;
    e0 = cudaFuncSetAttribute(k_lvc_p<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, lp_smem_bytes<64>());
    if (e0 != cudaSuccess) return e0;
    e0 = cudaFuncSetAttribute(k_lvc_p<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, lp_smem_bytes<256>());
    if (e0 != cudaSuccess) return e0;
    e0 = cudaFuncSetAttribute(k_lvc_layer_h<64, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, lh_smem_bytes<64, false, 2>());
    if (e0 != cudaSuccess) return e0;
    e0 = cudaFuncSetAttribute(k_lvc_layer_h<256, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, lh_smem_bytes<256, true, 2>());
    if (e0 != cudaSuccess) return e0;
    e0 = cudaFuncSetAttribute(k_lvc_layer_h<256, true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, lh_smem_bytes<256, true, 3>());
    if (e0 != cudaSuccess) return e0;
    cudaError_t e = cudaFuncSetAttribute(k_lvc_layer_tc<64, false, 1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, lt_smem_bytes<64, 1>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_lvc_layer_tc<256, true, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, lt_smem_bytes<256, 2>());
    return e;
}

#endif  // !FD_EMU

}  // namespace fd
