// K10 for LVC block 2 on the default path of mode tc_3xf16: lrelu -> ConvTranspose1d(32, 32, k = 8, stride 4, pad 2) (modules.py:205-206)
// + the block's first "x += audio_down" (skip = first_audio_conv(audio), FastDiff_model.py:85-88) -> the PIECE ROWS the first LVC layer
// consumes (fd_kernels_lvcp.cuh).  Replaces k_upsample_tc<4, true> there.  What is different (round-2 phase timeline, DESIGN.md 4c.5: that
// kernel spent 9,000 of a tile's 17,000 cycles in an epilogue whose instructions were half first_conv(audio) on FFMA, and 4,000 issuing
// 64 kind::tf32 MMAs):
//   * everything on kind::f16 pieces: the input rows are split once into [32 ch hi | 32 ch lo] rows of 16 * lrelu(x); per output phase
//     ph (t = 4 m + ph) the two taps share 128-byte weight rows [tap kk1 | tap kk1 + 4] with the hi and the lo weight pieces as separate
//     ROWS (merged-N, as in k_lvc_p): 8 MMAs of N = 64 per phase instead of 16 tf32 ones;
//   * THE SKIP IS FOUR MORE SMALL MMAs: for input row m all four phases read the audio window a_m = audio[4 m - 3 .. 4 m + 6], so ONE
//     im2col tile (128 rows x K = 16: the 10 samples, a constant one for the bias, zeros) serves the four phases against per-phase
//     Toeplitz weight tiles W_ph[i][co] = first_w[i - ph][co] (section FIRST_F16U), accumulated INTO the phase accumulators -- the
//     packer gives both weight tensors the same power-of-two scale (SCALES16[41]) so the sums share one scale;
//   * the epilogue is TMEM -> (hi-weight + lo-weight columns) * inv + bias -> lrelu -> pieces -> two 16-byte stores per phase.
// One tile = 128 input rows (+1 halo row either side) -> 512 output rows; 16 warps; 2 CTAs per SM (88 KB of shared memory, 256 TMEM columns).
// Rounding: up + skip are summed in the accumulator (the reference adds the skip to the rounded conv output): ~1e-7 relative, inside the
// stated tolerance; results do not depend on tiling or batch composition.  Compiles for the CPU fibre emulator too.
#pragma once

namespace fd {

constexpr int U4_AROWS = 136;                       // input rows m0-1 .. m0+128 (+ pad)
constexpr int U4_ATILE = U4_AROWS * 128;            // 17408
constexpr int U4_WBYTES = 4 * 8192;                 // per phase 64 rows (piece, co) x 128 B
constexpr int U4_SKBYTES = 128 * 128;               // skip tile: chunks 0-3 audio im2col (rows m), chunks 4-7 Toeplitz weights (rows ph * 32 + co)
constexpr int U4_AUW = 4 * 128 + 8;                 // audio window, positions 4 m0 - 4 .. 4 m0 + 515
constexpr int U4_SMEM_BYTES = 2 * U4_ATILE + U4_WBYTES + U4_SKBYTES + 2 * U4_AUW * 4 + 128 + 64 + 64 + 1024;

struct Up4Params {
    const float* w16;        // section LB2_UP_F16M: [4 ph][64 rows][128 B] merged-N weight tiles
    const float* first16u;   // section FIRST_F16U: [128 rows = ph * 32 + co][64 B]: K = 16 hi (32 B) | lo (32 B)
    const float* bias;       // [32] upsample bias
    const float* in;         // (B, Tin, 32) fp32 rows (output of LVC block 1)
    const float* audio;      // (B, 4 Tin)
    float* p_out;            // padded piece rows, B items of 4 Tin rows
    unsigned int* sat;
    int B, Tin;
    float inv;               // 1 / (S16_ACT * SCALES16[41])
};

#ifndef U4_STORE32
#define U4_STORE32 1   // 1: 256-bit stores of whole 32-byte sectors in the epilogue (thread = 16 channels x 2 phases); 0: 16-byte stores (8 channels x 4 phases)
#endif
// 32 bytes (two 16-byte chunks) to a 32-byte aligned global address
__device__ __forceinline__ void st_global_u8(float* p, const uint4 a, const uint4 b) {
#if FD_VEC256 && !defined(FD_EMU)
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
#else
    *reinterpret_cast<uint4*>(p) = a;
    *reinterpret_cast<uint4*>(p + 4) = b;
#endif
}
__global__ void __launch_bounds__(512, 2) k_upsample_p4(const Up4Params p) {
    pdl_trigger();

    FD_DYN_SMEM(unsigned char, smem_raw);
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* raw = smem;                           // bulk-copied fp32 rows: row ar <-> input row m0 - 1 + ar
    unsigned char* a_t = raw + U4_ATILE;                 // piece rows of 16 * lrelu(in)
    unsigned char* w_t = a_t + U4_ATILE;                 // [4 ph][64 rows][128 B]
    unsigned char* sk_t = w_t + U4_WBYTES;               // [128 rows][128 B]
    float* au_s = (float*)(sk_t + U4_SKBYTES);           // [2][U4_AUW]
    float* b_s = au_s + 2 * U4_AUW;                      // [32] bias
    uint64_t* bar = (uint64_t*)(b_s + 32);               // [0] MMAs, [1] loads
    uint32_t* tmem_base_s = (uint32_t*)(bar + 2);
    const int B = p.B, Tin = p.Tin, Tout = 4 * Tin;

    const int tid = threadIdx.x, gw = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init_fence(); }
    if (tid < 32) tmem_alloc(tmem_base_s, 256u);
    for (int i = tid; i < U4_WBYTES / 16; i += 512) reinterpret_cast<float4*>(w_t)[i] = reinterpret_cast<const float4*>(p.w16)[i];
    for (int i = tid; i < 128 * 8; i += 512) {           // skip tile: weights into logical chunks 4..7 of their rows, the rest zero
        const int row = i >> 3, c = i & 7;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c >= 4) v = reinterpret_cast<const float4*>(p.first16u)[row * 4 + (c - 4)];
        *reinterpret_cast<float4*>(sk_t + row * 128 + ((c ^ (row & 7)) << 4)) = v;
    }
    if (tid < C) b_s[tid] = p.bias[tid] * S16_ACT;
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;
    pdl_wait();   // programmatic dependent launch: everything above touched constants, shared memory and TMEM only
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const int gw_u = __shfl_sync(0xffffffffu, gw, 0);
    constexpr uint32_t idesc64 = umma_idesc_f16(128, 64), idesc32 = umma_idesc_f16(128, 32);
    float vmax = 0.f;

    const int ntt = (Tin + 127) / 128, total = B * ntt;
    auto issue_loads = [&](int tile, uint32_t buf) {
        const int b = tile / ntt, m0 = (tile % ntt) * 128;
        const int ar0 = m0 == 0 ? 1 : 0, ar1 = min(130, Tin - m0 + 1);   // rows inside [0, Tin)
        const uint32_t bytes = (uint32_t)(ar1 - ar0) * 128u;
        const int i0 = m0 == 0 ? 4 : 0, i1 = min(U4_AUW, 4 * (Tin - m0) + 4);   // audio positions inside [0, Tout): whole 16-byte groups
        const uint32_t abytes = (uint32_t)(i1 - i0) * 4u;
        mbar_expect_tx(&bar[1], bytes + abytes);
        bulk_g2s(raw + ar0 * 128, p.in + ((size_t)b * Tin + (m0 - 1 + ar0)) * C, bytes, &bar[1]);
        bulk_g2s(au_s + buf * U4_AUW + i0, p.audio + (size_t)b * Tout + (4 * m0 - 4 + i0), abytes, &bar[1]);
    };
    int tile = blockIdx.x;
    if (tile < total && gw_u == 0) { if (elect_one()) issue_loads(tile, 0u); __syncwarp(); }
    uint32_t parity = 0;
    for (; tile < total; tile += gridDim.x, parity ^= 1) {
        const int b = tile / ntt, m0 = (tile % ntt) * 128;
        float* au_t = au_s + parity * U4_AUW;
        if (m0 == 0 || 4 * m0 + U4_AUW - 4 > Tout)      // positions outside the utterance are zero (the first conv zero-pads); the copy never touches them
            for (int i = tid; i < U4_AUW; i += 512) {
                const int pos = 4 * m0 - 4 + i;
                if (pos < 0 || pos >= Tout) au_t[i] = 0.f;
            }
        mbar_wait(&bar[1], parity);
        if (m0 == 0 || 4 * m0 + U4_AUW - 4 > Tout) __syncthreads();   // the zero fill above is read by other threads below
        // ---- phase 1a: lrelu + x16 + fp16 split of the input rows into the A tile (8 lanes per row, 4 channels each) ----
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int ar = (tid >> 3) + i * 64, c4 = tid & 7, m = m0 - 1 + ar;
            if (ar < 130) {
                uint2 hi = make_uint2(0u, 0u), lo = make_uint2(0u, 0u);
                if (m >= 0 && m < Tin) {
                    const float4 x = *reinterpret_cast<const float4*>(raw + ar * 128 + c4 * 16);
                    const float y0 = lrelu02_s(x.x), y1 = lrelu02_s(x.y), y2 = lrelu02_s(x.z), y3 = lrelu02_s(x.w);
                    split4_f16_pre(y0, y1, y2, y3, hi, lo);
                    vmax = fmaxf(vmax, fmaxf(fmaxf(fabsf(y0), fabsf(y1)), fmaxf(fabsf(y2), fabsf(y3))));
                }
                const int sw = ar & 7;
                *reinterpret_cast<uint2*>(a_t + ar * 128 + (((c4 >> 1) ^ sw) << 4) + (c4 & 1) * 8) = hi;
                *reinterpret_cast<uint2*>(a_t + ar * 128 + (((4 + (c4 >> 1)) ^ sw) << 4) + (c4 & 1) * 8) = lo;
            }
        }
        // ---- phase 1b: audio im2col row of input row m (thread = row, quarter): a[i] = 16 audio[4 m - 3 + i] (i < 10), a[10] = 16 (bias tap), 0 ----
        {
            const int mr = tid >> 2, qd = tid & 3;               // qd 0: hi K 0..7; 1: hi K 8..15; 2: lo K 0..7; 3: lo K 8..15
            const int k0 = (qd & 1) * 8;
            float a[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int i = k0 + e;
                a[e] = i < 10 ? au_t[4 * mr + 1 + i] * LP_S_AU : (i == 10 ? LP_S_AU : 0.f);   // position 4 (m0 + mr) - 3 + i <-> window index 4 mr + 1 + i
            }
            uint4 hi, lo;
            lp_split8(a, hi, lo, vmax);
            *reinterpret_cast<uint4*>(sk_t + mr * 128 + ((qd ^ (mr & 7)) << 4)) = (qd & 2) ? lo : hi;
        }
        fence_async_smem();
        __syncthreads();
        // ---- phase 2: MMAs (one thread) ----
        if (gw_u == 0) {
            tc_fence_after();
            uint32_t at = smem_u32(a_t), wt = smem_u32(w_t);
            FD_OPAQUE2(at, wt);
            uint32_t st = smem_u32(sk_t);
            FD_OPAQUE(st);
            if (elect_one()) {
                const uint32_t a_lo = umma_desc_lo(at), w_lo = umma_desc_lo(wt), s_lo = umma_desc_lo(st);   // offsets in 16-byte units
#pragma unroll
                for (int ph = 0; ph < 4; ++ph) {
                    const int sh = (ph + 2) / 4;
                    const uint32_t d = tmem_u + ph * 64;
#pragma unroll
                    for (int tap = 0; tap < 2; ++tap) {       // tap 0: input m + sh (weights kk1); tap 1: input m + sh - 1 (weights kk1 + 4)
                        const uint32_t arow = (uint32_t)(1 + sh - tap) * 8u;
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint64_t dah = umma_desc_at(a_lo + arow + j * 2), dal = umma_desc_at(a_lo + arow + 4 + j * 2);
                            const uint64_t db = umma_desc_at(w_lo + ph * 512 + tap * 4 + j * 2);   // rows 0..31 W_hi(co), 32..63 W_lo(co)
                            umma_f16(d, dah, db, idesc64, (tap | j) ? 1u : 0u);
                            umma_f16(d, dal, db, idesc64, 1u);
                        }
                    }
                    // skip: a_m (K = 16) x W_ph, three piece passes into the hi-weight columns of the phase
                    const uint64_t sah = umma_desc_at(s_lo), sal = umma_desc_at(s_lo + 2);
                    const uint64_t sbh = umma_desc_at(s_lo + ph * 256 + 4), sbl = umma_desc_at(s_lo + ph * 256 + 6);
                    umma_f16(d, sah, sbh, idesc32, 1u);
                    umma_f16(d, sah, sbl, idesc32, 1u);
                    umma_f16(d, sal, sbh, idesc32, 1u);
                }
                tc_commit(&bar[0]);
            }
            __syncwarp();
        }
        mbar_wait(&bar[0], parity);
        tc_fence_after();
        if (gw_u == 0 && tile + (int)gridDim.x < total) { if (elect_one()) issue_loads(tile + gridDim.x, parity ^ 1u); __syncwarp(); }   // raw rows are free
#if U4_STORE32
        // ---- phase 3: epilogue: thread = (input row m, 16 channels, 2 of the 4 phases).  The two 8-channel chunks of a piece sit in one aligned
        //      32-byte pair of the swizzled row (chunk c at c ^ (t & 7): the pair base takes bits 1-2 of the XOR, bit 0 swaps the halves), so a
        //      thread writes whole 32-byte sectors with one 256-bit store per piece instead of half sectors with 16-byte stores ----
        {
            const int q = gw & 3, pair = (gw >> 2) & 1, phh = gw >> 3, m = m0 + q * 32 + lane;
            const float inv16 = p.inv * S16_ACT;
#pragma unroll
            for (int pi = 0; pi < 2; ++pi) {
                const int ph = phh * 2 + pi;
                uint4 hi[2], lo[2];
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t v[8], v2[8];
                    const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + ph * 64 + pair * 16 + hf * 8;
                    tmem_ld_32x32b_x8(ta, v);
                    tmem_ld_32x32b_x8(ta + 32, v2);
                    tmem_ld_wait();
                    float z[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float zz = fmaf(__uint_as_float(v[i]) + __uint_as_float(v2[i]), inv16, b_s[pair * 16 + hf * 8 + i]);   // 16 (up + skip + bias)
                        z[i] = fmaxf(zz, 0.2f * zz);
                    }
                    lp_split8(z, hi[hf], lo[hf], vmax);
                }
                if (m < Tin) {
                    const int t = 4 * m + ph, sw = t & 7, odd = sw & 1;
                    float* dst = p.p_out + lp_row_of(b, Tout, t) * C;            // 32 fp32-sized words = 128 bytes of pieces
                    const uint4 h_a = odd ? hi[1] : hi[0], h_b = odd ? hi[0] : hi[1];
                    const uint4 l_a = odd ? lo[1] : lo[0], l_b = odd ? lo[0] : lo[1];
                    st_global_u8(dst + (((2 * pair) ^ (sw & 6)) << 2), h_a, h_b);
                    st_global_u8(dst + (((4 + 2 * pair) ^ (sw & 6)) << 2), l_a, l_b);
                }
            }
        }
#else
        // ---- phase 3: epilogue: thread = (input row m, 8 channels); its 4 outputs are consecutive rows 4 m + ph ----
        {
            const int q = gw & 3, part = gw >> 2, m = m0 + q * 32 + lane;
            float bb[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) bb[i] = b_s[part * 8 + i];
            const float inv16 = p.inv * S16_ACT;
#pragma unroll
            for (int ph = 0; ph < 4; ++ph) {
                uint32_t v[8], v2[8];
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + ph * 64 + part * 8;
                tmem_ld_32x32b_x8(ta, v);
                tmem_ld_32x32b_x8(ta + 32, v2);
                tmem_ld_wait();
                if (m < Tin) {
                    const int t = 4 * m + ph;
                    float z[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float zz = fmaf(__uint_as_float(v[i]) + __uint_as_float(v2[i]), inv16, bb[i]);   // 16 (up + skip + bias)
                        z[i] = fmaxf(zz, 0.2f * zz);
                    }
                    uint4 hi, lo;
                    lp_split8(z, hi, lo, vmax);
                    uint4* dst = reinterpret_cast<uint4*>(p.p_out + lp_row_of(b, Tout, t) * C);
                    const int sw = t & 7;
                    dst[part ^ sw] = hi;
                    dst[(4 + part) ^ sw] = lo;
                }
            }
        }
#endif
        tc_fence_before();
        __syncthreads();
    }
    if (p.sat && vmax > F16_MAX) *p.sat = 1u;
    tc_fence_before();
    __syncthreads();
    if (tid < 32) {
        tc_fence_after();
        tmem_dealloc(*tmem_base_s, 256u);
    }
}

}  // namespace fd
