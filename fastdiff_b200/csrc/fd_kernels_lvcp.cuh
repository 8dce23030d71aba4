// K11 + K12 + K13 for LVC blocks 1 and 2, mode tc_3xf16, "P protocol": one warp-specialised, software-pipelined kernel per LVC layer
// (modules.py:208-217 -- x += skip; y = lrelu(conv_dil(lrelu(x))); y = LVC(y, kernels, bias); x = x + sigmoid(y[:32]) * tanh(y[32:])).
//
// What is different from k_lvc_layer_h (fd_kernels_tc.cuh), which it replaces on the default path:
//   * THE STATE TRAVELS AS OPERAND PIECES.  Between the layers of a block the activation z = x + skip is not stored as fp32 rows but as
//     the tcgen05 A operand of the next layer's dilated conv: per time step one 128-byte row [32 ch hi | 32 ch lo] of fp16 pieces of
//     16 * lrelu_0.2(z), 16-byte chunk c at position c ^ (t & 7) -- the SWIZZLE_128B image of a tile whose row 0 is a multiple of 8.
//     A layer therefore bulk-copies its input rows straight into the MMA tile: the whole "phase 1" of the old kernel (load raw rows,
//     add skip, lrelu, split, store: 39 % of its instructions, and its shared-memory bank conflicts) is gone, and the skip of block 2
//     (first_conv(audio), 7 taps) is evaluated once per produced row instead of once per loaded row (184 per 128).  The residual base
//     z of the gate epilogue is recovered from the same pieces: z = (hi + lo) / 16, times 5 where negative (lrelu is invertible);
//     hi + lo carries 22 significant bits of z, the 2 bits lost against fp32 cost ~2e-6 on eps (measured against fp64: at the fp32
//     noise floor of the reference itself, tests/test_f16_pieces.py).  The first layer's rows are written by the upsampling kernel
//     (k_upsample_tc<R, true>), the last layer of a block writes plain fp32 rows for its consumer.
//   * Rows outside [0, T) come from memory: the piece buffers carry 32 zero rows before the first item and 64 between items
//     (k_zero_pads), so the loads need no clamping and the conv sees the reference's zero padding.
//   * WARP SPECIALISATION.  26 warps: 8 conv-epilogue warps (TMEM -> lrelu -> pieces -> Y tile), 16 gate-epilogue warps (TMEM -> gate,
//     residual, skip, lrelu, pieces -> output), one loader (cp.async.bulk, 3 input stages + 2 kernel stages), one MMA issuer.  Conv
//     MMAs of tile n+1 are issued before the LVC MMAs of tile n; accumulators, Y tiles and output staging are double-buffered, so the
//     tensor pipe, both epilogues and the loads of three consecutive tiles overlap (the old kernel ran its five phases back to back
//     per 8-warp group: tensor pipe 16 % active, issue slots 42 %).
//   * Block 2 stages its output rows in shared memory and writes them with one cp.async.bulk per tile (whole 16 KB, no half-filled
//     sectors); block 1 (hop 64: two frames of predicted kernels per tile fill the shared memory) stores 16-byte chunks directly.
// Tile = 128 time steps, walked in DESCENDING order inside a CTA's contiguous chunk so that conv rows 128, 129 (needed by the last LVC
// taps) are Y rows 0, 1 of the tile processed just before (carried in 512 B of smem); chunk starts and utterance ends take them from a
// second MMA pass over A rows +128 -- the same instruction sequence, hence the same bits whatever the chunking.
// Compiles for the CPU fibre emulator too (tests/cudaemu).
#pragma once

namespace fd {

constexpr int LP_TT = 128;
constexpr int LP_HEAD_ROWS = 32;                 // zero rows before item 0
constexpr int LP_PAD_ROWS = 64;                  // zero rows after every item (32 tail + 32 head of the next item)
constexpr int LP_SLACK_ROWS = 192;               // readable (not necessarily zero) rows after the last pad: the last tile's window
constexpr int LP_STAGE_BYTES = 26624;            // 192 A rows (24576) | lbias 2 x 256 | audio window 136 floats (544) | pad -> multiple of 1024
template <int HOP> __host__ __device__ constexpr int lp_na() { return HOP == 256 ? 4 : 3; }   // input stages (hop 64: two frames of kernels per tile fill the smem)
constexpr int LP_Y_BYTES = 17408;                // 136 rows x 128 B (130 used)
constexpr int LP_CW_BYTES = 3 * C * 128;         // 12288
constexpr int LP_AU = 136;                       // audio positions t0-4 .. t0+131
constexpr int LP_THREADS = 864;                  // warps 0-15 gate epilogue, 16-23 conv epilogue, 24 loader, 25 MMA issuer, 26 output storer.  The hardware's warp
                                                 // arbiter favours HIGH warp ids: the roles on the critical path (MMA issue, loads, conv epilogue -- each
                                                 // tile's LVC MMAs wait for it) sit above the 16 gate warps, which otherwise starve them (ncu, round 2)

__host__ __device__ inline size_t lp_rows(int B, int T) { return (size_t)LP_HEAD_ROWS + (size_t)B * (T + LP_PAD_ROWS) + LP_SLACK_ROWS; }
__host__ __device__ inline size_t lp_row_of(int b, int T, int t) { return (size_t)LP_HEAD_ROWS + (size_t)b * (T + LP_PAD_ROWS) + t; }

template <int HOP> __host__ __device__ constexpr int lp_nf() { return HOP >= LP_TT ? 1 : LP_TT / HOP; }
constexpr int LP_SKA_BYTES = 16384;              // hop 256: audio im2col tile of the skip MMA, 128 rows x 128 B (64 B used)
constexpr int LP_SKB_BYTES = 4096;               // hop 256: first_audio_conv pieces (FIRST_F16), 32 rows x 128 B
constexpr float LP_S_AU = 16.f;                  // prescale of the audio pieces: the same range as the activations (|audio| < 4094 before saturation)
template <int HOP> __host__ __device__ constexpr int lp_smem_bytes() {
    return lp_na<HOP>() * LP_STAGE_BYTES + 2 * LP_Y_BYTES + 2 * lp_nf<HOP>() * 24576 + (HOP == 256 ? LP_SKA_BYTES + LP_SKB_BYTES : 0) +
           LP_CW_BYTES + C * 4 + 512 + 32 * 8 + 64 + 1024;
}

// Optional role timeline (-DLP_TIMELINE=1, GPU build only): CTA 0 of the block-2 launch with dilation LP_TL_DIL stamps clock64 at the
// protocol points of its first 24 tiles -- [role 0 loader | 1 MMA issuer | 2 conv epilogue (warp 16) | 3 gate epilogue (warp 0)][tile][8 slots];
// read with fd_debug_read("lp_timeline") (tests/gpu_scripts/lp_timeline.py).
#if defined(LP_TIMELINE) && !defined(FD_EMU)
#ifndef LP_TL_DIL
#define LP_TL_DIL 9
#endif
constexpr int LP_TL_TILES = 24;
__device__ unsigned long long g_lp_timeline[4 * LP_TL_TILES * 8];
#define LP_STAMP(role, tile, slot) do { if (tl_on && (tile) < LP_TL_TILES && (tile) >= 0) g_lp_timeline[((role) * LP_TL_TILES + (tile)) * 8 + (slot)] = clock64(); } while (0)
#else
#define LP_STAMP(role, tile, slot) do { } while (0)
#endif

struct LvcPParams {
    const float* cw16;       // [3 taps][32 co][128 B] SWIZZLE_128B image of this layer's dilated conv (LBn_CONV_F16)
    const float* conv_b;     // [32]
    const float* first16;    // hop 256: FIRST_F16 image (skip = first_conv(audio) as a K = 16 MMA: 7 taps + bias against a constant one)
    const float* p_in;       // padded piece rows of z = x + skip (input of this layer)
    const float* skip;       // hop 256: audio (B, T); hop 64: skip rows (B, T, 32) fp32
    const float* kern;       // this layer's slice of the predicted kernels: per (b, frame) at stride KCN floats
    float* p_out;            // padded piece rows of the next layer's input, or nullptr (last layer of a block)
    float* f_out;            // fp32 rows (B, T, 32) of the block output (last layer), or nullptr
    unsigned int* sat;       // sticky flag: an fp16 piece saturated (|16 * activation| > 65504); may be nullptr
    int B, T, Tm, dil;
    float inv_c, inv_l;
    float inv_sk;            // hop 256: 1 / (LP_S_AU * S(FIRST_F16))
};

// zero rows of a padded piece buffer: block 0 -> the 32 head rows, block i >= 1 -> the 64 rows after item i - 1
__global__ void __launch_bounds__(256) k_zero_pads(float* __restrict__ buf, int B, int T) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x;
    const size_t row0 = i == 0 ? 0 : (size_t)LP_HEAD_ROWS + (size_t)(i - 1) * (T + LP_PAD_ROWS) + T;
    const int nrows = i == 0 ? LP_HEAD_ROWS : LP_PAD_ROWS;
    float4* dst = reinterpret_cast<float4*>(buf + row0 * C);
    for (int k = threadIdx.x; k < nrows * 8; k += 256) dst[k] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// 8 values that already carry the x16 prescale -> one 16-byte chunk of hi pieces and one of lo pieces; m tracks max |v| (saturation guard)
__device__ __forceinline__ void lp_split8(const float (&v)[8], uint4& hi, uint4& lo, float& m) {
    uint2 h0, l0, h1, l1;
    split4_f16_pre(v[0], v[1], v[2], v[3], h0, l0);
    split4_f16_pre(v[4], v[5], v[6], v[7], h1, l1);
    hi = make_uint4(h0.x, h0.y, h1.x, h1.y);
    lo = make_uint4(l0.x, l0.y, l1.x, l1.y);
    m = fmaxf(m, fmaxf(fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3]))),
                       fmaxf(fmaxf(fabsf(v[4]), fabsf(v[5])), fmaxf(fabsf(v[6]), fabsf(v[7])))));
}

template <int HOP>
__global__ void __launch_bounds__(LP_THREADS, 1) k_lvc_p(const LvcPParams p) {
    pdl_trigger();

    constexpr int NF = lp_nf<HOP>();
    constexpr bool STAGE_OUT = (HOP == 256);
    constexpr int NA = lp_na<HOP>();
    constexpr bool SKIP_MMA = (HOP == 256);            // skip = first_conv(audio) on the tensor core; hop 64 loads skip rows from memory
    constexpr int W_BYTES = NF * 24576;
    // TMEM columns: conv accumulators 2 stages x 64 (hi-weight products | lo-weight products) at 0; second conv pass (rows +128, single-buffered,
    // rare) 64 at 128; skip MMA 2 x 32 at 192; LVC accumulators 2 stages x 128 at 256 (hop 256: hi-kernel | lo-kernel products of the 64
    // outputs; hop 64: the two frames of the tile, 64 outputs each)
    constexpr uint32_t C2ACC0 = 128, SKACC0 = 192, LACC0 = 256, LSTRIDE = 128;
    constexpr uint32_t TCOLS = 512;
    FD_DYN_SMEM(unsigned char, smem_raw);
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* a_st = smem;                                        // [3][LP_STAGE_BYTES]
    unsigned char* y_t = a_st + NA * LP_STAGE_BYTES;                   // [2][LP_Y_BYTES]
    unsigned char* w_t = y_t + 2 * LP_Y_BYTES;                         // [2][W_BYTES]
    unsigned char* sk_a = w_t + 2 * W_BYTES;                           // [128 rows][128 B] audio im2col pieces (SKIP_MMA), two column halves
    unsigned char* sk_b = sk_a + (SKIP_MMA ? LP_SKA_BYTES : 0);        // [32 rows][128 B] first conv pieces (SKIP_MMA)
    unsigned char* cw = sk_b + (SKIP_MMA ? LP_SKB_BYTES : 0);          // [3 taps][32 rows][128 B]
    float* cbs_s = (float*)(cw + LP_CW_BYTES);                         // [32] conv bias * S16_ACT
    unsigned char* carry = (unsigned char*)(cbs_s + C);                // [2][256 B]: Y rows 0, 1 of the previous tile
    uint64_t* bars = (uint64_t*)(carry + 512);
    uint64_t* a_full = bars;            // [NA] loader -> MMA, both epilogues (tx)
    uint64_t* a_free = bars + 4;        // [NA] gate epilogue (16 warps) -> loader
    uint64_t* w_full = bars + 8;        // [2] loader -> MMA (tx)
    uint64_t* w_free = bars + 10;       // [2] MMA commit -> loader
    uint64_t* cacc_full = bars + 12;    // [2] conv MMAs committed -> conv epilogue
    uint64_t* cacc_free = bars + 14;    // [2] conv epilogue (8 warps) -> MMA
    uint64_t* y_full = bars + 16;       // [2] conv epilogue (8 warps) -> MMA: Y tile, skip operand tile, prescaled LVC bias are in place
    uint64_t* lacc_full = bars + 18;    // [2] LVC (+ skip) MMAs committed -> gate epilogue; also "Y tile / skip tile free" for the conv epilogue
    uint64_t* lacc_free = bars + 20;    // [2] gate epilogue (16 warps) -> MMA
    uint64_t* out_full = bars + 24;     // [NA] (hop 256, layers 0-2) gate epilogue (16 warps) -> storer: the output rows are staged in the input stage
    uint32_t* tmem_base_s = (uint32_t*)(bars + 28);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        // staged output (hop 256, a next layer exists): the storer releases a stage once its bulk copy has read it; otherwise the 16 gate warps do
        const uint32_t free_cnt = (STAGE_OUT && p.p_out != nullptr) ? 1u : 16u;
        for (int i = 0; i < NA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_free[i], free_cnt); mbar_init(&out_full[i], 16); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&w_full[i], 1); mbar_init(&w_free[i], 1);
            mbar_init(&cacc_full[i], 1); mbar_init(&cacc_free[i], 8);
            mbar_init(&y_full[i], 8); mbar_init(&lacc_full[i], 1); mbar_init(&lacc_free[i], 16);
        }
        mbar_init_fence();
    }
    if (warp == 0) tmem_alloc(tmem_base_s, TCOLS);
    {
        const float4* src = reinterpret_cast<const float4*>(p.cw16);
        for (int i = tid; i < LP_CW_BYTES / 16; i += LP_THREADS) reinterpret_cast<float4*>(cw)[i] = src[i];
        if (SKIP_MMA) {
            const float4* s16 = reinterpret_cast<const float4*>(p.first16);
            for (int i = tid; i < LP_SKB_BYTES / 16; i += LP_THREADS) reinterpret_cast<float4*>(sk_b)[i] = s16[i];
            for (int i = tid; i < LP_SKA_BYTES / 16; i += LP_THREADS) reinterpret_cast<float4*>(sk_a)[i] = make_float4(0.f, 0.f, 0.f, 0.f);   // K 8..15 stay zero
        }
        if (tid < C) cbs_s[tid] = p.conv_b[tid] * S16_ACT;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;
    pdl_wait();   // programmatic dependent launch: everything above touched constants, shared memory and TMEM only
    const int B = p.B, T = p.T, Tm = p.Tm, dil = p.dil;
    const int ntt = (T + LP_TT - 1) / LP_TT, total = B * ntt;
    const int chunk = (total + (int)gridDim.x - 1) / (int)gridDim.x;
    const int tile_lo = (int)blockIdx.x * chunk, tile_hi = min(total, tile_lo + chunk);
    const int ntile = tile_hi > tile_lo ? tile_hi - tile_lo : 0;          // this CTA walks tiles tile_hi-1 .. tile_lo (descending)
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    // (b, tt) of the walk's first tile; every role then steps them down itself (no per-tile division)
    const int b_first = ntile ? (tile_hi - 1) / ntt : 0, tt_first = ntile ? (tile_hi - 1) % ntt : 0;
    const bool has_skip = (p.p_out != nullptr);   // the skip of the NEXT layer's "x += audio_down" is added to the rows this layer produces
#if defined(LP_TIMELINE) && !defined(FD_EMU)
    const bool tl_on = (HOP == 256) && blockIdx.x == 0 && p.dil == LP_TL_DIL;
#endif

    if (warp_u == 24) {
        // =========================================== loader ===========================================
        if (elect_one()) {
            int wc = 0, b = b_first, tt = tt_first, s = 0, sn = 0;     // s = n % NA, sn = n / NA
            bool new_frame = true;
            for (int n = 0; n < ntile; ++n) {
                const int t0 = tt * LP_TT;
                unsigned char* a = a_st + s * LP_STAGE_BYTES;
                float* lbias = (float*)(a + 24576);
                float* au = lbias + 128;
                LP_STAMP(0, n, 0);
                const int ar0 = 31 - dil, nrows = 130 + 2 * dil;
                if (n + 1 < ntile) {   // L2 prefetch of the NEXT tile's rows and kernels: its stage frees up about one tile time from now, and the
                                       // load latency sits on the critical cycle (stage free -> loads -> conv -> LVC -> gate -> stage free)
                    const int tt2 = tt > 0 ? tt - 1 : ntt - 1, b2 = tt > 0 ? b : b - 1, t02 = tt2 * LP_TT, f02 = t02 / HOP;
                    bulk_prefetch_l2(p.p_in + (lp_row_of(b2, T, t02 - 32 + ar0)) * C, (uint32_t)nrows * 128u);
                    if (HOP != 256 || tt == 0 || !(tt & 1)) {   // hop 256: only when the next tile starts a new frame
#pragma unroll
                        for (int fi = 0; fi < NF; ++fi)
                            if (f02 + fi < Tm) bulk_prefetch_l2(p.kern + ((size_t)b2 * Tm + f02 + fi) * KCN, (uint32_t)(KPL * 4));
                    }
                }
                mbar_wait(&a_free[s], (uint32_t)((sn & 1) ^ 1));
                LP_STAMP(0, n, 1);
                int i0 = 0, i1 = 0;
                if (SKIP_MMA && has_skip) { i0 = t0 == 0 ? 4 : 0; i1 = min(LP_AU, T - t0 + 4); }
                const int f0 = t0 / HOP;
                uint32_t bytes = (uint32_t)nrows * 128u + (uint32_t)(i1 - i0) * 4u;
#pragma unroll
                for (int fi = 0; fi < NF; ++fi) if (f0 + fi < Tm) bytes += 256u;
                mbar_expect_tx(&a_full[s], bytes);
                bulk_g2s(a + ar0 * 128, p.p_in + (lp_row_of(b, T, t0 - 32 + ar0)) * C, (uint32_t)nrows * 128u, &a_full[s]);
                if (i1 > i0) bulk_g2s(au + i0, p.skip + (size_t)b * T + (t0 - 4 + i0), (uint32_t)(i1 - i0) * 4u, &a_full[s]);
#pragma unroll
                for (int fi = 0; fi < NF; ++fi)
                    if (f0 + fi < Tm) bulk_g2s(lbias + fi * 64, p.kern + ((size_t)b * Tm + f0 + fi) * KCN + KK * LVC_OUT, 256u, &a_full[s]);
                // predicted kernels: hop 256 -> one frame serves two tiles (walked back to back); hop 64 -> two frames per tile
                if (HOP == 256) {
                    if (new_frame) {
                        const int ws = wc & 1;
                        mbar_wait(&w_free[ws], (uint32_t)(((wc >> 1) & 1) ^ 1));
                        mbar_expect_tx(&w_full[ws], 24576u);
                        bulk_g2s_once(w_t + ws * W_BYTES, p.kern + ((size_t)b * Tm + f0) * KCN, 24576u, &w_full[ws]);
                        ++wc;
                    }
                    new_frame = !(tt & 1);      // the next tile (b, tt - 1) shares this frame iff tt is odd
                } else {
                    const int ws = n & 1;
                    mbar_wait(&w_free[ws], (uint32_t)(((n >> 1) & 1) ^ 1));
                    uint32_t wb = 0;
#pragma unroll
                    for (int fi = 0; fi < NF; ++fi) if (f0 + fi < Tm) wb += 24576u;
                    mbar_expect_tx(&w_full[ws], wb);
                    // stage = [T01 hi rows f0 | f1 (8 KB each)][T01 lo rows f0 | f1][T2 f0 | f1]: every MMA covers both frames of the tile (N = 128)
#pragma unroll
                    for (int fi = 0; fi < NF; ++fi)
                        if (f0 + fi < Tm) {
                            const float* kf = p.kern + ((size_t)b * Tm + f0 + fi) * KCN;
#pragma unroll
                            for (int part = 0; part < 3; ++part)
                                bulk_g2s_once(w_t + ws * W_BYTES + part * 16384 + fi * 8192, kf + part * 2048, 8192u, &w_full[ws]);
                        }
                }
                if (--tt < 0) { tt = ntt - 1; --b; new_frame = true; }
                if (++s == NA) { s = 0; ++sn; }
            }
        }
        __syncwarp();
    } else if (warp_u == 25) {
        // =========================================== MMA issuer ===========================================
        constexpr uint32_t idesc_sk = umma_idesc_f16(128, 32), idesc_conv = umma_idesc_f16(128, 64), idesc_lvc = umma_idesc_f16(128, 128), idesc_l64 = umma_idesc_f16(128, 64);
        const uint32_t a_u = smem_u32(a_st), y_u = smem_u32(y_t), w_u = smem_u32(w_t), cw_u = smem_u32(cw);
        const uint32_t ska_u = smem_u32(sk_a), skb_u = smem_u32(sk_b);
        int wc = 0, cur_ws = 0, mb_tt = tt_first;                 // M2 runs one tile behind C: its own tile-in-item counter
        bool m_new_frame = true;
        auto lvc_mmas = [&](int m) {     // M2(m): LVC (+ skip) MMAs of the CTA's m-th tile
            const int ys = m & 1;
            LP_STAMP(1, m, 3);
            mbar_wait(&y_full[ys], (uint32_t)((m >> 1) & 1));
            LP_STAMP(1, m, 4);
            bool last_use = true;
            if (HOP == 256) {
                if (m_new_frame) {
                    cur_ws = wc & 1;
                    mbar_wait(&w_full[cur_ws], (uint32_t)((wc >> 1) & 1));
                    ++wc;
                }
                last_use = !((mb_tt & 1) && m + 1 < ntile);   // the next tile of the walk uses the same frame iff tt is odd (and it exists)
                m_new_frame = !(mb_tt & 1);
            } else {
                cur_ws = m & 1;
                mbar_wait(&w_full[cur_ws], (uint32_t)((m >> 1) & 1));
            }
            mbar_wait(&lacc_free[ys], (uint32_t)(((m >> 1) & 1) ^ 1));
            LP_STAMP(1, m, 5);
            tc_fence_after();
            uint32_t yt = y_u + (uint32_t)ys * LP_Y_BYTES, wt = w_u + (uint32_t)cur_ws * W_BYTES;
            FD_OPAQUE2(yt, wt);
            if (elect_one()) {
                // MERGED-N passes (kernel image: k_kc_gemm_tc2, exp_mask bit 64): taps 0 and 1 share 128-byte B rows [tap 0 | tap 1] with the hi and the
                // lo kernel pieces as separate ROWS, so one N = 128 MMA per activation piece forms both products (lo x lo rides along); tap 2 keeps
                // the three-pass form on its [hi | lo] rows.
                // DESCRIPTORS ARE ONE ADD EACH (umma_desc_lo / umma_desc_at): this single thread's instruction stream is what paces the tensor
                // pipe -- with descriptors rebuilt from addresses (shift, two masks, or: ~6 dependent uniform ops each) the issue loop of a tile
                // was ~360 instructions and took ~2,500 cycles whatever the number or shape of its MMAs (round-2 timelines).
                const uint32_t d = tmem_base + LACC0 + (uint32_t)ys * LSTRIDE;
                const uint32_t y_lo = umma_desc_lo(yt), w_lo = umma_desc_lo(wt);    // offsets below are in 16-byte units
#pragma unroll
                for (int k = 0; k < 3; ++k) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint64_t dah = umma_desc_at(y_lo + k * 8 + j * 2), dal = umma_desc_at(y_lo + k * 8 + 4 + j * 2);
                        const uint32_t acc = (k | j) ? 1u : 0u;
                        if (HOP == 256) {
                            if (k < 2) {   // rows 0..63 K_hi(o), 64..127 K_lo(o): D[:, o] += Y K_hi, D[:, 64 + o] += Y K_lo
                                const uint64_t db = umma_desc_at(w_lo + k * 4 + j * 2);
                                umma_f16(d, dah, db, idesc_lvc, acc);
                                umma_f16(d, dal, db, idesc_lvc, 1u);
                            } else {       // tap 2: D[:, o] += Y_hi K_hi + Y_hi K_lo + Y_lo K_hi
                                const uint64_t dbh = umma_desc_at(w_lo + 1024 + j * 2), dbl = umma_desc_at(w_lo + 1024 + 4 + j * 2);
                                umma_f16(d, dah, dbh, idesc_l64, 1u);
                                umma_f16(d, dah, dbl, idesc_l64, 1u);
                                umma_f16(d, dal, dbh, idesc_l64, 1u);
                            }
                        } else {           // hop 64: rows fi * 64 + o (both frames per MMA) -> D[:, fi * 64 + o], three passes
                            const uint64_t dbh = umma_desc_at(k < 2 ? w_lo + k * 4 + j * 2 : w_lo + 2048 + j * 2);
                            const uint64_t dbl = umma_desc_at(k < 2 ? w_lo + 1024 + k * 4 + j * 2 : w_lo + 2048 + 4 + j * 2);
                            umma_f16(d, dah, dbh, idesc_lvc, acc);
                            umma_f16(d, dah, dbl, idesc_lvc, 1u);
                            umma_f16(d, dal, dbh, idesc_lvc, 1u);
                        }
                    }
                }
                if (SKIP_MMA && has_skip) {   // skip[128 x 32] = audio im2col [128 x 16] x first conv [16 x 32], three piece passes
                    uint32_t sa = ska_u, sb = skb_u;
                    FD_OPAQUE2(sa, sb);
                    const uint32_t d2 = tmem_base + SKACC0 + (uint32_t)ys * 32;
                    const uint32_t sa_lo = umma_desc_lo(sa), sb_lo = umma_desc_lo(sb);
                    const uint64_t ah = umma_desc_at(sa_lo + ys * 4), al = umma_desc_at(sa_lo + ys * 4 + 2);   // column half ys of the operand tile
                    const uint64_t bh = umma_desc_at(sb_lo), bl = umma_desc_at(sb_lo + 2);
                    umma_f16(d2, ah, bh, idesc_sk, 0u);
                    umma_f16(d2, ah, bl, idesc_sk, 1u);
                    umma_f16(d2, al, bh, idesc_sk, 1u);
                }
                tc_commit(&lacc_full[ys]);
                if (last_use) tc_commit(&w_free[cur_ws]);
            }
            __syncwarp();
            LP_STAMP(1, m, 6);
            if (--mb_tt < 0) { mb_tt = ntt - 1; m_new_frame = true; }
        };
        int tt = tt_first, s = 0, sn = 0;
        for (int n = 0; n < ntile; ++n) {
            const int cs = n & 1;
            const bool have_carry = (n > 0) && (tt != ntt - 1);
            LP_STAMP(1, n, 0);
            mbar_wait(&a_full[s], (uint32_t)(sn & 1));
            mbar_wait(&cacc_free[cs], (uint32_t)(((n >> 1) & 1) ^ 1));
            LP_STAMP(1, n, 1);
            tc_fence_after();
            uint32_t at = a_u + (uint32_t)s * LP_STAGE_BYTES, cwt = cw_u;
            FD_OPAQUE2(at, cwt);
            // the second pass (rows +128; chunk starts and utterance ends only) has ONE accumulator: the conv epilogue must have read tile n-1's
            if (!have_carry && n >= 1) mbar_wait(&cacc_free[cs ^ 1], (uint32_t)(((n - 1) >> 1) & 1));
            if (elect_one()) {
                const uint32_t a_lo = umma_desc_lo(at), cw_lo = umma_desc_lo(cwt);   // offsets in 16-byte units: a 128-byte row = 8
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    // pass 1 (rows +128): its output rows 0, 1 are conv rows 128, 129; the other rows read past the loaded window and are never used
                    if (pass == 1 && have_carry) break;
                    const uint32_t d = tmem_base + (pass ? C2ACC0 : (uint32_t)cs * 64);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const uint32_t sh = (uint32_t)(pass * 128 + 31 + (k - 1) * dil) * 8u;
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint64_t dah = umma_desc_at(a_lo + sh + j * 2), dal = umma_desc_at(a_lo + sh + 4 + j * 2);
                            if (k < 2) {   // merged-N (image LBn_CONV_F16M): rows 0..31 W_hi(co), 32..63 W_lo(co) of [tap 0 | tap 1] -> D[:, co], D[:, 32 + co]
                                const uint64_t db = umma_desc_at(cw_lo + k * 4 + j * 2);
                                umma_f16(d, dah, db, idesc_conv, (k | j) ? 1u : 0u);
                                umma_f16(d, dal, db, idesc_conv, 1u);
                            } else {       // tap 2, three passes on its [hi | lo] rows -> D[:, co]
                                const uint64_t dbh = umma_desc_at(cw_lo + 512 + j * 2), dbl = umma_desc_at(cw_lo + 512 + 4 + j * 2);
                                umma_f16(d, dah, dbh, idesc_sk, 1u);
                                umma_f16(d, dah, dbl, idesc_sk, 1u);
                                umma_f16(d, dal, dbh, idesc_sk, 1u);
                            }
                        }
                    }
                }
                tc_commit(&cacc_full[cs]);
            }
            __syncwarp();
            LP_STAMP(1, n, 2);
            if (n >= 1) lvc_mmas(n - 1);
            if (--tt < 0) tt = ntt - 1;
            if (++s == NA) { s = 0; ++sn; }
        }
        if (ntile > 0) lvc_mmas(ntile - 1);
    } else if (warp_u == 26) {
        // =========================================== output storer ===========================================
        // One thread: waits until the 16 gate warps have staged a tile's 128 output rows (in place, rows 32..159 of the tile's input stage), writes
        // them with ONE bulk copy and releases the stage when the copy has read it.  (When a gate warp did this, its wait for the copy's read
        // held back all 16 warps at the next tile's barrier: ~1,200 of the gate loop's ~3,200 cycles per tile.)
        if (STAGE_OUT && p.p_out != nullptr && elect_one()) {
            int b = b_first, tt = tt_first, s = 0, sn = 0;
            for (int n = 0; n < ntile; ++n) {
                const int t0 = tt * LP_TT;
                mbar_wait(&out_full[s], (uint32_t)(sn & 1));
                const int rows = min(LP_TT, T - t0);
                bulk_s2g(p.p_out + lp_row_of(b, T, t0) * C, a_st + s * LP_STAGE_BYTES + 32 * 128, (uint32_t)rows * 128u);
                bulk_commit();
                bulk_wait_read0();
                mbar_arrive(&a_free[s]);
                if (--tt < 0) { tt = ntt - 1; --b; }
                if (++s == NA) { s = 0; ++sn; }
            }
            bulk_wait_all();
        }
        __syncwarp();
    } else if (warp_u >= 16) {
        // =========================================== conv epilogue (8 warps: 16..23) ===========================================
        const int q = warp & 3, part = (warp - 16) >> 2;  // TMEM lane quarter; which 16 of the 32 conv channels
        const int yr = q * 32 + lane;              // Y row of this thread <-> t = t0 - 1 + yr; (part 0) also skip-operand row <-> t = t0 + yr
        const float inv_cs = p.inv_c * S16_ACT;
        float vmax = 0.f;
        float cb[16];                              // this thread's 16 conv biases (x 16), loop-invariant
#pragma unroll
        for (int e = 0; e < 16; ++e) cb[e] = cbs_s[part * 16 + e];
        int tt = tt_first, s = 0, sn = 0;
        for (int n = 0; n < ntile; ++n) {
            const int t0 = tt * LP_TT;
            const int cs = n & 1;
            const bool have_carry = (n > 0) && (tt != ntt - 1);
            unsigned char* yt = y_t + cs * LP_Y_BYTES;
            unsigned char* a = a_st + s * LP_STAGE_BYTES;
            float* lbias = (float*)(a + 24576);
            float* au = lbias + 128;
            if (warp == 16 && lane == 0) LP_STAMP(2, n, 0);
            mbar_wait(&cacc_full[cs], (uint32_t)((n >> 1) & 1));
            if (warp == 16 && lane == 0) LP_STAMP(2, n, 1);
            tc_fence_after();
            uint32_t v[16], v2[16];   // products with the hi and with the lo pieces of the weights
            tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cs * 64 + part * 16, v);
            tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cs * 64 + 32 + part * 16, v2);
            tmem_ld_wait();
            if (n >= 2) mbar_wait(&lacc_full[cs], (uint32_t)(((n >> 1) - 1) & 1));   // the LVC MMAs of tile n-2 have read this Y tile
            if (warp == 16 && lane == 0) LP_STAMP(2, n, 2);
            // 16 accumulator columns of row `row` -> 16 lrelu(acc * inv + b) -> pieces: logical chunks 2 part, 2 part + 1 (hi) and 4 + ... (lo)
            auto emit_row = [&](const uint32_t (&acc)[16], const uint32_t (&acc2)[16], int row, unsigned char* copy_to) {
                const int t = t0 - 1 + row;
                const bool in = (t >= 0 && t < T);
                const int sw = row & 7;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    float y[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float x = fmaf(__uint_as_float(acc[cc * 8 + e]) + __uint_as_float(acc2[cc * 8 + e]), inv_cs, cb[cc * 8 + e]);
                        y[e] = fmaxf(x, 0.2f * x);
                    }
                    uint4 hi, lo;
                    lp_split8(y, hi, lo, vmax);
                    if (!in) hi = lo = make_uint4(0u, 0u, 0u, 0u);   // rows outside [0, T): the reference's zero padding of the LVC input (utterance ends only)
                    const int c = part * 2 + cc;
                    *reinterpret_cast<uint4*>(yt + row * 128 + ((c ^ sw) << 4)) = hi;
                    *reinterpret_cast<uint4*>(yt + row * 128 + (((4 + c) ^ sw) << 4)) = lo;
                    if (copy_to) {   // rows 0, 1 again for the next tile (a 2-row image in the same layout: 128 & 7 == 0, 129 & 7 == 1)
                        *reinterpret_cast<uint4*>(copy_to + row * 128 + ((c ^ sw) << 4)) = hi;
                        *reinterpret_cast<uint4*>(copy_to + row * 128 + (((4 + c) ^ sw) << 4)) = lo;
                    }
                }
            };
            emit_row(v, v2, yr, (q == 0 && lane < 2) ? carry + cs * 256 : nullptr);
            if (q == 0) {
                if (have_carry) {   // rows 128, 129 = rows 0, 1 of the previous tile: every warp copies back the four chunks per row it wrote itself
                    if (lane < 8) {
                        const int row = lane >> 2, c = part * 2 + (lane & 1) + ((lane & 2) ? 4 : 0), pc = c ^ row;   // row & 7 == row for rows 0, 1
                        reinterpret_cast<uint4*>(yt + 128 * 128)[row * 8 + pc] = reinterpret_cast<const uint4*>(carry + (cs ^ 1) * 256)[row * 8 + pc];
                    }
                } else {            // ... or rows 0, 1 of the second MMA pass (TMEM lanes 0, 1 of its accumulator)
                    tmem_ld_32x32b_x16(tmem_base + C2ACC0 + part * 16, v);
                    tmem_ld_32x32b_x16(tmem_base + C2ACC0 + 32 + part * 16, v2);
                    tmem_ld_wait();
                    if (lane < 2) emit_row(v, v2, 128 + lane, nullptr);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&cacc_free[cs]);   // both accumulators of this stage have been read
            if (warp == 16 && lane == 0) LP_STAMP(2, n, 3);
            // ---- work taken off the gate epilogue's hands: LVC bias prescaled by the gate's exponent constants (part 1); skip operand tile (part 0) ----
            mbar_wait(&a_full[s], (uint32_t)(sn & 1));
            if (warp == 16 && lane == 0) LP_STAMP(2, n, 4);
            if (part == 1) {
                if (yr < NF * 64) lbias[yr] *= ((yr & 63) < 32) ? -1.4426950408889634f : -2.8853900817779268f;   // sigmoid half: e^-a; tanh half: e^-2b
            } else if (SKIP_MMA && has_skip) {
                if (t0 == 0 || t0 + LP_AU - 4 > T) {   // audio positions outside [0, T) are zero (the first conv zero-pads); only utterance ends
                    for (int i = yr; i < LP_AU; i += 128) { const int pos = t0 - 4 + i; if (pos < 0 || pos >= T) au[i] = 0.f; }
                    group_sync(2, 128);
                }
                // the operand tile is double-buffered in its two 64-byte column halves (tile n -> half n & 1); the skip MMAs of tile n-2, which read
                // this half, completed with that tile's LVC MMAs (lacc_full, waited for above before the Y tile was rewritten)
                float x[8];
#pragma unroll
                for (int k = 0; k < 7; ++k) x[k] = au[yr + 1 + k] * LP_S_AU;    // audio position t + k - 3
                x[7] = LP_S_AU;                                                  // constant one: the bias rides as an eighth tap
                uint4 hi, lo;
                lp_split8(x, hi, lo, vmax);
                const int sw = yr & 7;
                *reinterpret_cast<uint4*>(sk_a + yr * 128 + (((cs * 4 + 0) ^ sw) << 4)) = hi;
                *reinterpret_cast<uint4*>(sk_a + yr * 128 + (((cs * 4 + 2) ^ sw) << 4)) = lo;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&y_full[cs]);
            if (warp == 16 && lane == 0) LP_STAMP(2, n, 5);
            if (--tt < 0) tt = ntt - 1;
            if (++s == NA) { s = 0; ++sn; }
        }
        if (p.sat && vmax > F16_MAX) *p.sat = 1u;
    } else {
        // =========================================== gate epilogue (16 warps: 0..15) ===========================================
        const int q = warp & 3, j = warp >> 2;                  // TMEM lane quarter, channel octet (gate channels 8j .. 8j+7)
        const int r = q * 32 + lane;                            // output row of this thread
        const int fi = (HOP >= LP_TT) ? 0 : r / HOP;            // warp-uniform (HOP is a multiple of 32)
        // everything below runs in the x16 domain of the pieces (exact: powers of two): z16 = 16 z, gate x 16, skip x 16
        const float c_s = p.inv_l * -1.4426950408889634f, c_t = p.inv_l * -2.8853900817779268f, c_sk = p.inv_sk * S16_ACT;
        float vmax = 0.f;
        int b = b_first, tt = tt_first, s = 0, sn = 0;
        for (int n = 0; n < ntile; ++n) {
            const int t0 = tt * LP_TT;
            const int ls = n & 1, t = t0 + r;
            const unsigned char* a = a_st + s * LP_STAGE_BYTES;
            const float* lbias = (const float*)(a + 24576);
            if (warp == 0 && lane == 0) LP_STAMP(3, n, 0);
            mbar_wait(&a_full[s], (uint32_t)(sn & 1));
            if (warp == 0 && lane == 0) LP_STAMP(3, n, 1);
            float sk[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) sk[c] = 0.f;
            if (!SKIP_MMA && has_skip && t < T) {   // hop 64: skip rows of the DBlock output
                const float4* sp = reinterpret_cast<const float4*>(p.skip + ((size_t)b * T + t) * C + j * 8);
                const float4 s0 = sp[0], s1 = sp[1];
                sk[0] = s0.x; sk[1] = s0.y; sk[2] = s0.z; sk[3] = s0.w; sk[4] = s1.x; sk[5] = s1.y; sk[6] = s1.z; sk[7] = s1.w;
            }
            // residual base z16 = 16 (x + skip) of this row, recovered from the pieces of 16 * lrelu(z) that fed the conv (lrelu^-1: x5 where negative)
            float z[8];
            {
                const int ar = 32 + r, sw = ar & 7;
                const uint4 h = *reinterpret_cast<const uint4*>(a + ar * 128 + ((j ^ sw) << 4));
                const uint4 l = *reinterpret_cast<const uint4*>(a + ar * 128 + (((4 + j) ^ sw) << 4));
                const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float2 hf = unpack_f16x2(hh[c]), lf = unpack_f16x2(ll[c]);
                    const float v0 = hf.x + lf.x, v1 = hf.y + lf.y;
                    z[2 * c] = fminf(v0, 5.f * v0);
                    z[2 * c + 1] = fminf(v1, 5.f * v1);
                }
            }
            if (warp == 0 && lane == 0) LP_STAMP(3, n, 2);
            mbar_wait(&lacc_full[ls], (uint32_t)((n >> 1) & 1));
            if (warp == 0 && lane == 0) LP_STAMP(3, n, 3);
            tc_fence_after();
            uint32_t zs[8], zt[8], za[8];
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + LACC0 + (uint32_t)ls * LSTRIDE + fi * 64 + j * 8;
            tmem_ld_32x32b_x8(ta, zs);
            tmem_ld_32x32b_x8(ta + 32, zt);
            if (HOP == 256) {   // merged-N: columns 64.. hold the products with the lo pieces of the kernels
                uint32_t zs2[8], zt2[8];
                tmem_ld_32x32b_x8(ta + 64, zs2);
                tmem_ld_32x32b_x8(ta + 96, zt2);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    zs[c] = __float_as_uint(__uint_as_float(zs[c]) + __uint_as_float(zs2[c]));
                    zt[c] = __float_as_uint(__uint_as_float(zt[c]) + __uint_as_float(zt2[c]));
                }
            }
            if (SKIP_MMA && has_skip) tmem_ld_32x32b_x8(tmem_base + ((uint32_t)(q * 32) << 16) + SKACC0 + (uint32_t)ls * 32 + j * 8, za);
            float lb[16];   // prescaled by the conv epilogue: sigmoid half x -log2(e), tanh half x -2 log2(e)
            {
                const float4* lp = reinterpret_cast<const float4*>(lbias + fi * 64 + j * 8);
                const float4 a0 = lp[0], a1 = lp[1], b0 = lp[8], b1 = lp[9];
                lb[0] = a0.x; lb[1] = a0.y; lb[2] = a0.z; lb[3] = a0.w; lb[4] = a1.x; lb[5] = a1.y; lb[6] = a1.z; lb[7] = a1.w;
                lb[8] = b0.x; lb[9] = b0.y; lb[10] = b0.z; lb[11] = b0.w; lb[12] = b1.x; lb[13] = b1.y; lb[14] = b1.z; lb[15] = b1.w;
            }
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&lacc_free[ls]);
            if (warp == 0 && lane == 0) LP_STAMP(3, n, 4);
            // sigmoid(a) tanh(b) = (1 - E) / ((1 + A)(1 + E)), A = e^-a, E = e^-2b (b clamped at -15: E finite, tanh(-15) = -1 in fp32)
            float xn[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float A = ex2_approx(fmaf(__uint_as_float(zs[c]), c_s, lb[c]));
                const float E = ex2_approx(fminf(fmaf(__uint_as_float(zt[c]), c_t, lb[8 + c]), 43.280851226668903f));
                const float e1 = 1.f + E;
                const float rinv = rcp_approx(fmaf(A, e1, e1));              // 1 / ((1 + A)(1 + E))
                xn[c] = fmaf(fmaf(E, -S16_ACT, S16_ACT), rinv, z[c]);        // 16 (x + gate)
            }
            if (warp == 0 && lane == 0) LP_STAMP(3, n, 5);
            if (p.f_out) {
                if (t < T) st_global_f8(p.f_out + ((size_t)b * T + t) * C + j * 8,
                                        make_float4(xn[0] * (1.f / S16_ACT), xn[1] * (1.f / S16_ACT), xn[2] * (1.f / S16_ACT), xn[3] * (1.f / S16_ACT)),
                                        make_float4(xn[4] * (1.f / S16_ACT), xn[5] * (1.f / S16_ACT), xn[6] * (1.f / S16_ACT), xn[7] * (1.f / S16_ACT)));
            } else {
                float v[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {   // (x + gate) + skip, one rounding like the reference's add; then 16 lrelu
                    const float zz = SKIP_MMA ? fmaf(__uint_as_float(za[c]), c_sk, xn[c]) : fmaf(sk[c], S16_ACT, xn[c]);
                    v[c] = fmaxf(zz, 0.2f * zz);
                }
                uint4 hi, lo;
                lp_split8(v, hi, lo, vmax);
                const int sw = r & 7;                                                   // == t & 7 (t0 is a multiple of 128)
                if (STAGE_OUT) {
                    // staged IN PLACE: the output pieces of row r, chunks j and 4 + j, go over the input pieces of the same row and chunks in the
                    // stage -- cells only this thread reads (z, above) -- and one bulk copy writes the 128 rows; the stage is released once
                    // the copy has read it (the shared memory of a separate staging tile buys the fourth input stage instead)
                    unsigned char* ot = const_cast<unsigned char*>(a) + 32 * 128;
                    *reinterpret_cast<uint4*>(ot + r * 128 + ((j ^ sw) << 4)) = hi;
                    *reinterpret_cast<uint4*>(ot + r * 128 + (((4 + j) ^ sw) << 4)) = lo;
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&out_full[s]);   // -> storer (which also releases the stage)
                } else if (t < T) {
                    uint4* dst = reinterpret_cast<uint4*>(p.p_out + lp_row_of(b, T, t) * C);
                    dst[j ^ sw] = hi;
                    dst[(4 + j) ^ sw] = lo;
                }
            }
            __syncwarp();
            if (!(STAGE_OUT && p.p_out != nullptr) && lane == 0) mbar_arrive(&a_free[s]);
            if (warp == 0 && lane == 0) LP_STAMP(3, n, 6);
            if (--tt < 0) { tt = ntt - 1; --b; }
            if (++s == NA) { s = 0; ++sn; }
        }
        if (p.sat && vmax > F16_MAX) *p.sat = 1u;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TCOLS);
    }
}

}  // namespace fd
