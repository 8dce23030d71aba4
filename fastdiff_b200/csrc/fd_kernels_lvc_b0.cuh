// LVC block 0 (hop 8) on tensor cores, mode tc_3xf16 (the default since round 2: 1.23 -> 0.82 ms per N = 4 call at config 2 on B200):
// k_lvc_layer_b0h, block 0 in swapped-operand form, fed by the kernel_conv GEMM writing block 0's kernels as fp16 pieces
// (k_kc_gemm_tc2<true, 16, true>).  Option "tc_b0" = 0 selects the SIMT kernel k_lvc_layer<8>.
// Included by fd_kernels_tc.cuh after k_lvc_layer_h, whose constants and helpers (LH_*, LvcHParams, split4_f16_pre, gate_st, ...) it uses.
#pragma once

namespace fd {

// Optional phase timeline (-DB0_TIMELINE=1, GPU build): thread 0 of CTA 0 stamps clock64 at the phase boundaries of its tiles (8 slots x 4 tiles; the
// last launch wins); fd_debug_read("b0_timeline").
#if defined(B0_TIMELINE) && !defined(FD_EMU)
__device__ unsigned long long g_b0_timeline[4 * 8];
#define B0_STAMP(slot) do { if (blockIdx.x == 0 && tid == 0 && tl_n < 4) g_b0_timeline[tl_n * 8 + (slot)] = clock64(); } while (0)
#else
#define B0_STAMP(slot) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------------------
// One LVC layer of block 0
// (hop 8) on tensor cores in SWAPPED-operand form.  With 8 samples per frame an M = 128 time-step tile would use 8 rows per
// predicted kernel; instead the kernels are the M side:
//   A = the predicted kernels of TWO consecutive frames stacked (rows 0-63: frame f's 64 output channels, rows 64-127: frame f+1's;
//       per tap the [32 i hi | 32 i lo] tiles of the fp16-piece image, two 8 KB bulk copies per tap),
//   B = the Y rows of the 16 time steps of the two frames shifted by the tap (N = 16: a plain row window of the Y tile),
//   D = (frame-in-pair, o) lanes x 16 step columns; the two diagonal 64 x 8 blocks are the result.
// 18 MMAs per frame pair; the 8 pairs of a 128-step tile stream through a 3-slot ring (48 KB per pair) fed by cp.async.bulk, so
// the kernel is bound by the HBM stream of the predicted kernels (24.8 KB per frame).  The dilated conv, the in-place A build, the
// Y epilogue and the skip placement are those of k_lvc_layer_h (block-1 flavour: skip rows from global); the two extra conv rows
// always come from the second MMA pass.  Gate epilogue: lanes o < 32 hold the sigmoid arguments, lanes o >= 32 the tanh arguments
// of the same (frame, step): the tanh warps publish tanh(z) through shared memory, the sigmoid warps (lane = channel) combine with
// the residual base and store coalesced rows.  One 16-warp group per CTA, persistent.
// ---------------------------------------------------------------------------------------------------------
constexpr int LB0_PAIR_BYTES = 3 * 2 * 8192;          // 48 KB: per tap [frame f: 64 rows x 128 B | frame f+1: 64 rows x 128 B]
constexpr int LB0_NSLOT = 3;
constexpr int LB0_SMEM_BYTES = 2 * LH_A_BYTES + 16384 + LH_CW_BYTES + LB0_NSLOT * LB0_PAIR_BYTES + 2 * C * 4 + 16 * 8 + 16 + 1024;

__global__ void __launch_bounds__(512, 1)
k_lvc_layer_b0h(LvcHParams p, const float* __restrict__ x_in, const float* __restrict__ skip, const float* __restrict__ kern,
                float* __restrict__ x_out, int B, int T, int Tm, int dil, float inv_c, float inv_l, int skip_in_rt, int skip_out_rt) {
    pdl_trigger();

    const bool skip_in = skip_in_rt != 0, skip_out = skip_out_rt != 0;
    FD_DYN_SMEM(unsigned char, smem_raw);
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* a_t = smem;                                  // A tile (raw x rows on arrival) | later: Y tile
    unsigned char* s_t = a_t + LH_A_BYTES;                      // raw skip rows (skip_in)
    float* ex_s = (float*)(s_t + LH_A_BYTES);                   // [128 steps][32 c] tanh values
    unsigned char* cw = (unsigned char*)ex_s + 16384;           // [3 taps][32 rows][128 B]
    unsigned char* ring = cw + LH_CW_BYTES;                     // [3 slots][3 taps][128 rows][128 B]
    float* cb_s = (float*)(ring + LB0_NSLOT * LB0_PAIR_BYTES);  // [32] conv bias
    float* cbs_s = cb_s + C;                                    // [32] conv bias * S16_ACT
    uint64_t* bars = (uint64_t*)(cbs_s + C);                    // [0] conv MMAs, [1] LVC MMAs, [2] row loads, [3..5] ring full, [6..8] ring empty
    uint32_t* tmem_base_s = (uint32_t*)(bars + 16);

    const int tid = threadIdx.x, lane = tid & 31;
    const int gw = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (tid == 0) {
        for (int i = 0; i < 9; ++i) mbar_init(&bars[i], 1);
        mbar_init_fence();
    }
    if (gw == 0) tmem_alloc(tmem_base_s, 256u);
    {
        const float4* src = reinterpret_cast<const float4*>(p.cw16);
        for (int i = tid; i < LH_CW_BYTES / 16; i += 512) reinterpret_cast<float4*>(cw)[i] = src[i];
        if (tid < C) { cb_s[tid] = p.conv_b[tid]; cbs_s[tid] = p.conv_b[tid] * S16_ACT; }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // TMEM columns: conv [0,32), second conv pass [32,64), LVC pair p at [64 + 16 p, +16)
    const uint32_t tmem_base = *tmem_base_s;
    pdl_wait();   // programmatic dependent launch: everything above touched constants, shared memory and TMEM only
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t smem_u = smem_u32(smem);
    constexpr uint32_t idesc_conv = umma_idesc_f16(128, 32), idesc_lvc = umma_idesc_f16(128, 16);
    const int c4 = tid & 7;
    // row within each 64-row block of the A transform: a warp takes rows r, r+4, r+8, r+12 (2 instead of 4 wavefronts per piece store)
    const int prow = ((gw >> 2) << 4) + (gw & 3) + ((lane >> 3) << 2);

    const int ntt = (T + LT_TT - 1) / LT_TT, total = B * ntt;
    const int r_lo = 27 - dil, r_hi = 157 + dil;
    uint64_t* bar_rows = &bars[2];
    uint64_t* ring_full = &bars[3];
    uint64_t* ring_empty = &bars[6];

    auto issue_rows = [&](int tile) {
        const int b = tile / ntt, t0 = (tile % ntt) * LT_TT;
        const int ar0 = max(r_lo, 28 - t0), ar1 = min(r_hi, T - t0 + 28);
        uint32_t bytes = 0;
        if (ar1 > ar0) bytes = (uint32_t)(ar1 - ar0) * 128u * (skip_in ? 2u : 1u);
        mbar_expect_tx(bar_rows, bytes);
        if (ar1 > ar0) {
            const size_t off = ((size_t)b * T + (t0 - 28 + ar0)) * C;
            bulk_g2s(a_t + ar0 * 128, x_in + off, (uint32_t)(ar1 - ar0) * 128u, bar_rows);
            if (skip_in) bulk_g2s(s_t + ar0 * 128, skip + off, (uint32_t)(ar1 - ar0) * 128u, bar_rows);
        }
    };
    // pair `pp` (0..7) of tile `tile` -> ring slot `slot`: frames f = t0/8 + 2 pp and f + 1 (missing frames past the end load nothing)
    auto issue_pair = [&](int tile, int pp, uint32_t slot) {
        const int b = tile / ntt, t0 = (tile % ntt) * LT_TT;
        const int f0 = t0 / 8 + 2 * pp;
        uint32_t bytes = 0;
        for (int h = 0; h < 2; ++h) if (f0 + h < Tm) bytes += 3 * 8192;
        mbar_expect_tx(&ring_full[slot], bytes);
        for (int h = 0; h < 2; ++h) {
            if (f0 + h >= Tm) continue;
            const float* src = kern + ((size_t)b * Tm + f0 + h) * KCN;
            for (int k = 0; k < 3; ++k)
                bulk_g2s_once(ring + slot * LB0_PAIR_BYTES + k * 16384 + h * 8192, src + k * 2048, 8192, &ring_full[slot]);
        }
    };

    // ring counters in pair loads (8 per tile), identical in every lane of warp 0.  The first LB0_NSLOT pairs of a tile are requested
    // together with its rows (one tile ahead: every slot is free once the previous tile's MMAs have completed), the rest in phase 4.
    uint32_t ld_issued = 0, ld_used = 0;
    auto issue_tile_head = [&](int tile, uint32_t li) {   // ONE thread
        issue_rows(tile);
        for (int pp = 0; pp < LB0_NSLOT; ++pp, ++li) {
            const uint32_t rs = li % LB0_NSLOT;
            if (li >= LB0_NSLOT) mbar_wait(&ring_empty[rs], ((li / LB0_NSLOT) - 1) & 1);
            issue_pair(tile, pp, rs);
        }
    };
    int tile = blockIdx.x;
    if (tile < total && gw == 0) { if (elect_one()) issue_tile_head(tile, 0); __syncwarp(); ld_issued = LB0_NSLOT; }
    uint32_t parity = 0;
    [[maybe_unused]] int tl_n = 0;
    for (; tile < total; tile += gridDim.x, parity ^= 1, ++tl_n) {
        B0_STAMP(0);
        const int b = tile / ntt, t0 = (tile % ntt) * LT_TT;
        // ---------------- phase 1: raw rows -> fp16 pieces, in place ----------------
        mbar_wait(bar_rows, parity);
        B0_STAMP(1);
        {
            float4 xv[3], sv[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int ar = r_lo + prow + i * 64, t = t0 - 28 + ar;
                xv[i] = make_float4(0.f, 0.f, 0.f, 0.f); sv[i] = xv[i];
                if (ar < r_hi && t >= 0 && t < T) {
                    xv[i] = *reinterpret_cast<const float4*>(a_t + ar * 128 + c4 * 16);
                    if (skip_in) sv[i] = *reinterpret_cast<const float4*>(s_t + ar * 128 + c4 * 16);
                }
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int ar = r_lo + prow + i * 64;
                const float4 pre = make_float4(xv[i].x + sv[i].x, xv[i].y + sv[i].y, xv[i].z + sv[i].z, xv[i].w + sv[i].w);
                uint2 hi, lo;
                split4_f16_pre(lrelu02_s(pre.x), lrelu02_s(pre.y), lrelu02_s(pre.z), lrelu02_s(pre.w), hi, lo);
                if (ar < r_hi) {
                    const int sw = ar & 7;
                    *reinterpret_cast<uint2*>(a_t + ar * 128 + (((c4 >> 1) ^ sw) << 4) + (c4 & 1) * 8) = hi;
                    *reinterpret_cast<uint2*>(a_t + ar * 128 + (((4 + (c4 >> 1)) ^ sw) << 4) + (c4 & 1) * 8) = lo;
                }
            }
        }
        fence_async_smem();
        __syncthreads();
        B0_STAMP(2);
        // ---------------- phase 2: dilated conv (two passes: rows 0..127 and the two extra rows) ----------------
        if (gw == 0) {
            tc_fence_after();
            uint32_t at = smem_u, cwt = smem_u32(cw);
            FD_OPAQUE2(at, cwt);
            if (elect_one()) {
                const uint32_t a_lo = umma_desc_lo(at), cw_lo = umma_desc_lo(cwt);   // one add per descriptor (offsets in 16-byte units)
#pragma unroll
                for (int pass = 0; pass < 2; ++pass)
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const uint32_t sh = (uint32_t)(pass * 128 + 27 + (k - 1) * dil) * 8u;
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint64_t dah = umma_desc_at(a_lo + sh + j * 2), dal = umma_desc_at(a_lo + sh + 4 + j * 2);
                            const uint64_t dbh = umma_desc_at(cw_lo + k * 256 + j * 2), dbl = umma_desc_at(cw_lo + k * 256 + 4 + j * 2);
                            umma_f16(tmem_u + pass * 32, dah, dbh, idesc_conv, (k | j) ? 1u : 0u);
                            umma_f16(tmem_u + pass * 32, dah, dbl, idesc_conv, 1u);
                            umma_f16(tmem_u + pass * 32, dal, dbh, idesc_conv, 1u);
                        }
                    }
                tc_commit(&bars[0]);
            }
            __syncwarp();
        }
        mbar_wait(&bars[0], parity);
        B0_STAMP(3);
        tc_fence_after();
        // ---------------- phase 3: y = lrelu(conv + b) -> pieces, rows of the Y tile (over the A tile) ----------------
        if (gw < 8) {
            const int q3 = gw & 3, part3 = gw >> 2;
            const float inv_cs = inv_c * S16_ACT;
            auto emit_row = [&](const uint32_t (&v)[16], int yr) {
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
                }
            };
            uint32_t v[16];
            tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q3 * 32) << 16) + part3 * 16, v);
            tmem_ld_wait();
            emit_row(v, q3 * 32 + lane);
            if (q3 == 0) {
                tmem_ld_32x32b_x16(tmem_base + 32 + part3 * 16, v);
                tmem_ld_wait();
                if (lane < 2) emit_row(v, 128 + lane);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        B0_STAMP(4);
        // ---------------- phase 4: location-variable conv, one accumulation group per frame pair, kernels through the ring ----------------
        if (gw == 0) {
            tc_fence_after();
            const uint32_t tile_base_ld = ld_used;   // 8 x the number of tiles this CTA has finished
            if (elect_one()) {
                uint32_t li = ld_issued, lu = ld_used;
                for (int pp = 0; pp < 8; ++pp) {
                    while (li < lu + LB0_NSLOT && li < tile_base_ld + 8) {   // keep the ring full (within this tile)
                        const uint32_t rs = li % LB0_NSLOT;
                        if (li >= LB0_NSLOT) mbar_wait(&ring_empty[rs], ((li / LB0_NSLOT) - 1) & 1);
                        issue_pair(tile, (int)(li - tile_base_ld), rs);
                        ++li;
                    }
                    const uint32_t rs = lu % LB0_NSLOT;
                    mbar_wait(&ring_full[rs], (lu / LB0_NSLOT) & 1);
                    tc_fence_after();
                    const uint32_t w_lo = umma_desc_lo(smem_u32(ring) + rs * LB0_PAIR_BYTES), y_lo = umma_desc_lo(smem_u);
                    const uint32_t d = tmem_u + 64 + pp * 16;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const uint32_t ysh = (uint32_t)(16 * pp + k) * 8u;
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint64_t dwh = umma_desc_at(w_lo + k * 1024 + j * 2), dwl = umma_desc_at(w_lo + k * 1024 + 4 + j * 2);
                            const uint64_t dyh = umma_desc_at(y_lo + ysh + j * 2), dyl = umma_desc_at(y_lo + ysh + 4 + j * 2);
                            umma_f16(d, dwh, dyh, idesc_lvc, (k | j) ? 1u : 0u);
                            umma_f16(d, dwh, dyl, idesc_lvc, 1u);
                            umma_f16(d, dwl, dyh, idesc_lvc, 1u);
                        }
                    }
                    tc_commit(&ring_empty[rs]);
                    ++lu;
                }
                tc_commit(&bars[1]);
            }
            __syncwarp();
            ld_used = tile_base_ld + 8;      // what the elected lane did, applied in every lane
            ld_issued = tile_base_ld + 8;
        }
        // ---------------- phase 5: gate + residual -> global ----------------
        {
            const int q = gw & 3, cg = gw >> 2;             // lane quarter / pairs 2 cg, 2 cg + 1
            const int h = q >> 1;                           // frame within the pair
            const bool tanh_part = (q & 1) != 0;            // lanes o >= 32
            float xs[16], so[16];
            if (!tanh_part) {                               // residual base of this lane's channel at its 16 steps
#pragma unroll
                for (int pi = 0; pi < 2; ++pi)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int r = 16 * (2 * cg + pi) + 8 * h + j, t = t0 + r;
                        float xv = 0.f, sk = 0.f;
                        if (t < T) {
                            const size_t e = ((size_t)b * T + t) * C + lane;
                            xv = x_in[e];
                            if (skip_in || skip_out) sk = skip[e];
                        }
                        xs[pi * 8 + j] = skip_in ? xv + sk : xv;
                        so[pi * 8 + j] = sk;
                    }
            }
            mbar_wait(&bars[1], parity);
            B0_STAMP(5);
            tc_fence_after();
            if (gw == 0 && tile + (int)gridDim.x < total) {   // the A/Y tile and every ring slot are free: request the next tile
                if (elect_one()) issue_tile_head(tile + gridDim.x, ld_issued);
                __syncwarp();
                ld_issued += LB0_NSLOT;
            }
            uint32_t z[2][8];   // this frame's 8 step columns of the two pairs (the diagonal blocks of D)
            tmem_ld_32x32b_x8(tmem_base + ((uint32_t)(q * 32) << 16) + 64 + cg * 32 + 8 * h, z[0]);
            tmem_ld_32x32b_x8(tmem_base + ((uint32_t)(q * 32) << 16) + 64 + cg * 32 + 16 + 8 * h, z[1]);
            tmem_ld_wait();
            // the LVC bias of this lane's (frame, o) comes straight from the kernel record (one coalesced 128-byte read per warp)
            float zz[16];
#pragma unroll
            for (int pi = 0; pi < 2; ++pi) {
                const int pp = 2 * cg + pi, f = t0 / 8 + 2 * pp + h;
                const float lbv = f < Tm ? kern[((size_t)b * Tm + f) * KCN + KK * LVC_OUT + (q & 1) * 32 + lane] : 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) zz[pi * 8 + j] = fmaf(__uint_as_float(z[pi][j]), inv_l, lbv);
            }
            if (tanh_part) {
#pragma unroll
                for (int pi = 0; pi < 2; ++pi)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int r = 16 * (2 * cg + pi) + 8 * h + j;
                        const float bc = fmaxf(zz[pi * 8 + j], -15.f);
                        const float E = ex2_approx(-2.8853900817779268f * bc);
                        ex_s[r * 32 + lane] = (1.f - E) * rcp_approx(1.f + E);
                    }
            }
            __syncthreads();
            if (!tanh_part) {
#pragma unroll
                for (int pi = 0; pi < 2; ++pi)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int r = 16 * (2 * cg + pi) + 8 * h + j, t = t0 + r;
                        if (t < T) {
                            const float sg = rcp_approx(1.f + ex2_approx(-1.4426950408889634f * zz[pi * 8 + j]));
                            float o = xs[pi * 8 + j] + sg * ex_s[r * 32 + lane];
                            if (skip_out) o = __fadd_rn(o, so[pi * 8 + j]);
                            x_out[((size_t)b * T + t) * C + lane] = o;
                        }
                    }
            }
        }
        tc_fence_before();
        __syncthreads();
        B0_STAMP(6);
    }
    tc_fence_before();
    __syncthreads();
    if (gw == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256u);
    }
}

}  // namespace fd
