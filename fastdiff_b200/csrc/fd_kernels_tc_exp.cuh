// Experimental tensor-core kernels of the location-variable-convolution layers: OFF in the default path, selected with fd_set_option
// (DESIGN.md section 9): k_lvc_layer_p ("lvc_pipe": block 2 with software-pipelined tiles) and k_lvc_layer_b0h + k_b0_panel_to_pieces
// ("tc_b0": block 0 in swapped-operand form).  Validated on the CPU model of tests/cudaemu (all asynchronous schedules), not yet timed.
// Included by fd_kernels_tc.cuh after k_lvc_layer_h, whose constants and helpers (LH_*, LvcHParams, split4_f16_pre, gate_st, ...) it uses.
#pragma once

namespace fd {

// ---------------------------------------------------------------------------------------------------------
// EXPERIMENTAL (option "lvc_pipe", default off; written against the round-1 ncu capture, validated on the CPU model only): LVC block 2
// (hop 256, skip = first_conv(audio) recomputed on the way in) with the phases of consecutive tiles of a group SOFTWARE-PIPELINED, so
// that a group's own SIMT work runs while its MMAs are in flight instead of the group sleeping on the mbarrier (18 % of the stall
// samples of k_lvc_layer_h<256, true, 2> are those two waits, 8 % the end-of-tile barrier):
//     iteration i:   P1(i) rows -> pieces   | issue conv MMAs(i) | P5(i-1) gate epilogue   | P3(i) conv -> Y pieces | issue LVC MMAs(i)
//                    [LVC MMAs(i-1) running]                      [conv MMAs(i) running]                              (run during P1(i+1))
// Everything a tile owns is double-buffered per group -- A/Y tile, xs rows, TMEM column set, lbias -- and the loads are issued at
// the point where their target is known to be free: the x rows / audio of tile i+1 and the kernels / biases of tile i right after
// LVC MMAs(i-1) have completed (the wait P5(i-1) starts with).  Two group barriers per tile (after P1, after P3) instead of four.
// Arithmetic, tile walk (descending chunks, carried halo rows, second-pass fallback) and results are those of k_lvc_layer_h.
// Hazards (the model cannot see them; argued here):  XS[s], TM[s], lbias[s] are written in iteration i and last read in iteration
// i+1 (P5(i)); their next writers run in iteration i+2, behind the two barriers of iteration i+1.  AY[s^1] receives the rows of tile
// i+1 only after LVC MMAs(i-1), its last reader, have completed.  The single audio buffer is read in P1(i) (before the first barrier
// of iteration i) and refilled after it.  LW is refilled after LVC MMAs(i-1) and awaited (bar 3) before LVC MMAs(i) are issued.
// ---------------------------------------------------------------------------------------------------------
constexpr int LP_SLOT = 2 * LH_A_BYTES + 2 * LH_XS_BYTES + LH_LW_BYTES;        // 106,496 B per group
constexpr int LP_SMALL = 2 * 256 + LT_AU * 4;                                  // lbias x 2 | audio
constexpr int LP_SMEM_BYTES = 2 * (LP_SLOT + LP_SMALL) + LH_CW_BYTES + (7 * C + C + C + C) * 4 + 2 * 512 + 2 * 8 * 8 + 16 + 1024;

// ROWS = true: the flavour for layers 1..3 under option "b2_skipbuf" -- the input rows already carry the skip (added by the previous
// layer's epilogue), so P1 is load / lrelu / split only, and the epilogue adds the (B,T,32) skip rows read from global memory for the
// next layer when skip_out is set.  ROWS = false: skip = the audio, first_conv recomputed on the way in (the default arithmetic).
template <bool ROWS>
__global__ void __launch_bounds__(512, 1)
k_lvc_layer_p(LvcHParams p, const float* __restrict__ x_in, const float* __restrict__ skip, const float* __restrict__ kern,
              float* __restrict__ x_out, int B, int T, int Tm, int dil, float inv_c, float inv_l, int skip_out_rt) {
    constexpr int HOP = 256, GROUPS = 2, GT = 256;
    const bool skip_out = ROWS && skip_out_rt != 0;
    FD_DYN_SMEM(unsigned char, smem_raw);
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* cw = smem + GROUPS * LP_SLOT;                 // [3 taps][32 rows][128 B]
    unsigned char* small0 = cw + LH_CW_BYTES;                    // [GROUPS][lbias 0 | lbias 1 | audio]
    float* fw_s = (float*)(small0 + GROUPS * LP_SMALL);          // [7][32]
    float* fb_s = fw_s + 7 * C;
    float* cb_s = fb_s + C;
    float* cbs_s = cb_s + C;                                     // conv bias * S16_ACT
    unsigned char* carry_s = (unsigned char*)(cbs_s + C);        // [GROUPS][2][256 B]
    uint64_t* bars = (uint64_t*)(carry_s + GROUPS * 512);        // [GROUPS][8]: 0 conv MMAs, 1 LVC MMAs, 2 rows + audio, 3 kernels + biases
    uint32_t* tmem_base_s = (uint32_t*)(bars + 8 * GROUPS);

    const int tid = threadIdx.x, g = tid / GT, gt = tid % GT, gw = gt >> 5, lane = tid & 31;
    unsigned char* slot = smem + g * LP_SLOT;
    unsigned char* lw = slot + 2 * LH_A_BYTES + 2 * LH_XS_BYTES;
    unsigned char* small = small0 + g * LP_SMALL;
    float* au_s = (float*)(small + 512);
    uint64_t* bar = bars + 8 * g;
    unsigned char* carry = carry_s + g * 512;

    if (tid == 0) {
        for (int i = 0; i < 8 * GROUPS; ++i) mbar_init(&bars[i], 1);
        mbar_init_fence();
    }
    if (tid < 32) tmem_alloc(tmem_base_s, 512u);
    {
        const float4* src = reinterpret_cast<const float4*>(p.cw16);
        for (int i = tid; i < LH_CW_BYTES / 16; i += GT * GROUPS) reinterpret_cast<float4*>(cw)[i] = src[i];
        if (tid < 7 * C) fw_s[tid] = ROWS ? 0.f : p.first_w[tid];
        if (tid < C) { fb_s[tid] = ROWS ? 0.f : p.first_b[tid]; cb_s[tid] = p.conv_b[tid]; cbs_s[tid] = p.conv_b[tid] * S16_ACT; }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // TMEM: group g at g * 256, column set s at + s * 128: conv [0,32), LVC [32,96), second conv pass [96,128)
    const uint32_t tmem_g = *tmem_base_s + g * 256;
    constexpr uint32_t idesc_conv = umma_idesc_f16(128, 32), idesc_lvc = umma_idesc_f16(128, 64);
    const int c4 = gt & 7;
    const int prow = ((gw >> 2) << 4) + (gw & 3) + ((lane >> 3) << 2);   // rows r, r+4, r+8, r+12 per warp (conflict-free piece stores)
    const int gw_u = __shfl_sync(0xffffffffu, gw, 0), g_u = __shfl_sync(0xffffffffu, g, 0);
    const uint32_t slot_u = smem_u32(smem) + (uint32_t)(g_u * LP_SLOT);
    const uint32_t cw_u = smem_u32(cw);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_g, 0);

    const int ntt = (T + LT_TT - 1) / LT_TT, total = B * ntt;
    const int r_lo = 27 - dil, r_hi = 157 + dil;
    const int ngroups = gridDim.x * GROUPS, chunk = (total + ngroups - 1) / ngroups;
    const int tile_lo = (blockIdx.x * GROUPS + g) * chunk, tile_hi = min(total, tile_lo + chunk) - 1;

    // ---- loads (ONE thread) ----
    auto issue_rows = [&](int tile, int s) {          // x rows -> AY[s], audio window -> au_s   (bar 2)
        const int b = tile / ntt, t0 = (tile % ntt) * LT_TT;
        const int ar0 = max(r_lo, 28 - t0), ar1 = min(r_hi, T - t0 + 28);
        const int i0 = max(0, 32 - t0), i1 = min(LT_AU, T - t0 + 32);
        uint32_t bytes = 0;
        if (ar1 > ar0) bytes += (uint32_t)(ar1 - ar0) * 128u;
        if (!ROWS && i1 > i0) bytes += (uint32_t)(i1 - i0) * 4u;
        mbar_expect_tx(&bar[2], bytes);
        if (ar1 > ar0) bulk_g2s(slot + s * LH_A_BYTES + ar0 * 128, x_in + ((size_t)b * T + (t0 - 28 + ar0)) * C, (uint32_t)(ar1 - ar0) * 128u, &bar[2]);
        if (!ROWS && i1 > i0) bulk_g2s(au_s + i0, skip + (size_t)b * T + (t0 - 32 + i0), (uint32_t)(i1 - i0) * 4u, &bar[2]);
    };
    auto issue_lw = [&](int tile, int s, bool keep_lw) {   // predicted kernels -> LW (unless the frame is already there), biases -> lbias[s]   (bar 3)
        const int b = tile / ntt, t0 = (tile % ntt) * LT_TT, f = t0 / HOP;
        uint32_t bytes = 0;
        if (f < Tm) bytes = (keep_lw ? 0u : (uint32_t)LH_LW_BYTES) + 256u;
        mbar_expect_tx(&bar[3], bytes);
        if (f < Tm) {
            const float* src = kern + ((size_t)b * Tm + f) * KCN;
            if (!keep_lw) bulk_g2s(lw, src, LH_LW_BYTES, &bar[3]);
            bulk_g2s(small + s * 256, src + KK * LVC_OUT, 256, &bar[3]);
        }
    };

    // ---- P1: raw x rows (+ first_conv(audio)) -> fp16 pieces in place, xs rows -> XS[s] ----
    auto phase1 = [&](int t0, int s) {
        unsigned char* a_t = slot + s * LH_A_BYTES;
        unsigned char* xs_t = slot + 2 * LH_A_BYTES + s * LH_XS_BYTES;
        float fwr[7][4], fbr[4];
        if (!ROWS) {
            if (gt < LT_AU) { const int pos = t0 - 32 + gt; if (pos < 0 || pos >= T) au_s[gt] = 0.f; }   // the first conv zero-pads
            group_sync(1 + g, GT);
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const float4 w4 = *reinterpret_cast<const float4*>(fw_s + k * C + c4 * 4);
                fwr[k][0] = w4.x; fwr[k][1] = w4.y; fwr[k][2] = w4.z; fwr[k][3] = w4.w;
            }
            const float4 b4 = *reinterpret_cast<const float4*>(fb_s + c4 * 4);
            fbr[0] = b4.x; fbr[1] = b4.y; fbr[2] = b4.z; fbr[3] = b4.w;
        }
        float4 xv[1536 / GT];
#pragma unroll
        for (int i = 0; i < 1536 / GT; ++i) {
            const int ar = r_lo + prow + i * (GT / 8), t = t0 - 28 + ar;
            xv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ar < r_hi && t >= 0 && t < T) xv[i] = *reinterpret_cast<const float4*>(a_t + ar * 128 + c4 * 16);
        }
        __syncwarp();   // the 8 lanes of a row sit in one warp: every raw chunk has been read before any row is overwritten
#pragma unroll
        for (int i = 0; i < 1536 / GT; ++i) {
            const int ar = r_lo + prow + i * (GT / 8), t = t0 - 28 + ar;
            const bool active = ar < r_hi;
            float4 pre = xv[i];   // zero outside [0,T)
            if (!ROWS && active && t >= 0 && t < T) {
                float4 sk = make_float4(fbr[0], fbr[1], fbr[2], fbr[3]);
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    const float a = au_s[ar + k + 1];
                    sk.x = fmaf(fwr[k][0], a, sk.x); sk.y = fmaf(fwr[k][1], a, sk.y);
                    sk.z = fmaf(fwr[k][2], a, sk.z); sk.w = fmaf(fwr[k][3], a, sk.w);
                }
                pre = make_float4(xv[i].x + sk.x, xv[i].y + sk.y, xv[i].z + sk.z, xv[i].w + sk.w);
            }
            uint2 hi, lo;
            split4_f16_pre(lrelu02_s(pre.x), lrelu02_s(pre.y), lrelu02_s(pre.z), lrelu02_s(pre.w), hi, lo);
            if (active) {
                const int sw = ar & 7;
                *reinterpret_cast<uint2*>(a_t + ar * 128 + (((c4 >> 1) ^ sw) << 4) + (c4 & 1) * 8) = hi;
                *reinterpret_cast<uint2*>(a_t + ar * 128 + (((4 + (c4 >> 1)) ^ sw) << 4) + (c4 & 1) * 8) = lo;
                if (ar >= 28 && ar < 28 + LT_TT)
                    *reinterpret_cast<float4*>(xs_t + (ar - 28) * 128 + ((c4 ^ ((ar - 28) & 7)) << 4)) = pre;
            }
        }
    };
    // ---- conv MMAs of a tile (ONE thread): A rows of AY[s] x conv weights -> TM[s].conv (+ second pass -> TM[s] + 96) ----
    auto issue_conv = [&](int s, bool have_carry, uint32_t at, uint32_t cwt) {
        const uint32_t d = tmem_u + s * 128;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 1 && have_carry) break;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const uint32_t sh = (uint32_t)(pass * 128 + 27 + (k - 1) * dil) * 128u;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const uint64_t dah = umma_desc_sw128(at + sh + j * 32), dal = umma_desc_sw128(at + sh + 64 + j * 32);
                    const uint64_t dbh = umma_desc_sw128(cwt + k * 4096 + j * 32), dbl = umma_desc_sw128(cwt + k * 4096 + 64 + j * 32);
                    umma_f16(d + pass * 96, dah, dbh, idesc_conv, (k | j) ? 1u : 0u);
                    umma_f16(d + pass * 96, dah, dbl, idesc_conv, 1u);
                    umma_f16(d + pass * 96, dal, dbh, idesc_conv, 1u);
                }
            }
        }
        tc_commit(&bar[0]);
    };
    // ---- P3: y = lrelu(conv + b) -> pieces, rows of the Y tile (over the A tile AY[s]) ----
    auto phase3 = [&](int t0, int s, bool have_carry, int it) {
        unsigned char* a_t = slot + s * LH_A_BYTES;
        const uint32_t tm = tmem_g + s * 128;
        const int q3 = gw & 3, part3 = gw >> 2;
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
                if (copy_to) {
                    *reinterpret_cast<uint4*>(copy_to + yr * 128 + ((chunk ^ sw) << 4)) = make_uint4(h0.x, h0.y, h1.x, h1.y);
                    *reinterpret_cast<uint4*>(copy_to + yr * 128 + (((4 + chunk) ^ sw) << 4)) = make_uint4(l0.x, l0.y, l1.x, l1.y);
                }
            }
        };
        uint32_t v[16];
        tmem_ld_32x32b_x16(tm + ((uint32_t)(q3 * 32) << 16) + part3 * 16, v);
        tmem_ld_wait();
        emit_row(v, q3 * 32 + lane, (q3 == 0 && lane < 2) ? carry + (it & 1) * 256 : nullptr);
        if (have_carry) {
            if (gt < 16) reinterpret_cast<uint4*>(a_t + 128 * 128)[gt] = reinterpret_cast<const uint4*>(carry + ((it & 1) ^ 1) * 256)[gt];
        } else if (q3 == 0) {
            tmem_ld_32x32b_x16(tm + 96 + part3 * 16, v);
            tmem_ld_wait();
            if (lane < 2) emit_row(v, 128 + lane, nullptr);
        }
    };
    // ---- LVC MMAs of a tile (ONE thread): Y rows of AY[s] x predicted kernels -> TM[s].lvc ----
    auto issue_lvc = [&](int s, uint32_t at, uint32_t lwb) {
        const uint32_t d = tmem_u + s * 128 + 32;
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint64_t dah = umma_desc_sw128(at + k * 128 + j * 32), dal = umma_desc_sw128(at + k * 128 + 64 + j * 32);
                const uint64_t dbh = umma_desc_sw128(lwb + k * 8192 + j * 32), dbl = umma_desc_sw128(lwb + k * 8192 + 64 + j * 32);
                umma_f16(d, dah, dbh, idesc_lvc, (k | j) ? 1u : 0u);
                umma_f16(d, dah, dbl, idesc_lvc, 1u);
                umma_f16(d, dal, dbh, idesc_lvc, 1u);
            }
        tc_commit(&bar[1]);
    };
    // ---- P5: gate + residual -> global (tile (b, t0), buffer set s, barrier parity ph); `after_wait` runs in warp 0 once the LVC MMAs are done ----
    auto phase5 = [&](int b, int t0, int s, uint32_t ph, auto&& after_wait) {
        const unsigned char* xs_t = slot + 2 * LH_A_BYTES + s * LH_XS_BYTES;
        const float* lbias = (const float*)(small + s * 256);
        const int q = gw & 3, part = gw >> 2;
        const int r = q * 32 + lane, t = t0 + r;
        const size_t row = ((size_t)b * T + (t < T ? t : 0)) * C + part * 16;
        float4 xs[4], so[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            xs[c] = *reinterpret_cast<const float4*>(xs_t + r * 128 + (((part * 4 + c) ^ (r & 7)) << 4));
            so[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (skip_out) so[c] = *reinterpret_cast<const float4*>(skip + row + c * 4);
        }
        mbar_wait(&bar[1], ph);
        tc_fence_after();
        after_wait();
        uint32_t zs[16], zt[16];
        const uint32_t ta = tmem_g + s * 128 + ((uint32_t)(q * 32) << 16) + 32 + part * 16;
        tmem_ld_32x32b_x16(ta, zs);
        tmem_ld_32x32b_x16(ta + 32, zt);
        tmem_ld_wait();
        if (t < T) {
            const float* lb = lbias + part * 16;
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
                *reinterpret_cast<float4*>(x_out + row + c * 4) = o4;
            }
        }
    };

    // ---- the pipelined tile walk of this group: tiles tile_hi, tile_hi - 1, ..., tile_lo ----
    if (tile_hi >= tile_lo && gw_u == 0) {
        if (elect_one()) { issue_rows(tile_hi, 0); issue_lw(tile_hi, 0, false); }
        __syncwarp();
    }
    int pb = 0, pt0 = 0;    // the previous tile (whose gate epilogue is still to come)
    int it = 0;
    for (int tile = tile_hi; tile >= tile_lo; --tile, ++it) {
        const int b = tile / ntt, tt = tile % ntt, t0 = tt * LT_TT;
        const int s = it & 1;
        const uint32_t ph = (uint32_t)(it & 1);
        const bool have_carry = (tile != tile_hi) && (tt != ntt - 1);
        // P1(i)
        mbar_wait(&bar[2], ph);
        phase1(t0, s);
        fence_async_smem();
        group_sync(1 + g, GT);
        // conv MMAs(i)
        if (gw_u == 0) {
            tc_fence_after();
            uint32_t at = slot_u + (uint32_t)(s * LH_A_BYTES), cwt = cw_u;
            FD_OPAQUE2(at, cwt);
            if (elect_one()) issue_conv(s, have_carry, at, cwt);
            __syncwarp();
        }
        // P5(i-1) while they run; its MMA wait also frees AY[s^1] / LW: request tile i+1's rows and this tile's kernels there
        auto loads_after_lvc = [&]() {
            if (gw_u == 0) {
                if (elect_one()) {
                    if (tile - 1 >= tile_lo) issue_rows(tile - 1, s ^ 1);
                    if (it > 0) issue_lw(tile, s, (tt & 1) == 0 && tt + 1 < ntt && tile != tile_hi);   // (b, tt + 1) was the previous tile: same frame
                }
                __syncwarp();
            }
        };
        if (it > 0) phase5(pb, pt0, s ^ 1, ph ^ 1, loads_after_lvc);
        else loads_after_lvc();
        // P3(i)
        mbar_wait(&bar[0], ph);
        tc_fence_after();
        phase3(t0, s, have_carry, it);
        fence_async_smem();
        tc_fence_before();
        group_sync(1 + g, GT);
        // LVC MMAs(i).  Every thread takes the kernels + biases barrier here (not only the issuing one): it is the acquire for the lbias
        // values its gate epilogue reads next iteration, and at this point the barrier cannot be a phase ahead (the next refill is
        // issued in the next iteration)
        mbar_wait(&bar[3], ph);
        if (gw_u == 0) {
            tc_fence_after();
            uint32_t at = slot_u + (uint32_t)(s * LH_A_BYTES), lwb = slot_u + (uint32_t)(2 * LH_A_BYTES + 2 * LH_XS_BYTES);
            FD_OPAQUE2(at, lwb);
            if (elect_one()) issue_lvc(s, at, lwb);
            __syncwarp();
        }
        pb = b; pt0 = t0;
    }
    if (it > 0) phase5(pb, pt0, (it - 1) & 1, (uint32_t)((it - 1) & 1), [] {});
    tc_fence_before();
    __syncthreads();
    if (tid < 32) {
        tc_fence_after();
        tmem_dealloc(*tmem_base_s, 512u);
    }
}

// ---------------------------------------------------------------------------------------------------------
// EXPERIMENTAL (option "tc_b0", default off; developed on the CPU model, to be measured in round 2): one LVC layer of block 0
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
                bulk_g2s(ring + slot * LB0_PAIR_BYTES + k * 16384 + h * 8192, src + k * 2048, 8192, &ring_full[slot]);
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
    for (; tile < total; tile += gridDim.x, parity ^= 1) {
        const int b = tile / ntt, t0 = (tile % ntt) * LT_TT;
        // ---------------- phase 1: raw rows -> fp16 pieces, in place ----------------
        mbar_wait(bar_rows, parity);
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
        // ---------------- phase 2: dilated conv (two passes: rows 0..127 and the two extra rows) ----------------
        if (gw == 0) {
            tc_fence_after();
            uint32_t at = smem_u, cwt = smem_u32(cw);
            FD_OPAQUE2(at, cwt);
            if (elect_one()) {
#pragma unroll
                for (int pass = 0; pass < 2; ++pass)
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const uint32_t sh = (uint32_t)(pass * 128 + 27 + (k - 1) * dil) * 128u;
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint64_t dah = umma_desc_sw128(at + sh + j * 32), dal = umma_desc_sw128(at + sh + 64 + j * 32);
                            const uint64_t dbh = umma_desc_sw128(cwt + k * 4096 + j * 32), dbl = umma_desc_sw128(cwt + k * 4096 + 64 + j * 32);
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
                    const uint32_t wt = smem_u32(ring) + rs * LB0_PAIR_BYTES;
                    const uint32_t d = tmem_u + 64 + pp * 16;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const uint32_t ysh = (uint32_t)(16 * pp + k) * 128u;
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint64_t dwh = umma_desc_sw128(wt + k * 16384 + j * 32), dwl = umma_desc_sw128(wt + k * 16384 + 64 + j * 32);
                            const uint64_t dyh = umma_desc_sw128(smem_u + ysh + j * 32), dyl = umma_desc_sw128(smem_u + ysh + 64 + j * 32);
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
    }
    tc_fence_before();
    __syncthreads();
    if (gw == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256u);
    }
}

// Block 0's predicted kernels are written by the GEMM in fp32 PANEL order [k][i/4][o][i%4] (the SIMT consumer's layout); this
// converter rewrites one (frame, layer) record in place into the fp16-piece image (staged through shared memory).  Experimental
// path only: in round 2 the GEMM epilogue writes the pieces directly, as it does for blocks 1 and 2.
__global__ void __launch_bounds__(256) k_b0_panel_to_pieces(float* __restrict__ kern, int n_frames) {
    __shared__ float w_s[KK * LVC_OUT];
    const int fr = blockIdx.x / LAYERS, l = blockIdx.x % LAYERS;
    if (fr >= n_frames) return;
    float* rec = kern + (size_t)fr * KCN + (size_t)l * KPL;
    for (int i = threadIdx.x; i < KK * LVC_OUT; i += 256) w_s[i] = rec[i];
    __syncthreads();
    uint16_t* out = reinterpret_cast<uint16_t*>(rec);
    for (int e = threadIdx.x; e < KK * LVC_OUT; e += 256) {
        const int ko = e >> 5, i = e & 31, k = ko >> 6, o = ko & 63;
        const float w = w_s[((k * 8 + (i >> 2)) * LVC_OUT + o) * 4 + (i & 3)];
        uint16_t hi, lo;
        f16_split(w, S16_KERN, hi, lo);
        out[ko * 64 + (((i >> 3) ^ (o & 7)) << 3) + (i & 7)] = hi;
        out[ko * 64 + (((4 + (i >> 3)) ^ (o & 7)) << 3) + (i & 7)] = lo;
    }
}

}  // namespace fd
