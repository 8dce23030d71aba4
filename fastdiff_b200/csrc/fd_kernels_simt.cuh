// fp32 SIMT kernels of the FastDiff denoiser + sampler update (FD_MODE_FP32_SIMT and all the
// small stages of every mode).  Activations are channels-last in HBM: (B, T, 32) fp32, one 128-byte
// row per time step, so a warp reads/writes whole rows and a time-tile with its halo is a
// contiguous block of rows.  Each kernel cites the reference stage it implements
// (/root/reference/modules/FastDiff/module/...).
#pragma once
#include "fd_common.cuh"

namespace fd {

// ------------------------------------------------------------------------------------------------
// K1+K2+K5  step embedding -> fc_t1 -> swish -> fc_t2 -> swish, then the three per-block fc_t
// (util.py:407-432, FastDiff_model.py:85-87, modules.py:202).  grid = B, block = 512.
// ------------------------------------------------------------------------------------------------
struct EmbedParams {
    const float* freq;                 // [64]
    const float* w1t; const float* b1; // [128][512], [512]
    const float* w2t; const float* b2; // [512][512], [512]
    const float* fct_wt[NBLK];         // [512][80]
    const float* fct_b[NBLK];          // [80]
};

// grid = (B, S): S = 1 with t_dev (fd_denoise: one diffusion step per item) or S reverse steps of a schedule at once (fd_sample: the
// steps' scalars are known before the loop, so the embedding MLP leaves the per-step critical path); slot s writes emb[s][B][512]
// and cnoise[s][3][B][80].
constexpr int EMB_SLOTS = 64;
struct EmbedSteps { float t[EMB_SLOTS]; };

__global__ void __launch_bounds__(512, 1) k_embed(EmbedParams p, const float* __restrict__ t_dev, EmbedSteps ts,
                                               float* __restrict__ emb_all, float* __restrict__ cnoise_all, int B) {
    pdl_trigger();
    pdl_wait();
    __shared__ float s_in[EMB_IN];
    __shared__ float s_mid[EMB_MID];
    __shared__ float s_out[EMB_OUT];
    const int b = blockIdx.x, slot = blockIdx.y, j = threadIdx.x;
    const float tv = t_dev ? t_dev[b] : ts.t[slot];
    float* emb = emb_all + (size_t)slot * B * EMB_OUT;
    float* cnoise = cnoise_all + (size_t)slot * NBLK * B * COND;
    if (j < EMB_IN / 2) {
        const float a = tv * p.freq[j];
        s_in[j] = sinf(a);
        s_in[EMB_IN / 2 + j] = cosf(a);
    }
    __syncthreads();
    {
        float acc = p.b1[j];
#pragma unroll 32
        for (int i = 0; i < EMB_IN; ++i) acc = fmaf(s_in[i], __ldg(p.w1t + i * EMB_MID + j), acc);   // 32 independent loads in flight (the chain is pure L2 latency), same summation order
        s_mid[j] = acc * sigmoidf_(acc);
    }
    __syncthreads();
    {
        float acc = p.b2[j];
#pragma unroll 32
        for (int i = 0; i < EMB_MID; ++i) acc = fmaf(s_mid[i], __ldg(p.w2t + i * EMB_OUT + j), acc);
        acc = acc * sigmoidf_(acc);
        s_out[j] = acc;
        emb[b * EMB_OUT + j] = acc;
    }
    __syncthreads();
    if (j < NBLK * COND) {
        const int blk = j / COND, c = j % COND;
        const float* w = FD_SEL3(p.fct_wt, blk);
        float acc = FD_SEL3(p.fct_b, blk)[c];
#pragma unroll 32
        for (int i = 0; i < EMB_OUT; ++i) acc = fmaf(s_out[i], __ldg(w + i * COND + c), acc);
        cnoise[(blk * B + b) * COND + c] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// K5+K6+K7  KernelPredictor hidden stack (modules.py:202-203, 328-329):
//   cond = mel + fc_t(e);  h0 = lrelu_.1(conv5(cond));  h = h0 + R(h0), R = 6 x [conv3 + lrelu_.1]
// One CTA = 32 output frames of one (block, batch item), halo (8 frames) recomputed.
// Output hk: (3, B, T'+2, 64) channels-last with one zero frame either side of every item, so the
// kernel_conv im2col row of frame f is the 192 contiguous floats starting at padded row f.  Written three
// times: raw fp32 (SIMT GEMM) and as tf32 pieces hi/lo -- or, with f16 != 0, as fp16 pieces of v*S16_HK in rows of 128 B
// (tensor-core GEMM operands).
// grid = (ceil(T'/32), B, 3), block = 256, dynamic smem = KP_SMEM_BYTES.
// ------------------------------------------------------------------------------------------------
struct KpParams {
    const float* in_w[NBLK];  const float* in_b[NBLK];   // [5][80][64], [64]
    const float* res_w[NBLK]; const float* res_b[NBLK];  // [6][3][64][64], [6][64]
};
constexpr int KP_FT = 32;               // output frames per CTA
constexpr int KP_R = 44;                // rows computed per layer (KP_FT + 2*6)
constexpr int KP_CR = 48;               // cond rows (KP_R + 4)
constexpr int KP_CS = 49, KP_HS = 47;   // smem row strides (odd: conflict-free column writes)
constexpr int KP_SMEM_BYTES = (COND * KP_CS + 3 * HID * KP_HS) * 4;

__global__ void __launch_bounds__(256) k_kp_hidden(KpParams p, const float* __restrict__ mel,
                                                   const float* __restrict__ cnoise, float* __restrict__ hk_all,
                                                   float* __restrict__ hk_hi_all, float* __restrict__ hk_lo_all, int B, int Tm, int f16) {
    FD_DYN_SMEM(float, sm);
    float* cond_s = sm;
    float* h0_s = cond_s + COND * KP_CS;
    float* ra_s = h0_s + HID * KP_HS;
    float* rb_s = ra_s + HID * KP_HS;
    const int blk = blockIdx.z, b = blockIdx.y, f0 = blockIdx.x * KP_FT, tid = threadIdx.x;

    for (int idx = tid; idx < COND * KP_CR; idx += 256) {
        const int ci = idx / KP_CR, cr = idx % KP_CR, f = f0 - 8 + cr;
        float v = 0.f;  // the conv zero-pads cond (= mel + noise), so outside [0,T') it is 0, not the noise
        if (f >= 0 && f < Tm) v = mel[((size_t)b * COND + ci) * Tm + f] + cnoise[(blk * B + b) * COND + ci];
        cond_s[ci * KP_CS + cr] = v;
    }
    if (tid < HID) {  // pad columns (row -1 and row 44) of the three hidden tiles
        h0_s[tid * KP_HS] = 0.f; h0_s[tid * KP_HS + KP_R + 1] = 0.f;
        ra_s[tid * KP_HS] = 0.f; ra_s[tid * KP_HS + KP_R + 1] = 0.f;
        rb_s[tid * KP_HS] = 0.f; rb_s[tid * KP_HS + KP_R + 1] = 0.f;
    }
    __syncthreads();

    const int co = tid & 63, g = tid >> 6, r0 = g * 11;
    float acc[11];
    {   // input_conv: 80 -> 64, k = 5
        const float* w = FD_SEL3(p.in_w, blk);
        const float bias = FD_SEL3(p.in_b, blk)[co];
#pragma unroll
        for (int r = 0; r < 11; ++r) acc[r] = bias;
        for (int ci = 0; ci < COND; ++ci) {
            float v[15];
#pragma unroll
            for (int r = 0; r < 15; ++r) v[r] = cond_s[ci * KP_CS + r0 + r];
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const float wv = w[(j * COND + ci) * HID + co];
#pragma unroll
                for (int r = 0; r < 11; ++r) acc[r] = fmaf(wv, v[r + j], acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < 11; ++r) {
            const int f = f0 - 6 + r0 + r;
            h0_s[co * KP_HS + 1 + r0 + r] = (f >= 0 && f < Tm) ? lrelu(acc[r], 0.1f) : 0.f;
        }
    }
    __syncthreads();

    const float* src = h0_s;
    const float* res_w = FD_SEL3(p.res_w, blk);
    const float* res_b = FD_SEL3(p.res_b, blk);
    for (int layer = 0; layer < 6; ++layer) {
        float* dst = (layer & 1) ? rb_s : ra_s;
        const float* w = res_w + layer * 3 * HID * HID;
        const float bias = res_b[layer * HID + co];
#pragma unroll
        for (int r = 0; r < 11; ++r) acc[r] = bias;
        for (int ci = 0; ci < HID; ++ci) {
            float v[13];
#pragma unroll
            for (int r = 0; r < 13; ++r) v[r] = src[ci * KP_HS + r0 + r];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float wv = w[(j * HID + ci) * HID + co];
#pragma unroll
                for (int r = 0; r < 11; ++r) acc[r] = fmaf(wv, v[r + j], acc[r]);
            }
        }
        if (layer < 5) {
#pragma unroll
            for (int r = 0; r < 11; ++r) {
                const int f = f0 - 6 + r0 + r;
                dst[co * KP_HS + 1 + r0 + r] = (f >= 0 && f < Tm) ? lrelu(acc[r], 0.1f) : 0.f;
            }
            __syncthreads();
            src = dst;
        }
    }
    // acc now holds the last residual conv: h = h0 + lrelu(acc)
    const size_t item = (size_t)(blk * B + b) * (Tm + 2) * HID;
#pragma unroll
    for (int r = 0; r < 11; ++r) {
        const int row = r0 + r, f = f0 + row - 6;
        if (row >= 6 && row < 6 + KP_FT && f < Tm) {
            const float v = h0_s[co * KP_HS + 1 + row] + lrelu(acc[r], 0.1f);
            const size_t o = item + (size_t)(1 + f) * HID + co;
            hk_all[o] = v;
            if (f16) {                         // fp16 pieces of v*S16_HK (3xFP16): rows of 64 values = 128 B
                uint16_t h16, l16;
                f16_split(v, S16_HK, h16, l16);
                reinterpret_cast<uint16_t*>(hk_hi_all)[o] = h16;
                reinterpret_cast<uint16_t*>(hk_lo_all)[o] = l16;
            } else {
                const float hi = tf32_rn(v);   // tf32 pieces for the tensor-core GEMM (3xTF32)
                hk_hi_all[o] = hi;
                hk_lo_all[o] = tf32_rn(v - hi);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K8+K9  kernel_conv + bias_conv as ONE fp32 GEMM (modules.py:330-331), SIMT version:
//   kern[(b,f)][n] = KC_B[n] + sum_kk hkpad[(p)*64 + kk] * KC_W[kk][n],   p = padded row of frame f - 1
// M runs over padded rows p in [0, B(T'+2)-2) (2 junk rows per item boundary are computed and
// dropped), so A is a plain strided matrix (row stride 64 floats, row length 192).
// 128x128x16 tiles, 8x8 register micro-tiles, register-prefetch double buffering.
// grid = (KCN/128, ceil(M/128), 3), block = 256.
// ------------------------------------------------------------------------------------------------
struct KcParams { const float* w[NBLK]; const float* b[NBLK]; };

__global__ void __launch_bounds__(256) k_kc_gemm_simt(KcParams p, const float* __restrict__ hk_all,
                                                      float* __restrict__ kern_all, int B, int Tm) {
    __shared__ __align__(16) float As[2][16][132];
    __shared__ __align__(16) float Bs[2][16][128];
    const int blk = blockIdx.z, tid = threadIdx.x;
    const float* hk = hk_all + (size_t)blk * B * (Tm + 2) * HID;
    float* kern = kern_all + (size_t)blk * B * Tm * KCN;
    const float* W = FD_SEL3(p.w, blk);
    const int M = B * (Tm + 2) - 2, m0 = blockIdx.y * 128, n0 = blockIdx.x * 128;

    const int a_row = tid >> 1, a_half = tid & 1;      // A loader: 128 rows x (2 x 8 floats)
    const int b_row = tid >> 4, b_col = (tid & 15) * 8; // B loader: 16 rows x (16 x 8 floats)
    const bool a_ok = (m0 + a_row) < M;
    const float4* a_src = reinterpret_cast<const float4*>(hk + (size_t)(m0 + a_row) * HID + a_half * 8);
    const float4* b_src = reinterpret_cast<const float4*>(W + (size_t)b_row * KCN + n0 + b_col);
    float4 ra0, ra1, rb0, rb1;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

#define FD_KC_GLOAD(kc)                                            \
    do {                                                           \
        ra0 = a_ok ? a_src[(kc) * 4] : zero4;                      \
        ra1 = a_ok ? a_src[(kc) * 4 + 1] : zero4;                  \
        rb0 = b_src[(size_t)(kc) * 16 * (KCN / 4)];                \
        rb1 = b_src[(size_t)(kc) * 16 * (KCN / 4) + 1];            \
    } while (0)
#define FD_KC_SSTORE(buf)                                                                     \
    do {                                                                                      \
        float* a = &As[buf][a_half * 8][a_row];                                               \
        a[0 * 132] = ra0.x; a[1 * 132] = ra0.y; a[2 * 132] = ra0.z; a[3 * 132] = ra0.w;       \
        a[4 * 132] = ra1.x; a[5 * 132] = ra1.y; a[6 * 132] = ra1.z; a[7 * 132] = ra1.w;       \
        float4* bd = reinterpret_cast<float4*>(&Bs[buf][b_row][b_col]);                       \
        bd[0] = rb0; bd[1] = rb1;                                                             \
    } while (0)

    const int ty = tid >> 4, tx = tid & 15;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    FD_KC_GLOAD(0);
    FD_KC_SSTORE(0);
    __syncthreads();
    constexpr int NKC = KCK / 16;
    for (int kc = 0; kc < NKC; ++kc) {
        const int buf = kc & 1;
        if (kc + 1 < NKC) FD_KC_GLOAD(kc + 1);
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kc + 1 < NKC) FD_KC_SSTORE(buf ^ 1);
        __syncthreads();
    }

    const float* bias = FD_SEL3(p.b, blk);
    const float4 bb0 = *reinterpret_cast<const float4*>(bias + n0 + tx * 4);
    const float4 bb1 = *reinterpret_cast<const float4*>(bias + n0 + 64 + tx * 4);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int prow = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (prow >= M) continue;
        const int center = prow + 1, bb = center / (Tm + 2), fp = center % (Tm + 2);
        if (fp < 1 || fp > Tm) continue;  // junk row straddling two batch items
        float* o = kern + ((size_t)bb * Tm + (fp - 1)) * KCN + n0;
        *reinterpret_cast<float4*>(o + tx * 4) =
            make_float4(acc[i][0] + bb0.x, acc[i][1] + bb0.y, acc[i][2] + bb0.z, acc[i][3] + bb0.w);
        *reinterpret_cast<float4*>(o + 64 + tx * 4) =
            make_float4(acc[i][4] + bb1.x, acc[i][5] + bb1.y, acc[i][6] + bb1.z, acc[i][7] + bb1.w);
    }
}

// ------------------------------------------------------------------------------------------------
// Shared inner loop: NT rows x 32 output channels of a 3-tap conv over a channels-last smem tile.
//   acc[n] += sum_{k,ci} tile[(row_n + (k-1)*dil)][ci] * w[k][ci][lane]
// lane = output channel (weights: conflict-free), rows are warp-uniform (inputs: float4 broadcast).
// ------------------------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ void conv3_acc(float (&acc)[NT], const float* __restrict__ tile, int row0, int dil,
                                          const float* __restrict__ w, int lane) {
    const float4* t4 = reinterpret_cast<const float4*>(tile);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int rk = row0 + (k - 1) * dil;
#pragma unroll 2
        for (int c4 = 0; c4 < 8; ++c4) {
            const float w0 = w[(k * C + c4 * 4 + 0) * C + lane];
            const float w1 = w[(k * C + c4 * 4 + 1) * C + lane];
            const float w2 = w[(k * C + c4 * 4 + 2) * C + lane];
            const float w3 = w[(k * C + c4 * 4 + 3) * C + lane];
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                const float4 v = t4[(rk + n) * 8 + c4];
                acc[n] = fmaf(v.x, w0, acc[n]);
                acc[n] = fmaf(v.y, w1, acc[n]);
                acc[n] = fmaf(v.z, w2, acc[n]);
                acc[n] = fmaf(v.w, w3, acc[n]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K3+K4  DiffusionDBlock (modules.py:127-138), optionally fused with first_audio_conv
// (FastDiff_model.py:89): xs = in[::F] (FIRST: in = first_conv(audio), evaluated only at the kept
// positions);  out = conv_d4(lrelu(conv_d2(lrelu(conv_d1(lrelu(xs)))))) + W1x1 xs + b.
// One CTA = 64 output positions, halo 7 recomputed.  grid = (ceil(To/64), B), block = 256.
// ------------------------------------------------------------------------------------------------
struct DbParams {
    const float* res_w; const float* res_b;    // [32][32], [32]
    const float* conv_w; const float* conv_b;  // [3][3][32][32], [3][32]
    const float* first_w; const float* first_b;  // [7][32], [32]  (FIRST only)
};
constexpr int DB_TO = 64, DB_HALO = 7, DB_R = DB_TO + 2 * DB_HALO, DB_PAD = 4, DB_ROWS = DB_R + 2 * DB_PAD + 2;
template <int F>
constexpr int db_smem_bytes() { return (2 * DB_ROWS * C + DB_TO * C + 3 * KK * C + C * C + (DB_R * F + 8)) * 4; }

template <int F, bool FIRST>
__global__ void __launch_bounds__(256) k_dblock(DbParams p, const float* __restrict__ in, float* __restrict__ out,
                                                int Tin, int To) {
    pdl_trigger();
    pdl_wait();
    FD_DYN_SMEM(float, sm);
    float* sa = sm;                      // [DB_ROWS][32], row index = r + DB_PAD
    float* sb = sa + DB_ROWS * C;
    float* raw = sb + DB_ROWS * C;       // [64][32]  xs (pre-activation) of the central rows
    float* cw = raw + DB_TO * C;         // [3][96][32]
    float* rw = cw + 3 * KK * C;         // [32][32]
    float* au = rw + C * C;              // [DB_R*F + 8] audio samples (FIRST)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int o0 = blockIdx.x * DB_TO, b = blockIdx.y;

    for (int i = tid; i < 3 * KK * C; i += 256) cw[i] = p.conv_w[i];
    for (int i = tid; i < C * C; i += 256) rw[i] = p.res_w[i];
    for (int i = tid; i < DB_ROWS * C; i += 256) { sa[i] = 0.f; sb[i] = 0.f; }
    if (FIRST) {
        for (int i = tid; i < DB_R * F + 8; i += 256) {
            const long pos = (long)(o0 - DB_HALO) * F - 3 + i;
            au[i] = (pos >= 0 && pos < Tin) ? in[(size_t)b * Tin + pos] : 0.f;
        }
    }
    __syncthreads();
    for (int idx = tid; idx < DB_R * C; idx += 256) {
        const int r = idx >> 5, c = idx & 31, o = o0 - DB_HALO + r;
        const bool valid = (o >= 0 && o < To);
        float v = 0.f;
        if (valid) {
            if (FIRST) {
                v = p.first_b[c];
#pragma unroll
                for (int k = 0; k < 7; ++k) v = fmaf(p.first_w[k * C + c], au[r * F + k], v);
            } else {
                v = in[((size_t)b * Tin + (size_t)o * F) * C + c];
            }
        }
        sa[(r + DB_PAD) * C + c] = valid ? lrelu(v, 0.2f) : 0.f;
        if (r >= DB_HALO && r < DB_HALO + DB_TO) raw[(r - DB_HALO) * C + c] = v;
    }
    __syncthreads();

    // conv 0 (dil 1): sa -> sb, conv 1 (dil 2): sb -> sa; 80 rows computed, rows >= 78 discarded.
    for (int layer = 0; layer < 2; ++layer) {
        const float* src = layer ? sb : sa;
        float* dst = layer ? sa : sb;
        const float bias = p.conv_b[layer * C + lane];
        {   // 80 rows = 8 warps x 10: one pass, every weight fetched from smem once per warp
            const int base = warp * 10;
            float acc[10];
#pragma unroll
            for (int n = 0; n < 10; ++n) acc[n] = bias;
            conv3_acc<10>(acc, src, base + DB_PAD, 1 << layer, cw + layer * KK * C, lane);
#pragma unroll
            for (int n = 0; n < 10; ++n) {
                const int r = base + n, o = o0 - DB_HALO + r;
                if (r < DB_R) dst[(r + DB_PAD) * C + lane] = (o >= 0 && o < To) ? lrelu(acc[n], 0.2f) : 0.f;
            }
        }
        __syncthreads();
    }
    // conv 2 (dil 4) on the central 64 rows + 1x1 residual on xs.
    {
        const float bias = p.conv_b[2 * C + lane] + p.res_b[lane];
        const float4* raw4 = reinterpret_cast<const float4*>(raw);
        {   // 64 rows = 8 warps x 8
            const int base = warp * 8;
            float acc[8];
#pragma unroll
            for (int n = 0; n < 8; ++n) acc[n] = bias;
            conv3_acc<8>(acc, sa, base + DB_HALO + DB_PAD, 4, cw + 2 * KK * C, lane);
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
                const float w0 = rw[(c4 * 4 + 0) * C + lane], w1 = rw[(c4 * 4 + 1) * C + lane];
                const float w2 = rw[(c4 * 4 + 2) * C + lane], w3 = rw[(c4 * 4 + 3) * C + lane];
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    const float4 v = raw4[(base + n) * 8 + c4];
                    acc[n] = fmaf(v.x, w0, acc[n]); acc[n] = fmaf(v.y, w1, acc[n]);
                    acc[n] = fmaf(v.z, w2, acc[n]); acc[n] = fmaf(v.w, w3, acc[n]);
                }
            }
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int o = o0 + base + n;
                if (o < To) out[((size_t)b * To + o) * C + lane] = acc[n];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K10  LVC-block upsample: lrelu_.2 -> ConvTranspose1d(32,32,k=2r,stride=r,pad=r/2)  (modules.py:205-206)
//   out[t][co] = b[co] + sum_{j in {j1-1, j1}} sum_ci lrelu(in[j][ci]) w[t + r/2 - j r][ci][co],  j1 = (t + r/2)/r
// Outputs of equal phase t mod r share both weight slices, so a warp works on 8 same-phase outputs.
// grid = (ceil(Tin/32), B), block = 256.
// ------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(256) k_upsample(const float* __restrict__ w, const float* __restrict__ bias,
                                                  const float* __restrict__ in, float* __restrict__ out, int Tin) {
    pdl_trigger();
    pdl_wait();
    constexpr int TI = 32;
    __shared__ __align__(16) float in_s[(TI + 2) * C];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int jb = blockIdx.x * TI, b = blockIdx.y;
    for (int idx = tid; idx < (TI + 2) * C; idx += 256) {
        const int r = idx >> 5, c = idx & 31, j = jb - 1 + r;
        in_s[idx] = (j >= 0 && j < Tin) ? lrelu(in[((size_t)b * Tin + j) * C + c], 0.2f) : 0.f;
    }
    __syncthreads();
    const float4* in4 = reinterpret_cast<const float4*>(in_s);
    const float bv = bias[lane];
    for (int item = warp; item < R * (TI / 8); item += 8) {
        const int ph = item / (TI / 8), m0 = (item % (TI / 8)) * 8;
        const int sh = (ph + R / 2) / R, kk1 = (ph + R / 2) % R, kk0 = kk1 + R;
        float acc[8];
#pragma unroll
        for (int n = 0; n < 8; ++n) acc[n] = bv;
#pragma unroll 2
        for (int c4 = 0; c4 < 8; ++c4) {
            float wa[4], wb[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                wa[q] = w[((size_t)kk1 * C + c4 * 4 + q) * C + lane];
                wb[q] = w[((size_t)kk0 * C + c4 * 4 + q) * C + lane];
            }
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const float4 v1 = in4[(m0 + n + sh + 1) * 8 + c4];  // input j1
                const float4 v0 = in4[(m0 + n + sh) * 8 + c4];      // input j1 - 1
                acc[n] = fmaf(v1.x, wa[0], acc[n]); acc[n] = fmaf(v1.y, wa[1], acc[n]);
                acc[n] = fmaf(v1.z, wa[2], acc[n]); acc[n] = fmaf(v1.w, wa[3], acc[n]);
                acc[n] = fmaf(v0.x, wb[0], acc[n]); acc[n] = fmaf(v0.y, wb[1], acc[n]);
                acc[n] = fmaf(v0.z, wb[2], acc[n]); acc[n] = fmaf(v0.w, wb[3], acc[n]);
            }
        }
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int j = jb + m0 + n;
            if (j < Tin) out[((size_t)b * Tin * R + (size_t)j * R + ph) * C + lane] = acc[n];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K11+K12+K13  one LVC layer (modules.py:208-217):
//   xs = x + skip;  y = lrelu_.2(conv_dil(lrelu_.2(xs)));  z = LVC(y; kernel, bias of frame t/hop);
//   x' = xs + sigmoid(z[:32]) * tanh(z[32:])
// skip is either a (B,T,32) buffer or (SKIP_FIRST) first_audio_conv(audio) recomputed on the fly.
// kern points at this layer's slice: kern[(b*T'+f)*KCN + {[96][64] weights, [64] bias}].
// One CTA = TT time steps (TT/HOP frames, >= 1); halo dil+1 re-read.  grid = (ceil(T/TT), B), block = 256.
// ------------------------------------------------------------------------------------------------
struct LvcParams {
    const float* conv_w; const float* conv_b;    // [3][32][32], [32]
    const float* first_w; const float* first_b;  // SKIP_FIRST only
};
constexpr int LVC_H = 28;  // max dilation 27 + 1
template <int HOP, int TT>
constexpr int lvc_smem_bytes() {
    return ((TT + 2 * LVC_H) * C + (TT + 2) * C + KK * C + (HOP >= 64 ? (TT / HOP) * KPL : 0) + (TT + 2 * LVC_H + 8)) * 4;
}

template <int HOP, int TT, bool SKIP_FIRST>
__global__ void __launch_bounds__(256) k_lvc_layer(LvcParams p, const float* __restrict__ x_in,
                                                   const float* __restrict__ skip, const float* __restrict__ kern,
                                                   float* __restrict__ x_out, int T, int Tm, int dil, int flags) {
    constexpr bool WL_SMEM = HOP >= 64;
    constexpr int NF = TT / HOP > 0 ? TT / HOP : 1;
    FD_DYN_SMEM(float, sm);
    float* a_s = sm;                                   // [(TT+2H)][32]  lrelu(x+skip), zero outside [0,T)
    float* y_s = a_s + (TT + 2 * LVC_H) * C;           // [(TT+2)][32]   conv output rows t0-1 .. t0+TT
    float* cw_s = y_s + (TT + 2) * C;                  // [96][32]
    float* wl_s = cw_s + KK * C;                       // [NF][KPL] (WL_SMEM)
    float* au_s = wl_s + (WL_SMEM ? NF * KPL : 0);     // [TT+2H+8] audio (SKIP_FIRST)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t0 = blockIdx.x * TT, b = blockIdx.y;

#ifndef FD_EMU
    if (!WL_SMEM && lane == 0 && (flags & 1)) {
        // hop 8: every warp reads its frame's 24.8 KB of predicted kernels straight from global in phase 3, a few loads at a time
        // (latency-bound).  Pull them into L2 now, all at once, so that phases 1-2 hide the HBM latency.
        const int f = (t0 + warp * 8) / HOP;
        if (f < Tm) {
            const float* src = kern + ((size_t)b * Tm + f) * KCN;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((uint32_t)(KPL * 4)) : "memory");
        }
    }
#endif
    for (int i = tid; i < KK * C; i += 256) cw_s[i] = p.conv_w[i];
    if (WL_SMEM) {
        for (int fi = 0; fi < NF; ++fi) {
            const int f = t0 / HOP + fi;
            if (f < Tm) {
                const float4* src = reinterpret_cast<const float4*>(kern + ((size_t)b * Tm + f) * KCN);
                float4* dst = reinterpret_cast<float4*>(wl_s + fi * KPL);
                for (int i = tid; i < KPL / 4; i += 256) dst[i] = src[i];
            }
        }
    }
    if (SKIP_FIRST) {
        for (int i = tid; i < TT + 2 * LVC_H + 6; i += 256) {
            const int pos = t0 - LVC_H - 3 + i;
            au_s[i] = (pos >= 0 && pos < T) ? skip[(size_t)b * T + pos] : 0.f;
        }
        __syncthreads();
    }
    {   // phase 1: a_s = lrelu(x + skip) on the rows the dilated conv will touch
        const int r_lo = LVC_H - dil - 1, r_hi = LVC_H + TT + dil + 1;
        float4* a4 = reinterpret_cast<float4*>(a_s);
        for (int idx = r_lo * 8 + tid; idx < r_hi * 8; idx += 256) {
            const int r = idx >> 3, c4 = idx & 7, t = t0 - LVC_H + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t >= 0 && t < T) {
                const float4 xv = reinterpret_cast<const float4*>(x_in)[((size_t)b * T + t) * 8 + c4];
                float4 sk;
                if (SKIP_FIRST) {
                    const float* fw = p.first_w + c4 * 4;
                    sk = *reinterpret_cast<const float4*>(p.first_b + c4 * 4);
#pragma unroll
                    for (int k = 0; k < 7; ++k) {
                        const float a = au_s[r + k];
                        sk.x = fmaf(fw[k * C + 0], a, sk.x); sk.y = fmaf(fw[k * C + 1], a, sk.y);
                        sk.z = fmaf(fw[k * C + 2], a, sk.z); sk.w = fmaf(fw[k * C + 3], a, sk.w);
                    }
                } else {
                    sk = reinterpret_cast<const float4*>(skip)[((size_t)b * T + t) * 8 + c4];
                }
                v.x = lrelu(xv.x + sk.x, 0.2f); v.y = lrelu(xv.y + sk.y, 0.2f);
                v.z = lrelu(xv.z + sk.z, 0.2f); v.w = lrelu(xv.w + sk.w, 0.2f);
            }
            a4[idx] = v;
        }
    }
    __syncthreads();
    {   // phase 2: y = lrelu(conv_dil(a) + b) for rows t0-1 .. t0+TT (zero outside [0,T): the LVC pads y with zeros)
        const float bias = p.conv_b[lane];
        for (int base = warp * 8; base < TT; base += 64) {
            float acc[8];
#pragma unroll
            for (int n = 0; n < 8; ++n) acc[n] = bias;
            conv3_acc<8>(acc, a_s, LVC_H - 1 + base, dil, cw_s, lane);
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int t = t0 - 1 + base + n;
                y_s[(base + n) * C + lane] = (t >= 0 && t < T) ? lrelu(acc[n], 0.2f) : 0.f;
            }
        }
        if (warp < 2) {
            const int yr = TT + warp, t = t0 - 1 + yr;
            float acc[1] = {bias};
            conv3_acc<1>(acc, a_s, LVC_H - 1 + yr, dil, cw_s, lane);
            y_s[yr * C + lane] = (t >= 0 && t < T) ? lrelu(acc[0], 0.2f) : 0.f;
        }
    }
    __syncthreads();
    {   // phase 3: location-variable conv + gated residual
        const float4* y4 = reinterpret_cast<const float4*>(y_s);
        for (int base = warp * 8; base < TT; base += 64) {
            if (t0 + base >= T) break;
            const int fi = base / HOP;
            const float* W = WL_SMEM ? (wl_s + fi * KPL) : (kern + ((size_t)b * Tm + (t0 + base) / HOP) * KCN);
            float a0[8], a1[8];
            const float b0 = W[KK * LVC_OUT + lane], b1 = W[KK * LVC_OUT + C + lane];
#pragma unroll
            for (int n = 0; n < 8; ++n) { a0[n] = b0; a1[n] = b1; }
            if (WL_SMEM) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
#pragma unroll 2
                    for (int c4 = 0; c4 < 8; ++c4) {
                        // SWIZZLE_128B tile order [k][o][(i/4) ^ (o&7)][i%4]: one 16-byte vector per lane = the 4 input channels
                        // of chunk c4 for output channel o (lane, lane+32: same o&7, so the same swizzled position)
                        const float4 wa = *reinterpret_cast<const float4*>(W + ((k * LVC_OUT + lane) * 8 + (c4 ^ (lane & 7))) * 4);
                        const float4 wb = *reinterpret_cast<const float4*>(W + ((k * LVC_OUT + C + lane) * 8 + (c4 ^ (lane & 7))) * 4);
                        const float w0[4] = {wa.x, wa.y, wa.z, wa.w}, w1[4] = {wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                        for (int n = 0; n < 8; ++n) {
                            const float4 v = y4[(base + n + k) * 8 + c4];
                            a0[n] = fmaf(v.x, w0[0], a0[n]); a0[n] = fmaf(v.y, w0[1], a0[n]);
                            a0[n] = fmaf(v.z, w0[2], a0[n]); a0[n] = fmaf(v.w, w0[3], a0[n]);
                            a1[n] = fmaf(v.x, w1[0], a1[n]); a1[n] = fmaf(v.y, w1[1], a1[n]);
                            a1[n] = fmaf(v.z, w1[2], a1[n]); a1[n] = fmaf(v.w, w1[3], a1[n]);
                        }
                    }
                }
            } else {
                // kernels straight from HBM (hop 8: each is used by 8 samples only).  Block 0's kernels are stored in PANEL order
                // [k][i/4][o][i%4] (fd_blob.h): consecutive lanes read consecutive 16-byte vectors -> 512 B coalesced per request.
#pragma unroll 1
                for (int k = 0; k < 3; ++k) {
                    float4 wreg[16];   // the whole tap (16 x 512 B per warp) in flight before the first FMA: latency-bound otherwise
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        wreg[2 * c4] = __ldg(reinterpret_cast<const float4*>(W + ((k * 8 + c4) * LVC_OUT + lane) * 4));
                        wreg[2 * c4 + 1] = __ldg(reinterpret_cast<const float4*>(W + ((k * 8 + c4) * LVC_OUT + C + lane) * 4));
                    }
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const float4 wa = wreg[2 * c4], wb = wreg[2 * c4 + 1];
#pragma unroll
                        for (int n = 0; n < 8; ++n) {
                            const float4 v = y4[(base + n + k) * 8 + c4];
                            a0[n] = fmaf(v.x, wa.x, a0[n]); a0[n] = fmaf(v.y, wa.y, a0[n]);
                            a0[n] = fmaf(v.z, wa.z, a0[n]); a0[n] = fmaf(v.w, wa.w, a0[n]);
                            a1[n] = fmaf(v.x, wb.x, a1[n]); a1[n] = fmaf(v.y, wb.y, a1[n]);
                            a1[n] = fmaf(v.z, wb.z, a1[n]); a1[n] = fmaf(v.w, wb.w, a1[n]);
                        }
                    }
                }
            }
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int t = t0 + base + n;  // T is a multiple of 8, so the whole chunk is in range
                float xs = x_in[((size_t)b * T + t) * C + lane];
                if (SKIP_FIRST) {
                    float sk = p.first_b[lane];
#pragma unroll
                    for (int k = 0; k < 7; ++k) sk = fmaf(p.first_w[k * C + lane], au_s[LVC_H + base + n + k], sk);
                    xs += sk;
                } else {
                    xs += skip[((size_t)b * T + t) * C + lane];
                }
                x_out[((size_t)b * T + t) * C + lane] = xs + sigmoidf_(a0[n]) * tanhf(a1[n]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller: on-device Gaussian draws for perf mode (the reference draws on the CPU
// generator, util.py:63-68; parity mode consumes that stream instead).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ float philox_normal(uint64_t elem, uint32_t draw, uint64_t seed) {
    uint32_t r[4];
    const uint64_t grp = elem >> 2;
    philox4x32_10((uint32_t)grp, (uint32_t)(grp >> 32), draw, 0x46443230u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const int sel = (int)(elem & 3);
    const float u1 = ((float)r[sel & 2] + 0.5f) * 2.3283064365386963e-10f;       // (0,1]
    const float u2 = ((float)r[(sel & 2) + 1] + 0.5f) * 2.3283064365386963e-10f;
    const float rad = sqrtf(-2.f * logf(u1 > 1e-37f ? u1 : 1e-37f));
    const float ang = 6.283185307179586f * u2;
    return rad * ((sel & 1) ? sinf(ang) : cosf(ang));
}

// Philox element index of local sample (b, t): b * win_L + win_off + t.  win_L = 0 means "the tensor itself" (b * L + t); a rank that
// holds the time window [win_off, win_off + L) of a longer utterance of win_L samples (time-shard mode) draws the SAME numbers the
// unsharded call would.  seed_ptr (optional) overrides `seed` with a device word: lets a captured graph be replayed with a new seed.
struct NoiseWin { long long win_L, win_off; };
__device__ __forceinline__ size_t noise_elem(const NoiseWin w, int b, int L, int t) {
    return w.win_L ? (size_t)b * (size_t)w.win_L + (size_t)w.win_off + t : (size_t)b * L + t;
}
__global__ void __launch_bounds__(256) k_fill_normal(float* __restrict__ out, int L, size_t n, uint64_t seed, uint32_t draw, NoiseWin win,
                                                     const unsigned long long* __restrict__ seed_ptr) {
    pdl_trigger();
    pdl_wait();
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) out[i] = philox_normal(noise_elem(win, (int)(i / L), L, (int)(i % L)), draw, seed_ptr ? (uint64_t)*seed_ptr : seed);
}
__global__ void k_set_u64(unsigned long long* dst, unsigned long long v) { *dst = v; }

// ------------------------------------------------------------------------------------------------
// K14+K15  final_conv (FastDiff_model.py:100) fused with the sampler update (util.py:219-229).
//   eps[t] = b + sum_{k<7,c} h[t+k-3][c] w[k][c]
//   mode 0: out = eps                       (FastDiff.forward)
//   mode 1: out = (x - coef*eps)/div (+ sigma*z)     rounding order as the reference's in-place ops
//   mode 2: out = c1*x + c2*eps + c3*eps    (the reference's ddim branch, reproduced as written)
// grid = (L/256, B), block = 256.
// ------------------------------------------------------------------------------------------------
struct FinalParams {
    float w[7 * C];
    float b;
    int mode;
    float coef, div, sigma, c1, c2, c3;
    int add_noise;
    uint32_t draw;
    uint64_t seed;
    NoiseWin win;
    const unsigned long long* seed_ptr;
};

#ifndef FINAL_BATCH_LOADS
#define FINAL_BATCH_LOADS 1   // measured on B200 (round 2): k_final 0.444 -> 0.301 ms per N=4 call at config 2
#endif
__global__ void __launch_bounds__(256) k_final(FinalParams p, const float* __restrict__ h, const float* __restrict__ x_t,
                                               const float* __restrict__ z, float* __restrict__ out,
                                               float* __restrict__ seq_out, int L) {
    pdl_trigger();
    pdl_wait();
    // h rows t0-3 .. t0+258 staged as float4 chunks, chunk c4 of row r at position c4 ^ (r & 7) (conflict-free both ways).
    // Thread (q = tid/4, cg = tid%4) accumulates the 4 outputs 4q..4q+3 over channels 8cg..8cg+7 from 10 rows (2 LDS.128 per row:
    // 2.8x less shared-memory traffic than one output per thread), the 4 channel-group partials are summed in a fixed order.
    __shared__ float4 s4[(256 + 6) * 8];
    __shared__ float part[4][256];
    const int tid = threadIdx.x, t0 = blockIdx.x * 256, b = blockIdx.y;
#if FINAL_BATCH_LOADS
    {   // experiment (off: not measured yet): all 9 loads of a thread in flight before the first store.  The loop below compiles to
        // LDG.128 -> STS.128 per iteration, one 512-byte request per warp at a time; ncu has the kernel at 2.1 TB/s (32 % of HBM peak).
        float4 v[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int idx = tid + i * 256, r = idx >> 3, c4 = idx & 7, t = t0 - 3 + r;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < (256 + 6) * 8 && t >= 0 && t < L) v[i] = *reinterpret_cast<const float4*>(h + ((size_t)b * L + t) * C + c4 * 4);
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int idx = tid + i * 256, r = idx >> 3, c4 = idx & 7;
            if (idx < (256 + 6) * 8) s4[r * 8 + (c4 ^ (r & 7))] = v[i];
        }
    }
#else
    for (int idx = tid; idx < (256 + 6) * 8; idx += 256) {
        const int r = idx >> 3, c4 = idx & 7, t = t0 - 3 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t >= 0 && t < L) v = *reinterpret_cast<const float4*>(h + ((size_t)b * L + t) * C + c4 * 4);
        s4[r * 8 + (c4 ^ (r & 7))] = v;
    }
#endif
    __syncthreads();
    {
        const int q = tid >> 2, cg = tid & 3;
        float wr[7][8];   // this channel group's taps in registers (p.w[...cg...] would be a lane-divergent constant-bank access per FMA)
#pragma unroll
        for (int k = 0; k < 7; ++k)
#pragma unroll
            for (int c = 0; c < 8; ++c) wr[k][c] = p.w[k * C + cg * 8 + c];
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int rr = 0; rr < 10; ++rr) {
            const int r = 4 * q + rr;
            const float4 va = s4[r * 8 + ((2 * cg) ^ (r & 7))], vb = s4[r * 8 + ((2 * cg + 1) ^ (r & 7))];
            const float v[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = rr - i;   // row r feeds output 4q+i through tap k
                if (k >= 0 && k < 7) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[i] = fmaf(v[c], wr[k][c], acc[i]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) part[cg][4 * q + i] = acc[i];
    }
    __syncthreads();
    const float eps = (((p.b + part[0][tid]) + part[1][tid]) + part[2][tid]) + part[3][tid];
    const size_t e = (size_t)b * L + t0 + tid;
    float r;
    if (p.mode == 0) {
        r = eps;
    } else if (p.mode == 1) {
        r = __fdiv_rn(__fsub_rn(x_t[e], __fmul_rn(p.coef, eps)), p.div);
        if (p.add_noise) {
            const float zz = z ? z[e] : philox_normal(noise_elem(p.win, b, L, t0 + tid), p.draw, p.seed_ptr ? (uint64_t)*p.seed_ptr : p.seed);
            r = __fadd_rn(r, __fmul_rn(p.sigma, zz));
        }
    } else {
        r = __fadd_rn(__fadd_rn(__fmul_rn(p.c1, x_t[e]), __fmul_rn(p.c2, eps)), __fmul_rn(p.c3, eps));
    }
    out[e] = r;
    if (seq_out) seq_out[e] = r;
}

// K14 + K15 again, as a streaming kernel without shared memory (fd_set_option("final_w", 1); OFF by default: measured slower than k_final --
// 0.116 vs 0.084 ms per launch at config 2 -- although it moves the same bytes with no staging: a warp is one long dependent chain here):
// a warp owns 32 consecutive outputs of one item, lane c = channel c.  It loads the 38 rows t0-3 .. t0+34 with one coalesced 128-byte
// request each (all in flight before the first use), every lane accumulates its channel's contribution to the 32 outputs over the 7 taps
// (224 FMA per lane), a 31-shuffle transpose-reduce leaves output t0 + l in lane l, and the lanes apply the reverse-step update and store
// coalesced.  k_final staged 262 rows per CTA through shared memory (33 KB written, 82 KB read back, two barriers) and reached 3 TB/s;
// the sums are associated differently (channel-major per lane, then across lanes), well inside the stated tolerance.
__global__ void __launch_bounds__(256) k_final_w(FinalParams p, const float* __restrict__ h, const float* __restrict__ x_t,
                                                 const float* __restrict__ z, float* __restrict__ out,
                                                 float* __restrict__ seq_out, int L) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, b = blockIdx.y;
    const int t0 = blockIdx.x * 256 + wid * 32;
    float w[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) w[k] = p.w[k * C + lane];
    const float* hb = h + (size_t)b * L * C + lane;
    float x[38];
#pragma unroll
    for (int r = 0; r < 38; ++r) {
        const int t = t0 - 3 + r;
        x[r] = (t >= 0 && t < L) ? hb[(size_t)t * C] : 0.f;
    }
    float v[32];
#pragma unroll
    for (int o = 0; o < 32; ++o) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < 7; ++k) a = fmaf(w[k], x[o + k], a);   // output t0 + o reads rows t0 + o + k - 3
        v[o] = a;
    }
    // transpose-reduce: after the step with stride s a lane keeps the half of its values whose index bit s equals its lane bit s
#pragma unroll
    for (int s2 = 16; s2 >= 1; s2 >>= 1) {
        const bool up = (lane & s2) != 0;
#pragma unroll
        for (int i = 0; i < s2; ++i) {
            const float send = up ? v[i] : v[i + s2];
            const float keep = up ? v[i + s2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s2);
        }
    }
    const float eps = p.b + v[0];
    const int t = t0 + lane;
    const size_t e = (size_t)b * L + t;
    float r;
    if (p.mode == 0) {
        r = eps;
    } else if (p.mode == 1) {
        r = __fdiv_rn(__fsub_rn(x_t[e], __fmul_rn(p.coef, eps)), p.div);
        if (p.add_noise) {
            const float zz = z ? z[e] : philox_normal(noise_elem(p.win, b, L, t), p.draw, p.seed_ptr ? (uint64_t)*p.seed_ptr : p.seed);
            r = __fadd_rn(r, __fmul_rn(p.sigma, zz));
        }
    } else {
        r = __fadd_rn(__fadd_rn(__fmul_rn(p.c1, x_t[e]), __fmul_rn(p.c2, eps)), __fmul_rn(p.c3, eps));
    }
    out[e] = r;
    if (seq_out) seq_out[e] = r;
}


// ------------------------------------------------------------------------------------------------
// One reverse-step update on its own (util.py:219-229), for denoisers other than FastDiff that share the sampler (SURVEY.md 8f.4:
// the network then is the caller's torch module, only the update runs here).  The same operation sequence as the tail of k_final:
//   ddim == 0:  x = (x - coef*eps) / div;  if (add_noise) x = x + sigma*z        ddim != 0:  x = c1*x + c2*eps + c3*eps
// z == nullptr with add_noise: Philox4x32-10 draw (element index, draw number, seed), as in fd_sample's device-noise mode.
// 4 elements per thread (float4 when the tensors are 16-byte aligned), one pass: 12-16 B read + 4 B written per element.
// ------------------------------------------------------------------------------------------------
struct UpdateParams {
    float coef, div, sigma, c1, c2, c3;
    int ddim, add_noise;
    uint32_t draw;
    uint64_t seed;
};

__device__ __forceinline__ float reverse_update_one(const UpdateParams& p, float x, float eps, const float* z, size_t e) {
    if (p.ddim) return __fadd_rn(__fadd_rn(__fmul_rn(p.c1, x), __fmul_rn(p.c2, eps)), __fmul_rn(p.c3, eps));
    float r = __fdiv_rn(__fsub_rn(x, __fmul_rn(p.coef, eps)), p.div);
    if (p.add_noise) r = __fadd_rn(r, __fmul_rn(p.sigma, z ? z[e] : philox_normal(e, p.draw, p.seed)));
    return r;
}

__global__ void __launch_bounds__(256) k_reverse_update(UpdateParams p, float* __restrict__ x, const float* __restrict__ eps,
                                                        const float* __restrict__ z, float* __restrict__ seq_out, size_t n, int vec) {
    const size_t i0 = ((size_t)blockIdx.x * 256 + threadIdx.x) * 4;
    if (i0 >= n) return;
    if (vec && i0 + 4 <= n) {
        float4 xv = *reinterpret_cast<const float4*>(x + i0);
        const float4 ev = *reinterpret_cast<const float4*>(eps + i0);
        xv.x = reverse_update_one(p, xv.x, ev.x, z, i0);
        xv.y = reverse_update_one(p, xv.y, ev.y, z, i0 + 1);
        xv.z = reverse_update_one(p, xv.z, ev.z, z, i0 + 2);
        xv.w = reverse_update_one(p, xv.w, ev.w, z, i0 + 3);
        *reinterpret_cast<float4*>(x + i0) = xv;
        if (seq_out) *reinterpret_cast<float4*>(seq_out + i0) = xv;
        return;
    }
    for (size_t e = i0; e < n && e < i0 + 4; ++e) {
        const float r = reverse_update_one(p, x[e], eps[e], z, e);
        x[e] = r;
        if (seq_out) seq_out[e] = r;
    }
}

// ------------------------------------------------------------------------------------------------
// The step after the path: wav / |wav|.max() per utterance (task/FastDiff.py:110) then the int16 encode of
// utils/audio.py:11-16 (wav *= 32767; astype(int16) = truncation toward zero), so only 2 bytes/sample leave the GPU.
// Same fp32 operation sequence as the reference (divide, then multiply, both round-to-nearest) -> bit-identical int16.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_absmax(const float* __restrict__ x, unsigned int* __restrict__ amax_bits, int L) {
    __shared__ float red[256];
    const int b = blockIdx.y;
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < (size_t)L; i += (size_t)gridDim.x * 256) m = fmaxf(m, fabsf(x[(size_t)b * L + i]));
    red[threadIdx.x] = m;
    __syncthreads();
    for (int s2 = 128; s2 > 0; s2 >>= 1) {
        if ((int)threadIdx.x < s2) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + s2]);
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicMax(&amax_bits[b], __float_as_uint(red[0]));   // non-negative floats order like their bit patterns
}

__global__ void __launch_bounds__(256) k_wav_int16(const float* __restrict__ x, const unsigned int* __restrict__ amax_bits,
                                                   int16_t* __restrict__ out, int L) {
    const int b = blockIdx.y;
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (size_t)L) return;
    const float amax = __uint_as_float(amax_bits[b]);
    const float v = __fmul_rn(__fdiv_rn(x[(size_t)b * L + i], amax), 32767.0f);
    out[(size_t)b * L + i] = (int16_t)(int)v;   // float -> int truncates toward zero, like numpy's astype
}

// ------------------------------------------------------------------------------------------------
// The step before the path: wav -> log10-mel, the reference's `process_utterance`
// (data_gen/tts/data_gen_utils.py:93-147): librosa 0.8.0 stft(n_fft 1024, hop 256, periodic Hann, centre padding with zeros),
// |.|, librosa.filters.mel(22050, 1024, 80, 80, 7600) (table + non-zero ranges built on the host), log10(max(1e-6, .)).
// One CTA = one frame: 1024 samples centred at f*256, window, radix-2 decimation-in-time FFT in shared memory (bit-reversed
// load, 10 stages of 512 butterflies), magnitudes of bins 0..512, 80 triangular filters.  grid = (T', B), block = 256.
// ------------------------------------------------------------------------------------------------
constexpr int MEL_NFFT = 1024, MEL_HOP = 256, MEL_BINS = 513, MEL_N = 80;

__device__ __forceinline__ int mel_bitrev10(int j) {
    int r = 0;
#pragma unroll
    for (int i = 0; i < 10; ++i) r |= ((j >> i) & 1) << (9 - i);
    return r;
}

__global__ void __launch_bounds__(256) k_mel_frontend(const float* __restrict__ wav, int n, const float* __restrict__ fb,
                                                      const int* __restrict__ fb_range, float* __restrict__ mel, int Tm) {
    __shared__ float re[MEL_NFFT], im[MEL_NFFT];
    __shared__ float twc[MEL_NFFT / 2], tws[MEL_NFFT / 2];
    __shared__ float mag[MEL_BINS + 3];
    const int tid = threadIdx.x, f = blockIdx.x, b = blockIdx.y;
    for (int i = tid; i < MEL_NFFT / 2; i += 256) {   // e^{-2 pi i k / 1024}
        float sn, cs;
        sincospif(-2.0f * (float)i / (float)MEL_NFFT, &sn, &cs);
        twc[i] = cs; tws[i] = sn;
    }
    for (int j = tid; j < MEL_NFFT; j += 256) {
        const long pos = (long)f * MEL_HOP - MEL_NFFT / 2 + j;
        const float v = (pos >= 0 && pos < n) ? wav[(size_t)b * n + pos] : 0.f;
        const float w = 0.5f - 0.5f * cospif(2.0f * (float)j / (float)MEL_NFFT);   // periodic Hann
        const int r = mel_bitrev10(j);
        re[r] = v * w; im[r] = 0.f;
    }
    __syncthreads();
    for (int half = 1; half < MEL_NFFT; half <<= 1) {
        for (int t = tid; t < MEL_NFFT / 2; t += 256) {
            const int k = t & (half - 1), i0 = ((t - k) << 1) + k, i1 = i0 + half;
            const int tw = k * (MEL_NFFT / 2 / half);
            const float c = twc[tw], s2 = tws[tw];
            const float tr = re[i1] * c - im[i1] * s2, ti = re[i1] * s2 + im[i1] * c;
            const float ur = re[i0], ui = im[i0];
            re[i1] = ur - tr; im[i1] = ui - ti;
            re[i0] = ur + tr; im[i0] = ui + ti;
        }
        __syncthreads();
    }
    for (int k = tid; k < MEL_BINS; k += 256) mag[k] = sqrtf(re[k] * re[k] + im[k] * im[k]);
    __syncthreads();
    if (tid < MEL_N) {
        const int lo = fb_range[2 * tid], hi = fb_range[2 * tid + 1];
        float acc = 0.f;
        for (int k = lo; k < hi; ++k) acc = fmaf(fb[tid * MEL_BINS + k], mag[k], acc);
        mel[((size_t)b * MEL_N + tid) * Tm + f] = log10f(fmaxf(1e-6f, acc));
    }
}

// ------------------------------------------------------------------------------------------------
// Debug/inspection gathers used by fd_debug_read (tests only): channels-last -> the reference's layouts.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cl_to_ncl(const float* __restrict__ in, float* __restrict__ out, int B, int T,
                                                   int Cn, int in_row_stride, int in_item_stride, int in_off) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (size_t)B * Cn * T) return;
    const int t = (int)(i % T), c = (int)((i / T) % Cn), b = (int)(i / ((size_t)T * Cn));
    out[i] = in[(size_t)b * in_item_stride + (size_t)t * in_row_stride + in_off + c];
}

__global__ void __launch_bounds__(256) k_kern_to_ref(const float* __restrict__ kern, float* __restrict__ out, int B, int Tm,
                                                     int want_bias, int panel) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (!want_bias) {  // (B,4,32,64,3,T')
        if (i >= (size_t)B * LAYERS * C * LVC_OUT * KS * Tm) return;
        size_t q = i;
        const int f = (int)(q % Tm); q /= Tm;
        const int k = (int)(q % KS); q /= KS;
        const int o = (int)(q % LVC_OUT); q /= LVC_OUT;
        const int ci = (int)(q % C); q /= C;
        const int l = (int)(q % LAYERS); q /= LAYERS;
        const int b = (int)q;
        if (panel == 3) {   // merged-N piece image (k_lvc_p; layout in k_kc_gemm_tc2): T01 rows (piece, o) x [tap 0 | tap 1], T2 rows o x [hi | lo] of tap 2
            const uint16_t* lp = reinterpret_cast<const uint16_t*>(kern + ((size_t)b * Tm + f) * KCN + l * KPL);
            float hi, lo;
            if (k < 2) {
                const int pos = ((((k << 2) + (ci >> 3)) ^ (o & 7)) << 3) + (ci & 7);
                hi = f16_bits_to_float(lp[o * 64 + pos]); lo = f16_bits_to_float(lp[(64 + o) * 64 + pos]);
            } else {
                const uint16_t* rowp = lp + 8192 + o * 64;
                hi = f16_bits_to_float(rowp[(((ci >> 3) ^ (o & 7)) << 3) + (ci & 7)]);
                lo = f16_bits_to_float(rowp[(((4 + (ci >> 3)) ^ (o & 7)) << 3) + (ci & 7)]);
            }
            out[i] = (hi + lo) * (1.f / S16_KERN);
            return;
        }
        if (panel == 2) {   // fp16 pieces of w*S16_KERN: row (k,o) = [32 i hi | 32 i lo], chunk c (8 values) at position c ^ (o & 7)
            const uint16_t* rowp = reinterpret_cast<const uint16_t*>(kern + ((size_t)b * Tm + f) * KCN + l * KPL) + (k * LVC_OUT + o) * 64;
            const float hi = f16_bits_to_float(rowp[(((ci >> 3) ^ (o & 7)) << 3) + (ci & 7)]);
            const float lo = f16_bits_to_float(rowp[(((4 + (ci >> 3)) ^ (o & 7)) << 3) + (ci & 7)]);
            out[i] = (hi + lo) * (1.f / S16_KERN);
            return;
        }
        const int within = panel ? ((k * 8 + (ci >> 2)) * LVC_OUT + o) * 4 + (ci & 3)
                                 : ((k * LVC_OUT + o) * 8 + ((ci >> 2) ^ (o & 7))) * 4 + (ci & 3);
        out[i] = kern[((size_t)b * Tm + f) * KCN + l * KPL + within];
    } else {  // (B,4,64,T')
        if (i >= (size_t)B * LAYERS * LVC_OUT * Tm) return;
        size_t q = i;
        const int f = (int)(q % Tm); q /= Tm;
        const int o = (int)(q % LVC_OUT); q /= LVC_OUT;
        const int l = (int)(q % LAYERS); q /= LAYERS;
        const int b = (int)q;
        out[i] = kern[((size_t)b * Tm + f) * KCN + l * KPL + KK * LVC_OUT + o];
    }
}

}  // namespace fd
