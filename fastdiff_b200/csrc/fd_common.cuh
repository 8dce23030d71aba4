// Shared constants, launch macro and small device helpers.
//
// The SIMT kernels and the host orchestration in this directory also compile as plain C++ when
// FD_EMU is defined (tests/cudaemu/cudaemu.h provides threadIdx/blockIdx/__syncthreads on OS
// threads).  That build is TEST INFRASTRUCTURE: it lets the CPU test-suite check the index
// arithmetic, the blob layout and the sampler logic of exactly this source against the oracle
// without a GPU.  It is never loaded by the product package.
#pragma once

#ifdef FD_EMU
#include "cudaemu.h"
#define FD_LAUNCH_PDL FD_LAUNCH
namespace fd { inline void pdl_wait() {} inline void pdl_trigger() {} }
#else
#include <cuda_runtime.h>
#define FD_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define FD_DYN_SMEM(type, name) extern __shared__ __align__(1024) unsigned char name##_raw_[]; type* name = reinterpret_cast<type*>(name##_raw_)
// Programmatic dependent launch: the kernels of a reverse step form one dependent chain of ~25 launches, each with a prologue that touches
// no data of its predecessor (barrier init, TMEM allocation, weight images into shared memory) and a tail in which most SMs idle.  Launched
// with the programmatic-serialization attribute, a kernel's CTAs are scheduled as soon as its predecessor's CTAs have all called
// pdl_trigger() (their first statement) and leave SM resources free; they run their prologue and then block in pdl_wait() until the
// predecessor has completed and flushed.  RULES: every kernel launched through FD_LAUNCH_PDL calls pdl_wait() before its first access
// to data another kernel writes or reads-then-overwrites, and writes nothing but shared memory / TMEM before it.
namespace fd {
static int g_fd_pdl = 0;   // fd_set_option("pdl", 1) sets the attribute.  OFF by default: measured on B200 (round 2, graph replay of the N=4 call)
                           // 1 s utterance 1.122 ms with / 1.052 ms without, config 2 9.80 / 9.70 ms -- the hand-over of a programmatic edge costs more
                           // than the overlapped prologues save (griddepcontrol.* are no-ops without the attribute)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t fd_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = g_fd_pdl ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
}
#define FD_LAUNCH_PDL(kern, grid, block, smem, stream, ...) fd::fd_launch_pdl(kern, dim3(grid), dim3(block), (smem), (stream), __VA_ARGS__)
#endif

#include <stdint.h>
#include <math.h>
#include <string.h>

namespace fd {

// The one architecture the reference ships (modules/FastDiff/config/base.yaml:21-33).
constexpr int C = 32;          // inner_channels
constexpr int COND = 80;       // cond_channels
constexpr int HID = 64;        // kpnet_hidden_channels
constexpr int LAYERS = 4;      // lvc_layers_each_block
constexpr int KS = 3;          // lvc_kernel_size
constexpr int NBLK = 3;        // len(upsample_ratios)
constexpr int EMB_IN = 128, EMB_MID = 512, EMB_OUT = 512;
constexpr int HOP_TOTAL = 256; // 8*8*4
constexpr int KK = C * KS;             // 96: contraction length of an LVC / dilated conv
constexpr int LVC_OUT = 2 * C;         // 64
constexpr int KPL = KK * LVC_OUT + LVC_OUT;  // 6208 floats per (frame, layer): [96][64] weights + 64 biases
constexpr int KCN = LAYERS * KPL;      // 24832 GEMM columns of kernel_conv + bias_conv
constexpr int KCK = HID * 3;           // 192 GEMM depth of kernel_conv

__host__ __device__ inline int ratio_of(int blk) { return blk == 2 ? 4 : 8; }        // upsample_ratios
__host__ __device__ inline int hop_of(int blk) { return blk == 0 ? 8 : (blk == 1 ? 64 : 256); }
__host__ __device__ inline int down_factor(int n) { return n == 0 ? 4 : 8; }         // upsample_ratios[2-n]

// static selection among three kernel-parameter pointers (dynamic indexing would spill the array to local memory)
#define FD_SEL3(arr, i) ((i) == 0 ? (arr)[0] : ((i) == 1 ? (arr)[1] : (arr)[2]))

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }
// round-to-nearest-even to tf32, kept in an fp32 container with the low 13 mantissa bits zero (== weights.tf32_round)
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t u = __float_as_uint(x);
    u = (u + 0xFFFu + ((u >> 13) & 1u)) & 0xFFFFE000u;
    return __uint_as_float(u);
}
// ---- fp16 pieces (mode tc_3xf16): v*S = hi + lo, hi = RN_f16(v*S), lo = RN_f16(v*S - hi).  hi and lo carry 11 significant
// bits each (the same as a tf32 pair), S is a power of two that keeps lo out of the fp16 subnormals for the magnitudes the
// network produces (full 22-bit precision for |v*S| >= 2^-3, absolute error <= 2^-25 below); |v*S| saturates at 65504.
// Software conversion = the bit-exact definition (round to nearest even, gradual underflow); the device build uses cvt.rn.f16.f32.
__host__ __device__ inline uint16_t f16_bits_rn_soft(float f) {
    uint32_t x; memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7FFFFFFFu;
    if (x >= 0x7F800000u) return (uint16_t)(sign | (x > 0x7F800000u ? 0x7E00u : 0x7C00u));
    if (x >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);                      // rounds to >= 65520 -> inf
    if (x < 0x33000001u) return (uint16_t)sign;                                   // <= 2^-25 -> 0 (ties to even)
    uint32_t e = x >> 23, m = (x & 0x7FFFFFu) | 0x800000u, shift, half;
    if (e >= 113) { shift = 13; half = (uint32_t)(e - 112) << 10; m &= 0x7FFFFFu; }   // normal: drop the implicit bit
    else { shift = 126 - e; half = 0; }                                            // subnormal: keep it and shift further
    const uint32_t q = m >> shift, rem = m & ((1u << shift) - 1u), mid = 1u << (shift - 1);
    uint32_t h = half + q;
    if (rem > mid || (rem == mid && (h & 1u))) ++h;                                // carries into the exponent correctly
    return (uint16_t)(sign | h);
}
__host__ __device__ inline float f16_bits_to_float_soft(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 31u, m = h & 0x3FFu;
    uint32_t x;
    if (e == 0) {
        if (m == 0) x = sign;
        else { int k = 0; uint32_t mm = m; while (!(mm & 0x400u)) { mm <<= 1; ++k; } x = sign | ((uint32_t)(113 - k) << 23) | ((mm & 0x3FFu) << 13); }
    } else if (e == 31) x = sign | 0x7F800000u | (m << 13);
    else x = sign | ((e + 112u) << 23) | (m << 13);
    float f; memcpy(&f, &x, 4);
    return f;
}
#ifdef FD_EMU
__device__ __forceinline__ uint16_t f16_bits_rn(float f) { return f16_bits_rn_soft(f); }
__device__ __forceinline__ float f16_bits_to_float(uint16_t h) { return f16_bits_to_float_soft(h); }
#else
__device__ __forceinline__ uint16_t f16_bits_rn(float f) { uint16_t h; asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(f)); return h; }
__device__ __forceinline__ float f16_bits_to_float(uint16_t h) { float f; asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"(h)); return f; }
#endif
constexpr float F16_MAX = 65504.f;
// (hi, lo) fp16 pieces of v*scale, saturated
__device__ __forceinline__ void f16_split(float v, float scale, uint16_t& hi, uint16_t& lo) {
    const float s = fminf(fmaxf(v * scale, -F16_MAX), F16_MAX);
    hi = f16_bits_rn(s);
    lo = f16_bits_rn(s - f16_bits_to_float(hi));
}
// fixed power-of-two prescales of the dynamic operands (weights carry per-tensor scales chosen by the packer, section SCALES16)
constexpr float S16_HK = 16.f;       // kernel-predictor hidden state (B operand of the kernel_conv GEMM)
constexpr float S16_ACT = 16.f;      // LVC-block activations (A operands of the dilated conv and of the location-variable conv)
constexpr float S16_KERN = 64.f;     // predicted LVC kernels (B operand of the location-variable conv), written by the kernel_conv GEMM

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

}  // namespace fd
