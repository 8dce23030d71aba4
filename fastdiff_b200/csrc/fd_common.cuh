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
#else
#include <cuda_runtime.h>
#define FD_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define FD_DYN_SMEM(type, name) extern __shared__ __align__(1024) unsigned char name##_raw_[]; type* name = reinterpret_cast<type*>(name##_raw_)
#endif

#include <stdint.h>
#include <math.h>

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
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

}  // namespace fd
