// Hardware probes for round-2 kernel work (measurement aid, not product code).  Run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I../../fastdiff_b200/csrc -I../../include -o tc_probe tc_probe.cu && ./tc_probe
// Probe 1: where do the rows of an M = 64 (cta_group::1) accumulator land in TMEM?  (needed for per-frame LVC MMAs in block 1 and
//          for the swapped-operand form of block 0)
// Probe 2: K-major SWIZZLE_64B operand addressing -- is chunk c of row r at position c ^ ((r >> 1) & 3) inside 512-byte groups of
//          8 rows of 64 bytes (descriptor layout type 4, SBO = 512)?  (needed for merged-N B tiles without padding)
#include <cstdio>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../../include/fastdiff_b200.h"
#include "fd_kernels_tc.cuh"
using namespace fd;

__device__ __forceinline__ uint64_t desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;    // SBO: 8 rows * 64 B
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;             // SWIZZLE_64B
    return d;
}

// mode 0: M = 64 row placement; mode 1: SWIZZLE_64B B operand, K-slice 0; mode 2: same, K-slice 1 (start + 32 B)
__global__ void __launch_bounds__(128, 1) k_probe(float* out, int mode) {
    extern __shared__ __align__(1024) unsigned char sm_raw[];
    unsigned char* sm = sm_raw + ((1024u - (smem_u32(sm_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_s;
    __half* A = reinterpret_cast<__half*>(sm);              // 128 rows x 128 B, SWIZZLE_128B
    unsigned char* Bt = sm + 16384;
    for (int i = threadIdx.x; i < 32768 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0u;
    __syncthreads();
    if (mode == 0) {
        // A[m][0] = m + 1 (k = 0 lives in chunk 0 -> position 0 ^ (m & 7)); B[n][0] = 1 for n < 8 (SW128 rows)
        for (int m = threadIdx.x; m < 128; m += 128) A[(m * 128 + ((0 ^ (m & 7)) << 4)) / 2] = __float2half((float)(m + 1));
        if (threadIdx.x < 8) reinterpret_cast<__half*>(Bt)[(threadIdx.x * 128 + ((0 ^ (threadIdx.x & 7)) << 4)) / 2] = __float2half(1.f);
    } else {
        // A[m][k] = 1 iff k == (m & 15) (+16 for slice 1): element k sits in chunk k >> 3 of the 128-byte row
        const int ks = mode == 2 ? 16 : 0;
        for (int m = threadIdx.x; m < 128; m += 128) {
            const int k = ks + (m & 15);
            A[(m * 128 + (((k >> 3) ^ (m & 7)) << 4)) / 2 + (k & 7)] = __float2half(1.f);
        }
        // B[n][k] = 32 n + k (n < 64, k < 32) under the ASSUMED SWIZZLE_64B layout
        for (int idx = threadIdx.x; idx < 64 * 32; idx += 128) {
            const int n = idx >> 5, k = idx & 31;
            const uint32_t off = (uint32_t)(n >> 3) * 512u + (uint32_t)(n & 7) * 64u + ((((uint32_t)k >> 3) ^ (((uint32_t)n >> 1) & 3u)) << 4) + (k & 7) * 2u;
            *reinterpret_cast<__half*>(Bt + off) = __float2half((float)(32 * n + k));
        }
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    if (threadIdx.x < 32) tmem_alloc(&tmem_s, 64u);
    fence_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tmem_s;
    if (threadIdx.x < 32) {
        if (elect_one()) {
            const uint32_t a0 = smem_u32(sm), b0 = smem_u32(Bt);
            if (mode == 0) {
                umma_f16(tm, umma_desc_sw128(a0), umma_desc_sw128(b0), umma_idesc_f16(64, 8), 0u);
            } else {
                const uint32_t ko = mode == 2 ? 32u : 0u;
                umma_f16(tm, umma_desc_sw128(a0 + ko), desc_sw64(b0 + ko), umma_idesc_f16(128, 64), 0u);
            }
            tc_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    {
        const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
        uint32_t v[32];
        tmem_ld_32x32b_x32(tm + ((uint32_t)(q * 32) << 16), v);
        tmem_ld_wait();
        for (int c = 0; c < 32; ++c) out[(q * 32 + lane) * 64 + c] = __uint_as_float(v[c]);
        tmem_ld_32x32b_x32(tm + ((uint32_t)(q * 32) << 16) + 32, v);
        tmem_ld_wait();
        for (int c = 0; c < 32; ++c) out[(q * 32 + lane) * 64 + 32 + c] = __uint_as_float(v[c]);
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 64u); }
}

int main() {
    float* d;
    cudaMalloc(&d, 128 * 64 * 4);
    std::vector<float> h(128 * 64);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    for (int mode = 0; mode < 3; ++mode) {
        cudaMemset(d, 0xFF, 128 * 64 * 4);
        k_probe<<<1, 128, 48 * 1024>>>(d, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        if (mode == 0) {
            printf("probe 1 (M = 64, N = 8): TMEM lane -> accumulator row (value - 1), column 0; '.' = untouched/other\n");
            for (int lane = 0; lane < 128; ++lane) {
                const float v = h[lane * 64];
                if (v >= 1.f && v <= 64.f) printf("  lane %3d <- row %2d\n", lane, (int)v - 1);
            }
        } else {
            const int ks = mode == 2 ? 16 : 0;
            int bad = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 64; ++n) bad += h[m * 64 + n] != (float)(32 * n + ks + (m & 15));
            printf("probe 2 (SWIZZLE_64B, K-slice %d): %d mismatches of %d  (D[3][5] = %.0f, expected %d)\n", mode - 1, bad, 128 * 64,
                   h[3 * 64 + 5], 32 * 5 + ks + 3);
        }
    }
    return 0;
}
