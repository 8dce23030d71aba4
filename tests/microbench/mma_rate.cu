// Microbenchmark (measurement aid, not product code): cycles per tcgen05.mma for small N, kind::tf32 / kind::f16,
// same vs rotating accumulators, SWIZZLE_128B K-major operands in smem (contents irrelevant).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../fastdiff_b200/csrc -o mma_rate mma_rate.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../include/fastdiff_b200.h"
#include "fd_kernels_tc.cuh"
using namespace fd;

// (umma_f16 comes from fd_kernels_tc.cuh)

template <int N, int KIND /*0 tf32, 1 bf16*/, int NACC, int SWZ>
__global__ void __launch_bounds__(128, 1) k_rate(unsigned long long* out, int iters) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_s;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) ((float*)sm)[i] = 0.f;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tmem_s;
    if (threadIdx.x < 32) {
        const uint32_t a0 = smem_u32(sm), b0 = smem_u32(sm) + 16384;
        const uint32_t idesc = KIND == 0 ? umma_idesc_tf32(128, N)
                                         : ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24));
        if (elect_one()) {
            const long long t0 = clock64();
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint64_t da = SWZ ? umma_desc_sw128(a0 + j * 32) : umma_desc_ns(a0 + 2 * j * 2064, 2064, 128);
                    const uint64_t db = SWZ ? umma_desc_sw128(b0 + j * 32) : umma_desc_ns(b0 + 2 * j * (N * 16), N * 16, 128);
                    const uint32_t d = tm + ((it * 4 + j) % NACC) * N;
                    if (KIND == 0) umma_tf32(d, da, db, idesc, 1u); else umma_f16(d, da, db, idesc, 1u);
                }
            }
            const long long t1 = clock64();
            tc_commit(&bar);
            mbar_wait(&bar, 0);
            const long long t2 = clock64();
            out[0] = t1 - t0; out[1] = t2 - t0;
        }
        __syncwarp();
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory"); }
}

template <int N, int KIND, int NACC, int SWZ>
void run(const char* name, unsigned long long* d, int grid) {
    cudaFuncSetAttribute(k_rate<N, KIND, NACC, SWZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 500;
    k_rate<N, KIND, NACC, SWZ><<<grid, 128, 64 * 1024>>>(d, iters);
    k_rate<N, KIND, NACC, SWZ><<<grid, 128, 64 * 1024>>>(d, iters);
    unsigned long long h[2];
    cudaError_t e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("%-46s grid %3d: issue %7.1f cyc/MMA, complete %7.1f cyc/MMA %s\n", name, grid, (double)h[0] / (iters * 4), (double)h[1] / (iters * 4),
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    unsigned long long* d; cudaMalloc(&d, 64);
    for (int grid : {1, 148}) {
        run<32, 0, 1, 1>("tf32 M128 N32  K8  same acc   sw128", d, grid);
        run<32, 0, 4, 1>("tf32 M128 N32  K8  4 accs     sw128", d, grid);
        run<64, 0, 1, 1>("tf32 M128 N64  K8  same acc   sw128", d, grid);
        run<64, 0, 4, 1>("tf32 M128 N64  K8  4 accs     sw128", d, grid);
        run<128, 0, 1, 1>("tf32 M128 N128 K8  same acc   sw128", d, grid);
        run<256, 0, 1, 1>("tf32 M128 N256 K8  same acc   sw128", d, grid);
        run<32, 0, 1, 0>("tf32 M128 N32  K8  same acc   no-swizzle", d, grid);
        run<64, 0, 1, 0>("tf32 M128 N64  K8  same acc   no-swizzle", d, grid);
        run<256, 0, 1, 0>("tf32 M128 N256 K8  same acc   no-swizzle", d, grid);
        run<32, 1, 1, 1>("bf16 M128 N32  K16 same acc   sw128", d, grid);
        run<64, 1, 1, 1>("bf16 M128 N64  K16 same acc   sw128", d, grid);
        run<256, 1, 1, 1>("bf16 M128 N256 K16 same acc   sw128", d, grid);
    }
    return 0;
}
