// Microbenchmark (not part of the product): what bounds the 2.05 GB store stream of the kernel_conv GEMM epilogue?
// Every variant writes the same number of bytes (FRAMES records of 99,328 B = the predicted-kernel tensor of config 2) from 148 CTAs x 512
// threads with the GEMM's item order (74 CTA pairs walk n-tiles fastest, 256-frame tiles slowest); only the store instruction form and
// the address pattern differ.  No arithmetic, no loads.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_patterns store_patterns.cu && ./store_patterns
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int KCN = 24832;                 // floats per frame record
constexpr size_t REC = (size_t)KCN * 4;    // 99,328 B
constexpr int FRAMES = 20736;              // 81 tiles of 256 (3 blocks x 6888 frames, rounded up)
constexpr int NPAIR = 97;                  // 256-channel (1 KB) tiles per record
constexpr int FT = FRAMES / 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int V>
__global__ void __launch_bounds__(512, 1) k_store(unsigned char* __restrict__ out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = blockIdx.x >> 1, rank = blockIdx.x & 1, npairs = gridDim.x >> 1;
    const int q = warp & 3, cpart = warp >> 2;
    if (V == 4 || V == 5) {
        for (int i = threadIdx.x; i < 16384 / 4; i += 512) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
    }
    const int total = NPAIR * FT;
    // V9: the same 2 x st.b16 form as V1, but every CTA pair walks a CONTIGUOUS range of the (frame tile, n tile) sequence -- it stays on one 256-frame
    // tile for ~97 items (the operand tile of those frames could then stay resident in shared memory) instead of the 74 pairs sharing one frame tile
    const int i_lo = V == 9 ? (int)((long long)total * pair / npairs) : pair, i_hi = V == 9 ? (int)((long long)total * (pair + 1) / npairs) : total;
    const int i_step = V == 9 ? 1 : npairs;
    for (int item = i_lo; item < i_hi; item += i_step) {
        const int ft = item / NPAIR, nt = item % NPAIR;
        if (V == 1 || V == 2 || V == 9) {          // the GEMM's form: a warp owns one 128-byte row of the record (q) for 64 frames (cpart)
            unsigned char* base = out + (size_t)(ft * 256 + cpart * 64) * REC + (size_t)nt * 1024 + rank * 512 + q * 128;
#pragma unroll 8
            for (int j = 0; j < 64; ++j) {
                unsigned char* r = base + (size_t)j * REC;
                if (V == 1 || V == 9) {
                    *reinterpret_cast<uint16_t*>(r + lane * 2) = (uint16_t)0x3c00;
                    *reinterpret_cast<uint16_t*>(r + 64 + lane * 2) = (uint16_t)0x3c00;
                } else {
                    *reinterpret_cast<uint32_t*>(r + lane * 4) = 0x3c003c00u;
                }
            }
        } else if (V == 3) {             // 16 bytes per lane: one warp store = the CTA's 512 contiguous bytes of one frame
            unsigned char* base = out + (size_t)(ft * 256 + warp * 16) * REC + (size_t)nt * 1024 + rank * 512 + lane * 16;
#pragma unroll 8
            for (int j = 0; j < 16; ++j) *reinterpret_cast<uint4*>(base + (size_t)j * REC) = make_uint4(1, 2, 3, 4);
        } else if (V == 4) {             // bulk shared -> global, 512 B per frame (staged epilogue with the record layout kept)
            if (lane < 16) {
                unsigned char* dst = out + (size_t)(ft * 256 + warp * 16 + lane) * REC + (size_t)nt * 1024 + rank * 512;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(sm + (warp * 16 + lane) * 32 % 8192)), "r"(512u) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
            }
        } else if (V == 5) {             // tile-major layout: the CTA's 256 frames x 512 B are one contiguous 128 KB run, written by 8 bulk copies of 16 KB
            if (warp < 8 && lane == 0) {
                unsigned char* dst = out + ((size_t)item * 2 + rank) * 131072 + (size_t)warp * 16384;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(sm)), "r"(16384u) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
            }
        } else if (V == 6) {             // tile-major layout, plain 16-byte stores (fully coalesced)
            unsigned char* dst = out + ((size_t)item * 2 + rank) * 131072 + threadIdx.x * 16;
#pragma unroll 8
            for (int j = 0; j < 16; ++j) *reinterpret_cast<uint4*>(dst + j * 8192) = make_uint4(1, 2, 3, 4);
        } else if (V == 7) {             // tile-major layout, but the GEMM's 2-byte stores (64 B per warp store): is the store size the limit?
            unsigned char* base = out + ((size_t)item * 2 + rank) * 131072 + (size_t)(cpart * 64) * 512 + q * 128;
#pragma unroll 8
            for (int j = 0; j < 64; ++j) {
                unsigned char* r = base + (size_t)j * 512;
                *reinterpret_cast<uint16_t*>(r + lane * 2) = (uint16_t)0x3c00;
                *reinterpret_cast<uint16_t*>(r + 64 + lane * 2) = (uint16_t)0x3c00;
            }
        } else if (V == 8) {             // record layout, 16 B per lane, a warp store covers 8 frames x 64 B (the "un-swapped" form: 32 lines touched per store)
            unsigned char* base = out + (size_t)(ft * 256 + warp * 16 + (lane >> 2)) * REC + (size_t)nt * 1024 + rank * 512 + (lane & 3) * 16;
#pragma unroll 8
            for (int j = 0; j < 16; ++j) *reinterpret_cast<uint4*>(base + (size_t)(j >> 3) * 8 * REC + (j & 7) * 64) = make_uint4(1, 2, 3, 4);
        }
    }
    if (V == 4 || V == 5) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int V>
static void run(const char* what, unsigned char* buf, size_t bytes) {
    cudaFuncSetAttribute(k_store<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        k_store<V><<<148, 512, 16384>>>(buf);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (rep && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("V%d %-98s %8.3f ms  %7.2f TB/s %s\n", V, what, best, bytes / (best * 1e-3) / 1e12, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    const size_t bytes = (size_t)FRAMES * REC;
    unsigned char* buf; cudaMalloc(&buf, bytes + (1 << 20));
    cudaMemset(buf, 0, bytes);
    printf("store-pattern microbenchmark: %.3f GB per launch, 148 CTAs x 512 threads\n", bytes / 1e9);
    {   // reference point: cudaMemset
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a); cudaMemsetAsync(buf, 1, bytes); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("cudaMemset %8.3f ms %7.2f TB/s\n", ms, bytes / (ms * 1e-3) / 1e12);
    }
    run<1>("record layout, 2 x st.b16 per value (the GEMM epilogue today: 64 B per warp store)", buf, bytes);
    run<2>("record layout, st.b32 (128 B per warp store)", buf, bytes);
    run<3>("record layout, st.v4 (512 B per warp store = the CTA's chunk of one frame)", buf, bytes);
    run<8>("record layout, st.v4, 8 frames x 64 B per warp store", buf, bytes);
    run<4>("record layout, cp.async.bulk shared->global 512 B per frame", buf, bytes);
    run<5>("tile-major layout, cp.async.bulk 16 KB (128 KB contiguous per CTA tile)", buf, bytes);
    run<6>("tile-major layout, st.v4 coalesced", buf, bytes);
    run<7>("tile-major layout, 2 x st.b16 per value", buf, bytes);
    run<9>("record layout, 2 x st.b16 per value, contiguous item range per CTA pair (frame tile resident)", buf, bytes);
    run<1>("record layout, 2 x st.b16 per value (again)", buf, bytes);
    return 0;
}
