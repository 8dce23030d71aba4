// Microbenchmark (hardware probe, not part of the product): issue rate of FFMA (3-register), FFMA with an immediate, and the packed FFMA2
// (fma.rn.f32x2) on sm_100a -- decides whether the LVC epilogues should be rewritten with packed fp32 math.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    asm volatile("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rd, {%0, %1}; fma.rn.f32x2 rd, ra, rb, rd; mov.b64 {%0, %1}, rd;}"
                 : "+f"(d0), "+f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
template <int MODE> __global__ void k(float* out, float a, float b, int iters) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
    float a2 = a * 1.0001f, b2 = b * 0.9999f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            if (MODE == 0) { x[i] = fmaf(x[i], a, b); x[i + 1] = fmaf(x[i + 1], a2, b2); }
            if (MODE == 1) { x[i] = fmaf(x[i], 1.0001f, b); x[i + 1] = fmaf(x[i + 1], 0.9999f, b2); }
            if (MODE == 2) { float t0 = x[i], t1 = x[i + 1]; fma2(t0, t1, t0, t1, a, a2); x[i] = t0; x[i + 1] = t1; }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> double run(const char* name, int warps) {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, warps * 32>>>(out, 1.0001f, 0.5f, 10);
    cudaEventRecord(e0);
    k<MODE><<<148, warps * 32>>>(out, 1.0001f, 0.5f, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fma = 148.0 * warps * 32 * 16.0 * iters;
    printf("%-28s warps/SM=%2d  %.3f ms  %.1f TFLOP/s  (%.1f fp32 FMA lanes per SM per clock at 1.965 GHz)\n", name, warps, ms, 2 * fma / ms / 1e9,
           fma / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
    return ms;
}
int main() {
    for (int w : {8, 16, 32}) {
        run<0>("FFMA r,r,r,r", w);
        run<1>("FFMA r,imm,r", w);
        run<2>("FFMA2 (f32x2)", w);
    }
    return 0;
}
