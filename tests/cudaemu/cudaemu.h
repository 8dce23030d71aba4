// Minimal CUDA execution-model shim for the CPU test-suite (TEST INFRASTRUCTURE, never shipped).
// Every CTA of a launch runs on ONE OS thread as blockDim cooperative fibres (ucontext).  A fibre runs until it blocks at a barrier
// (__syncthreads, bar.sync id/n, __syncwarp), yields in a spin-wait, or ends; barriers open when the expected number of fibres has
// arrived (all live ones for __syncthreads).  CTAs are handed out to a small pool of OS threads.
// "Device" memory = host memory, __shared__ = static thread_local (one CTA per OS thread at a time).  Enough for the SIMT kernels in
// fastdiff_b200/csrc/fd_kernels_simt.cuh (no warp shuffles, no tensor-core / TMA instructions).
#pragma once
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static thread_local

struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct float4 { float x, y, z, w; } __attribute__((aligned(16)));
static inline float4 make_float4(float a, float b, float c, float d) { float4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }
struct float2 { float x, y; };
struct uint2 { unsigned x, y; } __attribute__((aligned(8)));
struct uint4 { unsigned x, y, z, w; } __attribute__((aligned(16)));
static inline uint2 make_uint2(unsigned a, unsigned b) { uint2 r; r.x = a; r.y = b; return r; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { uint4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }
#define __grid_constant__
template <typename T> static inline T min(T a, T b) { return b < a ? b : a; }
template <typename T> static inline T max(T a, T b) { return a < b ? b : a; }
static inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }

namespace emu {
extern thread_local dim3 t_threadIdx, t_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern thread_local unsigned char* t_dyn_smem;
extern thread_local size_t t_dyn_smem_bytes;
extern thread_local unsigned t_linear_tid, t_cta_rank;
extern thread_local unsigned t_cta_serial;   // bumped for every CTA (pair) an OS thread starts: lets per-CTA model state expire
extern long long g_late_ops[2];               // how many bulk loads / MMAs actually took the late path (tests assert the schedule was exercised)
extern int g_bulk_late;                       // tcemu.h: bit 0 = global->shared bulk copies land when their mbarrier is first polled, bit 1 = MMAs run then
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body, unsigned cluster = 1);
unsigned cluster_rank();                  // CTA pairs (cluster = 2): both CTAs of a pair run as one fibre set on one OS thread
unsigned char* cluster_smem(unsigned rank);
void cluster_barrier();
void syncthreads();
void syncwarp();
void barrier(int key, int expected);   // key 1..15: named barrier (bar.sync key, expected)
void yield();                          // for spin-waits (mbarrier polling)
void cta_begin();                      // per-CTA state of the tensor-core model (tcemu.cpp): TMEM, mbarriers
void cta_end();
}
#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)
static inline void __syncthreads() { emu::syncthreads(); }

#define FD_LAUNCH(kern, grid, block, smem, stream, ...) emu::launch((grid), (block), (smem), [=]() { kern(__VA_ARGS__); })
#define FD_LAUNCH_CLUSTER2(kern, grid, block, smem, stream, ...) emu::launch((grid), (block), (smem), [=]() { kern(__VA_ARGS__); }, 2)
#define __cluster_dims__(...)
#define FD_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::t_dyn_smem)

// arithmetic intrinsics (compile with -ffp-contract=off so the _rn forms are honoured)
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }

template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline void sincospif(float x, float* s, float* c) { *s = (float)sin(3.14159265358979323846 * (double)x); *c = (float)cos(3.14159265358979323846 * (double)x); }
static inline float cospif(float x) { return (float)cos(3.14159265358979323846 * (double)x); }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static inline unsigned int atomicMax(unsigned int* a, unsigned int v) {
    unsigned int old = __atomic_load_n(a, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(a, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}

// runtime subset
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { return posix_memalign(p, 1024, n) ? 1 : cudaSuccess; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
template <typename K> static inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
// a real lane exchange (k_final_w's transpose-reduce): the warp's fibres publish their values, meet at the warp barrier, read the partner's
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
    static thread_local unsigned long long buf[2048];   // per OS thread = per CTA (pair) in flight
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    const unsigned me = emu::t_linear_tid + 1024u * emu::t_cta_rank;
    memcpy(&buf[me], &v, sizeof(T));
    emu::syncwarp();
    T r;
    memcpy(&r, &buf[me ^ (unsigned)lane_mask], sizeof(T));
    emu::syncwarp();
    return r;
}
