// tcgen05 / TMA path (placeholder until the tensor-core kernels land).
#pragma once
#include <string>
#include "fd_common.cuh"
namespace fd {
static inline int tc_init(void** state, int, const float*, const uint64_t*, std::string&) { *state = nullptr; return 0; }
static inline void tc_destroy(void*) {}
static inline bool tc_available(void*) { return false; }
static inline int tc_kc_gemm(void*, int, const float*, float*, int, int, cudaStream_t, std::string&, uint64_t*) { return -2; }
static inline int tc_lvc_layer(void*, int, int, int, const float*, const float*, const float*, float*, int, int, int, int,
                               cudaStream_t, std::string&, uint64_t*, bool* done) { *done = false; return 0; }
}
