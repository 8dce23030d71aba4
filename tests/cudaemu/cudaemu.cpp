// See cudaemu.h.  TEST INFRASTRUCTURE.
#include "cudaemu.h"
namespace emu {
thread_local dim3 t_threadIdx;
dim3 g_blockIdx, g_blockDim, g_gridDim;
pthread_barrier_t g_bar;
unsigned char* g_dyn_smem = nullptr;

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const unsigned nthreads = block.x * block.y * block.z;
    g_blockDim = block; g_gridDim = grid;
    void* sm = nullptr;
    if (posix_memalign(&sm, 1024, smem + 1024)) abort();
    g_dyn_smem = (unsigned char*)sm;
    pthread_barrier_init(&g_bar, nullptr, nthreads);
    std::vector<std::thread> th;
    th.reserve(nthreads);
    for (unsigned t = 0; t < nthreads; ++t) {
        th.emplace_back([=, &body]() {
            t_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            for (unsigned bz = 0; bz < grid.z; ++bz)
                for (unsigned by = 0; by < grid.y; ++by)
                    for (unsigned bx = 0; bx < grid.x; ++bx) {
                        if (t == 0) g_blockIdx = dim3(bx, by, bz);
                        pthread_barrier_wait(&g_bar);   // blockIdx visible; previous CTA fully retired
                        body();
                        pthread_barrier_wait(&g_bar);
                    }
        });
    }
    for (auto& x : th) x.join();
    pthread_barrier_destroy(&g_bar);
    free(sm);
    g_dyn_smem = nullptr;
}
}  // namespace emu
