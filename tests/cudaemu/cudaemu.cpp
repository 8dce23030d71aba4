// See cudaemu.h.  TEST INFRASTRUCTURE.
#include "cudaemu.h"
#include <sys/mman.h>
#include <ucontext.h>
#include <atomic>

namespace emu {
thread_local dim3 t_threadIdx, t_blockIdx;
dim3 g_blockDim, g_gridDim;
thread_local unsigned char* t_dyn_smem = nullptr;

namespace {
constexpr size_t kStack = 256 * 1024;

struct Worker {                       // one per OS thread: the fibers of the CTA it is running
    ucontext_t sched;
    std::vector<ucontext_t> ctx;
    std::vector<char> done;
    unsigned char* stacks = nullptr;
    unsigned nthreads = 0, cur = 0;
    const std::function<void()>* body = nullptr;
};
thread_local Worker* t_w = nullptr;

void fiber_entry() {
    Worker* w = t_w;
    (*w->body)();
    w->done[w->cur] = 1;
    swapcontext(&w->ctx[w->cur], &w->sched);   // never resumed
}

void run_cta(Worker& w, dim3 block) {
    const unsigned n = w.nthreads;
    for (unsigned t = 0; t < n; ++t) {
        getcontext(&w.ctx[t]);
        w.ctx[t].uc_stack.ss_sp = w.stacks + (size_t)t * kStack;
        w.ctx[t].uc_stack.ss_size = kStack;
        w.ctx[t].uc_link = nullptr;
        makecontext(&w.ctx[t], fiber_entry, 0);
        w.done[t] = 0;
    }
    unsigned live = n;
    while (live) {                    // one sweep = run every live fiber up to its next __syncthreads (or to its end)
        live = 0;
        for (unsigned t = 0; t < n; ++t) {
            if (w.done[t]) continue;
            w.cur = t;
            t_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            swapcontext(&w.sched, &w.ctx[t]);
            if (!w.done[t]) ++live;
        }
    }
}
}  // namespace

void syncthreads() {                  // yield to the scheduler; it resumes this fiber in the next sweep
    Worker* w = t_w;
    swapcontext(&w->ctx[w->cur], &w->sched);
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const unsigned nthreads = block.x * block.y * block.z;
    const size_t nblocks = (size_t)grid.x * grid.y * grid.z;
    g_blockDim = block; g_gridDim = grid;
    unsigned nw = std::thread::hardware_concurrency();
    if (nw == 0) nw = 4;
    if (nw > 16) nw = 16;
    if (nw > nblocks) nw = (unsigned)nblocks;
    std::atomic<size_t> next{0};
    auto work = [&]() {
        Worker w;
        w.nthreads = nthreads; w.body = &body;
        w.ctx.resize(nthreads); w.done.resize(nthreads);
        w.stacks = (unsigned char*)mmap(nullptr, (size_t)nthreads * kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (w.stacks == (unsigned char*)MAP_FAILED) abort();
        void* sm = nullptr;
        if (posix_memalign(&sm, 1024, smem + 1024)) abort();
        t_dyn_smem = (unsigned char*)sm;
        t_w = &w;
        for (size_t i = next.fetch_add(1); i < nblocks; i = next.fetch_add(1)) {
            t_blockIdx = dim3((unsigned)(i % grid.x), (unsigned)((i / grid.x) % grid.y), (unsigned)(i / ((size_t)grid.x * grid.y)));
            run_cta(w, block);
        }
        t_w = nullptr; t_dyn_smem = nullptr;
        free(sm);
        munmap(w.stacks, (size_t)nthreads * kStack);
    };
    std::vector<std::thread> th;
    for (unsigned i = 1; i < nw; ++i) th.emplace_back(work);
    work();
    for (auto& x : th) x.join();
}
}  // namespace emu
