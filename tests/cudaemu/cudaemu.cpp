// See cudaemu.h.  TEST INFRASTRUCTURE.
#include "cudaemu.h"
#include <stdio.h>
#include <sys/mman.h>
#include <ucontext.h>
#include <atomic>

namespace emu {
thread_local dim3 t_threadIdx, t_blockIdx;
dim3 g_blockDim, g_gridDim;
thread_local unsigned char* t_dyn_smem = nullptr;
thread_local size_t t_dyn_smem_bytes = 0;
thread_local unsigned t_linear_tid = 0;

namespace {
constexpr size_t kStack = 512 * 1024;
constexpr int kMaxBar = 64;           // 0: __syncthreads, 1..15: named barriers (bar.sync id, n), 32 + w: warp w (__syncwarp)

struct Worker {                       // one per OS thread: the fibres of the CTA it is running
    ucontext_t sched;
    std::vector<ucontext_t> ctx;
    std::vector<signed char> state;   // 0 runnable, 1 waiting at a barrier, 2 done
    std::vector<int> wait_key;
    int arrived[kMaxBar];
    unsigned char* stacks = nullptr;
    unsigned nthreads = 0, cur = 0, live = 0;
    const std::function<void()>* body = nullptr;
};
thread_local Worker* t_w = nullptr;

void release(Worker& w, int key) {
    for (unsigned t = 0; t < w.nthreads; ++t)
        if (w.state[t] == 1 && w.wait_key[t] == key) w.state[t] = 0;
    w.arrived[key] = 0;
}

void fiber_entry() {
    Worker* w = t_w;
    (*w->body)();
    w->state[w->cur] = 2;
    --w->live;
    if (w->arrived[0] > 0 && w->arrived[0] >= (int)w->live) release(*w, 0);   // __syncthreads counts the threads that are still alive
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
        w.state[t] = 0;
    }
    for (int k = 0; k < kMaxBar; ++k) w.arrived[k] = 0;
    w.live = n;
    while (w.live) {                  // round-robin over the runnable fibres; a fibre runs until it blocks, yields or ends
        bool ran = false;
        for (unsigned t = 0; t < n; ++t) {
            if (w.state[t] != 0) continue;
            ran = true;
            w.cur = t;
            t_linear_tid = t;
            t_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            swapcontext(&w.sched, &w.ctx[t]);
        }
        if (!ran) { fprintf(stderr, "cudaemu: deadlock (every live fibre waits at a barrier)\n"); abort(); }
    }
}
}  // namespace

void barrier(int key, int expected) {
    Worker* w = t_w;
    if (key == 0) expected = (int)w->live;
    if (++w->arrived[key] >= expected) { release(*w, key); return; }
    w->state[w->cur] = 1;
    w->wait_key[w->cur] = key;
    swapcontext(&w->ctx[w->cur], &w->sched);
}

void yield() {                        // spin-wait helper: let the other fibres of the CTA run
    Worker* w = t_w;
    swapcontext(&w->ctx[w->cur], &w->sched);
}

void syncthreads() { barrier(0, 0); }

void syncwarp() {
    Worker* w = t_w;
    const unsigned warp = w->cur >> 5, n = w->nthreads - warp * 32 < 32 ? w->nthreads - warp * 32 : 32;
    barrier(32 + (int)warp, (int)n);   // exited lanes are not expected by the kernels that use it
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const unsigned nthreads = block.x * block.y * block.z;
    const size_t nblocks = (size_t)grid.x * grid.y * grid.z;
    g_blockDim = block; g_gridDim = grid;
    if (nthreads > 32 * (kMaxBar - 32)) { fprintf(stderr, "cudaemu: block too large\n"); abort(); }
    unsigned nw = std::thread::hardware_concurrency();
    if (nw == 0) nw = 4;
    if (nw > 16) nw = 16;
    if (nw > nblocks) nw = (unsigned)nblocks;
    std::atomic<size_t> next{0};
    auto work = [&]() {
        Worker w;
        w.nthreads = nthreads; w.body = &body;
        w.ctx.resize(nthreads); w.state.resize(nthreads); w.wait_key.resize(nthreads);
        w.stacks = (unsigned char*)mmap(nullptr, (size_t)nthreads * kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (w.stacks == (unsigned char*)MAP_FAILED) abort();
        void* sm = nullptr;
        if (posix_memalign(&sm, 1024, smem + 1024)) abort();
        t_dyn_smem = (unsigned char*)sm;
        t_dyn_smem_bytes = smem;
        t_w = &w;
        for (size_t i = next.fetch_add(1); i < nblocks; i = next.fetch_add(1)) {
            t_blockIdx = dim3((unsigned)(i % grid.x), (unsigned)((i / grid.x) % grid.y), (unsigned)(i / ((size_t)grid.x * grid.y)));
            cta_begin();
            run_cta(w, block);
            cta_end();
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

namespace emu {
void cta_begin() {}
void cta_end() {}
}  // namespace emu
