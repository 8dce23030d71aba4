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
thread_local unsigned t_linear_tid = 0, t_cta_rank = 0;
thread_local unsigned t_cta_serial = 0;
int g_bulk_late = getenv("FD_EMU_SCHEDULE") ? atoi(getenv("FD_EMU_SCHEDULE")) & 3 : 0;   // see tcemu.h; the whole suite can be run under a late schedule this way
long long g_late_ops[2] = {0, 0};

namespace {
constexpr size_t kStack = 512 * 1024;
constexpr int kMaxBar = 64;           // per CTA: 0 = __syncthreads, 1..15 = bar.sync id, 32 + w = warp w (__syncwarp)
constexpr int kMaxRanks = 2;          // CTAs per cluster
constexpr int kClusterBar = kMaxRanks * kMaxBar;

struct Worker {                       // one per OS thread: the fibres of the CTA (or CTA pair) it is running
    ucontext_t sched;
    std::vector<ucontext_t> ctx;
    std::vector<signed char> state;   // 0 runnable, 1 waiting at a barrier, 2 done
    std::vector<int> wait_key;
    int arrived[kClusterBar + 1];
    unsigned char* stacks = nullptr;
    unsigned char* smem[kMaxRanks] = {nullptr, nullptr};
    unsigned per_cta = 0, nfib = 0, cur = 0, live[kMaxRanks] = {0, 0};
    size_t first_block = 0;
    dim3 grid, block;
    const std::function<void()>* body = nullptr;
};
thread_local Worker* t_w = nullptr;

void release(Worker& w, int key) {
    for (unsigned t = 0; t < w.nfib; ++t)
        if (w.state[t] == 1 && w.wait_key[t] == key) w.state[t] = 0;
    w.arrived[key] = 0;
}

void enter(Worker& w, unsigned f) {   // make fibre f the current one
    const unsigned rank = f / w.per_cta, t = f % w.per_cta;
    const size_t i = w.first_block + rank;
    w.cur = f;
    t_linear_tid = t;
    t_cta_rank = rank;
    t_dyn_smem = w.smem[rank];
    t_threadIdx = dim3(t % w.block.x, (t / w.block.x) % w.block.y, t / (w.block.x * w.block.y));
    t_blockIdx = dim3((unsigned)(i % w.grid.x), (unsigned)((i / w.grid.x) % w.grid.y), (unsigned)(i / ((size_t)w.grid.x * w.grid.y)));
}

void fiber_entry() {
    Worker* w = t_w;
    (*w->body)();
    const unsigned rank = w->cur / w->per_cta;
    w->state[w->cur] = 2;
    --w->live[rank];
    const int k0 = (int)rank * kMaxBar;   // __syncthreads counts the threads of the CTA that are still alive
    if (w->arrived[k0] > 0 && w->arrived[k0] >= (int)w->live[rank]) release(*w, k0);
    swapcontext(&w->ctx[w->cur], &w->sched);   // never resumed
}

void run_group(Worker& w, unsigned nranks) {
    w.nfib = w.per_cta * nranks;
    for (unsigned f = 0; f < w.nfib; ++f) {
        getcontext(&w.ctx[f]);
        w.ctx[f].uc_stack.ss_sp = w.stacks + (size_t)f * kStack;
        w.ctx[f].uc_stack.ss_size = kStack;
        w.ctx[f].uc_link = nullptr;
        makecontext(&w.ctx[f], fiber_entry, 0);
        w.state[f] = 0;
    }
    for (int k = 0; k <= kClusterBar; ++k) w.arrived[k] = 0;
    for (unsigned r = 0; r < kMaxRanks; ++r) w.live[r] = r < nranks ? w.per_cta : 0;
    while (w.live[0] + w.live[1]) {   // round-robin over the runnable fibres; a fibre runs until it blocks, yields or ends
        bool ran = false;
        for (unsigned f = 0; f < w.nfib; ++f) {
            if (w.state[f] != 0) continue;
            ran = true;
            enter(w, f);
            swapcontext(&w.sched, &w.ctx[f]);
        }
        if (!ran) { fprintf(stderr, "cudaemu: deadlock (every live fibre waits at a barrier)\n"); abort(); }
    }
}

void block_here(Worker* w, int key) {
    w->state[w->cur] = 1;
    w->wait_key[w->cur] = key;
    const unsigned me = w->cur;
    swapcontext(&w->ctx[me], &w->sched);
}
}  // namespace

void barrier(int key, int expected) {
    Worker* w = t_w;
    const unsigned rank = w->cur / w->per_cta;
    if (key == 0) expected = (int)w->live[rank];
    key += (int)rank * kMaxBar;
    if (++w->arrived[key] >= expected) { release(*w, key); return; }
    block_here(w, key);
}

void cluster_barrier() {
    Worker* w = t_w;
    if (++w->arrived[kClusterBar] >= (int)(w->live[0] + w->live[1])) { release(*w, kClusterBar); return; }
    block_here(w, kClusterBar);
}

void yield() {                        // spin-wait helper: let the other fibres run
    Worker* w = t_w;
    const unsigned me = w->cur;
    swapcontext(&w->ctx[me], &w->sched);
}

void syncthreads() { barrier(0, 0); }

void syncwarp() {
    Worker* w = t_w;
    const unsigned t = w->cur % w->per_cta, warp = t >> 5, n = w->per_cta - warp * 32 < 32 ? w->per_cta - warp * 32 : 32;
    barrier(32 + (int)warp, (int)n);   // exited lanes are not expected by the kernels that use it
}

unsigned cluster_rank() { return t_cta_rank; }
unsigned char* cluster_smem(unsigned rank) { return t_w->smem[rank]; }

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body, unsigned cluster) {
    const unsigned nthreads = block.x * block.y * block.z;
    const size_t nblocks = (size_t)grid.x * grid.y * grid.z;
    if (cluster < 1 || cluster > kMaxRanks || nblocks % cluster) { fprintf(stderr, "cudaemu: bad cluster size\n"); abort(); }
    const size_t ngroups = nblocks / cluster;
    g_blockDim = block; g_gridDim = grid;
    if (nthreads > 32 * (kMaxBar - 32)) { fprintf(stderr, "cudaemu: block too large\n"); abort(); }
    unsigned nw = std::thread::hardware_concurrency();
    if (nw == 0) nw = 4;
    if (nw > 16) nw = 16;
    if (nw > ngroups) nw = (unsigned)ngroups;
    std::atomic<size_t> next{0};
    auto work = [&]() {
        Worker w;
        w.per_cta = nthreads; w.body = &body; w.grid = grid; w.block = block;
        const unsigned nf = nthreads * cluster;
        w.ctx.resize(nf); w.state.resize(nf); w.wait_key.resize(nf);
        w.stacks = (unsigned char*)mmap(nullptr, (size_t)nf * kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (w.stacks == (unsigned char*)MAP_FAILED) abort();
        for (unsigned r = 0; r < cluster; ++r) {
            void* sm = nullptr;
            if (posix_memalign(&sm, 1024, smem + 1024)) abort();
            w.smem[r] = (unsigned char*)sm;
        }
        t_dyn_smem_bytes = smem;
        t_w = &w;
        for (size_t i = next.fetch_add(1); i < ngroups; i = next.fetch_add(1)) {
            w.first_block = i * cluster;
            cta_begin();
            run_group(w, cluster);
            cta_end();
        }
        t_w = nullptr; t_dyn_smem = nullptr;
        for (unsigned r = 0; r < cluster; ++r) free(w.smem[r]);
        munmap(w.stacks, (size_t)nf * kStack);
    };
    std::vector<std::thread> th;
    for (unsigned i = 1; i < nw; ++i) th.emplace_back(work);
    work();
    for (auto& x : th) x.join();
}

void cta_begin() { ++t_cta_serial; }
void cta_end() {}
}  // namespace emu

extern "C" long long cudaemu_late_ops(int which) { return __atomic_load_n(&emu::g_late_ops[which & 1], __ATOMIC_RELAXED); }
extern "C" void cudaemu_set_bulk_late(int mask) { emu::g_bulk_late = mask & 3; }   // bit 0: late bulk loads, bit 1: late MMAs (tcemu.h)
