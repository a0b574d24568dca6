// Kernel-logic emulator runtime (test infrastructure only; see cuda_emu.h).
#include "cuda_emu.h"

#include <mutex>

namespace emu {

thread_local BlockCtx* g_blk = nullptr;
thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
thread_local char* dyn_smem_ptr = nullptr;

static const size_t kStack = 256 * 1024;

void yield_() {
    BlockCtx* b = g_blk;
    Fiber* f = b->cur;
    swapcontext(&f->ctx, &b->sched);
}

void barrier_wait(Barrier& bar) {
    bar.arrived++;
    if (bar.arrived >= bar.expected) {
        bar.arrived = 0;
        bar.gen++;
        return;
    }
    unsigned long my = bar.gen;
    while (bar.gen == my) yield_();
}

static void fiber_entry() {
    BlockCtx* b = g_blk;
    b->body();
    b->cur->done = true;
    // a finished thread no longer takes part in barriers
    Fiber* f = b->cur;
    unsigned lin = f->tid.x + blockDim.x * (f->tid.y + blockDim.y * f->tid.z);
    Barrier* bars[2] = {&b->block_bar, &b->warp_bar[lin >> 5]};
    for (Barrier* bar : bars) {
        bar->expected--;
        if (bar->expected > 0 && bar->arrived >= bar->expected) { bar->arrived = 0; bar->gen++; }
    }
    swapcontext(&f->ctx, &b->sched);
}

static void run_block(BlockCtx& b, dim3 grid, dim3 block, uint3 bid, size_t smem) {
    g_blk = &b;
    gridDim = grid;
    blockDim = block;
    blockIdx = bid;
    unsigned nthreads = block.x * block.y * block.z;
    unsigned nwarps = (nthreads + 31) / 32;
    b.fibers.assign(nthreads, Fiber());
    if (b.stacks_size < (size_t)nthreads * kStack) {
        delete[] b.stacks;
        b.stacks_size = (size_t)nthreads * kStack;
        b.stacks = new char[b.stacks_size];   // uninitialised on purpose: only touched pages get committed
    }
    b.block_bar = Barrier();
    b.block_bar.expected = nthreads;
    b.warp_bar.assign(nwarps, Barrier());
    for (unsigned w = 0; w < nwarps; ++w) b.warp_bar[w].expected = std::min(32u, nthreads - w * 32);
    b.warp_scratch.assign((size_t)nwarps * 32, 0);
    if (b.dyn_smem.size() < smem + 64) b.dyn_smem.resize(smem + 64);
    dyn_smem_ptr = (char*)(((uintptr_t)b.dyn_smem.data() + 63) & ~(uintptr_t)63);
    for (unsigned t = 0; t < nthreads; ++t) {
        Fiber& f = b.fibers[t];
        f.tid.x = t % block.x;
        f.tid.y = (t / block.x) % block.y;
        f.tid.z = t / (block.x * block.y);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = b.stacks + (size_t)t * kStack;
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = &b.sched;
        makecontext(&f.ctx, (void (*)())fiber_entry, 0);
    }
    unsigned remaining = nthreads;
    while (remaining) {
        unsigned progressed = 0;
        for (unsigned t = 0; t < nthreads; ++t) {
            Fiber& f = b.fibers[t];
            if (f.done) continue;
            b.cur = &f;
            threadIdx = f.tid;
            swapcontext(&b.sched, &f.ctx);
            if (f.done) { --remaining; ++progressed; }
        }
        (void)progressed;
    }
    g_blk = nullptr;
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    size_t nblocks = (size_t)grid.x * grid.y * grid.z;
    if (nblocks == 0) return;
    unsigned nthreads = block.x * block.y * block.z;
    if (nthreads == 0 || nthreads > 1024) { fprintf(stderr, "emu: bad block size %u\n", nthreads); abort(); }
    unsigned hw = std::thread::hardware_concurrency();
    unsigned nworkers = (unsigned)std::min<size_t>(nblocks, hw ? hw : 1);
    const char* env = getenv("H2B_EMU_THREADS");
    if (env) nworkers = (unsigned)std::max(1, std::min<int>(atoi(env), (int)nblocks));
    std::atomic<size_t> next(0);
    auto worker = [&]() {
        static thread_local BlockCtx ctx;
        ctx.body = body;
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= nblocks) break;
            uint3 bid;
            bid.x = (unsigned)(i % grid.x);
            bid.y = (unsigned)((i / grid.x) % grid.y);
            bid.z = (unsigned)(i / ((size_t)grid.x * grid.y));
            run_block(ctx, grid, block, bid, smem);
        }
    };
    if (nworkers <= 1) { worker(); return; }
    std::vector<std::thread> th;
    for (unsigned w = 0; w < nworkers; ++w) th.emplace_back(worker);
    for (auto& t : th) t.join();
}

}  // namespace emu
