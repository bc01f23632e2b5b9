// emul_rt.cpp -- fibre scheduler for the kernel-logic emulator (tests only).
#include "emul_rt.h"

#include <ucontext.h>

#include <atomic>
#include <thread>
#include <vector>

namespace emu {
thread_local ThreadCtx* cur = nullptr;

namespace {
constexpr size_t kStack = 256 * 1024;

struct Fiber {
    ucontext_t ctx;
    ThreadCtx tc;
    bool done = false;
    unsigned char* stack = nullptr;
};

struct Worker {
    ucontext_t main_ctx;
    std::vector<Fiber> fibers;
    unsigned char* stacks = nullptr;   // malloc'ed: untouched pages cost nothing
    size_t stacks_n = 0;
    std::vector<unsigned char> smem;
    ~Worker() { std::free(stacks); }
    const std::function<void()>* body = nullptr;
    Fiber* running = nullptr;
};
thread_local Worker* wk = nullptr;

void trampoline() {
    Fiber* f = wk->running;
    (*wk->body)();
    f->done = true;
    swapcontext(&f->ctx, &wk->main_ctx);
}

void run_block(Worker& w, dim3 grid, dim3 block, size_t smem, uint3 bid) {
    const unsigned n = block.x * block.y * block.z;
    if (w.fibers.size() < n) w.fibers.resize(n);
    if (w.stacks_n < size_t(n) * kStack) {
        std::free(w.stacks);
        w.stacks_n = size_t(n) * kStack;
        w.stacks = static_cast<unsigned char*>(std::malloc(w.stacks_n));
    }
    if (w.smem.size() < smem + 64) w.smem.resize(smem + 64);
    unsigned char* sm = w.smem.data();
    sm += (16 - (reinterpret_cast<uintptr_t>(sm) & 15)) & 15;
    for (unsigned t = 0; t < n; ++t) {
        Fiber& f = w.fibers[t];
        f.done = false;
        f.stack = w.stacks + size_t(t) * kStack;
        f.tc.tid = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
        f.tc.bid = bid;
        f.tc.bdim = block;
        f.tc.gdim = grid;
        f.tc.smem = sm;
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = &w.main_ctx;
        makecontext(&f.ctx, trampoline, 0);
    }
    unsigned left = n;
    while (left) {
        for (unsigned t = 0; t < n; ++t) {
            Fiber& f = w.fibers[t];
            if (f.done) continue;
            w.running = &f;
            cur = &f.tc;
            swapcontext(&w.main_ctx, &f.ctx);
            if (f.done) --left;
        }
    }
    cur = nullptr;
}
}  // namespace

void syncthreads() {
    Fiber* f = wk->running;
    swapcontext(&f->ctx, &wk->main_ctx);
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const size_t nblocks = size_t(grid.x) * grid.y * grid.z;
    if (nblocks == 0) return;
    unsigned nthr = std::thread::hardware_concurrency();
    if (const char* e = std::getenv("LESGO_EMUL_THREADS")) nthr = std::atoi(e);
    if (nthr < 1) nthr = 1;
    if (nthr > nblocks) nthr = unsigned(nblocks);
    std::atomic<size_t> next{0};
    auto work = [&]() {
        Worker w;
        w.body = &body;
        wk = &w;
        for (;;) {
            size_t b = next.fetch_add(1);
            if (b >= nblocks) break;
            uint3 bid = {unsigned(b % grid.x), unsigned((b / grid.x) % grid.y), unsigned(b / (size_t(grid.x) * grid.y))};
            run_block(w, grid, block, smem, bid);
        }
        wk = nullptr;
    };
    if (nthr == 1) { work(); return; }
    std::vector<std::thread> ts;
    for (unsigned i = 0; i < nthr; ++i) ts.emplace_back(work);
    for (auto& t : ts) t.join();
}
}  // namespace emu
