// tests/emu/simt_emu.cpp - TEST INFRASTRUCTURE ONLY (see simt_emu.h).
#include "simt_emu.h"
#include <cstdio>
#include <thread>
#include <vector>
#include <atomic>

namespace emu {

thread_local Block *g_blk = nullptr;

// void emu_swap(void **save_sp, void *load_sp): save callee-saved registers on the current stack,
// store the stack pointer, switch to the other stack and restore.
extern "C" void emu_swap(void **save_sp, void *load_sp);
__asm__(
    ".text\n.globl emu_swap\n.type emu_swap,@function\nemu_swap:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
    ".size emu_swap,.-emu_swap\n");

static void fiber_main()
{
    Block *b = g_blk;
    (*b->body)();
    b->fibers[b->cur].done = true;
    emu_swap(&b->fibers[b->cur].sp, b->sched_sp);
    std::abort(); // never resumed
}

void yield()
{
    Block *b = g_blk;
    emu_swap(&b->fibers[b->cur].sp, b->sched_sp);
}

static const size_t kStack = 512 * 1024;

static void run_block(Block &b, std::vector<char> &stacks)
{
    g_blk = &b;
    b.bar_count = b.bar_gen = 0;
    std::memset(b.nbar_count, 0, sizeof b.nbar_count);
    std::memset(b.nbar_gen, 0, sizeof b.nbar_gen);
    std::memset(b.warps, 0, sizeof b.warps);
    for (int t = 0; t < b.nthreads; t++) {
        Fiber &f = b.fibers[t];
        f.stack = stacks.data() + (size_t) t * kStack;
        f.done = false;
        // initial frame: six callee-saved slots, then the entry address; at entry rsp % 16 == 8
        uintptr_t top = ((uintptr_t) f.stack + kStack) & ~(uintptr_t) 15;
        uint64_t *sp = (uint64_t *) top;
        *--sp = 0;                         // alignment pad -> entry sees rsp % 16 == 8 after ret
        *--sp = (uint64_t) (uintptr_t) &fiber_main;
        for (int k = 0; k < 6; k++) *--sp = 0;
        f.sp = sp;
    }
    int remaining = b.nthreads;
    long spins = 0;
    while (remaining > 0) {
        int progressed = 0;
        for (int t = 0; t < b.nthreads; t++) {
            Fiber &f = b.fibers[t];
            if (f.done) continue;
            b.cur = t;
            emu_swap(&b.sched_sp, f.sp);
            if (f.done) { remaining--; progressed = 1; }
        }
        if (!progressed && ++spins > 200000000L) { std::fprintf(stderr, "simt_emu: deadlock suspected\n"); std::abort(); }
    }
    g_blk = nullptr;
}

void launch(dim3_t grid, dim3_t block, size_t smem_bytes, const std::function<void()> &body)
{
    int const nblocks = (int) (grid.x * grid.y * grid.z);
    int const nthreads = (int) (block.x * block.y * block.z);
    unsigned hw = std::thread::hardware_concurrency();
    int nworkers = (int) (hw ? hw : 1);
    if (const char *e = std::getenv("LG_EMU_THREADS")) nworkers = std::atoi(e);
    if (nworkers > nblocks) nworkers = nblocks;
    if (nworkers < 1) nworkers = 1;
    std::atomic<int> next(0);
    auto worker = [&]() {
        std::vector<char> stacks((size_t) nthreads * kStack);
        std::vector<Fiber> fibers(nthreads);
        std::vector<unsigned char> smem(smem_bytes + 64);
        for (;;) {
            int const i = next.fetch_add(1);
            if (i >= nblocks) break;
            Block b;
            b.nthreads = nthreads;
            b.fibers = fibers.data();
            b.bdim = block;
            b.gdim = grid;
            b.bidx = dim3_t(i % grid.x, (i / grid.x) % grid.y, i / (grid.x * grid.y));
            b.smem = (unsigned char *) (((uintptr_t) smem.data() + 63) & ~(uintptr_t) 63);
            /* shared memory is NOT zeroed on a GPU: poison it so that reads of never-written words show up here too */
            std::memset(smem.data(), std::getenv("LG_EMU_SMEM_FILL") ? std::atoi(std::getenv("LG_EMU_SMEM_FILL")) : 0xCD, smem.size());
            b.body = &body;
            run_block(b, stacks);
        }
    };
    if (nworkers == 1) worker();
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < nworkers; i++) th.emplace_back(worker);
        for (auto &t : th) t.join();
    }
}

} // namespace emu
