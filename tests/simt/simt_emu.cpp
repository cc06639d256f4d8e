// simt_emu.cpp -- scheduler of the CPU SIMT emulation (tests only; see simt_emu.h)
#include "simt_emu.h"

namespace simt {

Block *g_blk = nullptr;
Fiber *g_cur = nullptr;
int g_schedule = 0;
long long g_count[8] = {0, 0, 0, 0, 0, 0, 0, 0};
unsigned long long g_clock = 0;

// x86-64 SysV context switch: callee-saved registers on the old stack, swap stack pointers
asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size simt_switch,.-simt_switch
)");

static void on_exit_fiber()
{
    Block &b = *g_blk;
    Fiber &f = *g_cur;
    f.done = true;
    b.progress++;
    // an exited thread is no longer waited for: complete what the others are parked at
    Warp &w = b.warps[f.warp];
    w.live--;
    if (w.arrived > 0 && w.arrived == w.live) {
        w.complete(w, w.out[w.gen & 1]);
        w.arrived = 0; w.present = 0; w.gen++;
    }
    b.live--;
    if (b.bar_arrived > 0 && b.bar_arrived == b.live) {
        b.bar_res[b.bar_gen & 1] = b.bar_acc; b.bar_rescnt[b.bar_gen & 1] = b.bar_cnt;
        b.bar_acc = 0; b.bar_cnt = 0; b.bar_arrived = 0; b.bar_gen++;
    }
    simt_switch(&f.sp, b.sched_sp);
    die("resumed a finished fiber");
}

static void trampoline()
{
    g_blk->body();
    on_exit_fiber();
}

static constexpr size_t kStack = 1u << 20;

static void run_block(Block &b, unsigned nthreads)
{
    g_blk = &b;
    b.fibers.assign(nthreads, Fiber{});
    b.warps.assign((nthreads + 31) / 32, Warp{});
    b.live = (int)nthreads;
    for (unsigned t = 0; t < nthreads; ++t) {
        Fiber &f = b.fibers[t];
        f.tidx = uint3{t, 0, 0}; f.warp = (int)(t / 32); f.lane = (int)(t % 32);
        f.stack = (unsigned char *)aligned_alloc(64, kStack);
        if (!f.stack) die("out of memory for fiber stacks");
        uintptr_t top = ((uintptr_t)f.stack + kStack) & ~(uintptr_t)15;
        void **sp = (void **)top;
        *--sp = nullptr;                       // alignment slot: rsp % 16 == 8 at function entry
        *--sp = (void *)&trampoline;           // return address of the first switch
        for (int i = 0; i < 6; ++i) *--sp = nullptr;   // rbp rbx r12 r13 r14 r15
        f.sp = sp;
    }
    for (size_t wi = 0; wi < b.warps.size(); ++wi) {
        const unsigned n = nthreads - (unsigned)wi * 32;
        b.warps[wi].live = n < 32 ? (int)n : 32;
    }
    std::vector<unsigned> order(nthreads);
    for (unsigned t = 0; t < nthreads; ++t) order[t] = g_schedule == 1 ? nthreads - 1 - t : t;
    uint64_t rng = 0x9E3779B97F4A7C15ull * (uint64_t)(g_schedule + 1);
    unsigned long long last = ~0ull; int idle = 0;
    while (b.live > 0) {
        if (g_schedule >= 2) {   // a fresh pseudo-random order every round
            for (unsigned i = nthreads - 1; i > 0; --i) {
                rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
                const unsigned j = (unsigned)(rng % (i + 1));
                std::swap(order[i], order[j]);
            }
        }
        for (unsigned k = 0; k < nthreads; ++k) {
            Fiber &f = b.fibers[order[k]];
            if (f.done) continue;
            g_cur = &f;
            simt_switch(&b.sched_sp, f.sp);
        }
        if (b.progress == last) { if (++idle > 4) die("deadlock: no thread of the block can make progress"); }
        else { idle = 0; last = b.progress; }
    }
    for (auto &f : b.fibers) free(f.stack);
    g_cur = nullptr; g_blk = nullptr;
}

void launch(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()> &body)
{
    for (unsigned bi = 0; bi < grid; ++bi) {
        Block b;
        b.bidx = uint3{bi, 0, 0}; b.bdim = uint3{block, 1, 1}; b.gdim = uint3{grid, 1, 1};
        b.smem = (unsigned char *)aligned_alloc(128, (smem_bytes + 127) & ~(size_t)127);
        memset(b.smem, 0xA5, smem_bytes);    // shared memory starts as garbage, as on the device
        b.body = body;
        run_block(b, block);
        free(b.smem);
    }
}

}  // namespace simt
