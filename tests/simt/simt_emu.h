// simt_emu.h -- a small CPU emulation of the CUDA execution model, for TESTS ONLY.
//
// It lets the unmodified source of the period kernel (r_mabm_b200/csrc/cats_period.cuh) be compiled with
// g++ and run on the CPU, so that the kernel's LOGIC -- commit rounds, replay order, compaction, the
// multi-warp hand-overs -- is checked against the oracle without a GPU, and under several lane
// schedules (a result that depends on the schedule is a missing __syncwarp / barrier: a race).
// Nothing under r_mabm_b200/ includes or loads this; the product path has no CPU fallback.
//
// Model: every CUDA thread of a block is a fiber (own stack, hand-written context switch); fibers run
// one at a time and yield only inside collectives (warp shuffles / votes / match / __syncwarp, block
// barriers, mbarrier waits).  A collective completes when every non-exited lane of the warp has
// arrived.  Blocks run one after the other.  fp64 arithmetic is IEEE on both sides (compile with
// -ffp-contract=off), so results are comparable bit for bit.
#pragma once
#include <cuda_runtime.h>   // vector types (double2, uint4, make_*) only: host mode declares no device intrinsics
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <vector>

#define CATS_EMU 1

namespace simt {

struct Fiber {
    void *sp = nullptr;
    unsigned char *stack = nullptr;
    uint3 tidx{0, 0, 0};
    int warp = 0, lane = 0;
    bool done = false;
};

struct Warp {
    int live = 32;          // non-exited lanes
    int arrived = 0;
    unsigned gen = 0;
    unsigned present = 0;   // lanes that have arrived at the pending collective
    uint64_t in[32], in2[32];
    uint64_t out[2][32];
    int kind = -1;          // kind of the pending collective (all arrivals must agree)
    void (*complete)(Warp &, uint64_t *) = nullptr;
};

struct Block {
    uint3 bidx{0, 0, 0}, bdim{1, 1, 1}, gdim{1, 1, 1};
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    unsigned char *smem = nullptr;
    int live = 0;
    int bar_arrived = 0;
    unsigned bar_gen = 0;
    int bar_acc = 0, bar_cnt = 0;     // __syncthreads_or / _count: the reduction of the barrier in progress
    int bar_res[2] = {0, 0}, bar_rescnt[2] = {0, 0};
    // named barriers (bar.sync id, count): arrival counts and generations
    int nb_arrived[16] = {0};
    unsigned nb_gen[16] = {0};
    unsigned long long progress = 0;
    void *sched_sp = nullptr;
    std::function<void()> body;
};

extern Block *g_blk;
extern Fiber *g_cur;
extern int g_schedule;   // 0 forward, 1 reverse, >= 2 pseudo-random (seed)
extern long long g_count[8];   // development counters (kernel code under CATS_EMU_COUNT bumps them; emu_counts reads them)
extern unsigned long long g_clock;

extern "C" void simt_switch(void **save_sp, void *load_sp);

inline void yield() { simt_switch(&g_cur->sp, g_blk->sched_sp); }

[[noreturn]] inline void die(const char *msg)
{
    fprintf(stderr, "simt_emu: %s (block %u, thread %u)\n", msg, g_blk ? g_blk->bidx.x : 0u, g_cur ? g_cur->tidx.x : 0u);
    abort();
}

enum { K_SYNC, K_SHFL, K_BALLOT, K_MATCH, K_REDUCE_OR };

inline uint64_t collective(int kind, uint64_t v, uint64_t v2, void (*complete)(Warp &, uint64_t *))
{
    Warp &w = g_blk->warps[g_cur->warp];
    const int lane = g_cur->lane;
    const unsigned gen = w.gen;
    if (w.arrived == 0) { w.kind = kind; w.complete = complete; }
    else if (w.kind != kind) die("lanes of one warp met at different kinds of collectives");
    w.in[lane] = v; w.in2[lane] = v2; w.present |= 1u << lane;
    if (++w.arrived == w.live) {
        complete(w, w.out[gen & 1]);
        w.arrived = 0; w.present = 0; w.gen = gen + 1; g_blk->progress++;
    } else {
        while (w.gen == gen) yield();
    }
    return w.out[gen & 1][lane];
}

template <class T> inline uint64_t to_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

inline void check_mask(unsigned mask) { if (mask != 0xffffffffu) die("only full-mask collectives are emulated"); }

inline void complete_shfl(Warp &w, uint64_t *out)
{
    for (int l = 0; l < 32; ++l) {
        if (!((w.present >> l) & 1u)) continue;
        const int src = (int)(int64_t)w.in2[l];
        out[l] = (src >= 0 && src < 32 && ((w.present >> src) & 1u)) ? w.in[src] : w.in[l];
    }
}

template <class T> inline T shfl_from(T v, int src)
{
    return from_bits<T>(collective(K_SHFL, to_bits(v), (uint64_t)(int64_t)src, complete_shfl));
}

void launch(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()> &body);

// the mbarrier object the kernels keep in shared memory (8 bytes).  Bulk copies are done at issue, so the
// transaction count is not modelled: expect_tx is an arrival.
struct MBar { uint16_t count, pending; uint32_t phase; };
inline void mbar_init(void *bar, unsigned count) { MBar *m = (MBar *)bar; m->count = m->pending = (uint16_t)count; m->phase = 0; }
inline void mbar_arrive(void *bar)
{
    MBar *m = (MBar *)bar;
    if (m->pending == 0) die("mbarrier: more arrivals than expected");
    if (--m->pending == 0) { m->pending = m->count; m->phase ^= 1u; g_blk->progress++; }
}
inline bool mbar_test(void *bar, unsigned parity) { return (((MBar *)bar)->phase & 1u) != (parity & 1u); }
inline void mbar_wait(void *bar, unsigned parity) { while (!mbar_test(bar, parity)) yield(); }

}  // namespace simt

// ---- CUDA spellings -------------------------------------------------------------------------------
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __grid_constant__
#define __grid_constant__
#define threadIdx (simt::g_cur->tidx)
#define blockIdx (simt::g_blk->bidx)
#define blockDim (simt::g_blk->bdim)
#define gridDim (simt::g_blk->gdim)

inline void __syncwarp(unsigned mask = 0xffffffffu)
{
    simt::check_mask(mask);
    simt::collective(simt::K_SYNC, 0, 0, [](simt::Warp &, uint64_t *) {});
}
template <class T> inline T __shfl_sync(unsigned mask, T v, int src) { simt::check_mask(mask); return simt::shfl_from(v, src & 31); }
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int off) { simt::check_mask(mask); return simt::shfl_from(v, simt::g_cur->lane ^ off); }
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned d) { simt::check_mask(mask); return simt::shfl_from(v, simt::g_cur->lane - (int)d); }
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned d)
{
    simt::check_mask(mask);
    const int s = simt::g_cur->lane + (int)d;
    return simt::shfl_from(v, s < 32 ? s : simt::g_cur->lane);
}
inline unsigned __ballot_sync(unsigned mask, int pred)
{
    simt::check_mask(mask);
    return (unsigned)simt::collective(simt::K_BALLOT, pred ? 1 : 0, 0, [](simt::Warp &w, uint64_t *out) {
        unsigned b = 0;
        for (int l = 0; l < 32; ++l) if (((w.present >> l) & 1u) && w.in[l]) b |= 1u << l;
        for (int l = 0; l < 32; ++l) out[l] = b;
    });
}
inline unsigned __reduce_or_sync(unsigned mask, unsigned v)
{
    simt::check_mask(mask);
    return (unsigned)simt::collective(simt::K_REDUCE_OR, v, 0, [](simt::Warp &w, uint64_t *out) {
        unsigned b = 0;
        for (int l = 0; l < 32; ++l) if ((w.present >> l) & 1u) b |= (unsigned)w.in[l];
        for (int l = 0; l < 32; ++l) out[l] = b;
    });
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0u; }
template <class T> inline unsigned __match_any_sync(unsigned mask, T v)
{
    simt::check_mask(mask);
    return (unsigned)simt::collective(simt::K_MATCH, simt::to_bits(v), 0, [](simt::Warp &w, uint64_t *out) {
        for (int l = 0; l < 32; ++l) {
            if (!((w.present >> l) & 1u)) continue;
            unsigned m = 0;
            for (int j = 0; j < 32; ++j) if (((w.present >> j) & 1u) && w.in[j] == w.in[l]) m |= 1u << j;
            out[l] = m;
        }
    });
}
inline int simt_block_barrier(int pred, bool count)
{
    simt::Block &b = *simt::g_blk;
    const unsigned gen = b.bar_gen;
    b.bar_acc |= pred ? 1 : 0; b.bar_cnt += pred ? 1 : 0;
    if (++b.bar_arrived == b.live) {
        b.bar_res[gen & 1] = b.bar_acc; b.bar_rescnt[gen & 1] = b.bar_cnt;
        b.bar_acc = 0; b.bar_cnt = 0; b.bar_arrived = 0; b.bar_gen = gen + 1; b.progress++;
    } else while (b.bar_gen == gen) simt::yield();
    return count ? b.bar_rescnt[gen & 1] : b.bar_res[gen & 1];
}
inline void __syncthreads() { simt_block_barrier(0, false); }
inline int __syncthreads_or(int pred) { return simt_block_barrier(pred, false); }
inline int __syncthreads_count(int pred) { return simt_block_barrier(pred, true); }
// bar.sync id, count
inline void simt_named_barrier(int id, int count)
{
    simt::Block &b = *simt::g_blk;
    const unsigned gen = b.nb_gen[id];
    if (++b.nb_arrived[id] == count) { b.nb_arrived[id] = 0; b.nb_gen[id] = gen + 1; b.progress++; }
    else while (b.nb_gen[id] == gen) simt::yield();
}

inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * (uint64_t)b) >> 32); }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { return (unsigned)(((((uint64_t)hi) << 32) | lo) >> (s & 31)); }
inline long long __double2ll_rn(double x) { return llrint(x); }   // default rounding mode: to nearest even
inline double __longlong_as_double(long long x) { return simt::from_bits<double>((uint64_t)x); }
inline long long __double_as_longlong(double x) { return (long long)simt::to_bits(x); }
template <class T> inline T __ldcg(const T *p) { return *p; }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __nanosleep(unsigned) { simt::yield(); }
inline long long clock64() { return (long long)++simt::g_clock; }
template <class T> inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <class T> inline T atomicAnd(T *p, T v) { T o = *p; *p = o & v; return o; }
template <class T> inline T atomicMin(T *p, T v) { T o = *p; *p = v < o ? v : o; return o; }
template <class T> inline T atomicMax(T *p, T v) { T o = *p; *p = v > o ? v : o; return o; }
template <class T> inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)p; }
