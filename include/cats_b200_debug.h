/* cats_b200_debug.h -- test and profiling hooks of libcats_b200.so.
 *
 * NOT part of the drop-in boundary (include/cats_b200.h): nothing here replaces a reference interface.  Every hook is a
 * process-wide setting (an atomic inside the library) that applies to LATER cats_step / cats_step_host calls of every
 * engine in the process; none of them changes results -- tests/ use them to run every kernel shape and data path
 * against the oracle, scripts/ to A/B them.  A product caller never needs them: the library picks per launch. */
#ifndef CATS_B200_DEBUG_H
#define CATS_B200_DEBUG_H
#include "cats_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Debug/profiling aid: when set to a device array of 16 int64, block 0 of every later cats_step
 * adds its clock64() cycles per phase group into it (slots: 0 firm decisions, 1 fire, 2 job search,
 * 3 produce, 4 capital market, 5 income+budget, 6 goods market, 7 accounting, 8 reward, 9 tracking,
 * 10 bankruptcy, 11 closing).  Pass NULL to switch it off. */
void cats_debug_phase_cycles(void *d_cycles16);

/* Debug aid: economies per thread block for later cats_step calls (1, 2, 4, 8, 12 or 14; 0 = chosen per
 * launch from batch size and working set, the default).  Results never depend on it: the warps of a
 * block share nothing but their start time; tests/ run every shape. */
void cats_debug_set_group(int economies_per_block);

/* Debug aid: economies of the reference's default size (W=1000, F=100, N=20) with all z <= 5 run a
 * kernel instantiation whose state layout is a compile-time constant (addresses are base + immediate,
 * loop bounds are known: ~20 % fewer instructions; it also needs z_all == {5, 2, 5}).  0 makes later cats_step calls use the generic
 * kernel for them too; 1 (default) restores the specialisation.  Results never depend on it. */
void cats_debug_set_static(int on);

/* Debug aid: warps per economy for later cats_step calls: 1 = the one-warp-per-economy kernel, 2 / 4 / 8 = the
 * multi-warp kernel (one thread block per economy, for launches that would leave the GPU under-filled), 0 = chosen per
 * launch from the batch size (default).  Results never depend on it. */
void cats_debug_set_warps(int warps_per_economy);

/* Debug aid: the goods market of the multi-warp kernel: 1 = producer / consumer pipeline (one warp commits with the
 * one-warp kernel's rounds, the others prepare households ahead of it), 0 = block-wide market, -1 = chosen per launch
 * (the pipeline; default).  Results never depend on it. */
void cats_debug_set_pipe(int mode);

/* Debug aid: how cats_step_host treats page-locked host buffers.  Bit 0: the kernel reads the inputs in place; bit 1:
 * it writes the outputs in place.  2 is the default (measured fastest); 0 stages both directions through d_scratch
 * like pageable memory.  Results never depend on it. */
void cats_debug_set_host_zero_copy(int mode);

#ifdef __cplusplus
}
#endif
#endif /* CATS_B200_DEBUG_H */
