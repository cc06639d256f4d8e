// cats_emu.cpp -- the period kernel's own source (r_mabm_b200/csrc/*.cuh) compiled for the CPU SIMT
// emulation.  TESTS ONLY: the pointers in cats_step_args are host pointers here.
#include "simt_emu.h"

#include "../../r_mabm_b200/csrc/cats_aux.cuh"
#include "../../r_mabm_b200/csrc/cats_period_mw.cuh"

using namespace cats;

namespace {

template <int ZM, int G, bool MASKED, int MODE, bool STRICT>
void run(const KArgs &a)
{
    const unsigned grid = (unsigned)((a.B + G - 1) / G);
    simt::launch(grid, 32u * G, (size_t)a.slice * G, [&]() { cats_period_kernel<ZM, G, 1, MASKED, MODE, STRICT>(a); });
}

template <int ZM, int NW, int MODE>
void run_mw(const KArgs &a, bool pipe)
{
    const size_t smem = (size_t)a.slice + MWLayout<NW>::total(a.L.F);
    if (pipe) simt::launch((unsigned)a.B, 32u * NW, smem, [&]() { cats_period_mw_kernel<ZM, NW, 1, MODE, true>(a); });
    else simt::launch((unsigned)a.B, 32u * NW, smem, [&]() { cats_period_mw_kernel<ZM, NW, 1, MODE, false>(a); });
}

}  // namespace

extern "C" {

// the multi-warp kernel: nw warps per economy; mode 1 (any size) or 2 (default configuration)
int emu_step_mw(const cats_step_args *s, int mode, int nw, int pipe)
{
    KArgs a;
    fill_kargs(s, a);
    if (s->d_tape || s->stop_phase || s->d_debug || s->d_active || (s->flags & CATS_FLAG_STRICT)) return -1;
    if (mode == 2) {
        if (s->W != DefaultLayout::kW || s->F != DefaultLayout::kF || s->N != DefaultLayout::kN) return -1;
        if (nw == 2) run_mw<5, 2, 2>(a, pipe); else if (nw == 4) run_mw<5, 4, 2>(a, pipe); else run_mw<5, 8, 2>(a, pipe);
    } else {
        if (nw == 2) run_mw<MAX_Z, 2, 1>(a, pipe); else if (nw == 4) run_mw<MAX_Z, 4, 1>(a, pipe); else run_mw<MAX_Z, 8, 1>(a, pipe);
    }
    return 0;
}


void emu_set_schedule(int s) { simt::g_schedule = s; }

// development counters (builds with -DCATS_EMU_COUNT): copies them out and clears them
void emu_counts(long long *out) { for (int i = 0; i < 8; ++i) { out[i] = simt::g_count[i]; simt::g_count[i] = 0; } }

int emu_init(unsigned char *blobs, long long B, int W, int F, int N, const double *params, long long params_stride,
             const unsigned char *active)
{
    const Layout L = make_layout(W, F, N);
    simt::launch((unsigned)B, 256, 0, [&]() { cats_init_kernel(blobs, B, L, params, params_stride, active); });
    return 0;
}

// mode: 0 debug/tape, 1 production, 2 default-configuration; g: economies per block (1 or 2)
int emu_step(const cats_step_args *s, int mode, int g)
{
    KArgs a;
    fill_kargs(s, a);
    const bool strict = (s->flags & CATS_FLAG_STRICT) != 0, masked = s->d_active != nullptr;
    if (mode == 2 && (s->W != DefaultLayout::kW || s->F != DefaultLayout::kF || s->N != DefaultLayout::kN)) return -1;
    if (mode != 0 && (s->d_tape || s->stop_phase || s->d_debug)) return -1;
    if (masked) { if (mode != 1 || strict) return -1; if (g == 2) run<MAX_Z, 2, true, 1, false>(a); else run<MAX_Z, 1, true, 1, false>(a); return 0; }
    if (strict) {
        if (mode == 0) run<MAX_Z, 1, false, 0, true>(a);
        else if (mode == 2) run<5, 1, false, 2, true>(a);
        else run<MAX_Z, 2, false, 1, true>(a);
        return 0;
    }
    if (mode == 0) { if (g == 2) run<MAX_Z, 2, false, 0, false>(a); else run<MAX_Z, 1, false, 0, false>(a); }
    else if (mode == 1) { if (g == 2 || s->max_z <= 0 || s->max_z > 5) run<MAX_Z, 2, false, 1, false>(a); else run<5, 1, false, 1, false>(a); }
    else { if (g == 2) run<5, 2, false, 2, false>(a); else run<5, 1, false, 2, false>(a); }
    return 0;
}

}  // extern "C"
