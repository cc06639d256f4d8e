// Instantiations of the multi-warp period kernel (one block = one economy, NW warps): cats_period_mw.cuh.
#include "cats_period_mw.cuh"
namespace cats {
template <int ZM, int MODE>
static KernelFn mw_pick(int nw)
{
    switch (nw) {
    case 2: return cats_period_mw_kernel<ZM, 2, 14, MODE>;
    case 4: return cats_period_mw_kernel<ZM, 4, 7, MODE>;
    default: return cats_period_mw_kernel<ZM, 8, 3, MODE>;
    }
}
KernelFn kernel_mw(int mode, int max_z, int nw)
{
    if (mode == 2) return mw_pick<5, 2>(nw);
    return (max_z > 0 && max_z <= 5) ? mw_pick<5, 1>(nw) : mw_pick<MAX_Z, 1>(nw);
}
int mw_extra_bytes(int nw, int F) { return nw == 2 ? MWLayout<2>::total(F) : nw == 4 ? MWLayout<4>::total(F) : MWLayout<8>::total(F); }
}  // namespace cats
