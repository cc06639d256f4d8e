// Instantiations of the multi-warp period kernel (one block = one economy, NW warps): cats_period_mw.cuh.
#include "cats_period_mw.cuh"
namespace cats {
template <int ZM, int MODE, bool PIPE>
static KernelFn mw_pick(int nw)
{
    switch (nw) {
    case 2: return cats_period_mw_kernel<ZM, 2, 14, MODE, PIPE>;
    case 4: return cats_period_mw_kernel<ZM, 4, 7, MODE, PIPE>;
    default: return cats_period_mw_kernel<ZM, 8, 3, MODE, PIPE>;
    }
}
// pipe: the goods market as a producer / consumer pipeline (few sellers) instead of the block-wide market (many)
KernelFn kernel_mw(int mode, int max_z, int nw, bool pipe)
{
    if (mode == 2) return pipe ? mw_pick<5, 2, true>(nw) : mw_pick<5, 2, false>(nw);
    if (max_z > 0 && max_z <= 5) return pipe ? mw_pick<5, 1, true>(nw) : mw_pick<5, 1, false>(nw);
    return pipe ? mw_pick<MAX_Z, 1, true>(nw) : mw_pick<MAX_Z, 1, false>(nw);
}
int mw_extra_bytes(int nw, int F) { return nw == 2 ? MWLayout<2>::total(F) : nw == 4 ? MWLayout<4>::total(F) : MWLayout<8>::total(F); }
}  // namespace cats
