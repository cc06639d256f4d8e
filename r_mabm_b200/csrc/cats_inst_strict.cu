// CATS_FLAG_STRICT instantiations of the period kernel (plain arithmetic, docs/MODEL.md 4.6).
#include "cats_period.cuh"
namespace cats {
KernelFn kernel_strict(int mode, int g)
{
    if (mode == 2) return select_group<5, false, 2, true>(g);
    if (mode == 1) return select_group<MAX_Z, false, 1, true>(g);
    return select_group<MAX_Z, false, 0, true>(g);
}
}  // namespace cats
