// MODE 0 instantiations of the period kernel: replay tape, debug stop after a phase, debug image.
#include "cats_period.cuh"
namespace cats {
KernelFn kernel_debug(int max_z, int g)
{
    return (max_z > 0 && max_z <= 5) ? select_group<5, false, 0>(g) : select_group<MAX_Z, false, 0>(g);
}
}  // namespace cats
