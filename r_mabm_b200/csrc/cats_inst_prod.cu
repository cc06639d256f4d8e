// MODE 1 instantiations of the period kernel: production path for economies of any size.
#include "cats_period.cuh"
namespace cats {
KernelFn kernel_prod(int max_z, int g)
{
    return (max_z > 0 && max_z <= 5) ? select_group<5, false, 1>(g) : select_group<MAX_Z, false, 1>(g);
}
}  // namespace cats
