// MODE 2 instantiations (default-size economies, compile-time layout) and the masked ones (resets).
#include "cats_period.cuh"
namespace cats {
KernelFn kernel_static(int g) { return select_group<5, false, 2>(g); }
KernelFn kernel_masked(int g) { return select_group<MAX_Z, true, 1>(g); }
}  // namespace cats
