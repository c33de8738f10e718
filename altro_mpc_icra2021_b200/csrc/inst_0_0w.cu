// Run-time sized solve kernel with the whole register file (one CTA per SM): the TMA-staged large-dimension layout.
#include "altro_kernels.cuh"

namespace altro {
const void *kernel_0_0_wide(int T)
{
    if (T == 256) return (const void *)altro_solve_kernel_wide<256>;
    return nullptr;
}
}  // namespace altro
