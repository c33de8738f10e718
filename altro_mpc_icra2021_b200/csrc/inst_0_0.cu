// Solve-kernel instantiations for n=0, m=0 (run-time dimensions); one translation unit per dimension pair
// so the library builds in parallel.
#include "altro_kernels.cuh"

namespace altro {
const void *kernel_0_0(int T)
{
    switch (T) {
    case 32: return (const void *)altro_solve_kernel<0, 0, 32>;
    case 64: return (const void *)altro_solve_kernel<0, 0, 64>;
    case 128: return (const void *)altro_solve_kernel<0, 0, 128>;
    case 256: return (const void *)altro_solve_kernel<0, 0, 256>;
    }
    return nullptr;
}
}  // namespace altro
