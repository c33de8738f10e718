// Solve-kernel instantiations for n=6, m=6 (grasp); one translation unit per dimension pair
// so the library builds in parallel.
#include "altro_kernels.cuh"

namespace altro {
const void *kernel_6_6(int T)
{
    switch (T) {
    case 32: return (const void *)altro_solve_kernel<6, 6, 32>;
    case 64: return (const void *)altro_solve_kernel<6, 6, 64>;
    case 128: return (const void *)altro_solve_kernel<6, 6, 128>;
    case 256: return (const void *)altro_solve_kernel<6, 6, 256>;
    }
    return nullptr;
}
}  // namespace altro
