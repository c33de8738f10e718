// Solve-kernel instantiations for n=12, m=3 (flexible satellite); one translation unit per dimension pair
// so the library builds in parallel.
#include "altro_kernels.cuh"

namespace altro {
const void *kernel_12_3(int T)
{
    switch (T) {
    case 32: return (const void *)altro_solve_kernel<12, 3, 32>;
    case 64: return (const void *)altro_solve_kernel<12, 3, 64>;
    case 128: return (const void *)altro_solve_kernel<12, 3, 128>;
    case 256: return (const void *)altro_solve_kernel<12, 3, 256>;
    }
    return nullptr;
}
}  // namespace altro
