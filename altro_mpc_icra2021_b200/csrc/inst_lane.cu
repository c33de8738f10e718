// Lane-per-instance solve kernels (altro_lane.cuh) for the small-dimension families: rocket (6, 3), grasp (6, 6).
#include "altro_lane.cuh"

namespace altro {
const void *lane_kernel(int n, int m)
{
    if (n == 6 && m == 3) return (const void *)altro_lane_kernel<6, 3, 1>;
    if (n == 6 && m == 6) return (const void *)altro_lane_kernel<6, 6, 1>;
    return nullptr;
}
}  // namespace altro
