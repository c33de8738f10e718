// Lane-per-instance solve kernels (altro_lane.cuh) for the small-dimension families: rocket (6, 3), grasp (6, 6).
#include "altro_lane.cuh"

namespace altro {

namespace {
template <int NX, int NU, int LPW>
cudaError_t launch(const LaneLaunch &a)
{
    LaneConst<NX, NU> C;
    for (int i = 0; i < NX * NX; ++i) C.A[i] = a.A[i];
    for (int i = 0; i < NX * NU; ++i) C.B[i] = a.B[i];
    for (int i = 0; i < NX; ++i) { C.d[i] = a.d[i]; C.Q[i] = a.Q[i]; C.Qf[i] = a.Qf[i]; }
    for (int i = 0; i < NU; ++i) C.R[i] = a.R[i];
    const void *k = (const void *)altro_lane_kernel<NX, NU, LPW>;
    if (a.query) { *a.query = k; return cudaSuccess; }
    const int grid = (a.P->B + LPW - 1) / LPW;
    altro_lane_kernel<NX, NU, LPW><<<grid, 32, a.smem, a.stream>>>(*a.P, C, a.L, a.ws, a.stride, a.scratch_per_lane);
    return cudaGetLastError();
}

template <int NX, int NU>
cudaError_t by_lpw(const LaneLaunch &a)
{
    switch (a.lpw) {
    case 8: return launch<NX, NU, 8>(a);
    case 32: return launch<NX, NU, 32>(a);
    }
    return cudaErrorInvalidValue;
}
}  // namespace

bool lane_supported(int n, int m) { return (n == 6 && m == 3) || (n == 6 && m == 6); }

// Launches (or, with a.query set, only resolves) the lane kernel for (n, m, a.lpw).
cudaError_t lane_launch(int n, int m, const LaneLaunch &a)
{
    if (n == 6 && m == 3) return by_lpw<6, 3>(a);
    if (n == 6 && m == 6) return by_lpw<6, 6>(a);
    return cudaErrorInvalidValue;
}

}  // namespace altro
