// altro_quadruped.cuh -- parameters of the quadruped pre-solve kernels (quadruped.cu).
#pragma once
#include <cuda_runtime.h>

namespace altro {

struct QuadrupedBody {
    double J[9], Jinv[9], mass;
};

struct QuadrupedGait {
    int num_phases;
    double contact[8 * 4];  // [phase][foot], 1.0 = stance
    double phase_times[8], phase_length, alpha, foot_radius;
    double nom_foot[12];    // body-frame foot positions at zero joint angles
};

// plain-double MRP rotation (the dual-number template in quadruped.cu covers the differentiated use)
__device__ inline void mrp_rot_d(const double p[3], double R[9])
{
    const double n2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
    const double den = (1.0 + n2) * (1.0 + n2), a = 8.0 / den, b = 4.0 * (1.0 - n2) / den;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i * 3 + j] = a * (p[i] * p[j]) - ((i == j) ? a * n2 : 0.0);
    R[0] += 1.0; R[4] += 1.0; R[8] += 1.0;
    R[1] -= b * p[2]; R[2] += b * p[1];
    R[3] += b * p[2]; R[5] -= b * p[0];
    R[6] -= b * p[1]; R[7] += b * p[0];
}

cudaError_t quadruped_linearize_launch(int B, int K, const double *xref, int xref_per_knot, const double *uref,
                                       int uref_per_knot, const double *foot, const double *contacts,
                                       const QuadrupedBody &body, double dt, double *A, double *Bm, double *d,
                                       cudaStream_t stream);
cudaError_t quadruped_gait_launch(int B, int K, const double *tnow, const double *xref, int xref_per_knot,
                                  const double *cur_foot, const QuadrupedGait &g, double dt, double *planner,
                                  double *contacts, double *foot, cudaStream_t stream);

}  // namespace altro
