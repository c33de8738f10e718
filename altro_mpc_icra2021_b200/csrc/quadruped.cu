// quadruped.cu -- the step immediately BEFORE the solve path of the quadruped benchmark, on the device (SURVEY.md 8f
// row f3):  per control tick the reference builds, on the host and with ForwardDiff,
//   * the contact pattern and the world foot positions of every knot      (footsteps.jl:29-84, gait.jl:1-15)
//   * A_k = I + A_c dt, B_k = B_c dt, d_k = d_c dt  with  A_c, B_c the Jacobians of the single-rigid-body dynamics at
//     (x_ref_k, u_ref_k) and d_c = f - A_c x_ref - B_c u_ref                 (linearized_dynamics.jl:1-66, altro_solver.jl:5-42)
// and writes them into the solver's model in place.  Here both run as kernels that write straight into the handle's
// per-instance, per-knot dynamics (altro_set_dynamics layout), so a control tick needs no host traffic at all.
//
// The Jacobians are taken the way the reference takes them -- forward-mode automatic differentiation of the
// nonlinear dynamics (a dual number with one tangent, one sweep per input direction) -- at a general orientation
// (MRP != 0) and a general u_ref, not only at the benchmark's rot = I point.
#include <cuda_runtime.h>

#include <string>

#include "altro_quadruped.cuh"

namespace altro {

namespace {

struct Dual {
    double v, d;
};
__device__ __forceinline__ Dual mk(double v, double d = 0.0) { return Dual{v, d}; }
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return Dual{a.v + b.v, a.d + b.d}; }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return Dual{a.v - b.v, a.d - b.d}; }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return Dual{a.v * b.v, a.d * b.v + a.v * b.d}; }
__device__ __forceinline__ Dual operator*(double a, Dual b) { return Dual{a * b.v, a * b.d}; }
__device__ __forceinline__ Dual operator/(Dual a, Dual b)
{
    const double q = a.v / b.v;
    return Dual{q, (a.d - q * b.d) / b.v};
}

// Rotation matrix of a modified Rodrigues parameter vector (Rotations.jl MRP -> 3 x 3, body to world):
//   R = I + (8 [p]x^2 + 4 (1 - p'p) [p]x) / (1 + p'p)^2
template <class S>
__device__ void mrp_rot(const S p[3], S R[9])
{
    const S n2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
    const S den = (mk(1.0) + n2) * (mk(1.0) + n2);
    const S a = mk(8.0) / den, b = 4.0 * (mk(1.0) - n2) / den;
    // [p]x^2 = p p' - (p'p) I
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i * 3 + j] = a * (p[i] * p[j]) - ((i == j) ? a * n2 : mk(0.0));
    R[0] = R[0] + mk(1.0); R[4] = R[4] + mk(1.0); R[8] = R[8] + mk(1.0);
    R[1] = R[1] - b * p[2]; R[2] = R[2] + b * p[1];
    R[3] = R[3] + b * p[2]; R[5] = R[5] - b * p[0];
    R[6] = R[6] - b * p[1]; R[7] = R[7] + b * p[0];
}

// NonLinearContinuousDynamics (linearized_dynamics.jl:1-36): x = (p, MRP, v, omega_body), u = 4 world-frame foot forces.
template <class S>
__device__ void srb_dynamics(const S x[12], const S u[12], const double *r /*[4][3] world*/, const double *c /*[4]*/,
                             const double *J /*[9]*/, const double *Jinv /*[9]*/, double mass, S out[12])
{
    S R[9];
    mrp_rot(x + 3, R);
    const S *p = x, *ph = x + 3, *v = x + 6, *w = x + 9;
    for (int i = 0; i < 3; ++i) out[i] = v[i];
    // Rotations.kinematics(MRP, omega) = 1/4 ((1 - p'p) w + 2 p x w + 2 (p'w) p)
    const S n2 = ph[0] * ph[0] + ph[1] * ph[1] + ph[2] * ph[2], pw = ph[0] * w[0] + ph[1] * w[1] + ph[2] * w[2];
    const S cr[3] = {ph[1] * w[2] - ph[2] * w[1], ph[2] * w[0] - ph[0] * w[2], ph[0] * w[1] - ph[1] * w[0]};
    for (int i = 0; i < 3; ++i) out[3 + i] = 0.25 * ((mk(1.0) - n2) * w[i] + 2.0 * cr[i] + 2.0 * (pw * ph[i]));
    S fs[3] = {mk(0.0), mk(0.0), mk(-9.81)}, ts[3] = {mk(0.0), mk(0.0), mk(0.0)};
    for (int f = 0; f < 4; ++f) {
        const S *uf = u + 3 * f;
        for (int i = 0; i < 3; ++i) fs[i] = fs[i] + (c[f] / mass) * uf[i];
        // r_b = R' (r - p);  torque += c [r_b]x R' u
        S a[3], rb[3], ub[3];
        for (int i = 0; i < 3; ++i) a[i] = mk(r[3 * f + i]) - p[i];
        for (int i = 0; i < 3; ++i) {
            rb[i] = R[0 * 3 + i] * a[0] + R[1 * 3 + i] * a[1] + R[2 * 3 + i] * a[2];
            ub[i] = R[0 * 3 + i] * uf[0] + R[1 * 3 + i] * uf[1] + R[2 * 3 + i] * uf[2];
        }
        ts[0] = ts[0] + c[f] * (rb[1] * ub[2] - rb[2] * ub[1]);
        ts[1] = ts[1] + c[f] * (rb[2] * ub[0] - rb[0] * ub[2]);
        ts[2] = ts[2] + c[f] * (rb[0] * ub[1] - rb[1] * ub[0]);
    }
    for (int i = 0; i < 3; ++i) out[6 + i] = fs[i];
    // inv(J) (-w x (J w) + torque)
    S Jw[3];
    for (int i = 0; i < 3; ++i) Jw[i] = J[i * 3 + 0] * w[0] + J[i * 3 + 1] * w[1] + J[i * 3 + 2] * w[2];
    const S g[3] = {ts[0] - (w[1] * Jw[2] - w[2] * Jw[1]), ts[1] - (w[2] * Jw[0] - w[0] * Jw[2]), ts[2] - (w[0] * Jw[1] - w[1] * Jw[0])};
    for (int i = 0; i < 3; ++i) out[9 + i] = Jinv[i * 3 + 0] * g[0] + Jinv[i * 3 + 1] * g[1] + Jinv[i * 3 + 2] * g[2];
}

// One thread per (instance, knot): 24 tangent sweeps + the value -> A_k, B_k, d_k (update_dynamics_matrices!).
__global__ void quadruped_linearize_kernel(int B, int K, const double *xref, int xref_per_knot, const double *uref,
                                           int uref_per_knot, const double *foot, const double *contacts,
                                           QuadrupedBody body, double dt, double *A, double *Bm, double *d)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * K) return;
    const int b = t / K, k = t - b * K;
    const double *xr = xref + ((size_t)b * (xref_per_knot ? K : 1) + (xref_per_knot ? k : 0)) * 12;
    const double *ur = uref ? uref + ((size_t)b * (uref_per_knot ? K : 1) + (uref_per_knot ? k : 0)) * 12 : nullptr;
    const double *r = foot + (size_t)t * 12, *c = contacts + (size_t)t * 4;
    double *Ak = A + (size_t)t * 144, *Bk = Bm + (size_t)t * 144, *dk = d + (size_t)t * 12;
    Dual x[12], u[12], f[12];
    for (int i = 0; i < 12; ++i) { x[i] = mk(xr[i]); u[i] = mk(ur ? ur[i] : 0.0); }
    double acc[12];
    srb_dynamics(x, u, r, c, body.J, body.Jinv, body.mass, f);
    for (int i = 0; i < 12; ++i) acc[i] = f[i].v;
    for (int j = 0; j < 24; ++j) {
        Dual &s = j < 12 ? x[j] : u[j - 12];
        s.d = 1.0;
        srb_dynamics(x, u, r, c, body.J, body.Jinv, body.mass, f);
        s.d = 0.0;
        const double zj = s.v;
        for (int i = 0; i < 12; ++i) {
            const double g = f[i].d;  // d f_i / d z_j
            acc[i] -= g * zj;         // d_c = f - A_c x_ref - B_c u_ref
            if (j < 12) Ak[i * 12 + j] = ((i == j) ? 1.0 : 0.0) + g * dt;
            else Bk[i * 12 + (j - 12)] = g * dt;
        }
    }
    for (int i = 0; i < 12; ++i) dk[i] = acc[i] * dt;
}

// foot_history! (footsteps.jl:29-84) with get_phase (gait.jl:1-9) and footstep_location (footsteps.jl:1-27), one thread
// per instance: contact pattern and world foot positions of knots 0 .. K-1 at times t, t + dt, ...
__global__ void quadruped_gait_kernel(int B, int K, const double *tnow, const double *xref, int xref_per_knot,
                                      const double *cur_foot /*[B][4][3] body frame*/, QuadrupedGait g, double dt,
                                      double *planner /*[B][4][3] in/out*/, double *contacts, double *foot)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    auto phase_of = [&](double t) {
        const double pt = fmod(t, g.phase_length);
        double s = 0.0;
        for (int i = 0; i < g.num_phases; ++i) {
            s += g.phase_times[i];
            if (pt < s) return i;
        }
        return g.num_phases - 1;
    };
    auto xr = [&](int k) { return xref + ((size_t)b * (xref_per_knot ? K : 1) + (xref_per_knot ? k : 0)) * 12; };
    double R[9], prev[12], plan[12];
    double t_i = tnow[b] + dt;
    int prev_phase = phase_of(tnow[b]);
    {
        const double *x = xr(0);
        mrp_rot_d(x + 3, R);
        for (int j = 0; j < 4; ++j)
            for (int i = 0; i < 3; ++i) {
                const double *cf = cur_foot + ((size_t)b * 4 + j) * 3;
                prev[3 * j + i] = x[i] + R[i * 3 + 0] * cf[0] + R[i * 3 + 1] * cf[1] + R[i * 3 + 2] * cf[2];
            }
    }
    for (int i = 0; i < 12; ++i) plan[i] = planner[(size_t)b * 12 + i];
    for (int j = 0; j < 4; ++j) contacts[((size_t)b * K + 0) * 4 + j] = g.contact[prev_phase * 4 + j];
    for (int i = 0; i < 12; ++i) foot[((size_t)b * K + 0) * 12 + i] = prev[i];
    for (int k = 1; k < K; ++k) {
        const int next_phase = phase_of(t_i);
        const double *x = xr(k);
        mrp_rot_d(x + 3, R);
        for (int j = 0; j < 4; ++j) {
            contacts[((size_t)b * K + k) * 4 + j] = g.contact[next_phase * 4 + j];
            if (g.contact[prev_phase * 4 + j] == 1.0) {
                if (g.contact[next_phase * 4 + j] == 0.0) {
                    // footstep_location: nominal stance under the body + alpha * t_next * v (k = 0, the third term of
                    // the reference's expression is a dangling statement and is not added), projected to the ground
                    const int np = next_phase + 1 == g.num_phases ? 0 : next_phase + 1;
                    const double t_next = g.phase_times[np];
                    const double *nf = g.nom_foot + 3 * j;
                    for (int i = 0; i < 2; ++i)
                        plan[3 * j + i] = x[i] + R[i * 3 + 0] * nf[0] + R[i * 3 + 1] * nf[1] + R[i * 3 + 2] * nf[2] +
                                          g.alpha * t_next * x[6 + i];
                    plan[3 * j + 2] = g.foot_radius;
                }
            } else if (g.contact[next_phase * 4 + j] == 1.0) {
                for (int i = 0; i < 3; ++i) prev[3 * j + i] = plan[3 * j + i];
            }
        }
        for (int i = 0; i < 12; ++i) foot[((size_t)b * K + k) * 12 + i] = prev[i];
        t_i += dt;
        prev_phase = next_phase;
    }
    for (int i = 0; i < 12; ++i) planner[(size_t)b * 12 + i] = plan[i];
}

}  // namespace

cudaError_t quadruped_linearize_launch(int B, int K, const double *xref, int xref_per_knot, const double *uref,
                                       int uref_per_knot, const double *foot, const double *contacts,
                                       const QuadrupedBody &body, double dt, double *A, double *Bm, double *d,
                                       cudaStream_t stream)
{
    const int threads = 128, blocks = (B * K + threads - 1) / threads;
    quadruped_linearize_kernel<<<blocks, threads, 0, stream>>>(B, K, xref, xref_per_knot, uref, uref_per_knot, foot,
                                                               contacts, body, dt, A, Bm, d);
    return cudaGetLastError();
}

cudaError_t quadruped_gait_launch(int B, int K, const double *tnow, const double *xref, int xref_per_knot,
                                  const double *cur_foot, const QuadrupedGait &g, double dt, double *planner,
                                  double *contacts, double *foot, cudaStream_t stream)
{
    const int threads = 128, blocks = (B + threads - 1) / threads;
    quadruped_gait_kernel<<<blocks, threads, 0, stream>>>(B, K, tnow, xref, xref_per_knot, cur_foot, g, dt, planner,
                                                          contacts, foot);
    return cudaGetLastError();
}

}  // namespace altro
