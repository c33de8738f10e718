// altro_lane.cuh -- lane-per-instance AL-iLQR (ALTRO) solve for small state / control dimensions.
//
// Same path as altro_kernels.cuh (reference call sites simple_rocket.jl:174,
// grasp_mpc.jl:55, random_linear_problem.jl:161), other mapping: ONE THREAD owns one MPC instance and a warp
// advances 32 instances in lock step.  For n, m <= 8 (rocket 6/3, grasp 6/6) the Riccati blocks of one instance
// fit a thread's registers, so the whole backward pass is straight-line FMA code with no barrier, no shuffle
// and no idle lane; the CTA-per-instance kernel spends most of its issue slots on index arithmetic and
// synchronisation for 6 x 6 blocks (148 k warp instructions per rocket solve, 18.6 of 32 lanes active).
//
// * Trajectories, gains and duals live in a per-handle global workspace laid out element-major
//   ws[element][instance]: every access of a warp is one coalesced 256-byte row, the working set of 4096
//   rocket instances (40 MB) stays in L2.
// * Control flow is a per-lane state machine (step()): every trip of the warp's loop executes at most one
//   backward pass and one line-search trial per lane, and a lane that finishes a solve moves on to its next
//   MPC step on its own.  Lanes therefore desynchronise across iterations, outer loops and MPC steps instead
//   of waiting for the slowest of the 32 at every solve; the dominant phase (backward pass) runs with nearly
//   all lanes active in every trip.
// STATUS (measured on B200, profiles/r2_lane_kernel.md): correct (bit-identical) but about 10x SLOWER than the
// CTA-per-instance kernel on the 4096-instance rocket batch, at every instances-per-warp setting.  With 4096
// instances per GPU there are only 128 .. 1024 such warps, each a serial instruction stream at 0.07 - 0.15 IPC
// (long-scoreboard and fixed-latency stalls, 250 k instructions per trip), and time = serial instructions x latency /
// instances in flight per SM comes out the same whatever the lanes-per-warp split.  Kept as an opt-in alternative
// (altro_set_kernel_mode(h, 2)) and as the host-testable restatement of the solver; never selected automatically.
//
// * Arithmetic is the oracle's, operation for operation (fma chains over ascending index from a stated initial
//   value, canonical 32-partial cost sums), so results are bit-identical to oracle/altro_oracle.c and to the
//   CTA kernel.  Everything per-lane is __host__ __device__: tests/native/lane_host.cu runs the same code on
//   the CPU against the oracle without a GPU.
#pragma once
#include <math.h>

#include "altro_kernels.cuh"

#ifdef __CUDACC__
#define ALTRO_HD __host__ __device__ __forceinline__
#define ALTRO_HDN __host__ __device__ __noinline__
#else
#define ALTRO_HD inline
#define ALTRO_HDN
#endif

namespace altro {

// Element offsets of one instance inside the element-major workspace.
struct LaneLayout {
    int X, U, Xb, Ub, K, dv, lam, xr, ur, total;
};

inline LaneLayout make_lane_layout(int n, int m, int N, int P)
{
    LaneLayout l{};
    int q = 0;
    auto take = [&](int c) { int at = q; q += c; return at; };
    l.X = take(N * n); l.U = take((N - 1) * m); l.Xb = take(N * n); l.Ub = take((N - 1) * m);
    l.K = take((N - 1) * m * n); l.dv = take((N - 1) * m); l.lam = take(P > 0 ? P : 1);
    l.xr = take(N * n); l.ur = take((N - 1) * m);
    l.total = q;
    return l;
}

constexpr int LANE_SCRATCH = 304;  // per-lane doubles of run-time indexed scratch, followed by N (1 + ncon) cost items

// Shared LTI model and diagonal weights, passed BY VALUE as a kernel parameter: every use has a compile-time index,
// so the operands come straight from the constant bank (no load instruction, no register).
template <int NX, int NU>
struct LaneConst {
    double A[NX * NX], B[NX * NU], d[NX], Q[NX], R[NU], Qf[NX];
};

// Arguments of one lane-kernel launch (host side; see inst_lane.cu).
struct LaneLaunch {
    const Params *P;
    const double *A, *B, *d, *Q, *R, *Qf;  // host copies of the shared LTI model and the diagonal weights
    LaneLayout L;
    double *ws;
    size_t stride;
    int lpw, scratch_per_lane;
    size_t smem;
    cudaStream_t stream;
    const void **query;  // non-null: only return the kernel's address (for attribute calls), do not launch
};

template <class Tp>
ALTRO_HD Tp *lane_global(Tp *p)
{
    // (no __builtin_assume(__isGlobal(p)) here: with it nvcc 12.9 generated code that gave wrong results on sm_100a)
    return p;
}
template <class Tp>
ALTRO_HD Tp *lane_shared(Tp *p)
{
    // (see lane_global)
    return p;
}
ALTRO_HD void lane_prefetch(const double *p)
{
#if defined(__CUDA_ARCH__) && defined(ALTRO_LANE_PREFETCH)  // measured: 12 % slower with the prefetches than without
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

enum { LP_STEP_BEGIN = 0, LP_OUTER_BEGIN, LP_BP, LP_TRIAL, LP_POST, LP_OUTER_END, LP_STEP_END, LP_DONE };

// SS: stride of the per-lane scratch = lanes per warp that carry an instance (device), 1 on the host.
template <int NX, int NU, int SS>
struct Lane {
    static constexpr int n = NX, m = NU, ss = SS;
    const Params &P;
    const LaneConst<NX, NU> &C;
    const LaneLayout L;
    double *ws;      // element-major workspace (global memory), this lane's column: element e at ws[e * stride]
    size_t stride;
    double *scr;     // run-time indexed scratch (shared memory), element i at scr[i * SS]
    int inst, N, ncon;
    const ConDesc *cd;  // descriptors (staged in shared memory on the device)
    // solver state
    int phase, st, steps, outer, iters, inner_it, trials, status, dJ_zero, ls_iter, kcur;
    double rho, drho, J, J_prev, Jls, dV1, dV2, alpha, z, cmax, pen_max, ctol, gtol;
    const int *sched;
    size_t dyn_base;
    int dyn_k;
    bool track;
    long long t0;

    // scratch map
    static constexpr int SC_V = 0;     // 32 partial sums
    static constexpr int SC_MU = 32;   // MAX_CON penalties
    static constexpr int SC_G = 48;    // g[w] + H (w*w dense, or w diagonal)
    static constexpr int SC_Y = 128;   // y / lb  [PMAX]
    static constexpr int SC_D = 136;   // D / q   [8]
    static constexpr int SC_EX = 144;  // state side of the knot in flight:   [Qx (8) | Qxx (8 x 8)]
    static constexpr int SC_EU = 216;  // control side of the knot in flight: [Qu (8) | Quu (8 x 8)]
    static constexpr int SC_ZX = 288;  // x_k of the knot in flight (constraint rows read their slice from here)
    static constexpr int SC_ZU = 296;  // u_k
    static constexpr int SC_ITM = 304; // cost items [N (1 + ncon)]

    ALTRO_HD double &W(int e) const { return lane_global(ws)[(size_t)e * stride]; }
    ALTRO_HD double &sc(int i) const { return lane_shared(scr)[i * SS]; }
    ALTRO_HD double &mu(int c) const { return lane_shared(scr)[(SC_MU + c) * SS]; }
    ALTRO_HD void prefetch(int e) const { lane_prefetch(ws + (size_t)e * stride); }

    ALTRO_HD Lane(const Params &P_, const LaneConst<NX, NU> &C_, const LaneLayout &L_, double *ws_col, size_t stride_,
                  double *scr_, const ConDesc *cd_, int inst_)
        : P(P_), C(C_), L(L_), ws(ws_col), stride(stride_), scr(scr_), inst(inst_)
    {
        N = P.N;
        ncon = P.ncon;
        cd = cd_;
        steps = P.steps > 0 ? P.steps : 1;
        st = 0;
        phase = LP_STEP_BEGIN;
        dyn_base = P.dyn_per_instance ? (size_t)inst * (P.dyn_sched ? (size_t)P.dyn_slots : (P.dyn_per_knot ? (size_t)(N - 1) : 1)) : 0;
        dyn_k = P.dyn_per_knot ? 1 : 0;
        sched = nullptr;
        kcur = P.kidx ? P.kidx[inst] : 0;
        track = P.steps > 0 && P.trackX != nullptr;
        set_step(0);
        outer = iters = inner_it = trials = dJ_zero = ls_iter = 0;
        status = ALTRO_UNSOLVED;
        rho = drho = J = J_prev = Jls = dV1 = dV2 = alpha = z = cmax = pen_max = ctol = gtol = 0.0;
        t0 = 0;
    }

    static ALTRO_HD double rcp(double x)
    {
#ifdef __CUDA_ARCH__
        return __drcp_rn(x);
#else
        return 1.0 / x;
#endif
    }
    static ALTRO_HD long long now_ns()
    {
#ifdef __CUDA_ARCH__
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        return t;
#else
        return 0;
#endif
    }

    ALTRO_HD void set_step(int s)
    {
        if (P.dyn_sched) {
            const int a = P.step0 + s, b = P.sched_len - N;
            sched = P.dyn_sched + (size_t)inst * P.sched_len + (a < b ? a : b);
        }
    }
    ALTRO_HD size_t con_idx(const ConDesc &c, int k) const
    {
        if (c.track) {
            const int r = kcur + k;
            return (size_t)(r < c.track - 1 ? r : c.track - 1);
        }
        size_t idx = c.per_instance ? (size_t)inst * (c.per_knot ? (size_t)(c.k1 - c.k0) : 1) : 0;
        return idx + (c.per_knot ? (size_t)(k - c.k0) : 0);
    }
    // reference window: the handle's per-instance copy, or (closed-loop runs) a slice of the padded track
    ALTRO_HD double xref(int k, int i) const
    {
        if (track) {
            const int k0 = P.kidx[inst] + st + 1;
            return P.trackX[(size_t)((k0 < P.Nt ? k0 : P.Nt) + k) * n + i];
        }
        return W(L.xr + k * n + i);
    }
    ALTRO_HD double uref(int k, int i) const
    {
        if (track) {
            const int k0 = P.kidx[inst] + st + 1;
            return P.trackU[(size_t)((k0 < P.Nt - 1 ? k0 : P.Nt - 1) + k) * m + i];
        }
        return W(L.ur + k * m + i);
    }

    // ---------------------------------------------------------------- load / store (instance-major host layout)
    ALTRO_HD void load()
    {
        if (P.steps > 0) {
            const double *gX = P.X + (size_t)inst * N * n;
            for (int i = 0; i < N * n; ++i) W(L.X + i) = gX[i];
        } else {
            const double *gx0 = P.x0 + (size_t)inst * n;
            for (int i = 0; i < n; ++i) W(L.X + i) = gx0[i];
        }
        const double *gU = P.U + (size_t)inst * (N - 1) * m;
        for (int i = 0; i < (N - 1) * m; ++i) W(L.U + i) = gU[i];
        if (!track) {
            const double *gxr = P.xref + (size_t)inst * N * n, *gur = P.uref + (size_t)inst * (N - 1) * m;
            for (int i = 0; i < N * n; ++i) W(L.xr + i) = gxr[i];
            for (int i = 0; i < (N - 1) * m; ++i) W(L.ur + i) = gur[i];
        }
        const double *gl = P.lam + (size_t)inst * P.P;
        const bool rd = P.o.reset_duals != 0;
        for (int i = 0; i < P.P; ++i) W(L.lam + i) = rd ? 0.0 : gl[i];
        for (int c = 0; c < MAX_CON; ++c) mu(c) = P.o.penalty_initial;
        if (P.t_ns) t0 = now_ns();
    }

    ALTRO_HD void store()
    {
        double *gX = P.X + (size_t)inst * N * n, *gU = P.U + (size_t)inst * (N - 1) * m;
        for (int i = 0; i < N * n; ++i) gX[i] = W(L.X + i);
        for (int i = 0; i < (N - 1) * m; ++i) gU[i] = W(L.U + i);
        double *gl = P.lam + (size_t)inst * P.P;
        for (int i = 0; i < P.P; ++i) gl[i] = W(L.lam + i);
        if (P.steps > 0) {
            for (int i = 0; i < n; ++i) P.x0[(size_t)inst * n + i] = W(L.X + i);
            if (track) {
                st = steps - 1;  // the window of the last step
                double *gxr = P.xref + (size_t)inst * N * n, *gur = P.uref + (size_t)inst * (N - 1) * m;
                for (int k = 0; k < N; ++k)
                    for (int i = 0; i < n; ++i) gxr[k * n + i] = xref(k, i);
                for (int k = 0; k < N - 1; ++k)
                    for (int i = 0; i < m; ++i) gur[k * m + i] = uref(k, i);
            }
        }
    }

    // ---------------------------------------------------------------- constraint values and costs (A.2, A.3)
    // x_k and u_k of a trajectory into registers (one batch of independent loads) and into the scratch slice the
    // constraint rows index at run time; the rows of the NEXT knot are prefetched into L1 meanwhile.
    ALTRO_HD void stage_knot(int k, int oXc, int oUc, double (&x)[NX], double (&u)[NU], int knext) const
    {
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] = W(oXc + k * NX + i);
        if (k < N - 1) {
#pragma unroll
            for (int i = 0; i < NU; ++i) u[i] = W(oUc + k * NU + i);
        }
        if (knext >= 0 && knext < N) {
#pragma unroll
            for (int i = 0; i < NX; ++i) prefetch(oXc + knext * NX + i);
            if (knext < N - 1) {
#pragma unroll
                for (int i = 0; i < NU; ++i) prefetch(oUc + knext * NU + i);
            }
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) sc(SC_ZX + i) = x[i];
        if (k < N - 1) {
#pragma unroll
            for (int i = 0; i < NU; ++i) sc(SC_ZU + i) = u[i];
        }
    }

    // Row r of c = G z[inds] + h on the staged knot (TO.evaluate)
    ALTRO_HD double row_value(const ConDesc &c, const double *G, const double *h, int zb, int r) const
    {
        if (c.rowsparse) return fma(lane_global(c.rs_coef)[r], sc(zb + c.inds[lane_global(c.rs_col)[r]]), h[r]);
        double acc = h[r];
        const double *g = G + r * c.w;
        for (int j = 0; j < c.w; ++j) acc = fma(g[j], sc(zb + c.inds[j]), acc);
        return acc;
    }

    // AL penalty term of block ci at the staged knot k
    ALTRO_HD double con_cost(int ci, int k) const
    {
        const ConDesc &c = cd[ci];
        if (k < c.k0 || k >= c.k1) return 0.0;
        const size_t di = con_idx(c, k);
        const double *G = lane_global(c.G) + di * c.p * c.w, *h = lane_global(c.h) + di * c.p;
        const int zb = c.side == ALTRO_STATE ? SC_ZX : SC_ZU;
        const int lo = L.lam + c.dual_off + (k - c.k0) * c.p;
        const double mu_c = mu(ci);
        double Jc = 0.0;
        if (c.sense == ALTRO_EQUALITY) {
            for (int r = 0; r < c.p; ++r) {
                const double v = row_value(c, G, h, zb, r);
                Jc += W(lo + r) * v + 0.5 * mu_c * v * v;
            }
        } else if (c.sense == ALTRO_INEQUALITY) {
            for (int r = 0; r < c.p; ++r) {
                const double v = row_value(c, G, h, zb, r), l = W(lo + r);
                const bool act = (v >= 0.0) || (l > 0.0);
                Jc += l * v + (act ? 0.5 * mu_c * v * v : 0.0);
            }
        } else {
            double a2 = 0.0, t = 0.0, nl = 0.0;
            for (int r = 0; r < c.p; ++r) {
                const double l = W(lo + r);
                const double lb = l - mu_c * row_value(c, G, h, zb, r);
                nl += l * l;
                if (r < c.p - 1) a2 += lb * lb;
                else t = lb;
            }
            const double a = sqrt(a2);
            double np;
            if (a <= -t) np = 0.0;
            else if (a <= t) np = a2 + t * t;
            else np = 0.5 * (a + t) * (a + t);
            Jc = (np - nl) / (2.0 * mu_c);
        }
        return Jc;
    }

    ALTRO_HD double stage_cost(int k, const double (&x)[NX], const double (&u)[NU]) const
    {
        double Jk = 0.0;
        if (k == N - 1) {
#pragma unroll
            for (int i = 0; i < NX; ++i) { const double e = x[i] - xref(k, i); Jk += 0.5 * C.Qf[i] * e * e; }
            return Jk;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) { const double e = x[i] - xref(k, i); Jk += 0.5 * C.Q[i] * e * e; }
#pragma unroll
        for (int i = 0; i < NU; ++i) { const double e = u[i] - uref(k, i); Jk += 0.5 * C.R[i] * e * e; }
        return Jk * P.dt;
    }

    // Canonical sum of the cost items (csum of the CTA kernel / oracle): 32 strided partials, then a binary tree.
    ALTRO_HD double item_sum(int count) const
    {
        for (int l = 0; l < 32; ++l) {
            double a = 0.0;
            for (int i = l; i < count; i += 32) a += sc(SC_ITM + i);
            sc(SC_V + l) = a;
        }
        for (int o = 16; o > 0; o >>= 1)
            for (int l = 0; l < o; ++l) sc(SC_V + l) = sc(SC_V + l) + sc(SC_V + l + o);
        return sc(SC_V);
    }

    // AL cost of a trajectory: items it = piece * N + k (piece 0 = stage cost, 1 + ci = block ci), evaluated knot by
    // knot (each knot is staged once) and summed in the canonical item order.
    ALTRO_HDN double al_cost(int oXc, int oUc) const
    {
        for (int k = 0; k < N; ++k) {
            double x[NX], u[NU];
            stage_knot(k, oXc, oUc, x, u, k + 1);
            sc(SC_ITM + k) = stage_cost(k, x, u);
            for (int ci = 0; ci < ncon; ++ci) sc(SC_ITM + (1 + ci) * N + k) = con_cost(ci, k);
        }
        return item_sum(N * (1 + ncon));
    }

    ALTRO_HD double objective_cost() const
    {
        for (int k = 0; k < N; ++k) {
            double x[NX], u[NU];
            stage_knot(k, L.X, L.U, x, u, k + 1);
            sc(SC_ITM + k) = stage_cost(k, x, u);
        }
        return item_sum(N);
    }

    ALTRO_HDN double max_violation() const
    {
        double v = 0.0;
        for (int k = 0; k < N; ++k) {
            double x[NX], u[NU];
            stage_knot(k, L.X, L.U, x, u, k + 1);
            for (int ci = 0; ci < ncon; ++ci) {
                const ConDesc &c = cd[ci];
                if (k < c.k0 || k >= c.k1) continue;
                const size_t di = con_idx(c, k);
                const double *G = lane_global(c.G) + di * c.p * c.w, *h = lane_global(c.h) + di * c.p;
                const int zb = c.side == ALTRO_STATE ? SC_ZX : SC_ZU;
                if (c.sense == ALTRO_EQUALITY) {
                    for (int r = 0; r < c.p; ++r) v = fmax(v, fabs(row_value(c, G, h, zb, r)));
                } else if (c.sense == ALTRO_INEQUALITY) {
                    for (int r = 0; r < c.p; ++r) v = fmax(v, row_value(c, G, h, zb, r));
                } else {
                    double a2 = 0.0, t = 0.0;
                    for (int r = 0; r < c.p; ++r) {
                        const double cv = row_value(c, G, h, zb, r);
                        if (r < c.p - 1) a2 += cv * cv;
                        else t = cv;
                    }
                    const double a = sqrt(a2);
                    if (!P.o.soc_viol_proj) {
                        v = fmax(v, a - t);
                    } else if (a <= -t) {
                        for (int r = 0; r < c.p; ++r) v = fmax(v, fabs(row_value(c, G, h, zb, r)));
                    } else if (a > t) {
                        const double cf = 0.5 * (1.0 + t / a);
                        for (int r = 0; r < c.p - 1; ++r) v = fmax(v, fabs((1.0 - cf) * row_value(c, G, h, zb, r)));
                        v = fmax(v, fabs(t - cf * a));
                    }
                }
            }
        }
        return v;
    }

    ALTRO_HDN void dual_update() const
    {
        for (int k = 0; k < N; ++k) {
            double x[NX], u[NU];
            stage_knot(k, L.X, L.U, x, u, k + 1);
            for (int ci = 0; ci < ncon; ++ci) {
                const ConDesc &c = cd[ci];
                if (k < c.k0 || k >= c.k1) continue;
                const double mu_c = mu(ci);
                const size_t di = con_idx(c, k);
                const double *G = lane_global(c.G) + di * c.p * c.w, *h = lane_global(c.h) + di * c.p;
                const int zb = c.side == ALTRO_STATE ? SC_ZX : SC_ZU;
                const int lo = L.lam + c.dual_off + (k - c.k0) * c.p;
                if (c.sense == ALTRO_EQUALITY) {
                    for (int r = 0; r < c.p; ++r)
                        W(lo + r) = fmin(fmax(W(lo + r) + mu_c * row_value(c, G, h, zb, r), -P.o.dual_max), P.o.dual_max);
                } else if (c.sense == ALTRO_INEQUALITY) {
                    for (int r = 0; r < c.p; ++r)
                        W(lo + r) = fmin(fmax(W(lo + r) + mu_c * row_value(c, G, h, zb, r), 0.0), P.o.dual_max);
                } else {
                    double a2 = 0.0, t = 0.0;
                    for (int r = 0; r < c.p; ++r) {
                        const double lb = W(lo + r) - mu_c * row_value(c, G, h, zb, r);
                        W(lo + r) = lb;
                        if (r < c.p - 1) a2 += lb * lb;
                        else t = lb;
                    }
                    const double a = sqrt(a2);
                    if (a <= -t) {
                        for (int r = 0; r < c.p; ++r) W(lo + r) = 0.0;
                    } else if (a > t) {
                        const double cf = 0.5 * (1.0 + t / a);
                        for (int r = 0; r < c.p - 1; ++r) W(lo + r) *= cf;
                        W(lo + c.p - 1) = cf * a;
                    }
                }
            }
        }
    }

    // ---------------------------------------------------------------- AL expansion of one (block, knot) (A.3)
    // g[w] and H (w x w dense symmetric, or w diagonal entries for row-sparse blocks) of the staged knot into SC_G.
    ALTRO_HD void expand_block(const ConDesc &c, int ci, int k) const
    {
        const double mu_c = mu(ci);
        const int w = c.w, p = c.p;
        const size_t di = con_idx(c, k);
        const double *G = lane_global(c.G) + di * p * w, *h = lane_global(c.h) + di * p;
        const int zb = c.side == ALTRO_STATE ? SC_ZX : SC_ZU;
        const int lo = L.lam + c.dual_off + (k - c.k0) * p;
        const int g = SC_G, H = SC_G + w;
        if (c.rowsparse) {
            for (int j = 0; j < 2 * w; ++j) sc(g + j) = 0.0;
            for (int r = 0; r < p; ++r) {
                const double v = row_value(c, G, h, zb, r), cf = lane_global(c.rs_coef)[r], l = W(lo + r);
                const bool act = c.sense == ALTRO_EQUALITY || (v >= 0.0) || (l > 0.0);
                const int col = lane_global(c.rs_col)[r];
                sc(g + col) += cf * (l + (act ? mu_c * v : 0.0));
                sc(H + col) += act ? cf * cf * mu_c : 0.0;
            }
        } else if (c.sense != ALTRO_SECOND_ORDER_CONE) {
            for (int r = 0; r < p; ++r) {
                const double v = row_value(c, G, h, zb, r), l = W(lo + r);
                const bool act = c.sense == ALTRO_EQUALITY || (v >= 0.0) || (l > 0.0);
                sc(SC_Y + r) = l + (act ? mu_c * v : 0.0);
                sc(SC_D + r) = act ? mu_c : 0.0;
            }
            for (int j = 0; j < w; ++j) {
                double acc = 0.0;
                for (int r = 0; r < p; ++r) acc = fma(G[r * w + j], sc(SC_Y + r), acc);
                sc(g + j) = acc;
            }
            for (int i = 0; i < w; ++i)
                for (int j = i; j < w; ++j) {
                    double acc = 0.0;
                    for (int r = 0; r < p; ++r) acc = fma(G[r * w + i] * sc(SC_D + r), G[r * w + j], acc);
                    sc(H + i * w + j) = acc;
                    sc(H + j * w + i) = acc;
                }
        } else {
            double a2 = 0.0;
            for (int r = 0; r < p; ++r) {
                const double lb = W(lo + r) - mu_c * row_value(c, G, h, zb, r);
                sc(SC_Y + r) = lb;
                if (r < p - 1) a2 += lb * lb;
            }
            const double t = sc(SC_Y + p - 1), a = sqrt(a2);
            const double *gt = G + (p - 1) * w;
            if (a <= -t) {
                for (int j = 0; j < w + w * w; ++j) sc(g + j) = 0.0;
            } else if (a <= t) {
                for (int j = 0; j < w; ++j) {
                    double acc = 0.0;
                    for (int r = 0; r < p; ++r) acc = fma(G[r * w + j], sc(SC_Y + r), acc);
                    sc(g + j) = -acc;
                }
                for (int i = 0; i < w; ++i)
                    for (int j = i; j < w; ++j) {
                        double acc = 0.0;
                        for (int r = 0; r < p; ++r) acc = fma(G[r * w + i], G[r * w + j], acc);
                        sc(H + i * w + j) = mu_c * acc;
                        sc(H + j * w + i) = mu_c * acc;
                    }
            } else {
                const double ia = 1.0 / a, cf = 0.5 * (1.0 + t * ia);
                const double cx = P.o.soc_hess_exact ? cf : cf * cf;
                for (int j = 0; j < w; ++j) {
                    double acc = 0.0;
                    for (int r = 0; r < p - 1; ++r) acc = fma(G[r * w + j], sc(SC_Y + r), acc);
                    sc(SC_D + j) = acc * ia;
                }
                for (int i = 0; i < w; ++i)
                    for (int j = i; j < w; ++j) {
                        double gg = 0.0;
                        for (int r = 0; r < p - 1; ++r) gg = fma(G[r * w + i], G[r * w + j], gg);
                        const double qi = sc(SC_D + i) + gt[i], qj = sc(SC_D + j) + gt[j];
                        const double hv = mu_c * (cx * (gg - sc(SC_D + i) * sc(SC_D + j)) + 0.5 * qi * qj);
                        sc(H + i * w + j) = hv;
                        sc(H + j * w + i) = hv;
                    }
                for (int j = 0; j < w; ++j) sc(g + j) = -cf * a * (sc(SC_D + j) + gt[j]);
            }
        }
    }

    // Adds the expansions of every block of `side` active at the staged knot k to the vector and the LD x LD matrix
    // that the caller initialised at scratch offset `at` ([vec (8) | mat]), in ascending block order (the oracle's
    // scatter_expansion).  The targets are run-time indices, hence scratch; the backward pass reads them back with
    // compile-time offsets.
    ALTRO_HD void add_expansions(int k, int side, int at, int LD) const
    {
        for (int ci = 0; ci < ncon; ++ci) {
            const ConDesc &c = cd[ci];
            if (c.side != side || k < c.k0 || k >= c.k1) continue;
            expand_block(c, ci, k);
            const int w = c.w;
            if (c.rowsparse) {
                for (int e = 0; e < w; ++e) {
                    const int zi = c.inds[e];
                    sc(at + zi) += sc(SC_G + e);
                    sc(at + 8 + zi * LD + zi) += sc(SC_G + w + e);
                }
            } else {
                for (int e = 0; e < w; ++e) sc(at + c.inds[e]) += sc(SC_G + e);
                for (int i = 0; i < w; ++i)
                    for (int j = 0; j < w; ++j) sc(at + 8 + c.inds[i] * LD + c.inds[j]) += sc(SC_G + w + i * w + j);
            }
        }
    }

    // prefetch the duals of every block at knot k (the expansion of that knot reads them next)
    ALTRO_HD void prefetch_duals(int k) const
    {
        for (int ci = 0; ci < ncon; ++ci) {
            const ConDesc &c = cd[ci];
            if (k < c.k0 || k >= c.k1) continue;
            const int lo = L.lam + c.dual_off + (k - c.k0) * c.p;
            for (int r = 0; r < c.p; ++r) prefetch(lo + r);
        }
    }

    // ---------------------------------------------------------------- backward pass (A.7)
    ALTRO_HD void reg_increase()
    {
        drho = fmax(drho * P.o.bp_reg_increase_factor, P.o.bp_reg_increase_factor);
        rho = fmax(rho * drho, P.o.bp_reg_min);
    }
    ALTRO_HD void reg_decrease()
    {
        drho = fmin(drho / P.o.bp_reg_increase_factor, 1.0 / P.o.bp_reg_increase_factor);
        const double r = rho * drho;
        rho = (r > P.o.bp_reg_min) ? r : 0.0;
    }

    // Returns false if Quu could not be made positive definite.
    ALTRO_HDN bool backward_pass()
    {
        const double dt = P.dt;
        for (;;) {
            double a1 = 0.0, a2 = 0.0;
            double S[NX * NX], s[NX];
            // terminal cost-to-go
            {
                double x[NX], u[NU];
                prefetch_duals(N - 1);
                stage_knot(N - 1, L.X, L.U, x, u, N - 2);
#pragma unroll
                for (int i = 0; i < NX; ++i) {
#pragma unroll
                    for (int j = 0; j < NX; ++j) sc(SC_EX + 8 + i * NX + j) = (i == j) ? C.Qf[i] : 0.0;
                    sc(SC_EX + i) = C.Qf[i] * (x[i] - xref(N - 1, i));
                }
                add_expansions(N - 1, ALTRO_STATE, SC_EX, NX);
#pragma unroll
                for (int i = 0; i < NX; ++i) {
#pragma unroll
                    for (int j = 0; j < NX; ++j) S[i * NX + j] = sc(SC_EX + 8 + i * NX + j);
                    s[i] = sc(SC_EX + i);
                }
            }
            bool bad = false;
            for (int k = N - 2; k >= 0; --k) {
                // this knot's state / control / duals: staged first, the previous knot's rows prefetched
                double xk[NX], uk[NU];
                prefetch_duals(k);
                stage_knot(k, L.X, L.U, xk, uk, k - 1);
                double SA[NX * NX], SB[NX * NU];
#pragma unroll
                for (int i = 0; i < NX; ++i) {
#pragma unroll
                    for (int j = 0; j < NX; ++j) {
                        double acc = 0.0;
#pragma unroll
                        for (int l = 0; l < NX; ++l) acc = fma(S[i * NX + l], C.A[l * NX + j], acc);
                        SA[i * NX + j] = acc;
                    }
#pragma unroll
                    for (int j = 0; j < NU; ++j) {
                        double acc = 0.0;
#pragma unroll
                        for (int l = 0; l < NX; ++l) acc = fma(S[i * NX + l], C.B[l * NU + j], acc);
                        SB[i * NU + j] = acc;
                    }
                }
                // cost expansion (diagonal LQR cost) + AL expansion
                double Qxx[NX * NX], Quu[NU * NU], Qx[NX], Qu[NU];
#pragma unroll
                for (int i = 0; i < NX; ++i) {
#pragma unroll
                    for (int j = 0; j < NX; ++j) sc(SC_EX + 8 + i * NX + j) = (i == j) ? dt * C.Q[i] : 0.0;
                    sc(SC_EX + i) = dt * C.Q[i] * (xk[i] - xref(k, i));
                }
#pragma unroll
                for (int i = 0; i < NU; ++i) {
#pragma unroll
                    for (int j = 0; j < NU; ++j) sc(SC_EU + 8 + i * NU + j) = (i == j) ? dt * C.R[i] : 0.0;
                    sc(SC_EU + i) = dt * C.R[i] * (uk[i] - uref(k, i));
                }
                add_expansions(k, ALTRO_STATE, SC_EX, NX);
                add_expansions(k, ALTRO_CONTROL, SC_EU, NU);
#pragma unroll
                for (int i = 0; i < NX; ++i) {
#pragma unroll
                    for (int j = 0; j < NX; ++j) Qxx[i * NX + j] = sc(SC_EX + 8 + i * NX + j);
                    Qx[i] = sc(SC_EX + i);
                }
#pragma unroll
                for (int i = 0; i < NU; ++i) {
#pragma unroll
                    for (int j = 0; j < NU; ++j) Quu[i * NU + j] = sc(SC_EU + 8 + i * NU + j);
                    Qu[i] = sc(SC_EU + i);
                }
                // action-value expansion: Qxx += A'SA, Qux = B'SA, Quu += B'SB, Qx += A's, Qu += B's
                double Qux[NU * NX];
#pragma unroll
                for (int i = 0; i < NX; ++i)
#pragma unroll
                    for (int j = 0; j < NX; ++j) {
                        double acc = Qxx[i * NX + j];
#pragma unroll
                        for (int l = 0; l < NX; ++l) acc = fma(C.A[l * NX + i], SA[l * NX + j], acc);
                        Qxx[i * NX + j] = acc;
                    }
#pragma unroll
                for (int i = 0; i < NU; ++i)
#pragma unroll
                    for (int j = 0; j < NX; ++j) {
                        double acc = 0.0;
#pragma unroll
                        for (int l = 0; l < NX; ++l) acc = fma(C.B[l * NU + i], SA[l * NX + j], acc);
                        Qux[i * NX + j] = acc;
                    }
#pragma unroll
                for (int i = 0; i < NU; ++i)
#pragma unroll
                    for (int j = 0; j < NU; ++j) {
                        double acc = Quu[i * NU + j];
#pragma unroll
                        for (int l = 0; l < NX; ++l) acc = fma(C.B[l * NU + i], SB[l * NU + j], acc);
                        Quu[i * NU + j] = acc;
                    }
#pragma unroll
                for (int i = 0; i < NX; ++i) {
                    double acc = Qx[i];
#pragma unroll
                    for (int l = 0; l < NX; ++l) acc = fma(C.A[l * NX + i], s[l], acc);
                    Qx[i] = acc;
                }
#pragma unroll
                for (int i = 0; i < NU; ++i) {
                    double acc = Qu[i];
#pragma unroll
                    for (int l = 0; l < NX; ++l) acc = fma(C.B[l * NU + i], s[l], acc);
                    Qu[i] = acc;
                }
                // LDL' of Quu + rho I: un-normalised lower factor, reciprocal pivots
                double Lf[NU * NU], ld[NU];
#pragma unroll
                for (int i = 0; i < NU; ++i)
#pragma unroll
                    for (int j = 0; j < NU; ++j) Lf[i * NU + j] = Quu[i * NU + j] + ((i == j) ? rho : 0.0);
#pragma unroll
                for (int j = 0; j < NU; ++j) {
                    if (!bad) {
#pragma unroll
                        for (int i = j; i < NU; ++i) {
                            double acc = Lf[i * NU + j];
#pragma unroll
                            for (int l = 0; l < j; ++l) acc = fma(-Lf[i * NU + l], Lf[j * NU + l] * ld[l], acc);
                            Lf[i * NU + j] = acc;
                        }
                        const double piv = Lf[j * NU + j];
                        if (!(piv > 0.0)) bad = true;
                        else ld[j] = rcp(piv);
                    }
                }
                if (bad) break;
                // gains [K | d] = -(Quu + rho I)^-1 [Qux | Qu]
                double Kk[NU * NX], dvk[NU];
#pragma unroll
                for (int c = 0; c <= NX; ++c) {
                    double b[NU];
#pragma unroll
                    for (int i = 0; i < NU; ++i) {
                        double acc = -((c < NX) ? Qux[i * NX + (c < NX ? c : 0)] : Qu[i]);
#pragma unroll
                        for (int l = 0; l < i; ++l) acc = fma(-Lf[i * NU + l], b[l] * ld[l], acc);
                        b[i] = acc;
                    }
#pragma unroll
                    for (int i = 0; i < NU; ++i) b[i] = b[i] * ld[i];
#pragma unroll
                    for (int i = NU - 1; i >= 0; --i) {
                        double acc2 = 0.0;
#pragma unroll
                        for (int l = NU - 1; l > i; --l) acc2 = fma(Lf[l * NU + i], b[l], acc2);
                        b[i] = fma(-ld[i], acc2, b[i]);
                    }
#pragma unroll
                    for (int i = 0; i < NU; ++i) {
                        if (c < NX) Kk[i * NX + (c < NX ? c : 0)] = b[i];
                        else dvk[i] = b[i];
                    }
                }
#pragma unroll
                for (int i = 0; i < NU * NX; ++i) W(L.K + k * NU * NX + i) = Kk[i];
#pragma unroll
                for (int i = 0; i < NU; ++i) W(L.dv + k * NU + i) = dvk[i];
                // cost-to-go with the unregularised Quu: T1 = Quu K + Qux, t1 = Quu d + Qu
                double T1[NU * NX], t1[NU];
#pragma unroll
                for (int i = 0; i < NU; ++i) {
#pragma unroll
                    for (int j = 0; j < NX; ++j) {
                        double acc = Qux[i * NX + j];
#pragma unroll
                        for (int l = 0; l < NU; ++l) acc = fma(Quu[i * NU + l], Kk[l * NX + j], acc);
                        T1[i * NX + j] = acc;
                    }
                    double acc = Qu[i];
#pragma unroll
                    for (int l = 0; l < NU; ++l) acc = fma(Quu[i * NU + l], dvk[l], acc);
                    t1[i] = acc;
                }
                // S' = Qxx + K'T1 + Qux'K (into SA), s = Qx + K't1 + Qux'd
#pragma unroll
                for (int i = 0; i < NX; ++i)
#pragma unroll
                    for (int j = 0; j < NX; ++j) {
                        double acc = Qxx[i * NX + j];
#pragma unroll
                        for (int l = 0; l < NU; ++l) acc = fma(Kk[l * NX + i], T1[l * NX + j], acc);
#pragma unroll
                        for (int l = 0; l < NU; ++l) acc = fma(Qux[l * NX + i], Kk[l * NX + j], acc);
                        SA[i * NX + j] = acc;
                    }
#pragma unroll
                for (int i = 0; i < NX; ++i) {
                    double acc = Qx[i];
#pragma unroll
                    for (int l = 0; l < NU; ++l) acc = fma(Kk[l * NX + i], t1[l], acc);
#pragma unroll
                    for (int l = 0; l < NU; ++l) acc = fma(Qux[l * NX + i], dvk[l], acc);
                    s[i] = acc;
                }
#pragma unroll
                for (int i = 0; i < NU; ++i) {
                    a1 = fma(dvk[i], Qu[i], a1);
                    a2 = fma(0.5 * dvk[i], t1[i] - Qu[i], a2);
                }
#pragma unroll
                for (int i = 0; i < NX; ++i)
#pragma unroll
                    for (int j = 0; j < NX; ++j) S[i * NX + j] = 0.5 * (SA[i * NX + j] + SA[j * NX + i]);
            }
            if (bad) {
                reg_increase();
                if (rho > P.o.bp_reg_max) return false;
                continue;
            }
            dV1 = a1;
            dV2 = a2;
            reg_decrease();
            return true;
        }
    }

    // ---------------------------------------------------------------- rollouts (A.6, A.8)
    ALTRO_HDN void rollout_open_loop() const
    {
        double x[NX];
#pragma unroll
        for (int j = 0; j < NX; ++j) x[j] = W(L.X + j);
        for (int k = 0; k < N - 1; ++k) {
            double u[NU], xn[NX];
#pragma unroll
            for (int j = 0; j < NU; ++j) u[j] = W(L.U + k * NU + j);
            if (k + 1 < N - 1) {
#pragma unroll
                for (int j = 0; j < NU; ++j) prefetch(L.U + (k + 1) * NU + j);
            }
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                double acc = C.d[i];
#pragma unroll
                for (int j = 0; j < NX; ++j) acc = fma(C.A[i * NX + j], x[j], acc);
#pragma unroll
                for (int j = 0; j < NU; ++j) acc = fma(C.B[i * NU + j], u[j], acc);
                xn[i] = acc;
                W(L.X + (k + 1) * NX + i) = acc;
            }
#pragma unroll
            for (int i = 0; i < NX; ++i) x[i] = xn[i];
        }
    }

    // 0: a state left the box, 2: bit-identical to (X, U), 1: otherwise
    ALTRO_HDN int rollout_alpha(double al) const
    {
        double xb[NX];
        bool bad = false, same = true;
#pragma unroll
        for (int i = 0; i < NX; ++i) { xb[i] = W(L.X + i); W(L.Xb + i) = xb[i]; }
        for (int k = 0; k < N - 1; ++k) {
            // one batch of independent loads for this knot, the next knot's rows prefetched
            double xo[NX], xo1[NX], uo[NU], dvo[NU], Kk[NU * NX];
#pragma unroll
            for (int j = 0; j < NX; ++j) { xo[j] = W(L.X + k * NX + j); xo1[j] = W(L.X + (k + 1) * NX + j); }
#pragma unroll
            for (int i = 0; i < NU; ++i) { uo[i] = W(L.U + k * NU + i); dvo[i] = W(L.dv + k * NU + i); }
#pragma unroll
            for (int i = 0; i < NU * NX; ++i) Kk[i] = W(L.K + k * NU * NX + i);
            if (k + 1 < N - 1) {
#pragma unroll
                for (int i = 0; i < NU * NX; ++i) prefetch(L.K + (k + 1) * NU * NX + i);
#pragma unroll
                for (int i = 0; i < NU; ++i) { prefetch(L.U + (k + 1) * NU + i); prefetch(L.dv + (k + 1) * NU + i); }
#pragma unroll
                for (int j = 0; j < NX; ++j) prefetch(L.X + (k + 2) * NX + j);
            }
            double dx[NX], ub[NU];
#pragma unroll
            for (int j = 0; j < NX; ++j) dx[j] = xb[j] - xo[j];
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                double acc = fma(al, dvo[i], uo[i]);
#pragma unroll
                for (int j = 0; j < NX; ++j) acc = fma(Kk[i * NX + j], dx[j], acc);
                ub[i] = acc;
                W(L.Ub + k * NU + i) = acc;
                if (!(acc == uo[i])) same = false;
            }
            double xn[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                double acc = C.d[i];
#pragma unroll
                for (int j = 0; j < NX; ++j) acc = fma(C.A[i * NX + j], xb[j], acc);
#pragma unroll
                for (int j = 0; j < NU; ++j) acc = fma(C.B[i * NU + j], ub[j], acc);
                xn[i] = acc;
                W(L.Xb + (k + 1) * NX + i) = acc;
                if (!(fabs(acc) <= P.o.max_state_value)) bad = true;
                else if (!(acc == xo1[i])) same = false;
            }
#pragma unroll
            for (int i = 0; i < NX; ++i) xb[i] = xn[i];
        }
        return bad ? 0 : (same ? 2 : 1);
    }

    ALTRO_HD void copy_traj(int oXd, int oUd, int oXs, int oUs) const
    {
#pragma unroll 6
        for (int i = 0; i < N * n; ++i) W(oXd + i) = W(oXs + i);
#pragma unroll 6
        for (int i = 0; i < (N - 1) * m; ++i) W(oUd + i) = W(oUs + i);
    }

    ALTRO_HD double gradient_todorov() const
    {
        for (int k = 0; k < N - 1; ++k) {
            double mx = 0.0;
#pragma unroll
            for (int i = 0; i < NU; ++i) mx = fmax(mx, fabs(W(L.dv + k * NU + i)) / (fabs(W(L.U + k * NU + i)) + 1.0));
            sc(SC_ITM + k) = mx;
        }
        return item_sum(N - 1) / (double)(N - 1);
    }

    // ---------------------------------------------------------------- warm-started MPC transition
    ALTRO_HDN void transition(int s)
    {
        const double *zz = P.noise ? lane_global(P.noise) + ((size_t)s * P.B + inst) * n : nullptr;
        double s0 = P.noise_w1, s1 = P.noise_w1;
        if (zz) {
            if (P.noise_mode == 1) {
                double mx = 0.0;
                for (int i = 0; i < n; ++i) mx = fmax(mx, fabs(W(L.X + n + i)));
                s0 = s1 = mx * P.noise_w1;
            } else if (P.noise_mode == 2) {
                double a = 0.0, b = 0.0;
                for (int i = 0; i < n / 2; ++i) a += W(L.X + n + i) * W(L.X + n + i);
                for (int i = n / 2; i < n; ++i) b += W(L.X + n + i) * W(L.X + n + i);
                s0 = sqrt(a) * P.noise_w1;
                s1 = sqrt(b) * P.noise_w2;
            }
        }
        for (int i = 0; i < n; ++i) {
            double v = W(L.X + n + i);
            if (zz) v += zz[i] * ((P.noise_mode == 2 && i >= n / 2) ? s1 : s0);
            W(L.X + i) = v;
        }
        if (P.shift) {
#pragma unroll 6
            for (int i = 0; i < (N - 2) * m; ++i) W(L.U + i) = W(L.U + i + m);
            for (int ci = 0; ci < ncon; ++ci) {
                const int cnt = (cd[ci].k1 - cd[ci].k0 - 1) * cd[ci].p, p = cd[ci].p, lo = L.lam + cd[ci].dual_off;
#pragma unroll 4
                for (int i = 0; i < cnt; ++i) W(lo + i) = W(lo + i + p);
            }
        }
    }

    // ---------------------------------------------------------------- solve! as a state machine
    // One trip: at most one backward pass and one line-search trial.  Returns when the lane has to wait for the
    // next trip (or is done); the warp's loop calls step() on every unfinished lane until all are LP_DONE.
    ALTRO_HD void step()
    {
        const altro_opts_t &o = P.o;
        if (phase == LP_STEP_BEGIN) {
            if (P.steps > 0) {
                set_step(st + 1);
                kcur = (P.kidx ? P.kidx[inst] : 0) + st + 1;
                transition(st);
                if (st > 0) {
                    if (o.reset_duals)
                        for (int i = 0; i < P.P; ++i) W(L.lam + i) = 0.0;
                    for (int c = 0; c < MAX_CON; ++c) mu(c) = o.penalty_initial;
                }
                if (P.x0_log)
                    for (int i = 0; i < n; ++i) P.x0_log[((size_t)st * P.B + inst) * n + i] = W(L.X + i);
            }
            iters = trials = 0;
            status = ALTRO_UNSOLVED;
            cmax = INFINITY;
            J = 0.0;
            pen_max = 0.0;
            outer = 1;
            phase = LP_OUTER_BEGIN;
        }
        if (phase == LP_OUTER_BEGIN) {
            const bool last = (outer == o.iterations_outer) || ncon == 0;
            ctol = last ? o.cost_tolerance : o.cost_tolerance_intermediate;
            gtol = last ? o.gradient_tolerance : o.gradient_tolerance_intermediate;
            rho = o.bp_reg_initial;
            drho = 0.0;
            dJ_zero = 0;
            rollout_open_loop();
            J_prev = al_cost(L.X, L.U);
            J = J_prev;
            if (o.first_step_unconditional) J_prev = INFINITY;
            inner_it = 0;
            phase = LP_BP;
        }
        if (phase == LP_BP) {
            if (!backward_pass()) {
                status = ALTRO_NOT_PD;
                phase = LP_OUTER_END;
            } else {
                Jls = INFINITY;
                alpha = 1.0;
                z = -1.0;
                ls_iter = 0;
                phase = LP_TRIAL;
            }
        }
        if (phase == LP_TRIAL) {
            // one evaluation of the body of  while ((z <= lo || z > hi) && J >= J_prev)
            if (ls_iter > o.iterations_linesearch) {
                copy_traj(L.Xb, L.Ub, L.X, L.U);
                Jls = al_cost(L.Xb, L.Ub);
                reg_increase();
                rho += o.bp_reg_fp;
                phase = LP_POST;
            } else {
                const int ok = rollout_alpha(alpha);
                ++trials;
                if (ok) {
                    Jls = al_cost(L.Xb, L.Ub);
                    const double expected = -alpha * (dV1 + alpha * dV2);
                    z = expected > 0.0 ? (J_prev - Jls) / expected : -1.0;
                }
                ++ls_iter;
                alpha *= 0.5;
                if (ok == 2) ls_iter = o.iterations_linesearch + 1;
                if (!((z <= o.line_search_lower_bound || z > o.line_search_upper_bound) && Jls >= J_prev)) phase = LP_POST;
                else if (ls_iter > o.iterations_linesearch) {  // failed: settle it in this trip, it costs one cost evaluation
                    copy_traj(L.Xb, L.Ub, L.X, L.U);
                    Jls = al_cost(L.Xb, L.Ub);
                    reg_increase();
                    rho += o.bp_reg_fp;
                    phase = LP_POST;
                }
            }
        }
        if (phase == LP_POST) {
            J = Jls;
            if (J > o.max_cost_value || !(J == J)) {
                status = ALTRO_MAXIMUM_COST;
                phase = LP_OUTER_END;
            } else {
                copy_traj(L.X, L.U, L.Xb, L.Ub);
                const double dJ = fabs(J - J_prev);
                J_prev = J;
                const double grad = gradient_todorov();
                ++iters;
                dJ_zero = (dJ == 0.0) ? dJ_zero + 1 : 0;
                const bool small = o.dj_zero_converges ? (dJ >= 0.0 && dJ < ctol) : (dJ > 0.0 && dJ < ctol);
                ++inner_it;
                if (small && grad < gtol) { status = ALTRO_SOLVE_SUCCEEDED; phase = LP_OUTER_END; }
                else if (iters >= o.iterations) { status = ALTRO_MAX_ITERATIONS; phase = LP_OUTER_END; }
                else if (dJ_zero > o.dJ_counter_limit) { status = ALTRO_NO_PROGRESS; phase = LP_OUTER_END; }
                else if (inner_it >= o.iterations_inner) phase = LP_OUTER_END;
                else phase = LP_BP;
            }
        }
        if (phase == LP_OUTER_END) {
            bool done = status > ALTRO_SOLVE_SUCCEEDED;
            if (!done) {
                cmax = max_violation();
                pen_max = 0.0;
                for (int c = 0; c < ncon; ++c) pen_max = fmax(pen_max, mu(c));
                if (cmax < o.constraint_tolerance) done = true;
                else if (o.kickout_max_penalty && pen_max >= o.penalty_max) done = true;
                else {
                    dual_update();
                    for (int c = 0; c < ncon; ++c) mu(c) = fmin(mu(c) * o.penalty_scaling, o.penalty_max);
                    if (outer == o.iterations_outer) { status = ALTRO_MAX_ITERATIONS_OUTER; done = true; }
                }
            }
            if (done) phase = LP_STEP_END;
            else { ++outer; phase = LP_OUTER_BEGIN; }
        }
        if (phase == LP_STEP_END) {
            cmax = max_violation();
            if (status <= ALTRO_SOLVE_SUCCEEDED)
                status = (cmax < o.constraint_tolerance) ? ALTRO_SOLVE_SUCCEEDED : ALTRO_UNSOLVED;
            const double Jobj = objective_cost();
            const size_t at = (size_t)st * P.B + inst;
            P.iters[at] = iters;
            P.outer[at] = outer;
            P.status[at] = status;
            P.trials[at] = trials;
            P.cost[at] = Jobj;
            P.cost_al[at] = J;
            P.cmax[at] = cmax;
            P.penmax[at] = pen_max;
            if (P.steps > 0 && P.u0_log)
                for (int i = 0; i < m; ++i) P.u0_log[((size_t)st * P.B + inst) * m + i] = W(L.U + i);
            if (P.t_ns) {
                const long long t1 = now_ns();
                P.t_ns[at] = t1 - t0;
                t0 = t1;
            }
            if (st + 1 < steps) { ++st; phase = LP_STEP_BEGIN; }
            else phase = LP_DONE;
        }
    }
};

#ifdef __CUDACC__

// One warp advances LPW instances (lanes 0 .. LPW-1; the other lanes idle): with a few thousand instances per GPU,
// fewer instances per warp means more warps, i.e. more of the 4 x 148 schedulers busy and more latency hidden, at
// the price of issuing each instruction for fewer instances.  One CTA = one warp; no communication between warps.
template <int NX, int NU, int LPW>
__global__ void __launch_bounds__(32) altro_lane_kernel(const __grid_constant__ Params P,
                                                        const __grid_constant__ LaneConst<NX, NU> C, const LaneLayout L,
                                                        double *ws, size_t stride, int scratch_per_lane)
{
    extern __shared__ __align__(16) double lane_smem[];
    const int lane = threadIdx.x;
    // descriptors first (shared by the warp's lanes), the per-lane scratch behind them
    ConDesc *cds = reinterpret_cast<ConDesc *>(lane_smem);
    {
        const int words = P.ncon * (int)(sizeof(ConDesc) / sizeof(int));
        const int *src = reinterpret_cast<const int *>(P.con);
        int *dst = reinterpret_cast<int *>(cds);
        for (int i = lane; i < words; i += 32) dst[i] = src[i];
    }
    __syncwarp();
    double *scr = lane_smem + (size_t)(P.ncon > 0 ? P.ncon : 1) * (sizeof(ConDesc) / sizeof(double)) + lane;
    (void)scratch_per_lane;
    const int inst = blockIdx.x * LPW + lane;
    const bool valid = lane < LPW && inst < P.B;
    Lane<NX, NU, LPW> ln(P, C, L, ws + (valid ? inst : 0), stride, scr, cds, valid ? inst + P.inst_offset : P.inst_offset);
    if (!valid) ln.phase = LP_DONE;
    else ln.load();
    while (__any_sync(0xffffffffu, ln.phase != LP_DONE)) {
        if (ln.phase != LP_DONE) ln.step();
    }
    if (valid) ln.store();
}

#endif

}  // namespace altro
