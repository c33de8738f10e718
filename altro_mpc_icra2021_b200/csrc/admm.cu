// admm.cu -- batched convex cross-check of the MPC problems on the device.
//
// The reference validates every ALTRO solve against an independent convex solver (OSQP for the QPs,
// random_linear_problem.jl:37-77,141-186; ECOS / COSMO / Mosek for the cone programs, simple_rocket.jl:184-192,
// grasp_mpc.jl:75-80) and records ||X_altro - X_other||_inf.  This file restores that check without leaving the GPU and
// without sharing any code path with the AL-iLQR kernels: an operator-splitting (ADMM) solver in the style of OSQP /
// COSMO specialised to optimal control (O'Donoghue, Stathopoulos, Boyd: "A splitting method for optimal control"),
// working on the SAME problem description (dynamics, diagonal tracking cost, affine conic blocks  c = G z[inds] + h).
//
//   minimise   sum_k 1/2 (x_k - xr_k)' Q~ (x_k - xr_k) + 1/2 (u_k - ur_k)' R~ (u_k - ur_k)      Q~ = Q dt, R~ = R dt, Qf
//   subject to x_{k+1} = A_k x_k + B_k u_k + d_k,  x_0 given,   y_b = G_b z[inds_b] + h_b  in  K_b   for every block b
//
//   z-step :  LQ problem with the extra terms rho/2 |G z + h - y + w|^2, dynamics eliminated exactly by ONE Riccati
//             sweep.  Its Hessian does not depend on (y, w): cost-to-go matrices, gains and the factors of Quu are
//             computed once per solve (the "factor caching" of OSQP); an iteration is an affine backward sweep, a
//             rollout, and
//   y-step :  y = Pi_K(G z + h + w)  (zero cone / non-positive orthant / second-order cone),   w += G z + h - y.
//
// One CTA per instance, run-time dimensions, per-instance workspace in global memory (L2-resident).  Plain FP64, no
// operation-order contract with anything: this is an independent check, agreement with ALTRO is to solver tolerance.
#include <cuda_runtime.h>

#include "altro_admm.cuh"

namespace altro {

namespace {

constexpr int TA = 64;

struct Acc {
    const AdmmParams &P;
    int inst, n, m, N;
    __device__ size_t dyn_index(int k) const
    {
        size_t base = P.dyn_per_instance ? (size_t)inst * (P.dyn_sched ? (size_t)P.dyn_slots : (P.dyn_per_knot ? (size_t)(N - 1) : 1)) : 0;
        if (P.dyn_sched) {
            const int s0 = min(P.step0, P.sched_len - N);
            return base + P.dyn_sched[(size_t)inst * P.sched_len + s0 + k];
        }
        return base + (P.dyn_per_knot ? (size_t)k : 0);
    }
    __device__ const double *A(int k) const { return P.A + dyn_index(k) * n * n; }
    __device__ const double *B(int k) const { return P.Bm + dyn_index(k) * n * m; }
    __device__ const double *d(int k) const { return P.d + dyn_index(k) * n; }
    __device__ size_t con_idx(const ConDesc &c, int k) const
    {
        if (c.track) return (size_t)min((P.kidx ? P.kidx[inst] : 0) + k, c.track - 1);
        size_t idx = c.per_instance ? (size_t)inst * (c.per_knot ? (size_t)(c.k1 - c.k0) : 1) : 0;
        return idx + (c.per_knot ? (size_t)(k - c.k0) : 0);
    }
};

// dense G row access that also serves row-sparse (bound) blocks: their G is stored densely as well
__device__ __forceinline__ double Gat(const double *G, const ConDesc &c, int r, int j) { return G[r * c.w + j]; }

__global__ void __launch_bounds__(TA) admm_kernel(const __grid_constant__ AdmmParams P)
{
    extern __shared__ double sm[];
    const int tid = threadIdx.x, n = P.n, m = P.m, N = P.N, ncon = P.ncon, inst = blockIdx.x;
    Acc a{P, inst, n, m, N};
    // workspace
    double *ws = P.ws + (size_t)inst * P.ws_doubles;
    double *S = ws;                              // [N][n][n]
    double *Kg = S + (size_t)N * n * n;          // [N-1][m][n]
    double *Lf = Kg + (size_t)(N - 1) * m * n;   // [N-1][m][m]  LDL' of Quu (unit lower + pivots on the diagonal)
    double *sv = Lf + (size_t)(N - 1) * m * m;   // [N][n]
    double *df = sv + (size_t)N * n;             // [N-1][m]
    double *X = df + (size_t)(N - 1) * m;        // [N][n]
    double *U = X + (size_t)N * n;               // [N-1][m]
    double *y = U + (size_t)(N - 1) * m;         // [Pd]
    double *w = y + P.Pd;                        // [Pd]
    double *T1 = w + P.Pd;                       // [n][n] scratch
    double *T2 = T1 + n * n;                     // [n][max(n,m)] scratch
    double *Qux = T2 + n * (n > m ? n : m);      // [m][n]
    double *vx = sm, *vu = sm + n, *tv = sm + n + m, *red = sm + 2 * n + m;  // small vectors in shared memory
    double rho = P.rho;  // residual-balanced (OSQP / Boyd 3.4.1): doubled or halved every few iterations, then refactored
    const double dt = P.dt;
    const double *xr = P.xref + (size_t)inst * N * n, *ur = P.uref + (size_t)inst * (N - 1) * m;

    // ---- scatter rho G'G of the blocks of `side` at knot k into a matrix (LD x LD), one thread per matrix entry pair
    auto add_GtG = [&](int k, int side, double *M, int LD) {
        for (int ci = 0; ci < ncon; ++ci) {
            const ConDesc &c = P.con[ci];
            if (c.side != side || k < c.k0 || k >= c.k1) continue;
            const double *G = c.G + a.con_idx(c, k) * c.p * c.w;
            for (int e = tid; e < c.w * c.w; e += TA) {
                const int i = e / c.w, j = e - i * c.w;
                double acc = 0.0;
                for (int r = 0; r < c.p; ++r) acc += Gat(G, c, r, i) * Gat(G, c, r, j);
                M[c.inds[i] * LD + c.inds[j]] += rho * acc;
            }
            __syncthreads();
        }
    };
    // ---- rho G'(h - y + w) of the blocks of `side` at knot k added to vec
    auto add_Gtr = [&](int k, int side, double *vec) {
        for (int ci = 0; ci < ncon; ++ci) {
            const ConDesc &c = P.con[ci];
            if (c.side != side || k < c.k0 || k >= c.k1) continue;
            const size_t di = a.con_idx(c, k);
            const double *G = c.G + di * c.p * c.w, *h = c.h + di * c.p;
            const int lo = c.dual_off + (k - c.k0) * c.p;
            for (int j = tid; j < c.w; j += TA) {
                double acc = 0.0;
                for (int r = 0; r < c.p; ++r) acc += Gat(G, c, r, j) * (h[r] - y[lo + r] + w[lo + r]);
                vec[c.inds[j]] += rho * acc;
            }
            __syncthreads();
        }
    };

    // ---- initial guess: the handle's current trajectory; y = Pi(G z + h), w = 0
    for (int i = tid; i < N * n; i += TA) X[i] = P.X[(size_t)inst * N * n + i];
    for (int i = tid; i < (N - 1) * m; i += TA) U[i] = P.U[(size_t)inst * (N - 1) * m + i];
    for (int i = tid; i < n; i += TA) X[i] = P.x0[(size_t)inst * n + i];
    for (int i = tid; i < 2 * P.Pd; i += TA) y[i] = 0.0;
    __syncthreads();

    // ---- factor phase: S_k, K_k and the factors of Quu (independent of y, w; repeated when rho changes)
    auto factor = [&]() {
        double *SN = S + (size_t)(N - 1) * n * n;
        for (int e = tid; e < n * n; e += TA) SN[e] = (e / n == e % n) ? P.Qf[e / n] : 0.0;
        __syncthreads();
        add_GtG(N - 1, ALTRO_STATE, SN, n);
        for (int k = N - 2; k >= 0; --k) {
            const double *A = a.A(k), *Bm = a.B(k), *Sn = S + (size_t)(k + 1) * n * n;
            double *Sk = S + (size_t)k * n * n, *Kk = Kg + (size_t)k * m * n, *L = Lf + (size_t)k * m * m;
            // T1 = Sn A (n x n), T2 = Sn B (n x m)
            for (int e = tid; e < n * n; e += TA) {
                const int i = e / n, j = e - i * n;
                double acc = 0.0;
                for (int l = 0; l < n; ++l) acc += Sn[i * n + l] * A[l * n + j];
                T1[e] = acc;
            }
            for (int e = tid; e < n * m; e += TA) {
                const int i = e / m, j = e - i * m;
                double acc = 0.0;
                for (int l = 0; l < n; ++l) acc += Sn[i * n + l] * Bm[l * m + j];
                T2[e] = acc;
            }
            __syncthreads();
            // Quu -> L, Qux, Qxx -> Sk
            for (int e = tid; e < m * m; e += TA) {
                const int i = e / m, j = e - i * m;
                double acc = (i == j) ? dt * P.R[i] : 0.0;
                for (int l = 0; l < n; ++l) acc += Bm[l * m + i] * T2[l * m + j];
                L[e] = acc;
            }
            for (int e = tid; e < m * n; e += TA) {
                const int i = e / n, j = e - i * n;
                double acc = 0.0;
                for (int l = 0; l < n; ++l) acc += Bm[l * m + i] * T1[l * n + j];
                Qux[e] = acc;
            }
            for (int e = tid; e < n * n; e += TA) {
                const int i = e / n, j = e - i * n;
                double acc = (i == j) ? dt * P.Q[i] : 0.0;
                for (int l = 0; l < n; ++l) acc += A[l * n + i] * T1[l * n + j];
                Sk[e] = acc;
            }
            __syncthreads();
            add_GtG(k, ALTRO_CONTROL, L, m);
            add_GtG(k, ALTRO_STATE, Sk, n);
            // in-place LDL' of Quu (lower triangle: unit L below the diagonal, D on it), thread 0: m is small
            if (tid == 0) {
                for (int j = 0; j < m; ++j) {
                    double dj = L[j * m + j];
                    for (int l = 0; l < j; ++l) dj -= L[j * m + l] * L[j * m + l] * L[l * m + l];
                    L[j * m + j] = dj;
                    for (int i = j + 1; i < m; ++i) {
                        double v = L[i * m + j];
                        for (int l = 0; l < j; ++l) v -= L[i * m + l] * L[j * m + l] * L[l * m + l];
                        L[i * m + j] = v / dj;
                    }
                }
            }
            __syncthreads();
            // K = -Quu^-1 Qux, one column per thread
            for (int c = tid; c < n; c += TA) {
                for (int i = 0; i < m; ++i) {
                    double v = -Qux[i * n + c];
                    for (int l = 0; l < i; ++l) v -= L[i * m + l] * Kk[l * n + c];
                    Kk[i * n + c] = v;
                }
                for (int i = 0; i < m; ++i) Kk[i * n + c] /= L[i * m + i];
                for (int i = m - 1; i >= 0; --i) {
                    double v = Kk[i * n + c];
                    for (int l = i + 1; l < m; ++l) v -= L[l * m + i] * Kk[l * n + c];
                    Kk[i * n + c] = v;
                }
            }
            __syncthreads();
            // S_k = Qxx + Qux' K, symmetrised
            for (int e = tid; e < n * n; e += TA) {
                const int i = e / n, j = e - i * n;
                double acc = Sk[e];
                for (int l = 0; l < m; ++l) acc += Qux[l * n + i] * Kk[l * n + j];
                T1[e] = acc;
            }
            __syncthreads();
            for (int e = tid; e < n * n; e += TA) Sk[e] = 0.5 * (T1[e] + T1[(e % n) * n + e / n]);
            __syncthreads();
        }
    };
    factor();

    // ---- iterations
    int it = 0;
    double rp = 0.0, rd = 0.0;
    for (it = 1; it <= P.max_iter; ++it) {
        // y-step and w-step on the current (X, U); residuals
        double lrp = 0.0, lrd = 0.0;
        for (int item = tid; item < ncon * N; item += TA) {
            const int ci = item / N, k = item - ci * N;
            const ConDesc &c = P.con[ci];
            if (k < c.k0 || k >= c.k1) continue;
            const size_t di = a.con_idx(c, k);
            const double *G = c.G + di * c.p * c.w, *h = c.h + di * c.p;
            const double *z = c.side == ALTRO_STATE ? X + k * n : U + k * m;
            const int lo = c.dual_off + (k - c.k0) * c.p;
            double cv[PMAX_ADMM], yn[PMAX_ADMM];
            for (int r0 = 0; r0 < c.p; r0 += PMAX_ADMM) {  // row-sparse (bound) blocks may have more than PMAX rows: chunks
                const int pr = min(PMAX_ADMM, c.p - r0);
                for (int r = 0; r < pr; ++r) {
                    double acc = h[r0 + r];
                    for (int j = 0; j < c.w; ++j) acc += Gat(G, c, r0 + r, j) * z[c.inds[j]];
                    cv[r] = acc;
                }
                if (c.sense == ALTRO_EQUALITY) {
                    for (int r = 0; r < pr; ++r) yn[r] = 0.0;
                } else if (c.sense == ALTRO_INEQUALITY) {
                    for (int r = 0; r < pr; ++r) yn[r] = fmin(cv[r] + w[lo + r0 + r], 0.0);
                } else {  // cones have p <= PMAX rows: a single chunk
                    double a2 = 0.0;
                    for (int r = 0; r < pr - 1; ++r) { const double v = cv[r] + w[lo + r]; a2 += v * v; }
                    const double t = cv[pr - 1] + w[lo + pr - 1], an = sqrt(a2);
                    if (an <= -t) { for (int r = 0; r < pr; ++r) yn[r] = 0.0; }
                    else if (an <= t) { for (int r = 0; r < pr; ++r) yn[r] = cv[r] + w[lo + r]; }
                    else {
                        const double cf = 0.5 * (1.0 + t / an);
                        for (int r = 0; r < pr - 1; ++r) yn[r] = cf * (cv[r] + w[lo + r]);
                        yn[pr - 1] = cf * an;
                    }
                }
                for (int r = 0; r < pr; ++r) {
                    lrp = fmax(lrp, fabs(cv[r] - yn[r]));
                    lrd = fmax(lrd, rho * fabs(yn[r] - y[lo + r0 + r]));
                    w[lo + r0 + r] += cv[r] - yn[r];
                    y[lo + r0 + r] = yn[r];
                }
            }
        }
        // block max of the residuals
        for (int o = 16; o > 0; o >>= 1) {
            lrp = fmax(lrp, __shfl_xor_sync(0xffffffffu, lrp, o));
            lrd = fmax(lrd, __shfl_xor_sync(0xffffffffu, lrd, o));
        }
        if ((tid & 31) == 0) { red[(tid >> 5) * 2] = lrp; red[(tid >> 5) * 2 + 1] = lrd; }
        __syncthreads();
        rp = fmax(red[0], red[2]);
        rd = fmax(red[1], red[3]);
        __syncthreads();
        if (it > 1 && rp < P.eps && rd < P.eps) break;
        if (P.adapt > 0 && it % P.adapt == 0 && it > 1) {  // residual balancing; w is the scaled dual lambda / rho
            double f = 1.0;
            if (rp > 10.0 * rd && rho < 1e6 * P.rho) f = 2.0;
            else if (rd > 10.0 * rp && rho > 1e-6 * P.rho) f = 0.5;
            if (f != 1.0) {
                rho *= f;
                for (int i = tid; i < P.Pd; i += TA) w[i] /= f;
                __syncthreads();
                factor();
            }
        }
        // z-step, affine backward sweep: s_N, then (df_k, s_k)
        double *sN = sv + (size_t)(N - 1) * n;
        for (int i = tid; i < n; i += TA) sN[i] = -P.Qf[i] * xr[(N - 1) * n + i];
        __syncthreads();
        add_Gtr(N - 1, ALTRO_STATE, sN);
        for (int k = N - 2; k >= 0; --k) {
            const double *A = a.A(k), *Bm = a.B(k), *dd = a.d(k), *Sn = S + (size_t)(k + 1) * n * n, *sn = sv + (size_t)(k + 1) * n;
            const double *Kk = Kg + (size_t)k * m * n, *L = Lf + (size_t)k * m * m;
            double *sk = sv + (size_t)k * n, *dk = df + (size_t)k * m;
            for (int i = tid; i < n; i += TA) {  // tv = s_{k+1} + S_{k+1} d_k
                double acc = sn[i];
                for (int l = 0; l < n; ++l) acc += Sn[i * n + l] * dd[l];
                tv[i] = acc;
            }
            __syncthreads();
            for (int i = tid; i < m; i += TA) {
                double acc = -dt * P.R[i] * ur[k * m + i];
                for (int l = 0; l < n; ++l) acc += Bm[l * m + i] * tv[l];
                vu[i] = acc;
            }
            for (int i = tid; i < n; i += TA) {
                double acc = -dt * P.Q[i] * xr[k * n + i];
                for (int l = 0; l < n; ++l) acc += A[l * n + i] * tv[l];
                vx[i] = acc;
            }
            __syncthreads();
            add_Gtr(k, ALTRO_CONTROL, vu);
            add_Gtr(k, ALTRO_STATE, vx);
            if (tid == 0) {  // df = -Quu^-1 Qu
                for (int i = 0; i < m; ++i) {
                    double v = -vu[i];
                    for (int l = 0; l < i; ++l) v -= L[i * m + l] * dk[l];
                    dk[i] = v;
                }
                for (int i = 0; i < m; ++i) dk[i] /= L[i * m + i];
                for (int i = m - 1; i >= 0; --i) {
                    double v = dk[i];
                    for (int l = i + 1; l < m; ++l) v -= L[l * m + i] * dk[l];
                    dk[i] = v;
                }
            }
            __syncthreads();
            for (int i = tid; i < n; i += TA) {  // s_k = Qx + K' Qu
                double acc = vx[i];
                for (int l = 0; l < m; ++l) acc += Kk[l * n + i] * vu[l];
                sk[i] = acc;
            }
            __syncthreads();
        }
        // rollout
        for (int k = 0; k < N - 1; ++k) {
            const double *A = a.A(k), *Bm = a.B(k), *dd = a.d(k), *Kk = Kg + (size_t)k * m * n;
            for (int i = tid; i < m; i += TA) {
                double acc = df[k * m + i];
                for (int l = 0; l < n; ++l) acc += Kk[i * n + l] * X[k * n + l];
                U[k * m + i] = acc;
            }
            __syncthreads();
            for (int i = tid; i < n; i += TA) {
                double acc = dd[i];
                for (int l = 0; l < n; ++l) acc += A[i * n + l] * X[k * n + l];
                for (int l = 0; l < m; ++l) acc += Bm[i * m + l] * U[k * m + l];
                X[(k + 1) * n + i] = acc;
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < N * n; i += TA) P.Xout[(size_t)inst * N * n + i] = X[i];
    for (int i = tid; i < (N - 1) * m; i += TA) P.Uout[(size_t)inst * (N - 1) * m + i] = U[i];
    if (tid == 0) {
        P.iters[inst] = it > P.max_iter ? P.max_iter : it;
        P.rprim[inst] = rp;
        P.rdual[inst] = rd;
    }
}

}  // namespace

size_t admm_workspace_doubles(int n, int m, int N, int Pd)
{
    const size_t mx = n > m ? n : m;
    return (size_t)N * n * n + (size_t)(N - 1) * m * n + (size_t)(N - 1) * m * m + (size_t)N * n + (size_t)(N - 1) * m +
           (size_t)N * n + (size_t)(N - 1) * m + 2 * (size_t)Pd + (size_t)n * n + (size_t)n * mx + (size_t)m * n;
}

cudaError_t admm_launch(const AdmmParams &P, cudaStream_t stream)
{
    const size_t smem = (size_t)(2 * P.n + P.m + 4) * sizeof(double);
    admm_kernel<<<P.B, TA, smem, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace altro
