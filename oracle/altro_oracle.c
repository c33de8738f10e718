/*
 * altro_oracle.c -- CPU restatement of the ALTRO AL-iLQR solve path.  See altro_oracle.h:
 * TEST INFRASTRUCTURE ONLY, PARITY UNPINNED bit for bit (un-vendored Altro.jl 0.2.0@socp /
 * TrajectoryOptimization.jl 0.3.2@socp; algorithm per SURVEY.md Appendix A).  What is pinned: the iteration
 * statistics of the reference's own saved benchmark runs (tests/golden/reference_stats.json, extracted from
 * benchmarks/**.jld2 by tests/golden/extract_reference_stats.py; tests/test_oracle_solutions.py).
 *
 * One instance is solved at a time by plain scalar loops; orc_solve_batch distributes
 * instances over pthreads pulling from a shared counter (the analogue of Threads.@threads over the batch).
 */
#include "altro_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>

/* optional per-iteration log, same columns as the product's altro_set_trace */
static double *orc_trace_buf = NULL;
static int orc_trace_rows = 0;
void orc_set_trace(double *buf, int rows) { orc_trace_buf = buf; orc_trace_rows = rows; }

void orc_default_opts(orc_opts_t *o)
{
    /* Altro.SolverOptions defaults */
    o->constraint_tolerance = 1e-6;
    o->cost_tolerance = 1e-4;
    o->cost_tolerance_intermediate = 1e-4;
    o->gradient_tolerance = 10.0;
    o->gradient_tolerance_intermediate = 1.0;
    o->penalty_initial = 1.0;
    o->penalty_scaling = 10.0;
    o->penalty_max = 1e8;
    o->dual_max = 1e8;
    o->line_search_lower_bound = 1e-8;
    o->line_search_upper_bound = 10.0;
    o->max_cost_value = 1e8;
    o->max_state_value = 1e8;
    o->bp_reg_initial = 0.0;
    o->bp_reg_increase_factor = 1.6;
    o->bp_reg_max = 1e8;
    o->bp_reg_min = 1e-8;
    o->bp_reg_fp = 10.0;
    o->iterations = 1000;
    o->iterations_inner = 300;
    o->iterations_outer = 30;
    o->iterations_linesearch = 20;
    o->dJ_counter_limit = 10;
    o->reset_duals = 1;
    o->reset_penalties = 1;
    o->kickout_max_penalty = 0;
    o->dj_zero_converges = 1;
    o->soc_hess_exact = 1;
    o->soc_viol_proj = 1;
    o->first_step_unconditional = 1;
}

int orc_dual_len(const orc_problem_t *pb)
{
    int P = 0;
    for (int c = 0; c < pb->ncon; ++c) P += (pb->con[c].k1 - pb->con[c].k0) * pb->con[c].p;
    return P;
}

/* ---------------------------------------------------------------- cones (A.3) */

/* Projection onto the second-order cone {(v,t): ||v|| <= t}, scalar last. */
void orc_soc_project(int p, const double *v, double *out)
{
    double a = 0.0, t = v[p - 1];
    for (int i = 0; i < p - 1; ++i) a += v[i] * v[i];
    a = sqrt(a);
    if (a <= -t) {
        for (int i = 0; i < p; ++i) out[i] = 0.0;
    } else if (a <= t) {
        for (int i = 0; i < p; ++i) out[i] = v[i];
    } else {
        double c = 0.5 * (1.0 + t / a);
        for (int i = 0; i < p - 1; ++i) out[i] = c * v[i];
        out[p - 1] = c * a;
    }
}

/* Jacobian of the projection (symmetric p x p). */
void orc_soc_project_jac(int p, const double *v, double *J)
{
    double a = 0.0, t = v[p - 1];
    for (int i = 0; i < p - 1; ++i) a += v[i] * v[i];
    a = sqrt(a);
    memset(J, 0, sizeof(double) * (size_t)p * p);
    if (a <= -t) return;
    if (a <= t) {
        for (int i = 0; i < p; ++i) J[i * p + i] = 1.0;
        return;
    }
    double c = 0.5 * (1.0 + t / a);
    double b = 0.5 * t / (a * a * a);
    for (int i = 0; i < p - 1; ++i) {
        for (int j = 0; j < p - 1; ++j) J[i * p + j] = -b * v[i] * v[j];
        J[i * p + i] += c;
        J[i * p + (p - 1)] = 0.5 * v[i] / a;
        J[(p - 1) * p + i] = 0.5 * v[i] / a;
    }
    J[(p - 1) * p + (p - 1)] = 0.5;
}

/* ---------------------------------------------------------------- workspace */

/* Arithmetic contract.  Every floating-point expression below is written in the same operation
 * order as the CUDA kernel (altro_mpc_icra2021_b200/csrc/altro_kernels.cuh) with fused multiply-adds
 * only where fma() is spelled out; this file is built with -ffp-contract=off and the kernel with
 * -fmad=false, so both sides evaluate identical IEEE-754 sequences and the parity tests can demand
 * bit-identical trajectories and iteration counts.  This matters because AL-iLQR on conic problems
 * is a semismooth Newton method: at a point that lies on a cone boundary to the last bit (a warm
 * start copied from an active reference) the Hessian branch is decided by one ulp. */

typedef struct {
    int n, m, N, P, EX, pmax, wmax, ncon;
    double *X, *U, *Xb, *Ub, *K, *dv;
    double *S, *s, *SA, *SB, *Qxx, *Qux, *Quu, *L, *Qx, *Qu, *T1, *t1, *ldiag;
    double *mu, *ex, *itm;
    int *off, *ex_off, *ex_stride, *rowsparse, *rs_col;
    double *rs_coef;
    int *rs_base;
} ws_t;

static double *dalloc(size_t k) { return (double *)calloc(k ? k : 1, sizeof(double)); }
static int *ialloc(size_t k) { return (int *)calloc(k ? k : 1, sizeof(int)); }

static ws_t *ws_new(const orc_problem_t *pb)
{
    ws_t *w = (ws_t *)calloc(1, sizeof(ws_t));
    int n = pb->n, m = pb->m, N = pb->N, nc = pb->ncon;
    w->n = n; w->m = m; w->N = N; w->ncon = nc;
    w->off = ialloc(nc + 1); w->ex_off = ialloc(nc + 1); w->ex_stride = ialloc(nc + 1);
    w->rowsparse = ialloc(nc + 1); w->rs_base = ialloc(nc + 1);
    int P = 0, EX = 0, rows = 0;
    for (int c = 0; c < nc; ++c) rows += pb->con[c].p;
    w->rs_col = ialloc(rows); w->rs_coef = dalloc(rows);
    rows = 0;
    for (int c = 0; c < nc; ++c) {
        const orc_con_t *cc = &pb->con[c];
        w->off[c] = P;
        w->ex_off[c] = EX;
        w->rs_base[c] = rows;
        /* row-sparse blocks (bounds): shared data, not a cone, every row has at most one nonzero */
        int rs = !cc->per_knot && !cc->per_instance && !cc->track && cc->sense != ORC_SOC;
        for (int r = 0; r < cc->p && rs; ++r) {
            int nz = 0;
            for (int j = 0; j < cc->w; ++j)
                if (cc->G[r * cc->w + j] != 0.0) { ++nz; w->rs_col[rows + r] = j; w->rs_coef[rows + r] = cc->G[r * cc->w + j]; }
            rs = nz <= 1;
        }
        w->rowsparse[c] = rs;
        w->ex_stride[c] = rs ? 2 * cc->w : cc->w + cc->w * cc->w;
        P += (cc->k1 - cc->k0) * cc->p;
        EX += (cc->k1 - cc->k0) * w->ex_stride[c];
        rows += cc->p;
    }
    w->P = P; w->EX = EX;
    w->X = dalloc((size_t)N * n); w->Xb = dalloc((size_t)N * n);
    w->U = dalloc((size_t)(N - 1) * m); w->Ub = dalloc((size_t)(N - 1) * m);
    w->K = dalloc((size_t)(N - 1) * m * n); w->dv = dalloc((size_t)(N - 1) * m);
    w->S = dalloc((size_t)n * n); w->s = dalloc(n);
    w->SA = dalloc((size_t)n * n); w->SB = dalloc((size_t)n * m);
    w->Qxx = dalloc((size_t)n * n); w->Qux = dalloc((size_t)m * n);
    w->Quu = dalloc((size_t)m * m); w->L = dalloc((size_t)m * m);
    w->Qx = dalloc(n); w->Qu = dalloc(m);
    w->T1 = dalloc((size_t)m * n); w->t1 = dalloc(m); w->ldiag = dalloc(m);
    w->mu = dalloc(nc); w->ex = dalloc(EX); w->itm = dalloc((size_t)N * (1 + nc));
    return w;
}

static void ws_free(ws_t *w)
{
    free(w->X); free(w->Xb); free(w->U); free(w->Ub); free(w->K); free(w->dv);
    free(w->S); free(w->s); free(w->SA); free(w->SB); free(w->Qxx); free(w->Qux);
    free(w->Quu); free(w->L); free(w->Qx); free(w->Qu); free(w->T1); free(w->t1); free(w->ldiag);
    free(w->mu); free(w->ex); free(w->itm);
    free(w->off); free(w->ex_off); free(w->ex_stride); free(w->rowsparse); free(w->rs_base);
    free(w->rs_col); free(w->rs_coef); free(w);
}

/* Canonical sum: 32 strided partials, then an xor-butterfly (the kernel's csum). */
static double csum(const double *arr, int count)
{
    double v[32], t[32];
    for (int l = 0; l < 32; ++l) {
        double a = 0.0;
        for (int i = l; i < count; i += 32) a += arr[i];
        v[l] = a;
    }
    for (int o = 16; o > 0; o >>= 1) {
        for (int l = 0; l < 32; ++l) t[l] = v[l] + v[l ^ o];
        memcpy(v, t, sizeof v);
    }
    return v[0];
}

/* ---------------------------------------------------------------- data access */

/* current MPC step of the instance being solved by this thread (gait-scheduled models only) */
static _Thread_local int orc_step = 0;

static void dyn_ptrs(const orc_problem_t *pb, int inst, int k, const double **A, const double **Bm,
                     const double **d)
{
    size_t idx = 0;
    if (pb->sched) {
        int s0 = orc_step < pb->sched_len - pb->N ? orc_step : pb->sched_len - pb->N;
        idx = (size_t)inst * pb->dyn_slots + (size_t)pb->sched[(size_t)inst * pb->sched_len + s0 + k];
    } else {
        if (pb->dyn_per_instance) idx = (size_t)inst * (pb->dyn_per_knot ? (size_t)(pb->N - 1) : 1);
        if (pb->dyn_per_knot) idx += (size_t)k;
    }
    *A = pb->A + idx * pb->n * pb->n;
    *Bm = pb->Bm + idx * pb->n * pb->m;
    *d = pb->d + idx * pb->n;
}

/* position of the instance being solved by this thread on the shared timelines */
static _Thread_local int orc_kcur = 0;

static void con_ptrs(const orc_con_t *c, int inst, int k, const double **G, const double **h)
{
    size_t idx = 0;
    if (c->track) {
        int r = orc_kcur + k;
        idx = (size_t)(r < c->track - 1 ? r : c->track - 1);
        *G = c->G + idx * c->p * c->w;
        *h = c->h + idx * c->p;
        return;
    }
    if (c->per_instance) idx = (size_t)inst * (c->per_knot ? (size_t)(c->k1 - c->k0) : 1);
    if (c->per_knot) idx += (size_t)(k - c->k0);
    *G = c->G + idx * c->p * c->w;
    *h = c->h + idx * c->p;
}

/* Row r of c = G z[inds] + h  (TO.evaluate) */
static double row_value(const orc_con_t *c, const double *G, const double *h, const double *z, int r)
{
    double acc = h[r];
    const double *g = G + r * c->w;
    for (int j = 0; j < c->w; ++j) acc = fma(g[j], z[c->inds[j]], acc);
    return acc;
}

/* x+ = A x + B u + d  (discrete_dynamics of an affine RD.LinearModel) */
static void dyn_step(int n, int m, const double *A, const double *Bm, const double *d, const double *x,
                     const double *u, double *xn)
{
    for (int i = 0; i < n; ++i) {
        double acc = d[i];
        for (int j = 0; j < n; ++j) acc = fma(A[i * n + j], x[j], acc);
        for (int j = 0; j < m; ++j) acc = fma(Bm[i * m + j], u[j], acc);
        xn[i] = acc;
    }
}

/* ---------------------------------------------------------------- cost (A.2, A.3) */

/* Stage / terminal tracking cost: dt*(1/2 dx'Q dx + 1/2 du'R du), terminal 1/2 dx'Qf dx.
 * Centred form of TO's 1/2x'Qx+q'x+c with q=-Q xref, c=1/2 xref'Q xref (same value). */
static double stage_cost(const orc_problem_t *pb, int inst, int k, const double *x, const double *u)
{
    int n = pb->n, m = pb->m, N = pb->N;
    const double *xr = pb->xref + ((size_t)inst * N + k) * n;
    double J = 0.0;
    if (k == N - 1) {
        for (int i = 0; i < n; ++i) { double e = x[i] - xr[i]; J += 0.5 * pb->Qf[i] * e * e; }
        return J;
    }
    for (int i = 0; i < n; ++i) { double e = x[i] - xr[i]; J += 0.5 * pb->Q[i] * e * e; }
    const double *ur = pb->uref + ((size_t)inst * (N - 1) + k) * m;
    for (int i = 0; i < m; ++i) { double e = u[i] - ur[i]; J += 0.5 * pb->R[i] * e * e; }
    return J * pb->dt;
}

/* AL penalty term of block ci at knot k (Altro cost!(J, conval)). */
static double con_cost(const orc_problem_t *pb, ws_t *w, int inst, int ci, int k, const double *X, const double *U,
                       const double *lam)
{
    const orc_con_t *c = &pb->con[ci];
    if (k < c->k0 || k >= c->k1) return 0.0;
    const double *G, *h;
    con_ptrs(c, inst, k, &G, &h);
    const double *z = c->side == ORC_STATE ? X + (size_t)k * pb->n : U + (size_t)k * pb->m;
    const double *l = lam + w->off[ci] + (k - c->k0) * c->p;
    const double mu = w->mu[ci];
    double J = 0.0;
    if (c->sense == ORC_EQ) {
        for (int r = 0; r < c->p; ++r) {
            double v = row_value(c, G, h, z, r);
            J += l[r] * v + 0.5 * mu * v * v;
        }
    } else if (c->sense == ORC_INEQ) {
        for (int r = 0; r < c->p; ++r) {
            double v = row_value(c, G, h, z, r);
            int act = (v >= 0.0) || (l[r] > 0.0);
            J += l[r] * v + (act ? 0.5 * mu * v * v : 0.0);
        }
    } else {
        /* 1/(2mu) (||Pi(lam - mu c)||^2 - ||lam||^2), with ||Pi(v,t)||^2 in closed form */
        double a2 = 0.0, t = 0.0, nl = 0.0;
        for (int r = 0; r < c->p; ++r) {
            double lb = l[r] - mu * row_value(c, G, h, z, r);
            nl += l[r] * l[r];
            if (r < c->p - 1) a2 += lb * lb;
            else t = lb;
        }
        double a = sqrt(a2), np;
        if (a <= -t) np = 0.0;
        else if (a <= t) np = a2 + t * t;
        else np = 0.5 * (a + t) * (a + t);
        J = (np - nl) / (2.0 * mu);
    }
    return J;
}

static double al_cost(const orc_problem_t *pb, ws_t *w, int inst, const double *X, const double *U,
                      const double *lam)
{
    int N = pb->N, items = N * (1 + pb->ncon);
    for (int it = 0; it < items; ++it) {
        int k = it % N, j = it / N;
        w->itm[it] = (j == 0)
            ? stage_cost(pb, inst, k, X + (size_t)k * pb->n, k < N - 1 ? U + (size_t)k * pb->m : NULL)
            : con_cost(pb, w, inst, j - 1, k, X, U, lam);
    }
    return csum(w->itm, items);
}

static double objective_cost(const orc_problem_t *pb, ws_t *w, int inst, const double *X, const double *U)
{
    for (int k = 0; k < pb->N; ++k)
        w->itm[k] = stage_cost(pb, inst, k, X + (size_t)k * pb->n, k < pb->N - 1 ? U + (size_t)k * pb->m : NULL);
    return csum(w->itm, pb->N);
}

/* max_violation (A.3): eq |c|, ineq max(0,c), SOC distance to the cone. */
static double max_violation(const orc_problem_t *pb, const orc_opts_t *o, int inst, const double *X,
                            const double *U)
{
    double v = 0.0;
    for (int ci = 0; ci < pb->ncon; ++ci) {
        const orc_con_t *c = &pb->con[ci];
        for (int k = c->k0; k < c->k1; ++k) {
            const double *G, *h;
            con_ptrs(c, inst, k, &G, &h);
            const double *z = c->side == ORC_STATE ? X + (size_t)k * pb->n : U + (size_t)k * pb->m;
            if (c->sense == ORC_EQ) {
                for (int r = 0; r < c->p; ++r) v = fmax(v, fabs(row_value(c, G, h, z, r)));
            } else if (c->sense == ORC_INEQ) {
                for (int r = 0; r < c->p; ++r) v = fmax(v, row_value(c, G, h, z, r));
            } else {
                double a2 = 0.0, t = 0.0;
                for (int r = 0; r < c->p; ++r) {
                    double cv = row_value(c, G, h, z, r);
                    if (r < c->p - 1) a2 += cv * cv;
                    else t = cv;
                }
                double a = sqrt(a2);
                if (!o->soc_viol_proj) {
                    v = fmax(v, a - t);
                } else if (a <= -t) { /* projection is 0: distance = |c|_inf */
                    for (int r = 0; r < c->p; ++r) v = fmax(v, fabs(row_value(c, G, h, z, r)));
                } else if (a > t) { /* c - Pi(c) = ((1-cf) v, t - cf a) */
                    double cf = 0.5 * (1.0 + t / a);
                    for (int r = 0; r < c->p - 1; ++r) v = fmax(v, fabs((1.0 - cf) * row_value(c, G, h, z, r)));
                    v = fmax(v, fabs(t - cf * a));
                }
            }
        }
    }
    return v;
}

/* ---------------------------------------------------------------- expansion (A.3) */

/* Gradient g[w] and Hessian (w x w, or w diagonal entries for row-sparse blocks) of the AL term of
 * every (block, knot) into the expansion scratch.  SOC blocks use the closed form of mu G' dPi G:
 *   inside  (a <= t):  mu G'G,                                   g = -G' lb
 *   outside (a >  t):  mu (cx (Gv'Gv - q q') + 1/2 (q+gt)(q+gt)'), g = -cf a (q + gt)
 * with lb = lam - mu c = (v, t), a = |v|, q = Gv' v / a, gt = last row of G, cf = (1 + t/a)/2 and
 * cx = cf (exact Hessian) or cf^2 (Gauss-Newton); orc_soc_project_jac is the dense cross-check. */
static void expand_constraints(const orc_problem_t *pb, const orc_opts_t *o, ws_t *w, int inst, const double *lam)
{
    for (int ci = 0; ci < pb->ncon; ++ci) {
        const orc_con_t *c = &pb->con[ci];
        const double mu = w->mu[ci];
        const int wd = c->w, p = c->p;
        for (int k = c->k0; k < c->k1; ++k) {
            const double *G, *h;
            con_ptrs(c, inst, k, &G, &h);
            const double *z = c->side == ORC_STATE ? w->X + (size_t)k * pb->n : w->U + (size_t)k * pb->m;
            const double *l = lam + w->off[ci] + (k - c->k0) * p;
            double *g = w->ex + w->ex_off[ci] + (k - c->k0) * w->ex_stride[ci];
            double *H = g + wd;
            if (w->rowsparse[ci]) {
                const int *col = w->rs_col + w->rs_base[ci];
                const double *coef = w->rs_coef + w->rs_base[ci];
                for (int j = 0; j < 2 * wd; ++j) g[j] = 0.0;
                for (int r = 0; r < p; ++r) {
                    double v = row_value(c, G, h, z, r), cf = coef[r];
                    int act = c->sense == ORC_EQ || (v >= 0.0) || (l[r] > 0.0);
                    g[col[r]] += cf * (l[r] + (act ? mu * v : 0.0));
                    H[col[r]] += act ? cf * cf * mu : 0.0;
                }
            } else if (c->sense != ORC_SOC) {
                double y[ORC_MAX_W], D[ORC_MAX_W];
                for (int r = 0; r < p; ++r) {
                    double v = row_value(c, G, h, z, r);
                    int act = c->sense == ORC_EQ || (v >= 0.0) || (l[r] > 0.0);
                    y[r] = l[r] + (act ? mu * v : 0.0);
                    D[r] = act ? mu : 0.0;
                }
                for (int j = 0; j < wd; ++j) {
                    double acc = 0.0;
                    for (int r = 0; r < p; ++r) acc = fma(G[r * wd + j], y[r], acc);
                    g[j] = acc;
                }
                for (int i = 0; i < wd; ++i)
                    for (int j = i; j < wd; ++j) {
                        double acc = 0.0;
                        for (int r = 0; r < p; ++r) acc = fma(G[r * wd + i] * D[r], G[r * wd + j], acc);
                        H[i * wd + j] = acc;
                        H[j * wd + i] = acc;
                    }
            } else {
                double lb[ORC_MAX_W] = {0.0}, q[ORC_MAX_W];  /* a cone has p >= 1 rows; the initialiser only quiets -Wmaybe-uninitialized */
                double a2 = 0.0;
                for (int r = 0; r < p; ++r) {
                    lb[r] = l[r] - mu * row_value(c, G, h, z, r);
                    if (r < p - 1) a2 += lb[r] * lb[r];
                }
                const double t = lb[p - 1], a = sqrt(a2);
                const double *gt = G + (p - 1) * wd;
                if (a <= -t) {
                    for (int j = 0; j < wd + wd * wd; ++j) g[j] = 0.0;
                } else if (a <= t) {
                    for (int j = 0; j < wd; ++j) {
                        double acc = 0.0;
                        for (int r = 0; r < p; ++r) acc = fma(G[r * wd + j], lb[r], acc);
                        g[j] = -acc;
                    }
                    for (int i = 0; i < wd; ++i)
                        for (int j = i; j < wd; ++j) {
                            double acc = 0.0;
                            for (int r = 0; r < p; ++r) acc = fma(G[r * wd + i], G[r * wd + j], acc);
                            H[i * wd + j] = mu * acc;
                            H[j * wd + i] = mu * acc;
                        }
                } else {
                    const double ia = 1.0 / a, cf = 0.5 * (1.0 + t * ia);
                    const double cx = o->soc_hess_exact ? cf : cf * cf;
                    for (int j = 0; j < wd; ++j) {
                        double acc = 0.0;
                        for (int r = 0; r < p - 1; ++r) acc = fma(G[r * wd + j], lb[r], acc);
                        q[j] = acc * ia;
                    }
                    for (int i = 0; i < wd; ++i)
                        for (int j = i; j < wd; ++j) {
                            double gg = 0.0;
                            for (int r = 0; r < p - 1; ++r) gg = fma(G[r * wd + i], G[r * wd + j], gg);
                            double qi = q[i] + gt[i], qj = q[j] + gt[j];
                            double hv = mu * (cx * (gg - q[i] * q[j]) + 0.5 * qi * qj);
                            H[i * wd + j] = hv;
                            H[j * wd + i] = hv;
                        }
                    for (int j = 0; j < wd; ++j) g[j] = -cf * a * (q[j] + gt[j]);
                }
            }
        }
    }
}

/* Add the expansions of every block of `side` active at knot k into (vec, mat[ld x ld]). */
static void scatter_expansion(const orc_problem_t *pb, ws_t *w, int k, int side, double *vec, double *mat, int ld)
{
    for (int ci = 0; ci < pb->ncon; ++ci) {
        const orc_con_t *c = &pb->con[ci];
        if (c->side != side || k < c->k0 || k >= c->k1) continue;
        const double *g = w->ex + w->ex_off[ci] + (k - c->k0) * w->ex_stride[ci];
        const int wd = c->w;
        if (w->rowsparse[ci]) {
            for (int e = 0; e < wd; ++e) {
                int zi = c->inds[e];
                vec[zi] += g[e];
                mat[zi * ld + zi] += g[wd + e];
            }
        } else {
            for (int e = 0; e < wd; ++e) vec[c->inds[e]] += g[e];
            for (int i = 0; i < wd; ++i)
                for (int j = 0; j < wd; ++j) mat[c->inds[i] * ld + c->inds[j]] += g[wd + i * wd + j];
        }
    }
}

/* ---------------------------------------------------------------- backward pass (A.7) */

static void reg_increase(const orc_opts_t *o, double *rho, double *drho)
{
    *drho = fmax(*drho * o->bp_reg_increase_factor, o->bp_reg_increase_factor);
    *rho = fmax(*rho * *drho, o->bp_reg_min);
}

static void reg_decrease(const orc_opts_t *o, double *rho, double *drho)
{
    *drho = fmin(*drho / o->bp_reg_increase_factor, 1.0 / o->bp_reg_increase_factor);
    double r = *rho * *drho;
    *rho = (r > o->bp_reg_min) ? r : 0.0;
}

/* Returns 0 ok, 1 if Quu could not be made positive definite. dV[2] = expected cost change. */
static int backward_pass(const orc_problem_t *pb, const orc_opts_t *o, ws_t *w, int inst, const double *lam,
                         double *rho, double *drho, double dV[2])
{
    const int n = pb->n, m = pb->m, N = pb->N;
    const double dt = pb->dt;
    expand_constraints(pb, o, w, inst, lam);
restart:;
    double a1 = 0.0, a2 = 0.0;
    /* terminal cost-to-go: S = Qf + state-side AL Hessian, s = Qf (x - xref) + AL gradient */
    const double *xrN = pb->xref + ((size_t)inst * N + (N - 1)) * n;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) w->S[i * n + j] = (i == j) ? pb->Qf[i] : 0.0;
    for (int i = 0; i < n; ++i) w->s[i] = pb->Qf[i] * (w->X[(size_t)(N - 1) * n + i] - xrN[i]);
    scatter_expansion(pb, w, N - 1, ORC_STATE, w->s, w->S, n);
    for (int k = N - 2; k >= 0; --k) {
        const double *A, *Bm, *dd;
        dyn_ptrs(pb, inst, k, &A, &Bm, &dd);
        const double *xr = pb->xref + ((size_t)inst * N + k) * n;
        const double *ur = pb->uref + ((size_t)inst * (N - 1) + k) * m;
        /* SA = S A, SB = S B */
        for (int i = 0; i < n; ++i) {
            for (int j = 0; j < n; ++j) {
                double acc = 0.0;
                for (int l = 0; l < n; ++l) acc = fma(w->S[i * n + l], A[l * n + j], acc);
                w->SA[i * n + j] = acc;
            }
            for (int j = 0; j < m; ++j) {
                double acc = 0.0;
                for (int l = 0; l < n; ++l) acc = fma(w->S[i * n + l], Bm[l * m + j], acc);
                w->SB[i * m + j] = acc;
            }
        }
        /* cost expansion (diagonal LQR cost) + AL expansion */
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) w->Qxx[i * n + j] = (i == j) ? dt * pb->Q[i] : 0.0;
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) w->Quu[i * m + j] = (i == j) ? dt * pb->R[i] : 0.0;
        for (int i = 0; i < n; ++i) w->Qx[i] = dt * pb->Q[i] * (w->X[(size_t)k * n + i] - xr[i]);
        for (int i = 0; i < m; ++i) w->Qu[i] = dt * pb->R[i] * (w->U[(size_t)k * m + i] - ur[i]);
        scatter_expansion(pb, w, k, ORC_STATE, w->Qx, w->Qxx, n);
        scatter_expansion(pb, w, k, ORC_CONTROL, w->Qu, w->Quu, m);
        /* action-value expansion (_calc_Q!): Qxx += A'SA, Qux = B'SA, Quu += B'SB, Qx += A's, Qu += B's.
         * Every chain starts from the value it accumulates into and runs over l ascending: this is exactly
         * what the kernel's FP64 tensor-core tiles (mma.sync m8n8k4) compute. */
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                double acc = w->Qxx[i * n + j];
                for (int l = 0; l < n; ++l) acc = fma(A[l * n + i], w->SA[l * n + j], acc);
                w->Qxx[i * n + j] = acc;
            }
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < n; ++j) {
                double acc = 0.0;
                for (int l = 0; l < n; ++l) acc = fma(Bm[l * m + i], w->SA[l * n + j], acc);
                w->Qux[i * n + j] = acc;
            }
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) {
                double acc = w->Quu[i * m + j];
                for (int l = 0; l < n; ++l) acc = fma(Bm[l * m + i], w->SB[l * m + j], acc);
                w->Quu[i * m + j] = acc;
            }
        for (int i = 0; i < n; ++i) {
            double acc = w->Qx[i];
            for (int l = 0; l < n; ++l) acc = fma(A[l * n + i], w->s[l], acc);
            w->Qx[i] = acc;
        }
        for (int i = 0; i < m; ++i) {
            double acc = w->Qu[i];
            for (int l = 0; l < n; ++l) acc = fma(Bm[l * m + i], w->s[l], acc);
            w->Qu[i] = acc;
        }
        /* control regularisation (_bp_reg!, :control) and LDL' of Quu + rho I, kept as the un-normalised lower
         * factor X (L = X diag(r)) and reciprocal pivots r = 1/D: one division per column. */
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) w->L[i * m + j] = w->Quu[i * m + j] + ((i == j) ? *rho : 0.0);
        int bad = 0;
        for (int j = 0; j < m && !bad; ++j) {
            for (int i = j; i < m; ++i) {
                double acc = w->L[i * m + j];
                for (int l = 0; l < j; ++l) acc = fma(-w->L[i * m + l], w->L[j * m + l] * w->ldiag[l], acc);
                w->L[i * m + j] = acc;
            }
            double piv = w->L[j * m + j];
            if (!(piv > 0.0)) { bad = 1; break; }
            w->ldiag[j] = 1.0 / piv; /* reciprocal pivot r_j */
        }
        if (bad) {
            reg_increase(o, rho, drho);
            if (*rho > o->bp_reg_max) return 1;
            goto restart;
        }
        /* gains (_calc_gains!): [K | d] = -(Quu + rho I)^-1 [Qux | Qu]:
         * forward y = L^-1 b, z = D^-1 y = y r;  backward x_i = z_i - r_i sum_{q>i} X[q][i] x_q (q descending) */
        double *K = w->K + (size_t)k * m * n, *dv = w->dv + (size_t)k * m;
        for (int c = 0; c <= n; ++c) {
            double *b = (c < n) ? K + c : dv;
            const double *src = (c < n) ? w->Qux + c : w->Qu;
            const int st = (c < n) ? n : 1;
            for (int i = 0; i < m; ++i) {
                double acc = -src[i * st];
                for (int l = 0; l < i; ++l) acc = fma(-w->L[i * m + l], b[l * st] * w->ldiag[l], acc);
                b[i * st] = acc; /* y_i */
            }
            for (int i = 0; i < m; ++i) b[i * st] = b[i * st] * w->ldiag[i]; /* z_i */
            for (int i = m - 1; i >= 0; --i) {
                double acc2 = 0.0;
                for (int l = m - 1; l > i; --l) acc2 = fma(w->L[l * m + i], b[l * st], acc2);
                b[i * st] = fma(-w->ldiag[i], acc2, b[i * st]);
            }
        }
        /* cost-to-go (_calc_ctg!), unregularised Quu:  T1 = Quu K + Qux,  t1 = Quu d + Qu */
        for (int i = 0; i < m; ++i) {
            for (int j = 0; j < n; ++j) {
                double acc = w->Qux[i * n + j];
                for (int l = 0; l < m; ++l) acc = fma(w->Quu[i * m + l], K[l * n + j], acc);
                w->T1[i * n + j] = acc;
            }
            double acc = w->Qu[i];
            for (int l = 0; l < m; ++l) acc = fma(w->Quu[i * m + l], dv[l], acc);
            w->t1[i] = acc;
        }
        /* S' = Qxx + K'T1 + Qux'K, s = Qx + K't1 + Qux'd */
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                double acc = w->Qxx[i * n + j];
                for (int l = 0; l < m; ++l) acc = fma(K[l * n + i], w->T1[l * n + j], acc);
                for (int l = 0; l < m; ++l) acc = fma(w->Qux[l * n + i], K[l * n + j], acc);
                w->SA[i * n + j] = acc;
            }
        for (int i = 0; i < n; ++i) {
            double acc = w->Qx[i];
            for (int l = 0; l < m; ++l) acc = fma(K[l * n + i], w->t1[l], acc);
            for (int l = 0; l < m; ++l) acc = fma(w->Qux[l * n + i], dv[l], acc);
            w->s[i] = acc;
        }
        /* expected change: [d'Qu, 1/2 d'Quu d] */
        for (int i = 0; i < m; ++i) {
            a1 = fma(dv[i], w->Qu[i], a1);
            a2 = fma(0.5 * dv[i], w->t1[i] - w->Qu[i], a2);
        }
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) w->S[i * n + j] = 0.5 * (w->SA[i * n + j] + w->SA[j * n + i]);
    }
    dV[0] = a1;
    dV[1] = a2;
    reg_decrease(o, rho, drho);
    return 0;
}

/* ---------------------------------------------------------------- forward pass (A.8) */

/* Closed-loop rollout with step alpha into Xb,Ub; returns 0 if a state leaves the box, 2 if the trial reproduces
 * the current trajectory bit for bit (every smaller step then does too: |alpha d| is below half an ulp of u at every
 * knot, so dx stays exactly 0), else 1. */
static int rollout_alpha(const orc_problem_t *pb, const orc_opts_t *o, ws_t *w, int inst, double alpha)
{
    int n = pb->n, m = pb->m, N = pb->N, ok = 1, same = 1;
    memcpy(w->Xb, w->X, sizeof(double) * n); /* x0 */
    for (int k = 0; k < N - 1; ++k) {
        const double *A, *Bm, *dd;
        dyn_ptrs(pb, inst, k, &A, &Bm, &dd);
        const double *K = w->K + (size_t)k * m * n, *dv = w->dv + (size_t)k * m;
        double *xb = w->Xb + (size_t)k * n, *ub = w->Ub + (size_t)k * m;
        const double *x = w->X + (size_t)k * n, *u = w->U + (size_t)k * m;
        for (int i = 0; i < m; ++i) {
            double acc = fma(alpha, dv[i], u[i]);
            for (int j = 0; j < n; ++j) acc = fma(K[i * n + j], xb[j] - x[j], acc);
            ub[i] = acc;
            if (!(acc == u[i])) same = 0;
        }
        dyn_step(n, m, A, Bm, dd, xb, ub, xb + n);
        for (int i = 0; i < n; ++i) {
            if (!(fabs(xb[n + i]) <= o->max_state_value)) ok = 0;
            if (!(xb[n + i] == x[n + i])) same = 0;
        }
    }
    return ok ? (same ? 2 : 1) : 0;
}

/* Returns the accepted cost J; *trials counts the rollouts evaluated. */
static double forward_pass(const orc_problem_t *pb, const orc_opts_t *o, ws_t *w, int inst, const double *lam,
                           const double dV[2], double J_prev, double *rho, double *drho, int *trials)
{
    int n = pb->n, m = pb->m, N = pb->N;
    double J = INFINITY, alpha = 1.0, z = -1.0;
    int iter = 0;
    while ((z <= o->line_search_lower_bound || z > o->line_search_upper_bound) && J >= J_prev) {
        if (iter > o->iterations_linesearch) {
            /* line search failed: keep the old trajectory and regularise */
            memcpy(w->Xb, w->X, sizeof(double) * (size_t)N * n);
            memcpy(w->Ub, w->U, sizeof(double) * (size_t)(N - 1) * m);
            J = al_cost(pb, w, inst, w->Xb, w->Ub, lam);
            reg_increase(o, rho, drho);
            *rho += o->bp_reg_fp;
            break;
        }
        int ok = rollout_alpha(pb, o, w, inst, alpha);
        ++*trials;
        if (!ok) { ++iter; alpha *= 0.5; continue; }
        J = al_cost(pb, w, inst, w->Xb, w->Ub, lam);
        double expected = -alpha * (dV[0] + alpha * dV[1]);
        z = expected > 0.0 ? (J_prev - J) / expected : -1.0;
        ++iter;
        alpha *= 0.5;
        /* the trial reproduced the current trajectory bit for bit: every remaining (smaller) step would too and
         * fail the same test, so the search is over -- same result as running them, without the rollouts */
        if (ok == 2) iter = o->iterations_linesearch + 1;
    }
    return J;
}

/* ---------------------------------------------------------------- AL updates (A.3, A.5) */

static void dual_update(const orc_problem_t *pb, const orc_opts_t *o, ws_t *w, int inst, double *lam)
{
    for (int ci = 0; ci < pb->ncon; ++ci) {
        const orc_con_t *c = &pb->con[ci];
        const double mu = w->mu[ci];
        for (int k = c->k0; k < c->k1; ++k) {
            const double *G, *h;
            con_ptrs(c, inst, k, &G, &h);
            const double *z = c->side == ORC_STATE ? w->X + (size_t)k * pb->n : w->U + (size_t)k * pb->m;
            double *l = lam + w->off[ci] + (k - c->k0) * c->p;
            if (c->sense == ORC_EQ) {
                for (int r = 0; r < c->p; ++r)
                    l[r] = fmin(fmax(l[r] + mu * row_value(c, G, h, z, r), -o->dual_max), o->dual_max);
            } else if (c->sense == ORC_INEQ) {
                for (int r = 0; r < c->p; ++r)
                    l[r] = fmin(fmax(l[r] + mu * row_value(c, G, h, z, r), 0.0), o->dual_max);
            } else { /* lam <- Pi(lam - mu c) */
                double a2 = 0.0, t = 0.0;
                for (int r = 0; r < c->p; ++r) {
                    double lb = l[r] - mu * row_value(c, G, h, z, r);
                    l[r] = lb;
                    if (r < c->p - 1) a2 += lb * lb;
                    else t = lb;
                }
                double a = sqrt(a2);
                if (a <= -t) {
                    for (int r = 0; r < c->p; ++r) l[r] = 0.0;
                } else if (a > t) {
                    double cf = 0.5 * (1.0 + t / a);
                    for (int r = 0; r < c->p - 1; ++r) l[r] *= cf;
                    l[c->p - 1] = cf * a;
                }
            }
        }
    }
}

/* ---------------------------------------------------------------- iLQR (A.6) + AL loop (A.5) */

typedef struct {
    int iters, outer, status, trials;
    double J, cmax, pen_max;
} res_t;

static void solve_instance(const orc_problem_t *pb, const orc_opts_t *o, ws_t *w, int inst, double *lam,
                           res_t *r)
{
    int n = pb->n, m = pb->m, N = pb->N;
    r->iters = r->outer = r->trials = 0;
    r->status = ORC_UNSOLVED;
    r->cmax = INFINITY;
    r->pen_max = 0.0;
    if (o->reset_duals) memset(lam, 0, sizeof(double) * (size_t)w->P);
    for (int c = 0; c < pb->ncon; ++c) w->mu[c] = o->penalty_initial;
    double J = 0.0;
    for (int outer = 1; outer <= o->iterations_outer; ++outer) {
        r->outer = outer;
        int last = (outer == o->iterations_outer) || pb->ncon == 0;
        double ctol = last ? o->cost_tolerance : o->cost_tolerance_intermediate;
        double gtol = last ? o->gradient_tolerance : o->gradient_tolerance_intermediate;
        /* ---- iLQR solve! */
        double rho = o->bp_reg_initial, drho = 0.0;
        int dJ_zero = 0;
        memcpy(w->X, pb->x0 + (size_t)inst * n, sizeof(double) * n);
        for (int k = 0; k < N - 1; ++k) { /* open-loop rollout!(solver) */
            const double *A, *Bm, *dd;
            dyn_ptrs(pb, inst, k, &A, &Bm, &dd);
            dyn_step(n, m, A, Bm, dd, w->X + (size_t)k * n, w->U + (size_t)k * m, w->X + (size_t)(k + 1) * n);
        }
        double J_prev = al_cost(pb, w, inst, w->X, w->U, lam);
        J = J_prev;
        /* first_step_unconditional: the cost of the initial rollout is not a line-search reference -- the first
         * forward pass of every iLQR solve compares against +inf, so its full step is always taken and the first
         * iteration can never be the converged one (see altro_oracle.h and DESIGN.md section 2 for the evidence) */
        if (o->first_step_unconditional) J_prev = INFINITY;
        for (int it = 0; it < o->iterations_inner; ++it) {
            double dV[2];
            if (backward_pass(pb, o, w, inst, lam, &rho, &drho, dV)) { r->status = ORC_NOT_PD; break; }
            J = forward_pass(pb, o, w, inst, lam, dV, J_prev, &rho, &drho, &r->trials);
            if (J > o->max_cost_value || !(J == J)) { r->status = ORC_MAXIMUM_COST; break; }
            memcpy(w->X, w->Xb, sizeof(double) * (size_t)N * n);
            memcpy(w->U, w->Ub, sizeof(double) * (size_t)(N - 1) * m);
            double dJ = fabs(J - J_prev);
            J_prev = J;
            for (int k = 0; k < N - 1; ++k) { /* gradient_todorov! */
                double mx = 0.0;
                for (int i = 0; i < m; ++i)
                    mx = fmax(mx, fabs(w->dv[(size_t)k * m + i]) / (fabs(w->U[(size_t)k * m + i]) + 1.0));
                w->itm[k] = mx;
            }
            double grad = csum(w->itm, N - 1) / (double)(N - 1);
            r->iters++;
            dJ_zero = (dJ == 0.0) ? dJ_zero + 1 : 0;
            if (orc_trace_buf && r->iters <= orc_trace_rows) {
                double *tr = orc_trace_buf + ((size_t)inst * orc_trace_rows + (r->iters - 1)) * 10;
                tr[0] = outer; tr[1] = r->iters; tr[2] = J; tr[3] = dJ; tr[4] = grad; tr[5] = rho;
                tr[6] = dV[0]; tr[7] = dV[1]; tr[8] = r->trials; tr[9] = NAN;
            }
            int small = o->dj_zero_converges ? (dJ >= 0.0 && dJ < ctol) : (dJ > 0.0 && dJ < ctol);
            if (small && grad < gtol) { r->status = ORC_SOLVE_SUCCEEDED; break; }
            if (r->iters >= o->iterations) { r->status = ORC_MAX_ITERATIONS; break; }
            if (dJ_zero > o->dJ_counter_limit) { r->status = ORC_NO_PROGRESS; break; }
        }
        if (r->status > ORC_SOLVE_SUCCEEDED) break;
        /* ---- AL outer loop bookkeeping */
        r->cmax = max_violation(pb, o, inst, w->X, w->U);
        if (orc_trace_buf && r->iters >= 1 && r->iters <= orc_trace_rows)
            orc_trace_buf[((size_t)inst * orc_trace_rows + (r->iters - 1)) * 10 + 9] = r->cmax;
        r->pen_max = 0.0;
        for (int c = 0; c < pb->ncon; ++c) r->pen_max = fmax(r->pen_max, w->mu[c]);
        if (r->cmax < o->constraint_tolerance) break;
        if (o->kickout_max_penalty && r->pen_max >= o->penalty_max) break;
        dual_update(pb, o, w, inst, lam);
        for (int c = 0; c < pb->ncon; ++c) w->mu[c] = fmin(w->mu[c] * o->penalty_scaling, o->penalty_max);
        if (outer == o->iterations_outer) r->status = ORC_MAX_ITERATIONS_OUTER;
    }
    r->cmax = max_violation(pb, o, inst, w->X, w->U);
    if (r->status <= ORC_SOLVE_SUCCEEDED)
        r->status = (r->cmax < o->constraint_tolerance) ? ORC_SOLVE_SUCCEEDED : ORC_UNSOLVED;
    r->J = J;
}

typedef struct {
    const orc_problem_t *pb;
    const orc_opts_t *o;
    const orc_run_t *run; /* NULL: one plain solve per instance */
    int i1;
    atomic_int *next;
    double *X, *U, *lam;
    int *iters, *iters_outer, *status, *ls_trials;
    double *cost, *cost_al, *cmax, *pen_max;
    double *x0_log, *u0_log;
} job_t;

/* Warm-started MPC transition of one instance (the kernel's Ctx::transition): x0 <- x_1 + noise, reference
 * window advanced along the track, primal (controls) and dual shift_fill. Writes the instance's rows of the
 * problem's x0 / xref / uref arrays, like the device does. */
static void transition(const orc_problem_t *pb, const orc_run_t *run, ws_t *w, int inst, int st, double *lam)
{
    const int n = pb->n, m = pb->m, N = pb->N;
    double *x0 = (double *)pb->x0 + (size_t)inst * n;
    const double *xo = w->X + n;
    const double *z = run->noise ? run->noise + ((size_t)st * pb->B + inst) * n : NULL;
    double s0 = run->w1, s1 = run->w1;
    if (z) {
        if (run->noise_mode == 1) {
            double mx = 0.0;
            for (int i = 0; i < n; ++i) mx = fmax(mx, fabs(xo[i]));
            s0 = s1 = mx * run->w1;
        } else if (run->noise_mode == 2) {
            double a = 0.0, b = 0.0;
            for (int i = 0; i < n / 2; ++i) a += xo[i] * xo[i];
            for (int i = n / 2; i < n; ++i) b += xo[i] * xo[i];
            s0 = sqrt(a) * run->w1;
            s1 = sqrt(b) * run->w2;
        }
    }
    for (int i = 0; i < n; ++i) {
        double v = xo[i];
        if (z) v += z[i] * ((run->noise_mode == 2 && i >= n / 2) ? s1 : s0);
        x0[i] = v;
    }
    if (run->trackX) {
        const int k0 = run->kidx[inst] + st + 1;
        double *xr = (double *)pb->xref + (size_t)inst * N * n, *ur = (double *)pb->uref + (size_t)inst * (N - 1) * m;
        for (int i = 0; i < N * n; ++i) {
            int k = k0 + i / n;
            if (k > run->Nt - 1) k = run->Nt - 1;
            xr[i] = run->trackX[(size_t)k * n + i % n];
        }
        for (int i = 0; i < (N - 1) * m; ++i) {
            int k = k0 + i / m;
            if (k > run->Nt - 2) k = run->Nt - 2;
            ur[i] = run->trackU[(size_t)k * m + i % m];
        }
    }
    if (run->shift) {
        if (N > 2) memmove(w->U, w->U + m, sizeof(double) * (size_t)(N - 2) * m);
        double *l = lam;
        for (int c = 0; c < pb->ncon; ++c) {
            int nk = pb->con[c].k1 - pb->con[c].k0, p = pb->con[c].p;
            if (nk > 1) memmove(l, l + p, sizeof(double) * (size_t)(nk - 1) * p);
            l += (size_t)nk * p;
        }
    }
}

/* Worker: pulls instance ids from a shared counter (dynamic schedule, chunk 1). */
static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    const orc_problem_t *pb = j->pb;
    int n = pb->n, m = pb->m, N = pb->N, B = pb->B;
    ws_t *w = ws_new(pb);
    const int steps = j->run ? j->run->steps : 1;
    for (;;) {
        int i = atomic_fetch_add(j->next, 1);
        if (i >= j->i1) break;
        res_t r;
        double *l = j->lam + (size_t)i * w->P;
        memcpy(w->U, j->U + (size_t)i * (N - 1) * m, sizeof(double) * (size_t)(N - 1) * m);
        if (j->run) memcpy(w->X, j->X + (size_t)i * N * n, sizeof(double) * (size_t)N * n);
        for (int st = 0; st < steps; ++st) {
            orc_step = pb->step0 + (j->run ? st + 1 : 0);
            {
                const int *ki = j->run && j->run->kidx ? j->run->kidx : pb->kidx;
                orc_kcur = (ki ? ki[i] : 0) + (j->run ? st + 1 : 0);
            }
            if (j->run) {
                transition(pb, j->run, w, i, st, l);
                if (j->x0_log) memcpy(j->x0_log + ((size_t)st * B + i) * n, pb->x0 + (size_t)i * n, sizeof(double) * n);
            }
            solve_instance(pb, j->o, w, i, l, &r);
            if (j->run && j->u0_log) memcpy(j->u0_log + ((size_t)st * B + i) * m, w->U, sizeof(double) * m);
            const size_t at = (size_t)st * B + i;
            if (j->iters) j->iters[at] = r.iters;
            if (j->iters_outer) j->iters_outer[at] = r.outer;
            if (j->status) j->status[at] = r.status;
            if (j->ls_trials) j->ls_trials[at] = r.trials;
            if (j->cost) j->cost[at] = objective_cost(pb, w, i, w->X, w->U);
            if (j->cost_al) j->cost_al[at] = r.J;
            if (j->cmax) j->cmax[at] = r.cmax;
            if (j->pen_max) j->pen_max[at] = r.pen_max;
        }
        memcpy(j->X + (size_t)i * N * n, w->X, sizeof(double) * (size_t)N * n);
        memcpy(j->U + (size_t)i * (N - 1) * m, w->U, sizeof(double) * (size_t)(N - 1) * m);
    }
    ws_free(w);
    return NULL;
}

static int run_jobs(job_t *j, int i0, int nthreads)
{
    const orc_problem_t *pb = j->pb;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    for (int c = 0; c < pb->ncon; ++c)
        if (pb->con[c].w > ORC_MAX_W || pb->con[c].p > ORC_MAX_W) return -1;
    atomic_int next;
    atomic_init(&next, i0);
    j->next = &next;
    pthread_t th[256];
    for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, worker, j);
    worker(j);
    for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
    return 0;
}

int orc_solve_batch(const orc_problem_t *pb, const orc_opts_t *o, int i0, int i1, int nthreads, double *X,
                    double *U, double *lam, int *iters, int *iters_outer, int *status, int *ls_trials,
                    double *cost, double *cost_al, double *cmax, double *pen_max)
{
    job_t j = {pb, o, NULL, i1, NULL, X, U, lam, iters, iters_outer, status, ls_trials, cost, cost_al, cmax, pen_max,
               NULL, NULL};
    return run_jobs(&j, i0, nthreads);
}

int orc_mpc_run(const orc_problem_t *pb, const orc_opts_t *o, const orc_run_t *run, int nthreads, double *X,
                double *U, double *lam, int *iters, int *iters_outer, int *status, int *ls_trials, double *cost,
                double *cmax, double *x0_log, double *u0_log)
{
    job_t j = {pb, o, run, pb->B, NULL, X, U, lam, iters, iters_outer, status, ls_trials, cost, NULL, cmax, NULL,
               x0_log, u0_log};
    return run_jobs(&j, 0, nthreads);
}

void orc_shift_fill(const orc_problem_t *pb, int primal, int dual, double *X, double *U, double *lam)
{
    int n = pb->n, m = pb->m, N = pb->N, P = orc_dual_len(pb);
    for (int i = 0; i < pb->B; ++i) {
        if (primal) {
            double *x = X + (size_t)i * N * n, *u = U + (size_t)i * (N - 1) * m;
            memmove(x, x + n, sizeof(double) * (size_t)(N - 1) * n);
            if (N > 2) memmove(u, u + m, sizeof(double) * (size_t)(N - 2) * m);
        }
        if (dual) {
            double *l = lam + (size_t)i * P;
            for (int c = 0; c < pb->ncon; ++c) {
                int nk = pb->con[c].k1 - pb->con[c].k0, p = pb->con[c].p;
                if (nk > 1) memmove(l, l + p, sizeof(double) * (size_t)(nk - 1) * p);
                l += (size_t)nk * p;
            }
        }
    }
}

void orc_evaluate(const orc_problem_t *pb, const orc_opts_t *o, const double *X, const double *U, double *cost,
                  double *cmax)
{
    ws_t *w = ws_new(pb);
    for (int i = 0; i < pb->B; ++i) {
        orc_kcur = pb->kidx ? pb->kidx[i] : 0;
        const double *x = X + (size_t)i * pb->N * pb->n, *u = U + (size_t)i * (pb->N - 1) * pb->m;
        if (cost) cost[i] = objective_cost(pb, w, i, x, u);
        if (cmax) cmax[i] = max_violation(pb, o, i, x, u);
    }
    ws_free(w);
}
