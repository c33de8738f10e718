/*
 * altro_oracle.c -- CPU restatement of the ALTRO AL-iLQR solve path.  See altro_oracle.h:
 * TEST INFRASTRUCTURE ONLY, PARITY UNPINNED (un-vendored Altro.jl 0.2.0@socp /
 * TrajectoryOptimization.jl 0.3.2@socp; algorithm per SURVEY.md Appendix A).
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

typedef struct {
    int n, m, N, P, pmax, wmax;
    double *X, *U, *Xb, *Ub, *K, *dv;
    double *S, *s, *SA, *SB, *Qxx, *Qux, *Quu, *L, *Qx, *Qu, *T1, *t1;
    double *lxx, *luu;
    double *mu;
    double *cv, *y, *lp, *D, *DG, *g, *H;
    int *off;
} ws_t;

static double *dalloc(size_t k) { return (double *)calloc(k ? k : 1, sizeof(double)); }

static ws_t *ws_new(const orc_problem_t *pb)
{
    ws_t *w = (ws_t *)calloc(1, sizeof(ws_t));
    int n = pb->n, m = pb->m, N = pb->N;
    w->n = n; w->m = m; w->N = N;
    w->P = orc_dual_len(pb);
    w->pmax = 1; w->wmax = 1;
    w->off = (int *)calloc((size_t)pb->ncon + 1, sizeof(int));
    int P = 0;
    for (int c = 0; c < pb->ncon; ++c) {
        w->off[c] = P;
        P += (pb->con[c].k1 - pb->con[c].k0) * pb->con[c].p;
        if (pb->con[c].p > w->pmax) w->pmax = pb->con[c].p;
        if (pb->con[c].w > w->wmax) w->wmax = pb->con[c].w;
    }
    int pm = w->pmax, wm = w->wmax;
    w->X = dalloc((size_t)N * n); w->Xb = dalloc((size_t)N * n);
    w->U = dalloc((size_t)(N - 1) * m); w->Ub = dalloc((size_t)(N - 1) * m);
    w->K = dalloc((size_t)(N - 1) * m * n); w->dv = dalloc((size_t)(N - 1) * m);
    w->S = dalloc((size_t)n * n); w->s = dalloc(n);
    w->SA = dalloc((size_t)n * n); w->SB = dalloc((size_t)n * m);
    w->Qxx = dalloc((size_t)n * n); w->Qux = dalloc((size_t)m * n);
    w->Quu = dalloc((size_t)m * m); w->L = dalloc((size_t)m * m);
    w->Qx = dalloc(n); w->Qu = dalloc(m);
    w->T1 = dalloc((size_t)m * n); w->t1 = dalloc(m);
    w->lxx = dalloc((size_t)n * n); w->luu = dalloc((size_t)m * m);
    w->mu = dalloc(pb->ncon);
    w->cv = dalloc(pm); w->y = dalloc(pm); w->lp = dalloc(pm);
    w->D = dalloc((size_t)pm * pm); w->DG = dalloc((size_t)pm * wm);
    w->g = dalloc(wm); w->H = dalloc((size_t)wm * wm);
    return w;
}

static void ws_free(ws_t *w)
{
    free(w->X); free(w->Xb); free(w->U); free(w->Ub); free(w->K); free(w->dv);
    free(w->S); free(w->s); free(w->SA); free(w->SB); free(w->Qxx); free(w->Qux);
    free(w->Quu); free(w->L); free(w->Qx); free(w->Qu); free(w->T1); free(w->t1);
    free(w->lxx); free(w->luu); free(w->mu);
    free(w->cv); free(w->y); free(w->lp); free(w->D); free(w->DG); free(w->g); free(w->H);
    free(w->off); free(w);
}

/* ---------------------------------------------------------------- data access */

static void dyn_ptrs(const orc_problem_t *pb, int inst, int k, const double **A, const double **Bm,
                     const double **d)
{
    size_t idx = 0;
    if (pb->dyn_per_instance) idx = (size_t)inst * (pb->dyn_per_knot ? (size_t)(pb->N - 1) : 1);
    if (pb->dyn_per_knot) idx += (size_t)k;
    *A = pb->A + idx * pb->n * pb->n;
    *Bm = pb->Bm + idx * pb->n * pb->m;
    *d = pb->d + idx * pb->n;
}

static void con_ptrs(const orc_con_t *c, int inst, int k, const double **G, const double **h)
{
    size_t idx = 0;
    if (c->per_instance) idx = (size_t)inst * (c->per_knot ? (size_t)(c->k1 - c->k0) : 1);
    if (c->per_knot) idx += (size_t)(k - c->k0);
    *G = c->G + idx * c->p * c->w;
    *h = c->h + idx * c->p;
}

/* c = G z[inds] + h  (TO.evaluate) */
static void con_eval(const orc_con_t *c, const double *G, const double *h, const double *z, double *cv)
{
    for (int r = 0; r < c->p; ++r) {
        double acc = h[r];
        for (int j = 0; j < c->w; ++j) acc += G[r * c->w + j] * z[c->inds[j]];
        cv[r] = acc;
    }
}

/* x+ = A x + B u + d  (discrete_dynamics of an affine RD.LinearModel) */
static void dyn_step(int n, int m, const double *A, const double *Bm, const double *d, const double *x,
                     const double *u, double *xn)
{
    for (int i = 0; i < n; ++i) {
        double acc = d[i];
        for (int j = 0; j < n; ++j) acc += A[i * n + j] * x[j];
        for (int j = 0; j < m; ++j) acc += Bm[i * m + j] * u[j];
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
    const double *ur = pb->uref + ((size_t)inst * (N - 1) + k) * m;
    for (int i = 0; i < n; ++i) { double e = x[i] - xr[i]; J += 0.5 * pb->Q[i] * e * e; }
    for (int i = 0; i < m; ++i) { double e = u[i] - ur[i]; J += 0.5 * pb->R[i] * e * e; }
    return J * pb->dt;
}

/* AL penalty term of one constraint block at one knot. */
static double con_cost(const orc_con_t *c, const double *cv, const double *lam, double mu, double *lp)
{
    double J = 0.0;
    if (c->sense == ORC_EQ) {
        for (int r = 0; r < c->p; ++r) J += lam[r] * cv[r] + 0.5 * mu * cv[r] * cv[r];
    } else if (c->sense == ORC_INEQ) {
        for (int r = 0; r < c->p; ++r) {
            int act = (cv[r] >= 0.0) || (lam[r] > 0.0);
            J += lam[r] * cv[r] + (act ? 0.5 * mu * cv[r] * cv[r] : 0.0);
        }
    } else {
        double nl = 0.0, np = 0.0;
        double lb[ORC_MAX_W + 1];
        for (int r = 0; r < c->p; ++r) { lb[r] = lam[r] - mu * cv[r]; nl += lam[r] * lam[r]; }
        orc_soc_project(c->p, lb, lp);
        for (int r = 0; r < c->p; ++r) np += lp[r] * lp[r];
        J = (np - nl) / (2.0 * mu);
    }
    return J;
}

static double al_cost(const orc_problem_t *pb, ws_t *w, int inst, const double *X, const double *U,
                      const double *lam)
{
    int n = pb->n, m = pb->m, N = pb->N;
    double J = 0.0;
    for (int k = 0; k < N; ++k) {
        double Jk = stage_cost(pb, inst, k, X + (size_t)k * n, k < N - 1 ? U + (size_t)k * m : NULL);
        for (int c = 0; c < pb->ncon; ++c) {
            const orc_con_t *cc = &pb->con[c];
            if (k < cc->k0 || k >= cc->k1) continue;
            const double *G, *h;
            con_ptrs(cc, inst, k, &G, &h);
            const double *z = cc->side == ORC_STATE ? X + (size_t)k * n : U + (size_t)k * m;
            con_eval(cc, G, h, z, w->cv);
            Jk += con_cost(cc, w->cv, lam + w->off[c] + (k - cc->k0) * cc->p, w->mu[c], w->lp);
        }
        J += Jk;
    }
    return J;
}

static double objective_cost(const orc_problem_t *pb, int inst, const double *X, const double *U)
{
    double J = 0.0;
    for (int k = 0; k < pb->N; ++k)
        J += stage_cost(pb, inst, k, X + (size_t)k * pb->n, k < pb->N - 1 ? U + (size_t)k * pb->m : NULL);
    return J;
}

/* max_violation (A.3): eq |c|, ineq max(0,c), SOC distance to the cone. */
static double max_violation(const orc_problem_t *pb, const orc_opts_t *o, ws_t *w, int inst, const double *X,
                            const double *U)
{
    double v = 0.0;
    for (int c = 0; c < pb->ncon; ++c) {
        const orc_con_t *cc = &pb->con[c];
        for (int k = cc->k0; k < cc->k1; ++k) {
            const double *G, *h;
            con_ptrs(cc, inst, k, &G, &h);
            const double *z = cc->side == ORC_STATE ? X + (size_t)k * pb->n : U + (size_t)k * pb->m;
            con_eval(cc, G, h, z, w->cv);
            if (cc->sense == ORC_EQ) {
                for (int r = 0; r < cc->p; ++r) v = fmax(v, fabs(w->cv[r]));
            } else if (cc->sense == ORC_INEQ) {
                for (int r = 0; r < cc->p; ++r) v = fmax(v, w->cv[r]);
            } else if (o->soc_viol_proj) {
                orc_soc_project(cc->p, w->cv, w->lp);
                for (int r = 0; r < cc->p; ++r) v = fmax(v, fabs(w->cv[r] - w->lp[r]));
            } else {
                double a = 0.0;
                for (int r = 0; r < cc->p - 1; ++r) a += w->cv[r] * w->cv[r];
                v = fmax(v, sqrt(a) - w->cv[cc->p - 1]);
            }
        }
    }
    return v;
}

/* ---------------------------------------------------------------- expansion (A.3) */

/* Gradient g[w] and Hessian H[w][w] of one constraint block's AL term w.r.t. z[inds]. */
static void con_expand(const orc_opts_t *o, ws_t *w, const orc_con_t *c, const double *G, const double *cv,
                       const double *lam, double mu)
{
    int p = c->p, wd = c->w;
    double *D = w->D, *DG = w->DG, *y = w->y;
    memset(D, 0, sizeof(double) * (size_t)p * p);
    if (c->sense == ORC_EQ) {
        for (int r = 0; r < p; ++r) { y[r] = lam[r] + mu * cv[r]; D[r * p + r] = mu; }
    } else if (c->sense == ORC_INEQ) {
        for (int r = 0; r < p; ++r) {
            int act = (cv[r] >= 0.0) || (lam[r] > 0.0);
            y[r] = lam[r] + (act ? mu * cv[r] : 0.0);
            D[r * p + r] = act ? mu : 0.0;
        }
    } else {
        double lb[ORC_MAX_W + 1];
        for (int r = 0; r < p; ++r) lb[r] = lam[r] - mu * cv[r];
        orc_soc_project(p, lb, w->lp);
        for (int r = 0; r < p; ++r) y[r] = -w->lp[r];
        orc_soc_project_jac(p, lb, D);
        if (!o->soc_hess_exact) { /* Gauss-Newton: dPi' dPi */
            double *T = (double *)malloc(sizeof(double) * (size_t)p * p);
            for (int i = 0; i < p; ++i)
                for (int j = 0; j < p; ++j) {
                    double acc = 0.0;
                    for (int l = 0; l < p; ++l) acc += D[l * p + i] * D[l * p + j];
                    T[i * p + j] = acc;
                }
            memcpy(D, T, sizeof(double) * (size_t)p * p);
            free(T);
        }
        for (int i = 0; i < p * p; ++i) D[i] *= mu;
    }
    /* g = G' y ; H = G' D G */
    for (int j = 0; j < wd; ++j) {
        double acc = 0.0;
        for (int r = 0; r < p; ++r) acc += G[r * wd + j] * y[r];
        w->g[j] = acc;
    }
    for (int r = 0; r < p; ++r)
        for (int j = 0; j < wd; ++j) {
            double acc = 0.0;
            for (int l = 0; l < p; ++l) acc += D[r * p + l] * G[l * wd + j];
            DG[r * wd + j] = acc;
        }
    for (int i = 0; i < wd; ++i)
        for (int j = 0; j < wd; ++j) {
            double acc = 0.0;
            for (int r = 0; r < p; ++r) acc += G[r * wd + i] * DG[r * wd + j];
            w->H[i * wd + j] = acc;
        }
}

/* Cost + AL expansion at knot k: lx[n], lxx[n][n], lu[m], luu[m][m] (cost_expansion!). */
static void knot_expansion(const orc_problem_t *pb, const orc_opts_t *o, ws_t *w, int inst, int k,
                           const double *lam, double *lx, double *lu)
{
    int n = pb->n, m = pb->m, N = pb->N;
    const double *x = w->X + (size_t)k * n;
    const double *xr = pb->xref + ((size_t)inst * N + k) * n;
    int term = (k == N - 1);
    double sc = term ? 1.0 : pb->dt;
    memset(w->lxx, 0, sizeof(double) * (size_t)n * n);
    for (int i = 0; i < n; ++i) {
        double q = term ? pb->Qf[i] : pb->Q[i];
        lx[i] = sc * q * (x[i] - xr[i]);
        w->lxx[i * n + i] = sc * q;
    }
    if (!term) {
        const double *u = w->U + (size_t)k * m;
        const double *ur = pb->uref + ((size_t)inst * (N - 1) + k) * m;
        memset(w->luu, 0, sizeof(double) * (size_t)m * m);
        for (int i = 0; i < m; ++i) {
            lu[i] = sc * pb->R[i] * (u[i] - ur[i]);
            w->luu[i * m + i] = sc * pb->R[i];
        }
    }
    for (int c = 0; c < pb->ncon; ++c) {
        const orc_con_t *cc = &pb->con[c];
        if (k < cc->k0 || k >= cc->k1) continue;
        const double *G, *h;
        con_ptrs(cc, inst, k, &G, &h);
        const double *z = cc->side == ORC_STATE ? x : w->U + (size_t)k * m;
        con_eval(cc, G, h, z, w->cv);
        con_expand(o, w, cc, G, w->cv, lam + w->off[c] + (k - cc->k0) * cc->p, w->mu[c]);
        double *gv = cc->side == ORC_STATE ? lx : lu;
        double *Hm = cc->side == ORC_STATE ? w->lxx : w->luu;
        int ld = cc->side == ORC_STATE ? n : m;
        for (int i = 0; i < cc->w; ++i) {
            gv[cc->inds[i]] += w->g[i];
            for (int j = 0; j < cc->w; ++j) Hm[cc->inds[i] * ld + cc->inds[j]] += w->H[i * cc->w + j];
        }
    }
}

/* ---------------------------------------------------------------- backward pass (A.7) */

/* In-place lower Cholesky of the m x m matrix L; returns 0 on success. */
static int cholesky(int m, double *L)
{
    for (int j = 0; j < m; ++j) {
        double dsum = L[j * m + j];
        for (int l = 0; l < j; ++l) dsum -= L[j * m + l] * L[j * m + l];
        if (!(dsum > 0.0)) return 1;
        double dj = sqrt(dsum);
        L[j * m + j] = dj;
        for (int i = j + 1; i < m; ++i) {
            double acc = L[i * m + j];
            for (int l = 0; l < j; ++l) acc -= L[i * m + l] * L[j * m + l];
            L[i * m + j] = acc / dj;
        }
    }
    return 0;
}

/* Solve (L L') x = b in place. */
static void chol_solve(int m, const double *L, double *b, int stride)
{
    for (int i = 0; i < m; ++i) {
        double acc = b[i * stride];
        for (int l = 0; l < i; ++l) acc -= L[i * m + l] * b[l * stride];
        b[i * stride] = acc / L[i * m + i];
    }
    for (int i = m - 1; i >= 0; --i) {
        double acc = b[i * stride];
        for (int l = i + 1; l < m; ++l) acc -= L[l * m + i] * b[l * stride];
        b[i * stride] = acc / L[i * m + i];
    }
}

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
    int n = pb->n, m = pb->m, N = pb->N;
restart:
    dV[0] = dV[1] = 0.0;
    /* terminal cost-to-go */
    knot_expansion(pb, o, w, inst, N - 1, lam, w->s, NULL);
    memcpy(w->S, w->lxx, sizeof(double) * (size_t)n * n);
    for (int k = N - 2; k >= 0; --k) {
        const double *A, *Bm, *dd;
        dyn_ptrs(pb, inst, k, &A, &Bm, &dd);
        knot_expansion(pb, o, w, inst, k, lam, w->Qx, w->Qu);
        /* action-value expansion (_calc_Q!) */
        for (int i = 0; i < n; ++i) {
            double acc = w->Qx[i];
            for (int l = 0; l < n; ++l) acc += A[l * n + i] * w->s[l];
            w->Qx[i] = acc;
        }
        for (int i = 0; i < m; ++i) {
            double acc = w->Qu[i];
            for (int l = 0; l < n; ++l) acc += Bm[l * m + i] * w->s[l];
            w->Qu[i] = acc;
        }
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                double acc = 0.0;
                for (int l = 0; l < n; ++l) acc += w->S[i * n + l] * A[l * n + j];
                w->SA[i * n + j] = acc;
            }
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < m; ++j) {
                double acc = 0.0;
                for (int l = 0; l < n; ++l) acc += w->S[i * n + l] * Bm[l * m + j];
                w->SB[i * m + j] = acc;
            }
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                double acc = w->lxx[i * n + j];
                for (int l = 0; l < n; ++l) acc += A[l * n + i] * w->SA[l * n + j];
                w->Qxx[i * n + j] = acc;
            }
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) {
                double acc = w->luu[i * m + j];
                for (int l = 0; l < n; ++l) acc += Bm[l * m + i] * w->SB[l * m + j];
                w->Quu[i * m + j] = acc;
            }
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < n; ++j) {
                double acc = 0.0;
                for (int l = 0; l < n; ++l) acc += Bm[l * m + i] * w->SA[l * n + j];
                w->Qux[i * n + j] = acc;
            }
        /* control regularisation (_bp_reg!, bp_reg_type = :control) and gains (_calc_gains!) */
        memcpy(w->L, w->Quu, sizeof(double) * (size_t)m * m);
        for (int i = 0; i < m; ++i) w->L[i * m + i] += *rho;
        if (cholesky(m, w->L)) {
            reg_increase(o, rho, drho);
            if (*rho > o->bp_reg_max) return 1;
            goto restart;
        }
        double *K = w->K + (size_t)k * m * n, *dv = w->dv + (size_t)k * m;
        for (int i = 0; i < m * n; ++i) K[i] = -w->Qux[i];
        for (int i = 0; i < m; ++i) dv[i] = -w->Qu[i];
        for (int j = 0; j < n; ++j) chol_solve(m, w->L, K + j, n);
        chol_solve(m, w->L, dv, 1);
        /* cost-to-go (_calc_ctg!), unregularised Quu:  T1 = Quu K + Qux,  t1 = Quu d + Qu */
        for (int i = 0; i < m; ++i) {
            for (int j = 0; j < n; ++j) {
                double acc = w->Qux[i * n + j];
                for (int l = 0; l < m; ++l) acc += w->Quu[i * m + l] * K[l * n + j];
                w->T1[i * n + j] = acc;
            }
            double acc = w->Qu[i];
            for (int l = 0; l < m; ++l) acc += w->Quu[i * m + l] * dv[l];
            w->t1[i] = acc;
        }
        /* s = Qx + K'(Quu d + Qu) + Qux' d ;  S = Qxx + K'(Quu K + Qux) + Qux' K, symmetrised */
        for (int i = 0; i < n; ++i) {
            double acc = w->Qx[i];
            for (int l = 0; l < m; ++l) acc += K[l * n + i] * w->t1[l];
            for (int l = 0; l < m; ++l) acc += w->Qux[l * n + i] * dv[l];
            w->s[i] = acc;
        }
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                double acc = w->Qxx[i * n + j];
                for (int l = 0; l < m; ++l) acc += K[l * n + i] * w->T1[l * n + j];
                for (int l = 0; l < m; ++l) acc += w->Qux[l * n + i] * K[l * n + j];
                w->SA[i * n + j] = acc;
            }
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) w->S[i * n + j] = 0.5 * (w->SA[i * n + j] + w->SA[j * n + i]);
        /* expected change: [d'Qu, 1/2 d'Quu d] */
        for (int i = 0; i < m; ++i) {
            dV[0] += dv[i] * w->Qu[i];
            double acc = 0.0;
            for (int l = 0; l < m; ++l) acc += w->Quu[i * m + l] * dv[l];
            dV[1] += 0.5 * dv[i] * acc;
        }
    }
    reg_decrease(o, rho, drho);
    return 0;
}

/* ---------------------------------------------------------------- forward pass (A.8) */

/* Closed-loop rollout with step alpha into Xb,Ub; returns 0 if a state leaves the box. */
static int rollout_alpha(const orc_problem_t *pb, const orc_opts_t *o, ws_t *w, int inst, double alpha)
{
    int n = pb->n, m = pb->m, N = pb->N;
    memcpy(w->Xb, w->X, sizeof(double) * n); /* x0 */
    for (int k = 0; k < N - 1; ++k) {
        const double *A, *Bm, *dd;
        dyn_ptrs(pb, inst, k, &A, &Bm, &dd);
        const double *K = w->K + (size_t)k * m * n, *dv = w->dv + (size_t)k * m;
        double *xb = w->Xb + (size_t)k * n, *ub = w->Ub + (size_t)k * m;
        const double *x = w->X + (size_t)k * n, *u = w->U + (size_t)k * m;
        for (int i = 0; i < m; ++i) {
            double acc = u[i] + alpha * dv[i];
            for (int j = 0; j < n; ++j) acc += K[i * n + j] * (xb[j] - x[j]);
            ub[i] = acc;
        }
        dyn_step(n, m, A, Bm, dd, xb, ub, xb + n);
        double mx = 0.0;
        for (int i = 0; i < n; ++i) mx = fmax(mx, fabs(xb[n + i]));
        if (!(mx <= o->max_state_value)) return 0;
    }
    return 1;
}

/* Returns the accepted cost J; *trials counts rollouts tried. */
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
    }
    return J;
}

/* ---------------------------------------------------------------- AL updates (A.3, A.5) */

static void dual_update(const orc_problem_t *pb, const orc_opts_t *o, ws_t *w, int inst, double *lam)
{
    for (int c = 0; c < pb->ncon; ++c) {
        const orc_con_t *cc = &pb->con[c];
        double mu = w->mu[c];
        for (int k = cc->k0; k < cc->k1; ++k) {
            const double *G, *h;
            con_ptrs(cc, inst, k, &G, &h);
            const double *z = cc->side == ORC_STATE ? w->X + (size_t)k * pb->n : w->U + (size_t)k * pb->m;
            con_eval(cc, G, h, z, w->cv);
            double *l = lam + w->off[c] + (k - cc->k0) * cc->p;
            if (cc->sense == ORC_EQ) {
                for (int r = 0; r < cc->p; ++r)
                    l[r] = fmin(fmax(l[r] + mu * w->cv[r], -o->dual_max), o->dual_max);
            } else if (cc->sense == ORC_INEQ) {
                for (int r = 0; r < cc->p; ++r) l[r] = fmin(fmax(l[r] + mu * w->cv[r], 0.0), o->dual_max);
            } else {
                double lb[ORC_MAX_W + 1];
                for (int r = 0; r < cc->p; ++r) lb[r] = l[r] - mu * w->cv[r];
                orc_soc_project(cc->p, lb, l);
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
        for (int it = 0; it < o->iterations_inner; ++it) {
            double dV[2];
            if (backward_pass(pb, o, w, inst, lam, &rho, &drho, dV)) { r->status = ORC_NOT_PD; break; }
            J = forward_pass(pb, o, w, inst, lam, dV, J_prev, &rho, &drho, &r->trials);
            if (J > o->max_cost_value || !(J == J)) { r->status = ORC_MAXIMUM_COST; break; }
            memcpy(w->X, w->Xb, sizeof(double) * (size_t)N * n);
            memcpy(w->U, w->Ub, sizeof(double) * (size_t)(N - 1) * m);
            double dJ = fabs(J - J_prev);
            J_prev = J;
            double grad = 0.0; /* gradient_todorov! */
            for (int k = 0; k < N - 1; ++k) {
                double mx = 0.0;
                for (int i = 0; i < m; ++i)
                    mx = fmax(mx, fabs(w->dv[(size_t)k * m + i]) / (fabs(w->U[(size_t)k * m + i]) + 1.0));
                grad += mx;
            }
            grad /= (double)(N - 1);
            r->iters++;
            dJ_zero = (dJ == 0.0) ? dJ_zero + 1 : 0;
#ifdef ORC_TRACE
            fprintf(stderr, "inst %d outer %d it %d J %.17g dJ %.3e grad %.3e rho %.3e dV %.3e %.3e trials %d\n", inst,
                    outer, r->iters, J, dJ, grad, rho, dV[0], dV[1], r->trials);
#endif
            int small = o->dj_zero_converges ? (dJ >= 0.0 && dJ < ctol) : (dJ > 0.0 && dJ < ctol);
            if (small && grad < gtol) { r->status = ORC_SOLVE_SUCCEEDED; break; }
            if (r->iters >= o->iterations) { r->status = ORC_MAX_ITERATIONS; break; }
            if (dJ_zero > o->dJ_counter_limit) { r->status = ORC_NO_PROGRESS; break; }
        }
        if (r->status > ORC_SOLVE_SUCCEEDED) break;
        /* ---- AL outer loop bookkeeping */
        r->cmax = max_violation(pb, o, w, inst, w->X, w->U);
#ifdef ORC_TRACE
        fprintf(stderr, "inst %d outer %d cmax %.3e mu %.3e status %d\n", inst, outer, r->cmax, w->mu[0], r->status);
#endif
        r->pen_max = 0.0;
        for (int c = 0; c < pb->ncon; ++c) r->pen_max = fmax(r->pen_max, w->mu[c]);
        if (r->cmax < o->constraint_tolerance) break;
        if (o->kickout_max_penalty && r->pen_max >= o->penalty_max) break;
        dual_update(pb, o, w, inst, lam);
        for (int c = 0; c < pb->ncon; ++c) w->mu[c] = fmin(w->mu[c] * o->penalty_scaling, o->penalty_max);
        if (outer == o->iterations_outer) r->status = ORC_MAX_ITERATIONS_OUTER;
    }
    if (r->status <= ORC_SOLVE_SUCCEEDED)
        r->status = (r->cmax < o->constraint_tolerance) ? ORC_SOLVE_SUCCEEDED : ORC_UNSOLVED;
    r->J = J;
}

typedef struct {
    const orc_problem_t *pb;
    const orc_opts_t *o;
    int i1;
    atomic_int *next;
    double *X, *U, *lam;
    int *iters, *iters_outer, *status, *ls_trials;
    double *cost, *cost_al, *cmax, *pen_max;
} job_t;

/* Worker: pulls instance ids from a shared counter (dynamic schedule, chunk 1). */
static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    const orc_problem_t *pb = j->pb;
    int n = pb->n, m = pb->m, N = pb->N;
    ws_t *w = ws_new(pb);
    for (;;) {
        int i = atomic_fetch_add(j->next, 1);
        if (i >= j->i1) break;
        res_t r;
        double *l = j->lam + (size_t)i * w->P;
        memcpy(w->U, j->U + (size_t)i * (N - 1) * m, sizeof(double) * (size_t)(N - 1) * m);
        solve_instance(pb, j->o, w, i, l, &r);
        memcpy(j->X + (size_t)i * N * n, w->X, sizeof(double) * (size_t)N * n);
        memcpy(j->U + (size_t)i * (N - 1) * m, w->U, sizeof(double) * (size_t)(N - 1) * m);
        if (j->iters) j->iters[i] = r.iters;
        if (j->iters_outer) j->iters_outer[i] = r.outer;
        if (j->status) j->status[i] = r.status;
        if (j->ls_trials) j->ls_trials[i] = r.trials;
        if (j->cost) j->cost[i] = objective_cost(pb, i, w->X, w->U);
        if (j->cost_al) j->cost_al[i] = r.J;
        if (j->cmax) j->cmax[i] = max_violation(pb, j->o, w, i, w->X, w->U);
        if (j->pen_max) j->pen_max[i] = r.pen_max;
    }
    ws_free(w);
    return NULL;
}

int orc_solve_batch(const orc_problem_t *pb, const orc_opts_t *o, int i0, int i1, int nthreads, double *X,
                    double *U, double *lam, int *iters, int *iters_outer, int *status, int *ls_trials,
                    double *cost, double *cost_al, double *cmax, double *pen_max)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    for (int c = 0; c < pb->ncon; ++c)
        if (pb->con[c].w > ORC_MAX_W || pb->con[c].p > ORC_MAX_W) return -1;
    atomic_int next;
    atomic_init(&next, i0);
    job_t j = {pb, o, i1, &next, X, U, lam, iters, iters_outer, status, ls_trials, cost, cost_al, cmax, pen_max};
    pthread_t th[256];
    for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, worker, &j);
    worker(&j);
    for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
    return 0;
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
        const double *x = X + (size_t)i * pb->N * pb->n, *u = U + (size_t)i * (pb->N - 1) * pb->m;
        if (cost) cost[i] = objective_cost(pb, i, x, u);
        if (cmax) cmax[i] = max_violation(pb, o, w, i, x, u);
    }
    ws_free(w);
}
