/*
 * altro_oracle.h -- CPU restatement of the ALTRO augmented-Lagrangian iLQR solve path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker or the timed CPU baseline -- never as the shipped compute path.
 *
 * PARITY UNPINNED.  The algorithm on this path lives in un-vendored Julia packages that are
 * absent from /root/reference and cannot be run here (no julia, no network):
 *   Altro.jl 0.2.0 @socp  tree 6b0c4be0713aec7af3e4ecf66dec417f69eec6dd
 *       (benchmarks/Manifest.toml:26-32; quadruped env pins tree 9cc5156b..., quadruped/Manifest.toml:15-21)
 *   TrajectoryOptimization.jl 0.3.2 @socp tree d8e9804f... (benchmarks/Manifest.toml:1030-1036)
 *   RobotDynamics.jl 0.2.2 (benchmarks/Manifest.toml:867-871)
 * The reference repo holds no tests and no golden trajectories for the path.
 * This file restates the published AL-iLQR algorithm of those packages
 * and anchors it on the reference's own call sites:
 *   solve!            random_linear_problem.jl:113,161  simple_rocket.jl:129,174  grasp_mpc.jl:55
 *                     altro_solver.jl:72  flexible_sat_mpc.jl:166,272
 *   SolverOptions     run_random_linear.jl:41-49  run_simple_rocket.jl:121-129  ALTROParams.jl:86-95
 *                     grasp_benchmark.jl:26-34  flexible_sat_mpc.jl:250-257
 *   shift_fill!       random_linear_problem.jl:136,139  simple_rocket.jl:78,81  altro_solver.jl:65,68
 * Its own correctness is established by closed-form LQR cases, KKT residuals, independent
 * convex solves (scipy) and the iteration-count statistics recovered from the reference's
 * saved results (tests/test_oracle_*.py).
 *
 * Conventions: all indices 0-based, row-major, FP64.  Knot k = 0..N-1; controls at k = 0..N-2.
 */
#ifndef ALTRO_ORACLE_H
#define ALTRO_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_W 64

enum { ORC_EQ = 0, ORC_INEQ = 1, ORC_SOC = 2 };
enum { ORC_STATE = 0, ORC_CONTROL = 1 };

/* Termination status, same order as Altro.TerminationStatus (simple_rocket.jl:144,181). */
enum {
    ORC_UNSOLVED = 0,
    ORC_SOLVE_SUCCEEDED = 1,
    ORC_MAX_ITERATIONS = 2,
    ORC_MAX_ITERATIONS_OUTER = 3,
    ORC_MAXIMUM_COST = 4,
    ORC_STATE_LIMIT = 5,
    ORC_CONTROL_LIMIT = 6,
    ORC_NO_PROGRESS = 7,
    ORC_COST_INCREASE = 8,
    ORC_NOT_PD = 9
};

/* One affine conic constraint block  c(z) = G z[inds] + h,  z = x_k (state side) or u_k
 * (control side), required to satisfy  c = 0 | c <= 0 | ||c[0:p-1]|| <= c[p-1]  on knots
 * [k0,k1).  Covers BoundConstraint / NormConstraint (rocket_landing_problem.jl:123),
 * NormConstraint2 (new_constraints.jl:72-116), LinearizedFrictionConstraint
 * (LinearizedFrictionConstraint.jl:14-25), the grasp torque/force rows (grasp_problem.jl:35-67). */
typedef struct {
    int sense, side;
    int k0, k1;
    int p, w;
    int inds[ORC_MAX_W];
    int per_knot, per_instance; /* data layout: G[inst?][knot-k0?][p][w], h[inst?][knot-k0?][p] */
    int track;                  /* > 0: G, h are a shared timeline [track][p][w]; knot k reads row min(kidx+k, track-1) */
    const double *G;
    const double *h;
} orc_con_t;

typedef struct {
    int n, m, N, B;
    double dt;
    int dyn_per_knot, dyn_per_instance; /* A[inst?][knot?][n][n], Bm[..][n][m], d[..][n] */
    const double *A, *Bm, *d;
    /* optional gait schedule (altro_set_dynamics_slots): A[B][dyn_slots][n][n].., knot k of the solve that follows
     * `step0` transitions uses slot sched[inst][min(step0, sched_len - N) + k]; sched == NULL disables */
    int dyn_slots, sched_len, step0;
    const int *sched;
    const int *kidx;            /* [B] current position of every instance on the shared timelines, or NULL (= 0) */
    const double *Q, *R, *Qf;   /* diagonal weights, shared: [n], [m], [n] */
    const double *xref, *uref;  /* [B][N][n], [B][N-1][m] tracking reference */
    const double *x0;           /* [B][n] */
    int ncon;
    const orc_con_t *con;
} orc_problem_t;

/* Mirrors Altro.SolverOptions. */
typedef struct {
    double constraint_tolerance;
    double cost_tolerance, cost_tolerance_intermediate;
    double gradient_tolerance, gradient_tolerance_intermediate;
    double penalty_initial, penalty_scaling, penalty_max, dual_max;
    double line_search_lower_bound, line_search_upper_bound;
    double max_cost_value, max_state_value;
    double bp_reg_initial, bp_reg_increase_factor, bp_reg_max, bp_reg_min, bp_reg_fp;
    int iterations, iterations_inner, iterations_outer, iterations_linesearch;
    int dJ_counter_limit;
    int reset_duals, reset_penalties, kickout_max_penalty;
    /* named switches for the recollection-uncertain details */
    int dj_zero_converges;  /* 1: 0<=dJ<tol converges (default), 0: 0<dJ<tol */
    int soc_hess_exact;     /* 1: mu*G'*dPi*G (= incl. second-order projection term), 0: Gauss-Newton dPi'dPi */
    int soc_viol_proj;      /* 1: ||c-Pi(c)||_inf, 0: max(0,||v||-t) */
    int first_step_unconditional; /* 1: J_prev = +inf at the first iteration of every iLQR solve (default), 0: the
                                     initial rollout's cost (Appendix A.6 as recollected) */
} orc_opts_t;

void orc_default_opts(orc_opts_t *o);

/* Number of dual variables per instance: sum_c (k1-k0)*p. Layout [con][knot-k0][p]. */
int orc_dual_len(const orc_problem_t *pb);

/* Solve instances [i0,i1) with `nthreads` host threads (dynamic schedule over instances).
 * In/out: U[B][N-1][m] warm start -> solution; lam[B][P] duals (reset if reset_duals).
 * Out: X[B][N][n]; per-instance stats arrays of length B (may be NULL). */
int orc_solve_batch(const orc_problem_t *pb, const orc_opts_t *o, int i0, int i1, int nthreads,
                    double *X, double *U, double *lam,
                    int *iters, int *iters_outer, int *status, int *ls_trials,
                    double *cost, double *cost_al, double *cmax, double *pen_max);

/* Closed-loop MPC run (the product's altro_mpc_run): per instance, `steps` x {transition; solve}, where the
 * transition is x0 <- x_1 of the last solution + noise (modes as altro_set_noise_model), the reference window
 * advanced along the track and the primal + dual shift_fill.  Instances run independently over the threads.
 * Writes pb's x0 / xref / uref rows in place (the arrays must be writable).  Per-step outputs [steps][B]. */
typedef struct {
    int steps, shift, noise_mode, Nt;
    double w1, w2;
    const double *noise;           /* [steps][B][n] or NULL */
    const double *trackX, *trackU; /* [Nt][n], [Nt-1][m] or NULL */
    const int *kidx;               /* [B] window start before the run */
} orc_run_t;

int orc_mpc_run(const orc_problem_t *pb, const orc_opts_t *o, const orc_run_t *run, int nthreads, double *X,
                double *U, double *lam, int *iters, int *iters_outer, int *status, int *ls_trials, double *cost,
                double *cmax, double *x0_log, double *u0_log);

/* Warm-start shifts (RD.shift_fill!, Altro.shift_fill!): z_k <- z_{k+1}, last kept. */
void orc_shift_fill(const orc_problem_t *pb, int primal, int dual, double *X, double *U, double *lam);

/* Objective value (no penalty terms) and max constraint violation of given trajectories. */
void orc_evaluate(const orc_problem_t *pb, const orc_opts_t *o, const double *X, const double *U,
                  double *cost, double *cmax);

/* Per-iteration log [B][rows][10] (outer, iter, J, dJ, grad, rho, dV1, dV2, trials, c_max); NULL disables. */
void orc_set_trace(double *buf, int rows);

/* Unit pieces exposed for tests. */
void orc_soc_project(int p, const double *v, double *out);
void orc_soc_project_jac(int p, const double *v, double *J /* p*p */);

#ifdef __cplusplus
}
#endif
#endif
