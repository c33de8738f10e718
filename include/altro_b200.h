/*
 * altro_b200.h -- C ABI of the B200-native batched ALTRO (augmented-Lagrangian iLQR) solver.
 *
 * The reference (RoboticExplorationLab/altro-mpc-icra2021) has no FFI on this path: its boundary
 * is the Julia API of Altro.jl / TrajectoryOptimization.jl / RobotDynamics.jl as used by the
 * benchmark scripts.  Every entry point below names the reference call it replaces (paths relative
 * to /root/reference/benchmarks).  A Julia shim (julia/AltroB200.jl) binds these with `ccall` and
 * re-exports the reference's names; altro_mpc_icra2021_b200/solver.py binds them with ctypes.
 *
 * Conventions
 *   - one handle = one batch of B structurally identical problems on one GPU and one stream;
 *     handles are independent and may be driven from different host threads; one handle is not
 *     re-entrant.
 *   - every pointer argument is a HOST pointer unless the name ends in _dev; buffers are caller
 *     owned; layout is instance-major, row-major, FP64, 0-based knots k = 0..N-1 (controls 0..N-2).
 *   - every function returns 0 on success or a negative altro_status_t error; altro_last_error()
 *     gives the message.  There is no CPU fallback: a missing GPU is an error.
 *   - setters enqueue an async H2D copy on the handle's stream (true DMA when the buffer was pinned
 *     with altro_host_register); altro_solve enqueues the solve; getters synchronise the stream.
 */
#ifndef ALTRO_B200_H
#define ALTRO_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct altro_handle_s *altro_handle_t;

enum { ALTRO_OK = 0, ALTRO_ERR_INVALID = -1, ALTRO_ERR_CUDA = -2, ALTRO_ERR_UNSUPPORTED = -3, ALTRO_ERR_STATE = -4 };

/* TrajectoryOptimization cone senses: Equality / Inequality / SecondOrderCone (scalar last). */
enum { ALTRO_EQUALITY = 0, ALTRO_INEQUALITY = 1, ALTRO_SECOND_ORDER_CONE = 2 };
enum { ALTRO_STATE = 0, ALTRO_CONTROL = 1 };

/* Altro.TerminationStatus (random_linear_problem.jl:166, simple_rocket.jl:144,181). */
enum {
    ALTRO_UNSOLVED = 0, ALTRO_SOLVE_SUCCEEDED = 1, ALTRO_MAX_ITERATIONS = 2, ALTRO_MAX_ITERATIONS_OUTER = 3,
    ALTRO_MAXIMUM_COST = 4, ALTRO_STATE_LIMIT = 5, ALTRO_CONTROL_LIMIT = 6, ALTRO_NO_PROGRESS = 7,
    ALTRO_COST_INCREASE = 8, ALTRO_NOT_PD = 9
};

/* Altro.SolverOptions, field for field (run_random_linear.jl:41-49, run_simple_rocket.jl:121-129,
 * ALTROParams.jl:86-95, grasp_benchmark.jl:26-34, flexible_sat_mpc.jl:250-257). */
typedef struct altro_opts_t {
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
    int dj_zero_converges; /* 1: 0<=dJ<tol converges, 0: 0<dJ<tol */
    int soc_hess_exact;    /* 1: exact projection Hessian, 0: Gauss-Newton */
    int soc_viol_proj;     /* 1: ||c-Pi(c)||_inf, 0: max(0,||v||-t) */
    int first_step_unconditional; /* 1: the first forward pass of every iLQR solve compares against J_prev = +inf
                                     (full step always taken, never the converged iteration); 0: against the cost of
                                     the initial rollout.  Default 1: reproduces the iteration counts and the
                                     one-rollout-per-iteration timing saved in grasp_benchmark_data.jld2 */
} altro_opts_t;

/* SolverOptions() defaults. */
int altro_default_options(altro_opts_t *opts);

/* ALTROSolver(prob, opts) (random_linear_problem.jl:87, simple_rocket.jl:128, ALTROParams.jl:96):
 * size(prob) = (n,m,N), batch B, dt = stage-cost scaling (prob.Z[1].dt). */
int altro_create(altro_handle_t *h, int device, int n, int m, int N, int batch, double dt);
int altro_destroy(altro_handle_t h);
const char *altro_last_error(altro_handle_t h); /* h may be NULL: last error of a failed create */

/* Use a caller-provided cudaStream_t (e.g. torch's current stream) instead of the handle's own. */
int altro_set_stream(altro_handle_t h, void *cuda_stream);

/* SolverOptions(...) / set_options!(solver; ...) (flexible_sat_mpc.jl:163,250-257, grasp_mpc.jl:36). */
int altro_set_options(altro_handle_t h, const altro_opts_t *opts);

/* RD.LinearModel(A,B[,d]) and its in-place update opt.model.A[i] = ... (altro_solver.jl:35-37):
 * x+ = A x + B u + d.  A[inst?][knot?][n][n], B[..][n][m], d[..][n] (d may be NULL = 0). */
int altro_set_dynamics(altro_handle_t h, int per_knot, int per_instance, const double *A, const double *B,
                       const double *d);

/* Gait-scheduled LTV models (quadruped, altro_solver.jl:5-42 + gait.jl:1-9): instead of re-uploading A_k,B_k,d_k at
 * every control tick, upload per instance the `nslots` distinct models A[B][nslots][n][n].. (one per gait phase) and
 * the schedule sched[B][sched_len] of slot indices along absolute time: knot k of the MPC step that follows s
 * transitions uses slot sched[inst][s + k].  altro_mpc_transition / altro_mpc_run advance s. */
int altro_set_dynamics_slots(altro_handle_t h, int nslots, const double *A, const double *B, const double *d,
                             const int *sched, int sched_len);

/* LQRObjective / TrackingObjective diagonal weights (mpc.jl:26-29, ALTROParams.jl:46-47,81). */
int altro_set_cost_diag(altro_handle_t h, const double *Q, const double *R, const double *Qf);

/* TO.update_trajectory!(obj, Z_track, k) (random_linear_problem.jl:133, simple_rocket.jl:75):
 * Xref[B][N][n], Uref[B][N-1][m]; the cost is centred on it (q = -Q xref, r = -R uref). */
int altro_set_reference(altro_handle_t h, const double *Xref, const double *Uref);

/* TO.add_constraint!(cons, con, inds) (random_linear_problem.jl:24, rocket_landing_problem.jl:96-167,
 * ALTROParams.jl:65-78, grasp_problem.jl:35-67) for one affine conic block
 *   c(z) = G z[inds] + h,  z = x_k or u_k,  knots [k0,k1),  p rows, w indices.
 * G[inst?][knot?][p][w], h[..][p].  Must precede the first solve.  Returns the block id. */
int altro_add_constraint(altro_handle_t h, int sense, int side, int k0, int k1, int p, int w, const int *inds,
                         int per_knot, int per_instance, const double *G, const double *hvec, int *con_id);

/* Time-varying constraint data along a shared timeline (the grasp benchmark's torque-balance / grasp-force /
 * friction-cone data, rewritten every MPC step by grasp_mpc_helpers.jl:26-55): G[Nt][p][w], h[Nt][p]; knot k of an
 * instance whose timeline position is kidx reads row min(kidx + k, Nt - 1).  kidx is the per-instance index that
 * altro_set_track / altro_set_track_index set and altro_mpc_transition / altro_mpc_run advance. */
int altro_add_track_constraint(altro_handle_t h, int sense, int side, int k0, int k1, int p, int w, const int *inds,
                               const double *G, const double *hvec, int Nt, int *con_id);
int altro_set_track_index(altro_handle_t h, const int *kidx);

/* In-place constraint data update cons[1].A[i] = ... (grasp_mpc_helpers.jl:46-55). */
int altro_update_constraint_data(altro_handle_t h, int con_id, const double *G, const double *hvec);

/* TO.set_initial_state!(prob, x0) / problem.x0 .= x0 (random_linear_problem.jl:130, flexible_sat_mpc.jl:271). */
int altro_set_x0(altro_handle_t h, const double *x0);

/* initial_states!/initial_controls! and states()/controls() (altro_solver.jl:70-71,78-79).
 * X[B][N][n], U[B][N-1][m]; either may be NULL. */
int altro_set_trajectory(altro_handle_t h, const double *X, const double *U);
int altro_get_trajectory(altro_handle_t h, double *X, double *U);

/* Altro.get_duals (run_random_linear.jl:88): lam[B][P], P = sum over blocks of (k1-k0)*p. */
int altro_dual_len(altro_handle_t h, int *P);
int altro_set_duals(altro_handle_t h, const double *lam);
int altro_get_duals(altro_handle_t h, double *lam);

/* RD.shift_fill!(Z) and Altro.shift_fill!(conSet) (random_linear_problem.jl:136,139,
 * simple_rocket.jl:78,81, altro_solver.jl:65,68): z_k <- z_{k+1}, last knot kept. On device. */
int altro_shift_fill(altro_handle_t h, int primal, int dual);

/* solve!(solver) (random_linear_problem.jl:113,161 ... ): enqueues the batched AL-iLQR solve. */
int altro_solve(altro_handle_t h);
int altro_sync(altro_handle_t h);

/* iterations(s), status(s), s.stats, cost(s), max_violation(s) (random_linear_problem.jl:166-171,
 * altro_solver.jl:75-76, simple_rocket.jl:178-198).  Arrays of length B; any may be NULL. */
int altro_get_stats(altro_handle_t h, int *iterations, int *iterations_outer, int *status, int *ls_trials,
                    double *cost, double *cost_al, double *c_max, double *penalty_max);

/* s.stats.tsolve: device time of the last altro_solve in ms (CUDA events on the solve stream);
 * per_instance_ns[B] (optional) = completion time of each instance since kernel start. */
int altro_get_timing(altro_handle_t h, double *device_ms, long long *per_instance_ns);

/* SolverOptions(verbose=...) (run_simple_rocket.jl:66,92): per-iteration log of every instance,
 * max_rows iLQR iterations x 10 columns {outer, iteration, J, dJ, gradient, rho, dV1, dV2,
 * line-search trials so far, c_max (NaN unless the iteration closed an outer loop)}; 0 disables. */
int altro_set_trace(altro_handle_t h, int max_rows);
int altro_get_trace(altro_handle_t h, double *out /* [B][max_rows][10] */);

/* Profiling aid: SM-clock cycles per instance accumulated since enabling, out[B][8] = {initial rollout + cost,
 * backward pass incl. expansion, forward pass, whole solve, expansion, line-search rollouts, line-search costs, 0}.
 * enable != 0 (re)starts the counters after the optional read; enable == 0 stops them. */
int altro_get_phase_cycles(altro_handle_t h, int enable, long long *out);

/* benchmark_solve!(solver) restore-and-resolve semantics (random_linear_problem.jl:161):
 * snapshot / restore, on the device, of everything a solve or a closed-loop run mutates: X, U, duals, x0, the
 * reference window, the per-instance track index and the noise-bank position. */
int altro_snapshot(altro_handle_t h);
int altro_restore(altro_handle_t h);

/* Device-side MPC transition between two solves (random_linear_problem.jl:121-139,
 * simple_rocket.jl:59-82): x0 <- A_0 x_0 + B_0 u_0 + d_0 + noise_dev[inst][n] (noise may be NULL),
 * then reference window advanced by one knot along a device-resident track
 * (track_X_dev[Nt][n], track_U_dev[Nt-1][m], per-instance start index kept on the device),
 * then primal + dual shift_fill.  Registers the track with altro_set_track. */
int altro_set_track(altro_handle_t h, const double *track_X, const double *track_U, int Nt, const int *k_start);
int altro_mpc_transition(altro_handle_t h, const double *noise /* host [B][n] or NULL */, int shift);
/* How the noise sample z is applied to the propagated state x (the reference's per-benchmark formulas):
 * 0: x += w1*z;  1: x += z*|x|_inf*w1 (random_linear_problem.jl:129);
 * 2: positions x[0:n/2] += z*|x[0:n/2]|_2*w1, velocities x[n/2:n] += z*|x[n/2:n]|_2*w2 (simple_rocket.jl:63-70). */
int altro_set_noise_model(altro_handle_t h, int mode, double w1, double w2);
int altro_get_x0(altro_handle_t h, double *x0);
/* Optional device-resident noise for `steps` transitions, noise[steps][B][n]: used in turn (cyclically)
 * whenever altro_mpc_transition is called with noise == NULL, so a closed-loop run has no host traffic. */
int altro_set_noise_bank(altro_handle_t h, const double *noise, int steps);

/* Closed-loop MPC run, entirely on the device: `steps` x {altro_mpc_transition; altro_solve} per instance inside
 * ONE launch -- the body of run_MPC (random_linear_problem.jl:121-187) / run_Rocket_MPC (simple_rocket.jl:159-203)
 * for every instance of the batch.  Instances advance independently (no lock-step between steps), noise comes
 * from the noise bank, references from the track.  Starts from the current solution (an altro_solve must have
 * run).  Per-step results: statistics [steps][B], closed-loop states x0_log[steps][B][n] (X_traj of the
 * reference) and applied controls u0_log[steps][B][m], per-step device time per instance t_ns[steps][B]. */
int altro_mpc_run(altro_handle_t h, int steps, int shift);
/* Sizes the per-step statistics and closed-loop log buffers for runs of up to `steps` steps ahead of time, so that
 * altro_mpc_run does not have to synchronise and reallocate when a longer run follows a shorter one. */
int altro_reserve_steps(altro_handle_t h, int steps);
int altro_get_run_results(altro_handle_t h, int steps, int *iterations, int *iterations_outer, int *status,
                          int *ls_trials, double *cost, double *c_max, double *x0_log, double *u0_log,
                          long long *t_ns);

/* Independent convex cross-check on the device: the reference validates every ALTRO solve against
 * OSQP / ECOS / COSMO / Mosek (random_linear_problem.jl:37-77,141-186, simple_rocket.jl:184-192, grasp_mpc.jl:75-80).
 * altro_admm_solve solves the handle's CURRENT problems (x0, reference, dynamics, constraint blocks as the next
 * altro_solve would see them) with a batched operator-splitting solver that shares no code path with the AL-iLQR
 * kernels (csrc/admm.cu): fixed penalty rho, stops when the primal and dual residuals are below eps or after
 * max_iter iterations.  The handle's own trajectories and duals are not touched.  X[B][N][n], U[B][N-1][m],
 * per-instance iterations and final residuals; any output may be NULL. */
int altro_admm_solve(altro_handle_t h, double rho, double eps, int max_iter, double *X, double *U, int *iterations,
                     double *r_prim, double *r_dual);

/* Quadruped, the step before the solve path, on the device.
 * altro_quadruped_linearize replaces update_dynamics_matrices! (altro_solver.jl:5-42): A_k = I + A_c dt, B_k = B_c dt,
 * d_k = (f - A_c x_ref - B_c u_ref) dt with A_c, B_c the Jacobians of NonLinearContinuousDynamics
 * (linearized_dynamics.jl:1-66, forward-mode AD like the reference's ForwardDiff) at x_ref[B][(N-1)?][12],
 * u_ref[B][(N-1)?][12] (NULL = 0), world foot positions foot[B][N-1][4][3], contact flags contacts[B][N-1][4],
 * body inertia J[3][3] and sprung mass.  The result is written into the handle's per-instance, per-knot model
 * (opt.model.A[i] = ..., altro_solver.jl:35-37) without passing through the host.
 * altro_quadruped_tick additionally builds the contact pattern and the foot positions of the horizon on the device
 * from the tick time t[B] and the current body-frame foot positions cur_foot[B][4][3]: foot_history! (footsteps.jl:29-84)
 * with get_phase (gait.jl:1-9) and footstep_location (footsteps.jl:1-27); contact_phases[num_phases][4],
 * phase_times[num_phases], alpha / foot_radius / nom_foot[4][3] as in GaitParams.jl:38-49, Woofer.yaml.  The footstep
 * planner's state (planner_foot_loc) lives in the handle.  altro_quadruped_get_schedule / altro_get_dynamics read the
 * results back (tests, logging). */
int altro_quadruped_linearize(altro_handle_t h, const double *x_ref, int x_ref_per_knot, const double *u_ref,
                              int u_ref_per_knot, const double *foot, const double *contacts, const double *J,
                              double mass);
int altro_quadruped_tick(altro_handle_t h, const double *t, const double *x_ref, int x_ref_per_knot,
                         const double *cur_foot, int num_phases, const double *contact_phases,
                         const double *phase_times, double alpha, double foot_radius, const double *nom_foot,
                         const double *J, double mass);
int altro_quadruped_get_schedule(altro_handle_t h, double *contacts, double *foot);
int altro_get_dynamics(altro_handle_t h, double *A, double *B, double *d);

/* Pin / unpin a caller buffer so setters and getters DMA directly (cudaHostRegister). */
int altro_host_register(void *ptr, size_t bytes);
int altro_host_unregister(void *ptr);

/* Launch geometry of the solve kernel: threads per instance, dynamic shared memory bytes per
 * instance (CTA), registers per thread, resident CTAs per SM. threads=0 in the setter = automatic. */
int altro_set_launch_config(altro_handle_t h, int threads_per_instance);
/* Which kernel runs solve!: 1 = one CTA per instance (any dimensions; the default), 2 = one thread per instance, a
 * warp advancing 8 instances (small dimensions with a shared LTI model: rocket 6/3, grasp 6/6; an experiment that
 * measured slower, see csrc/altro_lane.cuh), 0 = automatic (= 1).  Results are bit-identical either way.  The getter
 * reports the kernel in use. */
int altro_set_kernel_mode(altro_handle_t h, int mode);
/* Scheduling of altro_mpc_run: steps_per_item >= 1 (default 1) runs it on a persistent grid whose CTAs pull
 * (instance, steps_per_item consecutive steps) work items from an atomic queue, so that no slot idles while another
 * still owns long chains; 0 = one CTA per instance for the whole run.  Same results either way. */
int altro_set_run_queue(altro_handle_t h, int steps_per_item);
int altro_get_kernel_mode(altro_handle_t h, int *mode, int *lane_regs_per_thread, int *lane_smem_bytes,
                          int *instances_per_warp);
int altro_get_launch_info(altro_handle_t h, int *threads_per_instance, int *smem_bytes, int *regs_per_thread,
                          int *ctas_per_sm, int *num_sms);

/* Line search of the forward pass (Altro ilqr/forwardpass.jl, SURVEY.md A.8: alpha = 1, 1/2, 1/4, ... until the
 * first step passes the expected-decrease test). speculative = 1: each warp of the instance rolls out and costs
 * one of the next threads/32 trial steps at the same time and the acceptance test is replayed over them in
 * order -- same accepted step, trajectory and trial count, lower latency; 0: one trial after the other;
 * -1 (default): speculative when the extra candidate buffers do not cost resident CTAs. Set before the first
 * solve. The getter reports what the finalized handle uses. */
int altro_set_line_search_mode(altro_handle_t h, int speculative);
int altro_get_line_search_mode(altro_handle_t h, int *speculative);

/* FP64 roofline denominators measured on `device`: dependent-free DFMA stream and
 * mma.sync m8n8k4 f64 stream (TFLOP/s), and a device copy (GB/s). Any may be NULL. */
int altro_measure_peaks(int device, double *dfma_tflops, double *dmma_tflops, double *copy_gbs);

#ifdef __cplusplus
}
#endif
#endif /* ALTRO_B200_H */
