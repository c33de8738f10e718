"""Parity tests proper: the CUDA path, called through the C ABI (ctypes -> libaltro_b200.so), against the CPU
oracle on the same seeded inputs and against the committed golden vectors.

Bar: BIT-IDENTICAL FP64 trajectories, duals, costs, and identical iteration / line-search / status counts.
(north_star asks for 1e-6 relative and equal iteration counts; AL-iLQR on conic problems is a semismooth Newton
method whose Hessian branch at a cone boundary is decided by one ulp, so equal iteration counts are only
guaranteed by identical arithmetic -- see DESIGN.md 'Parity contract'.)"""
import copy
import ctypes
import os

import numpy as np
import pytest

from altro_mpc_icra2021_b200.problem import (ConstraintList, Equality, GoalConstraint, Inequality, LinearConstraint,
                                             LinearModel, LQRObjective, Problem, SecondOrderCone, SolverOptions, CONTROL)
from altro_mpc_icra2021_b200.problems import mpc, quadruped, random_linear, rocket
from tests.golden import cases
from tests.helpers import OracleSolver, assert_bit_identical, lqr_problem

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gpu_solver(prob, opts, **kw):
    from altro_mpc_icra2021_b200.solver import ALTROSolver

    return ALTROSolver(prob, opts, **kw)


def solve_both(prob, opts, **kw):
    pg = copy.deepcopy(prob)
    o = OracleSolver(prob, opts).solve()
    g = gpu_solver(pg, opts, **kw).solve()
    assert_bit_identical(pg, g.stats, g.get_duals(), o.stats)
    return pg, g, o


# ------------------------------------------------------------------ golden vectors and MPC loops

def test_rocket_cold_solve_golden():
    gold = np.load(os.path.join(GOLD, "rocket_cold.npz"))
    prob, opts = cases.rocket_track()
    g = gpu_solver(prob, opts).solve()
    assert np.array_equal(prob.X[0], gold["X"]) and np.array_equal(prob.U[0], gold["U"])
    assert np.array_equal(g.stats.iterations, gold["iters"]) and g.stats.status[0] == 1


def test_rocket_mpc_golden():
    cold = np.load(os.path.join(GOLD, "rocket_cold.npz"))
    out = cases.run_case(gpu_solver, *cases.case_rocket_mpc(cold["X"], cold["U"]))
    gold = np.load(os.path.join(GOLD, "rocket_mpc.npz"))
    for k in gold.files:
        assert np.array_equal(out[k], gold[k]), k


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_family_golden(name):
    out = cases.run_case(gpu_solver, *cases.CASES[name]())
    gold = np.load(os.path.join(GOLD, f"{name}.npz"))
    for k in gold.files:
        assert np.array_equal(out[k], gold[k]), k


@pytest.mark.parametrize("family", ["rocket", "random_linear", "quadruped_lin", "quadruped_soc"])
def test_mpc_loop_against_live_oracle(family):
    """256 instances, 4 warm-started MPC steps, GPU and oracle advanced with identical host updates."""
    B = 256
    if family == "rocket":
        cold = np.load(os.path.join(GOLD, "rocket_cold.npz"))
        prob, opts, steps, adv = cases.case_rocket_mpc(cold["X"], cold["U"], batch=B)
        _, _, _, adv2 = cases.case_rocket_mpc(cold["X"], cold["U"], batch=B)
    elif family == "random_linear":
        prob, opts, steps, adv = cases.case_random_linear(batch=B)
        _, _, _, adv2 = cases.case_random_linear(batch=B)
    else:
        prob, opts, steps, adv = cases.case_quadruped(family.endswith("lin"), batch=B)
        _, _, _, adv2 = cases.case_quadruped(family.endswith("lin"), batch=B)
    pg = copy.deepcopy(prob)
    o, g = OracleSolver(prob, opts, nthreads=8), gpu_solver(pg, opts)
    for st in range(4):
        o.solve()
        g.solve()
        assert_bit_identical(pg, g.stats, g.get_duals(), o.stats, f"{family} step {st}")
        assert np.all(g.stats.status == 1)
        adv(prob, o, st)
        adv2(pg, g, st)


# ------------------------------------------------------------------ kernel variants

@pytest.mark.parametrize("threads", [32, 64, 128, 256])
def test_threads_per_instance_do_not_change_bits(threads):
    prob, opts, _, _ = cases.case_quadruped(False, batch=16)
    solve_both(prob, opts, threads_per_instance=threads)


@pytest.mark.parametrize("threads", [64, 128, 256])
@pytest.mark.parametrize("spec", [False, True])
@pytest.mark.parametrize("family", ["rocket", "quadruped_soc", "random_linear"])
def test_speculative_line_search_does_not_change_bits(family, spec, threads):
    """One line-search trial per warp (forward_pass_spec) vs one after the other: same accepted step, same
    trajectories, same trial counts -- against the oracle, which only knows the sequential search."""
    if family == "rocket":
        cold = np.load(os.path.join(GOLD, "rocket_cold.npz"))
        prob, opts, _, _ = cases.case_rocket_mpc(cold["X"], cold["U"], batch=32)
    elif family == "random_linear":
        prob, opts, _, _ = cases.case_random_linear(batch=32)
    else:
        prob, opts, _, _ = cases.case_quadruped(False, batch=32)
    _, g, o = solve_both(prob, opts, threads_per_instance=threads, speculative_line_search=spec)
    assert g.launch_info()["speculative_line_search"] == spec
    assert np.array_equal(g.stats.ls_trials, o.stats.ls_trials)


@pytest.mark.parametrize("batch", [1, 7, 33, 200])
def test_lane_kernel_matches_oracle_and_cta_kernel(batch):
    """One thread per instance (altro_lane.cuh) vs one CTA per instance vs the oracle: same bits, any batch size
    (partial warps included)."""
    cold = np.load(os.path.join(GOLD, "rocket_cold.npz"))
    prob, opts, _, _ = cases.case_rocket_mpc(cold["X"], cold["U"], batch=batch)
    prob.set_initial_state(prob.x0 + 0.05 * mpc.rng_for(2, batch).standard_normal(prob.x0.shape))
    pc = copy.deepcopy(prob)
    pl, gl, o = solve_both(prob, opts, kernel="lane")
    gc = gpu_solver(pc, opts, kernel="cta").solve()
    assert gl.launch_info()["kernel"] == "lane" and gc.launch_info()["kernel"] == "cta"
    assert_bit_identical(pc, gc.stats, gc.get_duals(), o.stats, "cta")
    assert o.stats.iterations.max() > 2


def test_lane_kernel_is_opt_in_and_limited_to_small_shared_lti_problems():
    cold = np.load(os.path.join(GOLD, "rocket_cold.npz"))
    prob, opts, _, _ = cases.case_rocket_mpc(cold["X"], cold["U"], batch=64)
    p2 = copy.deepcopy(prob)  # (a solver consumes the problem's dirty flags: one problem object per solver)
    assert gpu_solver(prob, opts).launch_info()["kernel"] == "cta"
    assert gpu_solver(p2, opts, kernel="lane").launch_info()["kernel"] == "lane"
    pq, oq, _, _ = cases.case_quadruped(True, batch=4)
    pq2 = copy.deepcopy(pq)
    assert gpu_solver(pq, oq).launch_info()["kernel"] == "cta"
    from altro_mpc_icra2021_b200.solver import AltroError
    with pytest.raises(AltroError, match="lane-per-instance"):
        gpu_solver(pq2, oq, kernel="lane").solve()


def test_speculative_line_search_on_long_horizon_family():
    """Flexible satellite (n = 12, m = 3, N = 40 here): 4 warps per instance, 4 trial steps at a time."""
    prob, opts, _, _ = cases.case_flexsat(batch=6)
    for spec in (False, True):
        _, g, o = solve_both(copy.deepcopy(prob), opts, threads_per_instance=128, speculative_line_search=spec)
        assert np.array_equal(g.stats.ls_trials, o.stats.ls_trials)


def test_repeated_runs_are_identical_and_logs_survive_reallocation():
    """The same closed-loop run from the same state twice, the second one after the log buffers were re-sized:
    identical bits, no stale or zeroed log (guards the null-stream memset race fixed in dalloc)."""
    cold = np.load(os.path.join(GOLD, "rocket_cold.npz"))
    runs = []
    B = 64
    ks = mpc.rng_for(11, 0).integers(0, cold["X"].shape[0] - 21 - 110, size=B)
    noise = mpc.rng_for(3, 3).standard_normal((7, B, 6))
    for reserve in (0, 12, 0):
        prob, opts, _, _ = cases.case_rocket_mpc(cold["X"], cold["U"], batch=B)
        sv = gpu_solver(prob, opts).solve()
        sv.set_track(cold["X"], cold["U"], ks)
        sv.set_noise_model(2, 1e-3, 1e-2)
        sv.set_noise_bank(noise)
        if reserve:
            sv.reserve_steps(reserve)
        sv.mpc_run(2)
        runs.append(sv.mpc_run(5))
        sv.close()
    for k in ("iterations", "ls_trials", "status", "cost", "c_max", "x0", "u0"):
        assert np.array_equal(runs[0][k], runs[1][k]) and np.array_equal(runs[0][k], runs[2][k]), k
    assert np.abs(runs[0]["u0"]).max() > 0


def test_reusable_page_locked_result_buffers_return_the_same_results():
    """run_results(reuse_buffers=True) copies into one page-locked set kept by the solver (what a caller that reads
    results every tick does): same numbers as the fresh-array path, views valid until the next reuse call."""
    cold = np.load(os.path.join(GOLD, "rocket_cold.npz"))
    B = 48
    ks = mpc.rng_for(5, 1).integers(0, cold["X"].shape[0] - 21 - 40, size=B)
    noise = mpc.rng_for(4, 4).standard_normal((6, B, 6))
    prob, opts, _, _ = cases.case_rocket_mpc(cold["X"], cold["U"], batch=B)
    sv = gpu_solver(prob, opts, pin=True).solve()
    sv.set_track(cold["X"], cold["U"], ks)
    sv.set_noise_model(2, 1e-3, 1e-2)
    sv.set_noise_bank(noise)
    sv.snapshot()
    fresh = sv.mpc_run(6)
    sv.restore()
    sv.reserve_host_results(6)
    again = sv.mpc_run(6, reuse_buffers=True)
    for k in ("iterations", "ls_trials", "status", "cost", "c_max", "x0", "u0"):
        assert np.array_equal(fresh[k], again[k]), k
    sv.restore()
    short = sv.mpc_run(2, reuse_buffers=True)  # a shorter run reuses the same buffers
    assert short["x0"].shape[0] == 2 and np.array_equal(short["x0"], fresh["x0"][:2])
    sv.close()


def test_runtime_dimension_kernel_matches_compiled_dimensions(monkeypatch):
    prob, opts, _, _ = cases.case_random_linear(batch=8)
    ref = copy.deepcopy(prob)
    a = gpu_solver(ref, opts).solve()
    monkeypatch.setenv("ALTRO_B200_GENERIC", "1")
    gen = copy.deepcopy(prob)
    b = gpu_solver(gen, opts).solve()
    assert np.array_equal(ref.X, gen.X) and np.array_equal(ref.U, gen.U)
    assert np.array_equal(a.stats.iterations, b.stats.iterations)


@pytest.mark.parametrize("n,m,N", [(2, 2, 21), (5, 2, 9), (15, 2, 21), (30, 6, 21), (3, 1, 2)])
def test_odd_dimensions_use_the_runtime_kernel(n, m, N):
    solve_both(lqr_problem(n=n, m=m, N=N, batch=7, seed=n + m, u_bnd=0.4), SolverOptions(constraint_tolerance=1e-6))


@pytest.mark.parametrize("n,m,N", [(64, 8, 11), (100, 20, 9), (200, 25, 6)])
def test_large_state_dimensions_use_the_global_workspace(n, m, N):
    """n >= 64: S no longer fits in shared memory; the n-sized matrices and the gains live in a per-instance global
    workspace (make_layout_big) and the Riccati products are real FP64 tensor-core GEMM tiles.  Same bits."""
    prob = lqr_problem(n=n, m=m, N=N, batch=3, seed=n + m, u_bnd=0.4)
    _, g, _ = solve_both(prob, SolverOptions(constraint_tolerance=1e-6))
    assert g.launch_info()["smem_bytes"] < 227 * 1024


@pytest.mark.parametrize("n,m,N", [(64, 16, 7), (96, 7, 6), (100, 25, 5), (160, 20, 5), (200, 25, 4)])
def test_staged_panel_gemm_and_direct_tiles_give_the_same_bits(n, m, N, monkeypatch):
    """The large-dimension Riccati products through TMA-staged operand panels (panel_gemm_fn: cp.async.bulk + mbarrier
    ring, producer warp, chained products, fused symmetrisation) and through the direct L2-operand tiles: both equal the
    oracle bit for bit.  Covers m not a multiple of 4 (partial last k-step of the K = m products), odd m (scalar
    prologue / epilogue of the m-wide outputs) and even m (16-byte path), one and several column windows."""
    info = {}
    for tma in ("1", "0"):
        monkeypatch.setenv("ALTRO_B200_TMA", tma)
        prob = lqr_problem(n=n, m=m, N=N, batch=3, seed=n + m, u_bnd=0.4)
        _, g, _ = solve_both(prob, SolverOptions(constraint_tolerance=1e-6))
        info[tma] = g.launch_info()
        g.close()
    assert info["1"]["smem_bytes"] > info["0"]["smem_bytes"]  # the panel stages are part of the staged layout only
    assert info["1"]["ctas_per_sm"] == 1


def test_staged_panel_path_in_a_queued_closed_loop_run(monkeypatch):
    """Closed-loop run of a large-dimension batch through the work queue (workspace bound to the CTA slot, state
    carried through global memory between the items of an instance) against the oracle's run."""
    from oracle.oracle import OracleProblem

    monkeypatch.setenv("ALTRO_B200_TMA", "1")
    n, m, K, B = 96, 12, 3, 5
    prob, Xt, Ut, ks = random_linear.mpc_problem(n, m, 9, batch=B, seed=31)
    opts = random_linear.mpc_options()
    pg = copy.deepcopy(prob)
    sv = gpu_solver(pg, opts)
    sv.set_track(Xt, Ut, ks)
    noise = mpc.rng_for(n, m).standard_normal((K, B, n))
    sv.set_noise_model(1, 0.01, 0.0)
    sv.set_noise_bank(noise)
    sv.solve()
    rg = sv.mpc_run(K)
    op = OracleProblem(prob)
    op.solve(opts, 4)
    ro = op.mpc_run(opts, K, noise, (1, 0.01, 0.0), (Xt, Ut), ks, True, 4)
    for k in ro:
        assert np.array_equal(rg[k], ro[k]), k
    assert np.array_equal(pg.X, prob.X) and np.array_equal(pg.U, prob.U)


# ------------------------------------------------------------------ edge cases

def test_single_instance_and_unconstrained():
    solve_both(lqr_problem(batch=1, seed=3), SolverOptions())


def test_empty_constraint_list_large_batch():
    pg, g, o = solve_both(lqr_problem(n=6, m=3, N=21, batch=300, seed=4), SolverOptions())
    assert np.all(g.stats.status == 1) and np.all(g.stats.c_max == 0.0)


def test_goal_equality_and_state_bounds():
    from altro_mpc_icra2021_b200.problem import BoundConstraint

    prob = lqr_problem(n=4, m=2, N=15, batch=9, seed=8)
    prob.constraints.add_constraint(GoalConstraint(np.zeros(4)), 14)
    prob.constraints.add_constraint(BoundConstraint(4, 2, x_min=-2.0, x_max=2.0, u_min=-1.0, u_max=1.0), (1, 15))
    opts = SolverOptions(constraint_tolerance=1e-6, penalty_initial=10.0)
    pg, g, o = solve_both(prob, opts)
    ok = g.stats.status == 1  # some random instances cannot reach the goal inside the box: they must report failure
    assert ok.sum() >= 4 and np.abs(pg.X[ok, -1]).max() < 1e-5 and np.all(g.stats.c_max[~ok] > 1e-6)


def test_per_knot_per_instance_constraint_data_and_update():
    """Time-varying, per-instance affine rows (the grasp benchmark's torque-balance data) and their in-place
    update between solves (grasp_mpc_helpers.jl:46-55)."""
    n, m, N, B = 4, 3, 10, 6
    prob = lqr_problem(n=n, m=m, N=N, batch=B, seed=21)
    rng = np.random.default_rng(21)
    A = rng.standard_normal((B, N - 1, 2, m))
    b = 0.1 * rng.standard_normal((B, N - 1, 2))
    prob.constraints.add_constraint(LinearConstraint(n, m, A, b, Equality, ":control", per_knot=True, per_instance=True),
                                    (0, N - 1))
    Ac = rng.standard_normal((N - 1, 3, m))
    Ac[:, -1] = np.abs(Ac[:, -1]) + 1.0
    prob.constraints.add_constraint(LinearConstraint(n, m, Ac, np.zeros((N - 1, 3)) - np.array([0, 0, 1.0]),
                                                     SecondOrderCone, ":control", per_knot=True), (0, N - 1))
    opts = SolverOptions(constraint_tolerance=1e-5, penalty_initial=10.0, reset_duals=False)
    pg = copy.deepcopy(prob)
    o, g = OracleSolver(prob, opts).solve(), gpu_solver(pg, opts).solve()
    assert_bit_identical(pg, g.stats, g.get_duals(), o.stats, "first solve")
    for p_ in (prob, pg):
        p_.set_constraint_data(0, h=p_.constraints.flat[0].h * 0.5)
        p_.set_initial_state(p_.x0 * 0.9)
    o.solve()
    g.solve()
    assert_bit_identical(pg, g.stats, g.get_duals(), o.stats, "after data update")


def test_iteration_caps_and_status_codes():
    prob, opts, _, _ = cases.case_quadruped(False, batch=12)
    o2 = opts.copy()
    o2.iterations = 1
    pg, g, o = solve_both(prob, o2)
    assert set(g.stats.status) <= {0, 1, 2}
    prob, opts, _, _ = cases.case_quadruped(False, batch=12)
    o3 = opts.copy()
    o3.iterations_outer, o3.constraint_tolerance = 1, 1e-12
    pg, g, o = solve_both(prob, o3)
    assert np.all(g.stats.iterations_outer == 1)


@pytest.mark.parametrize("switch", ["dj_zero_converges", "soc_hess_exact", "soc_viol_proj", "reset_duals",
                                    "first_step_unconditional"])
def test_option_switches_flip_consistently(switch):
    prob, opts, _, _ = cases.case_quadruped(False, batch=10)
    o2 = opts.copy()
    setattr(o2, switch, not getattr(o2, switch))
    solve_both(prob, o2)


# ------------------------------------------------------------------ device-side helpers of the MPC loop

def test_device_mpc_transition_equals_host_update():
    cold = np.load(os.path.join(GOLD, "rocket_cold.npz"))
    Xt, Ut = cold["X"], cold["U"]
    B = 64
    prob, opts, _, _ = cases.case_rocket_mpc(Xt, Ut, batch=B)
    ks = np.zeros(B, dtype=np.int64)
    ks[:] = mpc.rng_for(11, 0).integers(0, Xt.shape[0] - 21 - 110, size=B)  # same draw as rocket.mpc_problem
    pg = copy.deepcopy(prob)
    o, g = OracleSolver(prob, opts), gpu_solver(pg, opts)
    g.set_track(Xt, Ut, ks)
    o.solve()
    g.solve()
    rng = mpc.rng_for(5, 5)
    for st in range(3):
        nz = rocket.noise(prob.X[:, 1, :], rng)
        ks = ks + 1
        prob.set_initial_state(prob.X[:, 1, :] + nz)
        prob.update_trajectory(*mpc.window_reference(Xt, Ut, ks, prob.N))
        o.shift_fill(True, True)
        g.mpc_transition(nz, shift=True)  # plant step + noise + reference window + shifts, all on the device
        o.solve()
        g.solve()
        assert_bit_identical(pg, g.stats, g.get_duals(), o.stats, f"step {st}")
        assert np.array_equal(pg.X[:, 0], prob.x0)


def test_snapshot_restore_and_benchmark_solve():
    prob, opts, _, _ = cases.case_random_linear(batch=32)
    g = gpu_solver(prob, opts)
    g.solve()
    prob.set_initial_state(prob.X[:, 1, :] * 1.01)
    g.shift_fill(True, True)
    times = g.benchmark_solve(samples=2, evals=2)
    X1, it1 = prob.X.copy(), g.stats.iterations.copy()
    assert times.shape == (4,) and np.all(times > 0)
    g._ck(g.lib.altro_restore(g.h))
    g.solve()
    assert np.array_equal(prob.X, X1) and np.array_equal(g.stats.iterations, it1)  # restore-and-resolve repeats


def test_snapshot_restore_on_a_large_unconstrained_batch():
    """P = 0: the dual buffers hold one element per instance at most; snapshot / restore / benchmark_solve must not
    copy past them (B * max(P, 1) doubles were copied over an 8-byte allocation)."""
    prob = lqr_problem(n=4, m=2, N=15, batch=2048, seed=3)
    pg = copy.deepcopy(prob)
    opts = SolverOptions()
    o = OracleSolver(prob, opts, nthreads=8).solve()
    g = gpu_solver(pg, opts)
    times = g.benchmark_solve(samples=2, evals=1)
    assert times.shape == (2,)
    assert np.array_equal(pg.X, prob.X) and np.array_equal(g.stats.iterations, o.stats.iterations)
    g.snapshot()
    g.solve()
    X1 = pg.X.copy()
    g.restore()
    g.solve()
    assert np.array_equal(pg.X, X1)  # restore-and-resolve repeats


def test_two_handles_same_dimensions_different_horizons_interleaved():
    """The dynamic shared-memory limit is per kernel, process wide: a second handle with a shorter horizon must not
    lower it under the first one."""
    pa, opts, _, _ = cases.case_random_linear(batch=8)
    pb_full, X, U, ks = random_linear.mpc_problem(12, 6, 11, batch=8)
    oa, ob = copy.deepcopy(pa), copy.deepcopy(pb_full)
    ra, rb = OracleSolver(oa, opts).solve(), OracleSolver(ob, opts).solve()
    ga = gpu_solver(pa, opts)
    ga.solve()                      # long horizon first: sets the larger limit
    gb = gpu_solver(pb_full, opts)
    gb.solve()                      # shorter horizon finalised later
    assert_bit_identical(pb_full, gb.stats, gb.get_duals(), rb.stats, "short horizon")
    pa.set_initial_state(pa.x0 * 1.0)
    ga.solve()                      # the first handle launches again with its larger request
    gb.solve()
    ga.solve()
    assert ga.launch_info()["smem_bytes"] > gb.launch_info()["smem_bytes"]
    assert np.all(ga.stats.status == 1)


def test_reset_penalties_false_is_refused_not_ignored():
    from altro_mpc_icra2021_b200.solver import AltroError
    prob, opts, _, _ = cases.case_random_linear(batch=2)
    o2 = opts.copy()
    o2.reset_penalties = False
    with pytest.raises(AltroError):
        gpu_solver(prob, o2).solve()


def test_run_results_are_bounded_by_the_last_run_and_stats_follow_it():
    from altro_mpc_icra2021_b200.solver import AltroError
    prob, opts, _, _ = cases.case_random_linear(batch=8)
    g = gpu_solver(prob, opts)
    g.solve()
    g.set_noise_model(1, 0.01, 0.0)
    g.set_noise_bank(mpc.rng_for(1, 1).standard_normal((6, 8, 12)))
    r = g.mpc_run(6, shift=True)
    s = g.fetch()  # statistics of the run's final solve, not of its first step
    assert np.array_equal(s.iterations, r["iterations"][-1]) and np.array_equal(s.cost, r["cost"][-1])
    g.mpc_run(3, shift=True)
    with pytest.raises(AltroError):
        g.run_results(6)            # the last run had 3 steps
    g.mpc_transition(None, shift=True)
    with pytest.raises(AltroError):
        g.mpc_run(2, shift=True)    # a transition is pending: its solve comes first
    g.solve()
    g.mpc_run(2, shift=True)


def test_trace_matches_oracle_trace():
    prob, opts, _, _ = cases.case_quadruped(True, batch=4)
    pg = copy.deepcopy(prob)
    o = OracleSolver(prob, opts, nthreads=1)
    tro = o.op.set_trace(16)
    o.solve()
    o.op.set_trace(0)
    g = gpu_solver(pg, opts)
    g.set_trace(16)
    g.solve()
    trg = g.get_trace()
    assert np.array_equal(np.nan_to_num(trg, nan=-1.0), np.nan_to_num(tro, nan=-1.0))


# ------------------------------------------------------------------ error behaviour of the C ABI

def test_abi_rejects_bad_arguments():
    from altro_mpc_icra2021_b200 import solver as S

    lib = S.load_library()
    h = ctypes.c_void_p()
    assert lib.altro_create(ctypes.byref(h), 0, 0, 3, 21, 4, 0.1) != 0
    assert lib.altro_create(ctypes.byref(h), 99, 6, 3, 21, 4, 0.1) != 0
    assert lib.altro_create(ctypes.byref(h), 0, 6, 3, 21, 4, 0.1) == 0
    inds = np.arange(3, dtype=np.int32)
    G, hv = np.eye(4, 3), np.zeros(4)
    cid = ctypes.c_int()
    add = lib.altro_add_constraint
    assert add(h, 2, 1, 0, 21, 4, 3, S._p(inds), 0, 0, S._p(G), S._p(hv), ctypes.byref(cid)) == -1  # control block past N-1
    assert b"knot range" in lib.altro_last_error(h)
    bad = np.array([0, 1, 7], dtype=np.int32)
    assert add(h, 2, 1, 0, 20, 4, 3, S._p(bad), 0, 0, S._p(G), S._p(hv), ctypes.byref(cid)) == -1
    assert add(h, 2, 1, 0, 20, 4, 3, S._p(inds), 0, 0, S._p(G), S._p(hv), ctypes.byref(cid)) == 0 and cid.value == 0
    assert lib.altro_solve(h) == -4 and b"dynamics and cost" in lib.altro_last_error(h)  # nothing uploaded yet
    assert lib.altro_destroy(h) == 0


def test_long_horizon_with_wide_gains_uses_the_global_workspace():
    """n = 40, m = 30, N = 101: the gains alone (100 x 30 x 40 doubles) exceed shared memory -> workspace layout."""
    solve_both(lqr_problem(n=40, m=30, N=101, batch=2, seed=1, u_bnd=0.4), SolverOptions(constraint_tolerance=1e-6))


def test_oversized_problem_is_an_error_not_a_fallback():
    # the trajectories themselves (2 x N x (n + m) doubles) must fit in shared memory even with the global workspace
    prob = lqr_problem(n=200, m=20, N=121, batch=2, seed=1)
    from altro_mpc_icra2021_b200.solver import AltroError

    with pytest.raises(AltroError, match="shared memory"):
        gpu_solver(prob, SolverOptions()).solve()


# ------------------------------------------------------------------ closed-loop run in one launch

@pytest.mark.parametrize("family", ["rocket", "random_linear"])
def test_closed_loop_run_matches_oracle_and_stepwise_path(family):
    """altro_mpc_run (steps x {transition; solve} per instance in ONE launch) against the oracle's orc_mpc_run and
    against the same steps issued one launch at a time."""
    B, steps = 96, 5
    if family == "rocket":
        cold = np.load(os.path.join(GOLD, "rocket_cold.npz"))
        Xt, Ut = cold["X"], cold["U"]
        prob, opts, _, _ = cases.case_rocket_mpc(Xt, Ut, batch=B)
        ks = mpc.rng_for(11, 0).integers(0, Xt.shape[0] - 21 - 110, size=B)
        model = (2, 1e-3, 1e-2)
    else:
        prob, Xt, Ut, ks = random_linear.mpc_problem(12, 6, 21, batch=B, seed=12)
        opts, model = random_linear.mpc_options(), (1, 0.01, 0.0)
    noise = mpc.rng_for(3, 3).standard_normal((steps, B, prob.n))
    pg, ps = copy.deepcopy(prob), copy.deepcopy(prob)
    o = OracleSolver(prob, opts, nthreads=8).solve()
    g, s = gpu_solver(pg, opts).solve(), gpu_solver(ps, opts).solve()
    for sv in (g, s):
        sv.set_track(Xt, Ut, ks)
        sv.set_noise_model(*model)
        sv.set_noise_bank(noise)
    rg = g.mpc_run(steps)
    ro = o.op.mpc_run(opts, steps, noise, model, (Xt, Ut), ks, True, nthreads=8)
    for k in ro:
        assert np.array_equal(rg[k], ro[k]), k
    assert np.array_equal(pg.X, prob.X) and np.array_equal(pg.U, prob.U) and np.array_equal(g.get_duals(), o.op.lam)
    assert np.array_equal(g.get_x0(), prob.x0)
    assert np.all(rg["status"] == 1)
    for st in range(steps):  # the same run, one transition + one solve launch per step
        s.mpc_transition(None, shift=True)
        s.solve()
        assert np.array_equal(s.stats.iterations, rg["iterations"][st]), st
        assert np.array_equal(ps.X[:, 0], rg["x0"][st]) and np.array_equal(ps.U[:, 0], rg["u0"][st])
    assert np.array_equal(ps.X, pg.X) and np.array_equal(ps.U, pg.U)
    # and the run can be continued: two runs of 2 + 3 steps equal one of 5
    pc = copy.deepcopy(prob)  # prob was advanced by the oracle run; rebuild the starting point instead


@pytest.mark.parametrize("chunk", [0, 1, 3, 100])
def test_closed_loop_run_scheduling_does_not_change_bits(chunk):
    """Persistent grid + work queue of (instance, chunk of steps) items vs one CTA per instance for the whole run."""
    B, steps = 300, 7
    cold = np.load(os.path.join(GOLD, "rocket_cold.npz"))
    Xt, Ut = cold["X"], cold["U"]
    prob, opts, _, _ = cases.case_rocket_mpc(Xt, Ut, batch=B)
    ks = mpc.rng_for(11, 0).integers(0, Xt.shape[0] - 21 - 110, size=B)
    noise = mpc.rng_for(3, 3).standard_normal((steps, B, prob.n))
    pg = copy.deepcopy(prob)
    o = OracleSolver(prob, opts, nthreads=8).solve()
    g = gpu_solver(pg, opts).solve()
    g.set_run_queue(chunk)
    g.set_track(Xt, Ut, ks)
    g.set_noise_model(2, 1e-3, 1e-2)
    g.set_noise_bank(noise)
    rg = g.mpc_run(steps)
    ro = o.op.mpc_run(opts, steps, noise, (2, 1e-3, 1e-2), (Xt, Ut), ks, True, nthreads=8)
    for k in ro:
        assert np.array_equal(rg[k], ro[k]), k
    assert np.array_equal(pg.X, prob.X) and np.array_equal(pg.U, prob.U) and np.array_equal(g.get_duals(), o.op.lam)
    rg2 = g.mpc_run(2)  # and again from where it stopped
    ro2 = o.op.mpc_run(opts, 2, noise[:2], (2, 1e-3, 1e-2), (Xt, Ut), ks + steps, True, nthreads=8)
    assert np.array_equal(rg2["x0"], ro2["x0"]) and np.array_equal(rg2["iterations"], ro2["iterations"])


@pytest.mark.parametrize("linearized", [True, False])
def test_quadruped_closed_loop_run_with_gait_schedule(linearized):
    B, steps = 64, 5
    prob, _ = quadruped.mpc_problem(B, linearized_friction=linearized, seed=41, gait_slots=steps + 2)
    opts = quadruped.mpc_options()
    noise = mpc.rng_for(6, 6).standard_normal((steps, B, 12))
    pg = copy.deepcopy(prob)
    o, g = OracleSolver(prob, opts, nthreads=8).solve(), gpu_solver(pg, opts).solve()
    assert_bit_identical(pg, g.stats, g.get_duals(), o.stats, "initial")
    g.set_noise_model(0, 1e-3, 0.0)
    g.set_noise_bank(noise)
    rg = g.mpc_run(steps)
    ro = o.op.mpc_run(opts, steps, noise, (0, 1e-3, 0.0), None, None, True, nthreads=8)
    for k in ro:
        assert np.array_equal(rg[k], ro[k]), k
    assert np.array_equal(pg.X, prob.X) and np.array_equal(pg.U, prob.U) and np.array_equal(g.get_duals(), o.op.lam)
    assert np.all(rg["status"] == 1)
    # one more tick through the per-step API continues the same schedule
    prob.set_initial_state(prob.X[:, 1, :].copy())
    o.op.shift_fill(True, True)
    o.op.step_abs = steps + 1
    o.solve()
    g.set_noise_bank(None)
    g.mpc_transition(None, shift=True)
    g.solve()
    assert np.array_equal(pg.X, prob.X) and np.array_equal(g.stats.iterations, o.stats.iterations)


def test_grasp_cold_solve_and_closed_loop_run():
    """Grasp family: cold solve N=251 (expansion blocks spill to global memory), then a closed-loop MPC run whose
    constraint data follow each instance along shared timelines."""
    from altro_mpc_icra2021_b200.problems import grasp

    cold = grasp.cold_problem()
    cg = copy.deepcopy(cold)
    oc = OracleSolver(cold, grasp.cold_options(), nthreads=1).solve()
    gc = gpu_solver(cg, grasp.cold_options()).solve()
    assert_bit_identical(cg, gc.stats, gc.get_duals(), oc.stats, "cold")
    assert gc.stats.status[0] == 1
    Xt, Ut = cold.X[0].copy(), cold.U[0].copy()
    B, steps = 48, 4
    prob, ks = grasp.mpc_problem(cold, Xt, Ut, 21, batch=B, seed=8)
    pg = copy.deepcopy(prob)
    opts = grasp.mpc_options()
    o, g = OracleSolver(prob, opts, nthreads=8).solve(), gpu_solver(pg, opts)
    g.set_track(Xt, Ut, ks)
    g.solve()
    assert_bit_identical(pg, g.stats, g.get_duals(), o.stats, "first MPC solve")
    noise = mpc.rng_for(9, 9).standard_normal((steps, B, 6))
    g.set_noise_model(1, 0.01, 0.0)
    g.set_noise_bank(noise)
    rg = g.mpc_run(steps)
    ro = o.op.mpc_run(opts, steps, noise, (1, 0.01, 0.0), (Xt, Ut), None, True, nthreads=8)
    for k in ro:
        assert np.array_equal(rg[k], ro[k]), k
    assert np.array_equal(pg.X, prob.X) and np.array_equal(g.get_duals(), o.op.lam) and np.all(rg["status"] == 1)
    assert np.array_equal(pg.kidx, prob.kidx) and np.array_equal(pg.kidx, ks + steps)
