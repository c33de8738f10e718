"""Small deterministic cases shared by the golden-vector generator, the oracle test and the GPU parity test.

Each case returns (problem, options, n_steps, advance) where advance(prob, solver, step) performs the
reference's between-solve MPC update on the host (same code for the oracle and the CUDA path)."""
import numpy as np

from altro_mpc_icra2021_b200.problems import flexsat, mpc, quadruped, random_linear, rocket


def _track_advance(Xt, Ut, k0, noise_fn, seed):
    rng = mpc.rng_for(seed, 99)
    state = {"k": np.array(k0, dtype=np.int64).copy()}

    def advance(prob, solver, step):
        x0 = prob.X[:, 1, :] + noise_fn(prob.X[:, 1, :], rng)  # x_1 of the solution = plant step with u_0
        state["k"] += 1
        prob.set_initial_state(x0)
        prob.update_trajectory(*mpc.window_reference(Xt, Ut, state["k"], prob.N))
        solver.shift_fill(True, True)

    return advance


def rocket_track():
    """Cold-solved landing trajectory (run_simple_rocket.jl:31-67) -- itself a golden case."""
    return rocket.cold_problem(), rocket.cold_options()


def case_rocket_mpc(Xt, Ut, batch=6):
    cold = rocket.cold_problem()
    pm, ks = rocket.mpc_problem(cold, Xt, Ut, 21, batch=batch, seed=11)
    return pm, rocket.mpc_options(), 3, _track_advance(Xt, Ut, ks, rocket.noise, 11)


def case_random_linear(batch=5):
    pm, X, U, ks = random_linear.mpc_problem(12, 6, 21, batch=batch, seed=12)
    return pm, random_linear.mpc_options(), 3, _track_advance(X, U, ks, random_linear.noise, 12)


def case_quadruped(linearized, batch=5):
    pq, st = quadruped.mpc_problem(batch, linearized_friction=linearized, seed=13)
    rng = mpc.rng_for(13, 98)

    def advance(prob, solver, step):
        quadruped.advance(prob, st, rng)
        solver.shift_fill(True, True)

    return pq, quadruped.mpc_options(), 3, advance


def case_flexsat(batch=3):
    pf = flexsat.mpc_problem(40, batch=batch, seed=14)
    rng = mpc.rng_for(14, 97)

    def advance(prob, solver, step):  # no shifting (flexible_sat_mpc.jl:271-276)
        prob.set_initial_state(prob.X[:, 1, :] + flexsat.noise(prob.X[:, 1, :], rng))

    return pf, flexsat.mpc_options(), 2, advance


CASES = {
    "random_linear": case_random_linear,
    "quadruped_lin": lambda: case_quadruped(True),
    "quadruped_soc": lambda: case_quadruped(False),
    "flexsat": case_flexsat,
}


def run_case(make_solver, prob, opts, steps, advance):
    """Runs the MPC loop and returns the per-step results to be pinned."""
    solver = make_solver(prob, opts)
    out = {}
    for st in range(steps):
        solver.solve()
        s = solver.stats
        out[f"X{st}"], out[f"U{st}"] = prob.X.copy(), prob.U.copy()
        out[f"lam{st}"] = solver.get_duals()
        out[f"iters{st}"], out[f"outer{st}"] = s.iterations.copy(), s.iterations_outer.copy()
        out[f"status{st}"], out[f"ls{st}"] = s.status.copy(), s.ls_trials.copy()
        out[f"cost{st}"], out[f"cmax{st}"] = s.cost.copy(), s.c_max.copy()
        if st + 1 < steps:
            advance(prob, solver, st)
    return out
