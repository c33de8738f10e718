"""Rocket soft-landing family (benchmarks/rocket_landing).

RocketModel: rocket_landing_problem.jl:17-40 (exponential discretisation of the affine double
integrator).  RocketProblem: :44-186 with the script's values run_simple_rocket.jl:32-62.
Cold-solve options: run_simple_rocket.jl:39-50.  MPC options: :121-129.  Noise: simple_rocket.jl:59-71.
"""
from __future__ import annotations

import numpy as np

from ..problem import (ConstraintList, GoalConstraint, LinearModel, LQRObjective, NormConstraint, NormConstraint2,
                       Problem, SecondOrderCone, SolverOptions)
from .mpc import gen_tracking_problem, rng_for


def rocket_model(mass: float, grav, dt: float) -> LinearModel:
    """Exact ZOH of xdd = u/mass + g (omega_planet = 0): A=[I dt I;0 I], B=[dt^2/2m I; dt/m I]."""
    I3, Z3 = np.eye(3), np.zeros((3, 3))
    A = np.block([[I3, dt * I3], [Z3, I3]])
    B = np.vstack([0.5 * dt * dt / mass * I3, dt / mass * I3])
    g = np.asarray(grav, dtype=float)
    d = np.concatenate([0.5 * dt * dt * g, dt * g])
    return LinearModel(A, B, d, dt=dt)


def cold_problem(N: int = 301, dt: float = 0.05, x0=(4.0, 2.0, 20.0, -3.0, 2.0, -5.0), Qk=1e-2, Qfk=1e4, Rk=1.0,
                 mass=10.0, gravity=(0.0, 0.0, -9.81), per_weight_max=2.0, theta_thrust_max=5.0,
                 theta_glideslope=45.0, glide_recover_k=8, include_goal=True) -> Problem:
    n, m = 6, 3
    model = rocket_model(mass, gravity, dt)
    obj = LQRObjective(np.full(n, Qk), np.full(m, Rk), np.full(n, Qfk), np.zeros(n), N)
    cons = ConstraintList(n, m, N)
    if include_goal:
        cons.add_constraint(GoalConstraint(np.zeros(n)), N - 1)
    u_bnd = mass * abs(gravity[2]) * per_weight_max
    cons.add_constraint(NormConstraint(n, m, u_bnd, SecondOrderCone, ":control"), (0, N - 1), "max_thrust")
    a_max = np.tan(np.deg2rad(theta_thrust_max))
    cons.add_constraint(NormConstraint2(n, m, np.diag([1.0, 1.0, 0.0]), np.array([0.0, 0.0, a_max]), SecondOrderCone,
                                        ":control"), (0, N - 1), "thrust_angle")
    a_gl = np.tan(np.deg2rad(theta_glideslope))
    Ag = np.zeros((6, 6))
    Ag[0, 0] = Ag[1, 1] = 1.0
    cg = np.zeros(6)
    cg[2] = a_gl
    cons.add_constraint(NormConstraint2(n, m, Ag, cg, SecondOrderCone, ":state"), (glide_recover_k - 1, N - 1),
                        "glideslope")
    U0 = np.broadcast_to(-mass * np.asarray(gravity), (N - 1, m)).copy()  # hover
    return Problem(model, obj, N, x0=np.asarray(x0, float), constraints=cons, U0=U0)


def cold_options() -> SolverOptions:
    return SolverOptions(cost_tolerance_intermediate=1e-4, penalty_scaling=500.0, penalty_initial=1e-2,
                         projected_newton=False, constraint_tolerance=1e-5, iterations=5000, iterations_inner=100,
                         iterations_linesearch=100, iterations_outer=500)


def mpc_options() -> SolverOptions:
    return SolverOptions(cost_tolerance=1e-4, cost_tolerance_intermediate=1e-4, constraint_tolerance=1e-4,
                         reset_duals=False, penalty_initial=1000.0, penalty_scaling=10.0, projected_newton=False)


def noise(x0: np.ndarray, rng: np.random.Generator, wp: float = 1e-3, wv: float = 1e-2) -> np.ndarray:
    pos = np.linalg.norm(x0[..., :3], axis=-1, keepdims=True)
    vel = np.linalg.norm(x0[..., 3:], axis=-1, keepdims=True)
    return np.concatenate([rng.standard_normal(x0[..., :3].shape) * pos * wp,
                           rng.standard_normal(x0[..., 3:].shape) * vel * wv], axis=-1)


def mpc_problem(cold: Problem, X_track, U_track, N_mpc: int = 21, batch: int = 1, seed: int = 0xA1720 + 2,
                spread_starts: bool = True, x0_sigma=(0.0,) * 6):
    """gen_tracking_problem(prob, 21) (run_simple_rocket.jl:130-131).  Instances start at different
    indices along the cold-solved track (spread_starts) and may get a perturbed initial state."""
    rng = rng_for(seed, 0)
    Nl = X_track.shape[0]
    k_start = rng.integers(0, Nl - N_mpc - 110, size=batch) if (spread_starts and batch > 1) else np.zeros(batch, int)
    pm = gen_tracking_problem(cold, X_track, U_track, N_mpc, batch=batch, k_start=k_start)
    if any(s > 0 for s in x0_sigma):
        pm.set_initial_state(pm.x0 + rng.standard_normal(pm.x0.shape) * np.asarray(x0_sigma))
    return pm, k_start
