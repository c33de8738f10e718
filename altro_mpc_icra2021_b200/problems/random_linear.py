"""Random linear MPC family (benchmarks/random_linear_mpc).

gendiscrete: random_linear.jl:26-41.  gen_random_linear: random_linear_problem.jl:5-32.
gen_trajectory: run_random_linear.jl:29-39.  Options: run_random_linear.jl:41-49.
Julia's MersenneTwister streams cannot be reproduced without Julia; a Philox stream is used instead.
"""
from __future__ import annotations

import numpy as np

from ..problem import BoundConstraint, ConstraintList, LinearModel, LQRObjective, Problem, SolverOptions
from .mpc import gen_tracking_problem, rng_for


def gendiscrete(n: int, m: int, rng: np.random.Generator, tol: float = 1e-4):
    """A = Q diag(v/(|v|_inf+tol)) Q', Q from QR of a Gaussian matrix; B Gaussian."""
    v = rng.standard_normal(n)
    v = v / (np.abs(v).max() + tol)
    Qm, _ = np.linalg.qr(rng.standard_normal((n, n)))
    return Qm @ np.diag(v) @ Qm.T, rng.standard_normal((n, m))


def gen_random_linear(n: int, m: int, N: int, dt: float = 0.1, rng=None, u_bnd: float = 3.0) -> Problem:
    rng = rng_for(0xA1720, 100) if rng is None else rng
    A, B = gendiscrete(n, m, rng)
    model = LinearModel(A, B, dt=dt)
    Q = 10.0 * rng.random(n)
    obj = LQRObjective(Q, 0.1 * np.ones(m), Q * (N - 1), np.zeros(n), N)
    cons = ConstraintList(n, m, N)
    cons.add_constraint(BoundConstraint(n, m, u_min=-u_bnd, u_max=u_bnd), (0, N - 1))
    return Problem(model, obj, N, x0=np.zeros(n), constraints=cons)


def gen_trajectory(n: int, m: int, N: int, dt: float = 0.1, rng=None):
    """Reference to track: U ~ N(0,1), X rolled out from 0 (run_random_linear.jl:29-39)."""
    rng = rng_for(0xA1720, 100) if rng is None else rng
    prob = gen_random_linear(n, m, N, dt, rng)
    U = rng.standard_normal((N - 1, m))
    X = np.zeros((N, n))
    for k in range(N - 1):
        X[k + 1] = prob.model.A @ X[k] + prob.model.B @ U[k]
    return prob, X, U


def mpc_options() -> SolverOptions:
    return SolverOptions(cost_tolerance=1e-4, cost_tolerance_intermediate=1e-4, constraint_tolerance=1e-4,
                         penalty_initial=1000.0, penalty_scaling=100.0, reset_duals=False, projected_newton=False)


def noise(x0: np.ndarray, rng: np.random.Generator) -> np.ndarray:
    """1 % noise: randn(n) * |x0|_inf / 100 (random_linear_problem.jl:129)."""
    return rng.standard_normal(x0.shape) * np.abs(x0).max(axis=-1, keepdims=True) / 100.0


def mpc_problem(n=12, m=6, N_mpc=21, batch=1, N_track=1101, dt=0.1, seed=0xA1720 + 1, spread_starts=True):
    """Tracking MPC batch: instance i tracks the same reference from start index k_start[i]."""
    rng = rng_for(seed, 0)
    prob, X, U = gen_trajectory(n, m, N_track, dt, rng)
    if spread_starts and batch > 1:
        k_start = rng.integers(0, max(1, N_track - N_mpc - 200), size=batch)
    else:
        k_start = np.zeros(batch, dtype=np.int64)
    pm = gen_tracking_problem(prob, X, U, N_mpc, batch=batch, k_start=k_start)
    return pm, X, U, k_start
