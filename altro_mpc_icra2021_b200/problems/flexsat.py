"""Flexible spacecraft MPC family (benchmarks/flexible_satellite/flexible_sat_mpc.jl).

generate_AB: :72-130 (ZOH via c2d :59-70).  Problem: :133-161.  MPC options: :250-257.
Per-step update (no shifting): :264-272.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import expm

from ..problem import BoundConstraint, ConstraintList, LinearModel, LQRObjective, Problem, SolverOptions


def c2d(A, B, dt):
    n, p = A.shape[0], B.shape[1]
    M = np.zeros((n + p, n + p))
    M[:n, :n], M[:n, n:] = A * dt, B * dt
    E = expm(M)
    return E[:n, :n], E[:n, n:]


def generate_AB(dt: float = 0.5):
    J = np.diag([1.0, 2.0, 3.0])
    B_sc = np.eye(3)
    delta = np.array([[0, 0, 1], [0, 1, 0], [-0.7, 0.1, 0.1]], dtype=float)
    T = np.linalg.inv(J - delta.T @ delta)
    zeta = np.array([0.001, 0.001, 0.001])
    Delta = np.array([0.05, 0.2, 0.125]) * (2 * np.pi)
    Cm, Km = np.diag(2 * zeta * Delta), np.diag(Delta ** 2)
    Z3 = np.zeros((3, 3))
    A = np.block([[Z3, 0.25 * np.eye(3), Z3, Z3],
                  [Z3, Z3, T @ delta.T @ Km, T @ delta.T @ Cm],
                  [Z3, Z3, Z3, np.eye(3)],
                  [Z3, Z3, -Km - delta @ T @ delta.T @ Km, -Cm - delta @ T @ delta.T @ Cm]])
    B = np.vstack([Z3, -T @ B_sc, Z3, delta @ T @ B_sc])
    return c2d(A, B, dt)


def mpc_options(tol: float = 1e-4) -> SolverOptions:
    return SolverOptions(constraint_tolerance=tol, cost_tolerance=tol, cost_tolerance_intermediate=tol,
                         penalty_initial=100.0, penalty_scaling=100.0, projected_newton=False)


def noise(x0, rng):
    return 2e-4 * rng.standard_normal(x0.shape)


def mpc_problem(N: int = 80, batch: int = 1, seed: int = 0xA1720 + 4, x0_sigma: float = 1e-2) -> Problem:
    Ad, Bd = generate_AB()
    n, m = Bd.shape
    model = LinearModel(Ad, Bd, dt=0.1)  # dt only scales the stage cost (flexible_sat_mpc.jl:144-146)
    obj = LQRObjective(10.0 * np.ones(n), 0.1 * np.ones(m), 10.0 * np.ones(n), np.zeros(n), N)
    cons = ConstraintList(n, m, N)
    cons.add_constraint(BoundConstraint(n, m, u_min=-0.01, u_max=0.01), (0, N))
    x0 = np.zeros((batch, n))
    x0[:, :3] = 0.1
    if batch > 1:
        from .mpc import rng_for
        x0 += x0_sigma * rng_for(seed, 0).standard_normal((batch, n))
    return Problem(model, obj, N, x0=x0, constraints=cons, batch=batch)
