"""Two-finger grasp of a rotating box (benchmarks/grasp_optimization).

SquareObject / orientation trajectory: src/grasp_model.jl:4-55, src/utils.jl:15-31.  Dynamics (double integrator
+ gravity, both finger forces): src/grasp_model.jl:74-92.  GraspProblem: src/grasp_problem.jl:1-107.  Options and
tracking weights: grasp_benchmark.jl:19-34,72-83.  Per-step constraint-data update: src/grasp_mpc_helpers.jl:26-55
(here: TrackConstraint timelines indexed by each instance's position, no rewrite needed).
"""
from __future__ import annotations

import numpy as np

from ..problem import (ConstraintList, Equality, GoalConstraint, Inequality, LinearModel, LQRObjective, Problem,
                       SecondOrderCone, SolverOptions, TrackConstraint, CONTROL)
from .mpc import gen_tracking_problem, rng_for

MU, MASS, F_MAX, G = 0.5, 0.2, 3.0, np.array([0.0, 0.0, -9.81])


def rot3(th):
    c, s = np.cos(th), np.sin(th)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def skew(a):
    return np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])


def orientation_trajectory(dt, tf, th0=0.0, thf=np.pi / 4, thd0=0.0, thdf=0.15):
    """Cubic theta(t) with the given end conditions; contact points p_i, inward normals v_i, torque maps skew(p_i)."""
    t0 = 0.0
    Am = np.array([[t0 ** 3, t0 ** 2, t0, 1], [tf ** 3, tf ** 2, tf, 1], [3 * t0 ** 2, 2 * t0, 1, 0],
                   [3 * tf ** 2, 2 * tf, 1, 0]])
    c = np.linalg.solve(Am, np.array([th0, thf, thd0, thdf]))
    ts = np.arange(0.0, tf + 0.5 * dt, dt)
    th = c[0] * ts ** 3 + c[1] * ts ** 2 + c[2] * ts + c[3]
    thdd = 6 * c[0] * ts + 2 * c[1]
    p0 = [np.array([0.0, -1.0, 0.0]), np.array([0.0, 1.0, 0.0])]
    v0 = [np.array([0.0, 1.0, 0.0]), np.array([0.0, -1.0, 0.0])]
    p = np.array([[rot3(a) @ p0[i] for a in th] for i in range(2)])  # (2, T, 3)
    v = np.array([[rot3(a) @ v0[i] for a in th] for i in range(2)])
    return th, thdd, p, v


def grasp_model(dt: float) -> LinearModel:
    I3, Z3 = np.eye(3), np.zeros((3, 3))
    A = np.block([[I3, dt * I3], [Z3, I3]])
    B = np.vstack([np.hstack([0.5 * dt * dt / MASS * I3] * 2), np.hstack([dt / MASS * I3] * 2)])
    d = np.concatenate([0.5 * dt * dt * G, dt * G])
    return LinearModel(A, B, d, dt=dt)


def constraint_timelines(thdd, p, v):
    """Per-time-step data of the four stage constraints, rows 0..T-1 of the timelines."""
    T = thdd.shape[0]
    At = np.zeros((T, 3, 6)); bt = np.zeros((T, 3))
    Ag = np.zeros((T, 2, 6)); bg = np.full((T, 2), F_MAX)
    Af = np.zeros((2, T, 4, 3))
    for k in range(T):
        At[k] = np.hstack([skew(p[0, k]), skew(p[1, k])])
        bt[k, 0] = thdd[k]
        Ag[k, 0, :3], Ag[k, 1, 3:] = v[0, k], v[1, k]
        for i in range(2):
            vv = v[i, k]
            Af[i, k, :3] = np.eye(3) - np.outer(vv, vv)
            Af[i, k, 3] = MU * vv
    return At, bt, Ag, bg, Af


def cold_problem(N: int = 251, tf: float = 6.0, x0=(0.0, 3.0, 3.0, 0.0, 0.0, 0.0)) -> Problem:
    n = m = 6
    dt = tf / (N - 1)
    _, thdd, p, v = orientation_trajectory(dt, tf)
    At, bt, Ag, bg, Af = constraint_timelines(thdd, p, v)
    obj = LQRObjective(np.full(n, 1e-3), np.ones(m), np.full(n, 10.0), np.zeros(n), N)
    cons = ConstraintList(n, m, N)
    cons.add_constraint(GoalConstraint(np.zeros(n)), N - 1)
    cons.add_constraint(TrackConstraint(n, m, At, bt, Equality, ":control"), (0, N - 1), "torque_balance")
    cons.add_constraint(TrackConstraint(n, m, Ag, bg, Inequality, ":control"), (0, N - 1), "max_grasp_force")
    for i in range(2):
        cons.add_constraint(TrackConstraint(n, m, Af[i], np.zeros((Af.shape[1], 4)), SecondOrderCone,
                                            (CONTROL, np.arange(3 * i, 3 * i + 3))), (0, N - 1), f"friction{i + 1}")
    u0 = np.array([0, -1.5, MASS * 9.81 / 2, 0, 1.5, MASS * 9.81 / 2])
    return Problem(grasp_model(dt), obj, N, x0=np.asarray(x0, float), constraints=cons,
                   U0=np.broadcast_to(u0, (N - 1, m)))


def cold_options() -> SolverOptions:
    return SolverOptions(projected_newton=False, cost_tolerance=1e-6, cost_tolerance_intermediate=1e-4,
                         constraint_tolerance=1e-6)


def mpc_options() -> SolverOptions:
    return SolverOptions(cost_tolerance=1e-4, cost_tolerance_intermediate=1e-3, constraint_tolerance=1e-4,
                         projected_newton=False, penalty_initial=1e4, penalty_scaling=100.0)


def noise(x0, rng):
    return rng.standard_normal(x0.shape) * np.abs(x0).max(axis=-1, keepdims=True) / 100.0


def mpc_problem(cold: Problem, X_track, U_track, N_mpc: int = 21, batch: int = 1, seed: int = 0xA1720 + 5,
                max_steps: int = 110):
    """gen_tracking_problem(prob_cold, N_mpc, Qk=1e3, Rk=1, Qfk=10) (grasp_benchmark.jl:79-80); instance i starts at
    timeline position k_start[i]."""
    rng = rng_for(seed, 0)
    Nl = X_track.shape[0]
    hi = max(1, Nl - N_mpc - max_steps)
    k_start = rng.integers(0, hi, size=batch) if batch > 1 else np.zeros(batch, dtype=np.int64)
    pm = gen_tracking_problem(cold, X_track, U_track, N_mpc, Qk=1e3, Rk=1.0, Qfk=10.0, batch=batch, k_start=k_start)
    return pm, k_start
