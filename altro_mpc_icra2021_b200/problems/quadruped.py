"""Quadruped (Woofer) convex-MPC family (benchmarks/quadruped/Woofer/MPCControl).

AltroParams: Structs/ALTROParams.jl:32-108.  Dynamics: linearized_dynamics.jl:1-66 linearised at
(x_des, u_ref=0) and discretised A_k = I + A_c dt, B_k = B_c dt, d_k = d_c dt (altro_solver.jl:25-37).
Friction: LinearizedFrictionConstraint.jl:14-25 / FrictionConstraint.jl:1-8.  Weights, mu, force
limits, N, dt: MPC.yaml.  Mass / inertia / geometry: Woofer.yaml, Config.jl:56-62.
Trot schedule: Structs/GaitParams.jl:38-49, gait.jl:1-9.  MuJoCo is replaced by a synthetic batch:
per-instance gait-phase offset, state noise and foot-position noise.
"""
from __future__ import annotations

import numpy as np

from ..problem import (BoundConstraint, ConstraintList, Inequality, LinearConstraint, LinearModel, LQRObjective,
                       NormConstraint2, Problem, SecondOrderCone, SolverOptions, CONTROL)
from .mpc import rng_for

MASS = 3.0 + 4 * 1.033 + 8 * 0.070  # sprung mass, Config.jl:58
J_BODY = np.diag([0.025, 0.854, 0.897])
N_HORIZON, DT, MU, F_MIN, F_MAX, STANCE_H = 15, 0.03, 0.5, 0.0, 133.0, 0.28
Q_DIAG = np.array([1.0, 1.0, 500.0, 5000.0, 5000.0, 1000.0, 500.0, 1000.0, 1000.0, 500.0, 500.0, 100.0])
R_DIAG = np.array([1.0, 1.0, 0.001] * 4)
_LZ = -np.sqrt(0.32 ** 2 - 0.18 ** 2)
NOM_FOOT = np.array([[0.23, -0.173, _LZ], [0.23, 0.173, _LZ], [-0.23, -0.173, _LZ], [-0.23, 0.173, _LZ]])  # FR FL BR BL
# trot(): rows = feet, columns = phases (GaitParams.jl:40-45); 0.2 s per phase (MPC.yaml:2-5)
TROT = np.array([[1, 1, 1, 0], [1, 0, 1, 1], [1, 0, 1, 1], [1, 1, 1, 0]], dtype=float)
PHASE_T = np.array([0.2, 0.2, 0.2, 0.2])
X_DES = np.array([0, 0, STANCE_H, 0, 0, 0, 0, 0, 0, 0, 0, 0], dtype=float)
U_HOVER = 9.81 * MASS / 4 * np.array([0, 0, 1.0] * 4)


def skew(v):
    z = np.zeros(v.shape[:-1])
    return np.stack([np.stack([z, -v[..., 2], v[..., 1]], -1), np.stack([v[..., 2], z, -v[..., 0]], -1),
                     np.stack([-v[..., 1], v[..., 0], z], -1)], -2)


def contact_schedule(t: np.ndarray, N: int = N_HORIZON, dt: float = DT) -> np.ndarray:
    """contacts[b,k,foot] for knot k at time t + k dt (footsteps.jl:39-56, gait.jl:1-9)."""
    tk = (t[:, None] + np.arange(N - 1)[None, :] * dt) % PHASE_T.sum()
    phase = np.searchsorted(np.cumsum(PHASE_T), tk, side="right")
    return TROT.T[phase]  # (B, N-1, 4)


def linearized_dynamics(contacts: np.ndarray, foot_rel: np.ndarray, dt: float = DT):
    """A_k, B_k, d_k at x_ref = x_des (rot = I, omega = 0), u_ref = 0.
    contacts (B,N-1,4); foot_rel (B,4,3) = world foot position minus body position."""
    Bn, Nk = contacts.shape[:2]
    Ac = np.zeros((12, 12))
    Ac[0:3, 6:9] = np.eye(3)
    Ac[3:6, 9:12] = 0.25 * np.eye(3)  # MRP kinematics at phi = 0
    A = np.broadcast_to(np.eye(12) + Ac * dt, (Bn, Nk, 12, 12)).copy()
    Bc = np.zeros((Bn, Nk, 12, 12))
    Jinv = np.linalg.inv(J_BODY)
    sk = Jinv @ skew(foot_rel)  # (B,4,3,3)
    for i in range(4):
        c = contacts[:, :, i][:, :, None, None]
        Bc[:, :, 6:9, 3 * i:3 * i + 3] = c * np.eye(3) / MASS
        Bc[:, :, 9:12, 3 * i:3 * i + 3] = c * sk[:, None, i]
    d = np.zeros((Bn, Nk, 12))
    d[:, :, 8] = -9.81 * dt
    return A, Bc * dt, d


def friction_rows(mu: float = MU):
    """(fx - mu fz, -mu fz - fx, fy - mu fz, -mu fz - fy) <= 0 on one foot's (fx,fy,fz)."""
    return np.array([[1, 0, -mu], [-1, 0, -mu], [0, 1, -mu], [0, -1, -mu]], dtype=float)


def mpc_options(tol: float = 1e-4) -> SolverOptions:
    return SolverOptions(cost_tolerance=tol, cost_tolerance_intermediate=1e-3, constraint_tolerance=tol,
                         projected_newton=False, penalty_initial=10.0, penalty_scaling=100.0, reset_duals=False,
                         static_bp=True)


def sample_batch(batch: int, seed: int = 0xA1720 + 3):
    """Per-instance gait time offset, current state and foot positions."""
    rng = rng_for(seed, 0)
    t0 = rng.random(batch) * PHASE_T.sum()
    sig = np.array([0.01] * 3 + [0.01] * 3 + [0.05] * 3 + [0.1] * 3)
    x_curr = X_DES + rng.standard_normal((batch, 12)) * sig
    foot_rel = NOM_FOOT[None] + 0.01 * rng.standard_normal((batch, 4, 3))
    return t0, x_curr, foot_rel


def gait_slot_model(t0: np.ndarray, foot_rel: np.ndarray, ticks: int, N: int = N_HORIZON, dt: float = DT):
    """The same LTV models as `linearized_dynamics`, stored once per gait phase: at the benchmark's linearisation point
    A_k and d_k are constant and B_k depends on the knot only through the contact pattern, i.e. through the gait phase.
    Returns A, B, d with shape (B, 4, ...) and the schedule sched[b, j] = phase at time t0_b + j*dt (update_dt == dt,
    MPC.yaml:22,52), for `ticks` control ticks of an N-knot horizon."""
    Bn = t0.shape[0]
    contacts = np.broadcast_to(TROT.T[None], (Bn, 4, 4))  # [b, phase, foot]
    A, Bm, d = linearized_dynamics(contacts, foot_rel, dt)
    tj = (t0[:, None] + np.arange(ticks + N)[None, :] * dt) % PHASE_T.sum()
    sched = np.searchsorted(np.cumsum(PHASE_T), tj, side="right").astype(np.int32)
    return A, Bm, d, sched


def mpc_problem(batch: int = 1, linearized_friction: bool = True, seed: int = 0xA1720 + 3, N: int = N_HORIZON,
                gait_slots: int = 0):
    """gait_slots = number of control ticks to schedule: the dynamics are then stored per gait phase with a schedule
    (closed-loop runs on the device); 0 = materialised per-knot A_k, B_k, d_k rebuilt by `advance` every tick."""
    n = m = 12
    t0, x_curr, foot_rel = sample_batch(batch, seed)
    if gait_slots:
        A, Bm, d, sched = gait_slot_model(t0, foot_rel, gait_slots, N)
        model = LinearModel(A, Bm, d, dt=DT, sched=sched)
    else:
        A, Bm, d = linearized_dynamics(contact_schedule(t0, N), foot_rel)
        model = LinearModel(A, Bm, d, dt=DT)
    cons = ConstraintList(n, m, N)
    for i in range(4):
        idx = (CONTROL, np.arange(3 * i, 3 * i + 3))
        if linearized_friction:
            cons.add_constraint(LinearConstraint(n, m, friction_rows(), np.zeros(4), Inequality, idx), (0, N - 1),
                                f"friction{i}")
        else:
            cons.add_constraint(NormConstraint2(n, m, np.diag([1.0, 1.0, 0.0]), MU * np.array([0, 0, 1.0]),
                                                SecondOrderCone, idx, compact=False), (0, N - 1), f"friction_soc{i}")
    u_min = np.array([-np.inf, -np.inf, F_MIN] * 4)
    u_max = np.array([np.inf, np.inf, F_MAX] * 4)
    cons.add_constraint(BoundConstraint(n, m, u_min=u_min, u_max=u_max), (0, N), "fz_bound")
    obj = LQRObjective(Q_DIAG, R_DIAG, Q_DIAG, X_DES, N)
    prob = Problem(model, obj, N, x0=x_curr, constraints=cons, batch=batch,
                   X0=np.broadcast_to(X_DES, (N, n)), U0=np.broadcast_to(U_HOVER, (N - 1, m)))
    return prob, dict(t0=t0, foot_rel=foot_rel)


def advance(prob: Problem, state: dict, rng: np.random.Generator, update_dt: float = 0.03, sigma: float = 1e-3):
    """One control tick (altro_solver.jl:44-72 with the MuJoCo plant replaced by the linear model):
    plant step with the first control + noise, new contact schedule -> new B_k, set_initial_state!."""
    mdl = prob.model
    x = np.einsum("bij,bj->bi", mdl.A[:, 0], prob.X[:, 0]) + np.einsum("bij,bj->bi", mdl.B[:, 0], prob.U[:, 0]) \
        + mdl.d[:, 0]
    x = x + sigma * rng.standard_normal(x.shape)
    state["t0"] = state["t0"] + update_dt
    A, Bm, d = linearized_dynamics(contact_schedule(state["t0"], prob.N), state["foot_rel"])
    prob.set_dynamics(A, Bm, d)
    prob.set_initial_state(x)


# ---------------------------------------------------------------------------------------------------------------
# numpy restatement of the quadruped's pre-solve step (checker of the device kernels in csrc/quadruped.cu)

FOOT_RADIUS = 0.02  # Woofer.yaml:19
GAIT_ALPHA = 0.5    # GaitParams.jl:33


def mrp_rotation(p):
    """Rotation matrix (body -> world) of a modified Rodrigues parameter vector, Rotations.MRP."""
    p = np.asarray(p)
    n2 = p @ p
    S = np.array([[0, -p[2], p[1]], [p[2], 0, -p[0]], [-p[1], p[0], 0]], dtype=p.dtype)
    return np.eye(3, dtype=p.dtype) + (8.0 * (S @ S) + 4.0 * (1.0 - n2) * S) / (1.0 + n2) ** 2


def nonlinear_dynamics(x, u, r, contacts, J=J_BODY, mass=MASS):
    """NonLinearContinuousDynamics, linearized_dynamics.jl:1-36 (works on complex arguments: complex-step derivative)."""
    x, u = np.asarray(x), np.asarray(u)
    R = mrp_rotation(x[3:6])
    p, ph, v, w = x[0:3], x[3:6], x[6:9], x[9:12]
    kin = 0.25 * ((1.0 - ph @ ph) * w + 2.0 * np.cross(ph, w) + 2.0 * (ph @ w) * ph)  # Rotations.kinematics(MRP, w)
    fs = np.array([0.0, 0.0, -9.81], dtype=x.dtype if np.iscomplexobj(x) else u.dtype)
    ts = np.zeros(3, dtype=fs.dtype)
    for i in range(4):
        ui = u[3 * i:3 * i + 3]
        fs = fs + contacts[i] / mass * ui
        rb = R.T @ (r[i] - p)
        ts = ts + contacts[i] * np.cross(rb, R.T @ ui)
    wd = np.linalg.inv(J) @ (-np.cross(w, J @ w) + ts)
    return np.concatenate([v, kin, fs, wd])


def linearize_reference(x_ref, u_ref, foot, contacts, dt=DT, J=J_BODY, mass=MASS):
    """update_dynamics_matrices!, altro_solver.jl:5-42, for one knot: Jacobians by the complex-step derivative (exact to
    roundoff for this rational function; the reference uses ForwardDiff)."""
    x_ref, u_ref = np.asarray(x_ref, float), np.asarray(u_ref, float)
    h = 1e-30
    Ac, Bc = np.zeros((12, 12)), np.zeros((12, 12))
    for j in range(12):
        xz = x_ref.astype(complex)
        xz[j] += 1j * h
        Ac[:, j] = nonlinear_dynamics(xz, u_ref.astype(complex), foot, contacts, J, mass).imag / h
        uz = u_ref.astype(complex)
        uz[j] += 1j * h
        Bc[:, j] = nonlinear_dynamics(x_ref.astype(complex), uz, foot, contacts, J, mass).imag / h
    dc = nonlinear_dynamics(x_ref, u_ref, foot, contacts, J, mass) - Ac @ x_ref - Bc @ u_ref
    return np.eye(12) + Ac * dt, Bc * dt, dc * dt


def foot_history_reference(t, x_ref, cur_foot, planner, K, dt=DT, contact_phases=None, phase_times=PHASE_T,
                           nom_foot=NOM_FOOT, alpha=GAIT_ALPHA, foot_radius=FOOT_RADIUS):
    """foot_history! (footsteps.jl:29-84) + get_phase (gait.jl:1-9) + footstep_location (footsteps.jl:1-27) of one
    instance.  x_ref (12,) or (K,12); cur_foot (4,3) body frame; planner (4,3) in/out.  Returns contacts (K,4),
    foot (K,4,3) world and the updated planner state."""
    cp = TROT.T if contact_phases is None else np.asarray(contact_phases, float)  # (num_phases, 4)
    x_ref = np.asarray(x_ref, float)
    xr = (lambda k: x_ref[k]) if x_ref.ndim == 2 else (lambda k: x_ref)
    total = float(np.sum(phase_times))

    def phase_of(tt):
        pt, s = np.fmod(tt, total), 0.0
        for i, d in enumerate(phase_times):
            s += d
            if pt < s:
                return i
        return len(phase_times) - 1

    planner = np.array(planner, float)
    contacts, foot = np.zeros((K, 4)), np.zeros((K, 4, 3))
    prev_phase = phase_of(t)
    x = xr(0)
    prev = x[0:3] + (mrp_rotation(x[3:6]) @ np.asarray(cur_foot, float).T).T
    contacts[0], foot[0] = cp[prev_phase], prev
    t_i = t + dt
    for k in range(1, K):
        nxt = phase_of(t_i)
        x = xr(k)
        R = mrp_rotation(x[3:6])
        contacts[k] = cp[nxt]
        for j in range(4):
            if cp[prev_phase][j] == 1:
                if cp[nxt][j] == 0:
                    t_next = phase_times[(nxt + 1) % len(phase_times)]
                    loc = x[0:3] + R @ nom_foot[j] + alpha * t_next * x[6:9]
                    planner[j] = np.array([loc[0], loc[1], foot_radius])
            elif cp[nxt][j] == 1:
                prev[j] = planner[j]
        foot[k] = prev
        t_i += dt
        prev_phase = nxt
    return contacts, foot, planner
