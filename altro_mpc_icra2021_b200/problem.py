"""Host-side problem description: the Python mirror of the Julia data model the reference drives
ALTRO through (TrajectoryOptimization.Problem / ConstraintList / Objective, RobotDynamics.LinearModel).

Reference usage mirrored here (paths relative to /root/reference/benchmarks):
  Problem(model, obj, xf, tf, x0=..., constraints=...)   random_linear_mpc/random_linear_problem.jl:28-29
  ConstraintList(n,m,N) + add_constraint!(cons, con, inds)  random_linear_problem.jl:22-24, mpc.jl:32-40
  BoundConstraint(n,m,u_min=,u_max=)                      random_linear_problem.jl:23, ALTROParams.jl:75-78
  GoalConstraint(xf)                                      rocket_landing/rocket_landing_problem.jl:96
  NormConstraint(n,m,val,SecondOrderCone(),:control)      rocket_landing_problem.jl:123
  NormConstraint2(n,m,A,c,sense,inds)  ||A y|| <= c'y     grasp_optimization/src/new_constraints.jl:72-116
  LinearConstraint-style rows (A y - b)                    grasp_optimization/src/grasp_problem.jl:35-67
  LQRObjective(Q,R,Qf,xf,N) / TrackingObjective(Q,R,Z,Qf=) rocket_landing_problem.jl:83, mpc.jl:29
  TO.set_initial_state! / TO.update_trajectory!           random_linear_problem.jl:130,133

Every problem is a *batch* of B independent instances with identical structure (dimensions,
constraint list, weights); per-instance data are x0, the tracking reference, optionally the
dynamics and the constraint data.  All indices are 0-based here (Julia's 1:N-1 -> range(0, N-1)).
Only data are described -- user callbacks of the Julia API (TO.evaluate / TO.jacobian!) become
affine descriptors c(z) = G z[inds] + h, which is what every constraint in the reference is.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

EQUALITY, INEQUALITY, SECOND_ORDER_CONE = 0, 1, 2
STATE, CONTROL = 0, 1


class Equality:
    code = EQUALITY


class Inequality:
    code = INEQUALITY


class SecondOrderCone:
    code = SECOND_ORDER_CONE


def _sense_code(sense) -> int:
    if isinstance(sense, int):
        return sense
    return sense.code


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _bcast(a, shape) -> np.ndarray:
    """Writable C-contiguous float64 copy of `a` broadcast to `shape`."""
    return np.array(np.broadcast_to(np.asarray(a, dtype=np.float64), shape), dtype=np.float64, order="C")


# --------------------------------------------------------------------------- dynamics


class LinearModel:
    """Discrete affine dynamics x+ = A x + B u + d (RD.LinearModel, already discretised).

    A may be (n,n) shared LTI, (N-1,n,n) shared LTV, (B,n,n) per-instance LTI or (B,N-1,n,n)
    per-instance LTV; `per_instance` disambiguates the two 3-D cases.
    """

    def __init__(self, A, B, d=None, dt: float = 0.0, per_instance: Optional[bool] = None, sched=None):
        """`sched` (B, L) int32: gait-scheduled models -- A is (B, nslots, n, n) and knot k of the solve that follows
        s transitions uses slot sched[i, s + k] (altro_set_dynamics_slots)."""
        self.sched = None if sched is None else np.ascontiguousarray(sched, dtype=np.int32)
        A = _f64(A)
        Bm = _f64(B)
        n, m = A.shape[-1], Bm.shape[-1]
        if d is None:
            d = np.zeros(A.shape[:-2] + (n,))
        d = _f64(d)
        if A.ndim == 2:
            self.per_knot, self.per_instance = False, False
        elif A.ndim == 3:
            self.per_instance = bool(per_instance)
            self.per_knot = not self.per_instance
        elif A.ndim == 4:
            self.per_knot, self.per_instance = True, True
        else:
            raise ValueError("A must have 2, 3 or 4 dimensions")
        assert A.shape[-2:] == (n, n) and Bm.shape[-2:] == (n, m) and d.shape[-1] == n
        assert A.shape[:-2] == Bm.shape[:-2] == d.shape[:-1], "A, B, d leading dims must agree"
        self.A, self.B, self.d, self.dt, self.n, self.m = A, Bm, d, float(dt), n, m


# --------------------------------------------------------------------------- objective


class Objective:
    """Diagonal quadratic tracking cost (TO.LQRObjective / TO.TrackingObjective):
    J = sum_{k<N-1} dt (1/2 (x-xr)'Q(x-xr) + 1/2 (u-ur)'R(u-ur)) + 1/2 (x_N-xr_N)'Qf(x_N-xr_N)."""

    def __init__(self, Q, R, Qf, Xref, Uref):
        self.Q, self.R, self.Qf = _f64(Q), _f64(R), _f64(Qf)
        self.Xref, self.Uref = _f64(Xref), _f64(Uref)  # (N,n)/(B,N,n), (N-1,m)/(B,N-1,m)


def _diag(v, k):
    v = np.asarray(v, dtype=np.float64)
    if v.ndim == 0:
        return np.full(k, float(v))
    if v.ndim == 2:
        assert np.allclose(v, np.diag(np.diag(v))), "only diagonal weights are supported"
        return np.diag(v).copy()
    return v.copy()


def LQRObjective(Q, R, Qf, xf, N: int) -> Objective:
    xf = _f64(xf)
    n = xf.shape[-1]
    R_ = np.asarray(R, dtype=np.float64)
    m = R_.shape[-1] if R_.ndim else None
    assert m is not None, "R must be a vector or matrix so the control dimension is known"
    Xref = np.broadcast_to(xf, (N, n)).copy()
    return Objective(_diag(Q, n), _diag(R, m), _diag(Qf, n), Xref, np.zeros((N - 1, m)))


def TrackingObjective(Q, R, Xref, Uref, Qf=None) -> Objective:
    Xref, Uref = _f64(Xref), _f64(Uref)
    n, m = Xref.shape[-1], Uref.shape[-1]
    Qd = _diag(Q, n)
    return Objective(Qd, _diag(R, m), Qd if Qf is None else _diag(Qf, n), Xref, Uref)


# --------------------------------------------------------------------------- constraints


@dataclass
class FlatConstraint:
    """One affine conic block c = G z[inds] + h on knots [k0,k1) of one side (state|control)."""

    sense: int
    side: int
    k0: int
    k1: int
    inds: np.ndarray  # (w,) int32
    G: np.ndarray  # (p,w) | (nk,p,w) | (B,p,w) | (B,nk,p,w)
    h: np.ndarray  # (p,)  | ...
    per_knot: bool = False
    per_instance: bool = False
    name: str = ""
    track: bool = False  # G (Nt,p,w), h (Nt,p): shared timeline; knot k of instance i reads row min(kidx_i + k, Nt-1)

    @property
    def p(self) -> int:
        return self.G.shape[-2]

    @property
    def w(self) -> int:
        return self.G.shape[-1]


class StageConstraint:
    """Base: lower(n, m) returns [(side, inds, G, h, sense, per_knot, per_instance)]."""

    def lower(self, n, m):  # pragma: no cover - interface
        raise NotImplementedError


def _side_inds(n, m, inds):
    """Julia-style index sets over z=[x;u] (0-based here), ':state', ':control', or explicit
    (side, indices)."""
    if isinstance(inds, str):
        if inds in (":state", "state"):
            return STATE, np.arange(n)
        if inds in (":control", "control"):
            return CONTROL, np.arange(m)
        raise ValueError(inds)
    side, idx = inds
    return side, np.asarray(idx, dtype=np.int64)


class BoundConstraint(StageConstraint):
    """z_min <= z <= z_max on finite entries; rows [z - z_max; z_min - z] (TO.BoundConstraint)."""

    def __init__(self, n, m, x_min=-np.inf, x_max=np.inf, u_min=-np.inf, u_max=np.inf):
        self.x_min = np.broadcast_to(np.asarray(x_min, float), (n,)).copy()
        self.x_max = np.broadcast_to(np.asarray(x_max, float), (n,)).copy()
        self.u_min = np.broadcast_to(np.asarray(u_min, float), (m,)).copy()
        self.u_max = np.broadcast_to(np.asarray(u_max, float), (m,)).copy()

    @staticmethod
    def _rows(lo, hi):
        up = np.flatnonzero(np.isfinite(hi))
        dn = np.flatnonzero(np.isfinite(lo))
        idx = np.union1d(up, dn)
        if idx.size == 0:
            return None
        pos = {j: i for i, j in enumerate(idx)}
        G = np.zeros((up.size + dn.size, idx.size))
        h = np.zeros(up.size + dn.size)
        for r, j in enumerate(up):
            G[r, pos[j]] = 1.0
            h[r] = -hi[j]
        for r, j in enumerate(dn):
            G[up.size + r, pos[j]] = -1.0
            h[up.size + r] = lo[j]
        return idx, G, h

    def lower(self, n, m):
        out = []
        for side, lo, hi in ((STATE, self.x_min, self.x_max), (CONTROL, self.u_min, self.u_max)):
            rows = self._rows(lo, hi)
            if rows is not None:
                out.append((side, rows[0], rows[1], rows[2], INEQUALITY, False, False))
        return out


class GoalConstraint(StageConstraint):
    """x_N = xf (TO.GoalConstraint), optionally on a subset of state indices."""

    def __init__(self, xf, inds=None):
        self.xf = _f64(xf)
        self.inds = np.arange(self.xf.shape[-1]) if inds is None else np.asarray(inds)

    def lower(self, n, m):
        w = self.inds.size
        return [(STATE, self.inds, np.eye(w), -self.xf[self.inds], EQUALITY, False, False)]


class NormConstraint(StageConstraint):
    """||z[inds]|| <= val as SOC value [z[inds]; val]  (TO.NormConstraint, rocket_landing_problem.jl:123),
    or with Inequality sense the scalar row z'z - val^2 is NOT supported (non-affine)."""

    def __init__(self, n, m, val, sense=SecondOrderCone, inds=":control"):
        assert _sense_code(sense) == SECOND_ORDER_CONE, "only the SecondOrderCone NormConstraint is affine"
        self.val, self.inds = float(val), inds

    def lower(self, n, m):
        side, idx = _side_inds(n, m, self.inds)
        w = idx.size
        G = np.vstack([np.eye(w), np.zeros((1, w))])
        h = np.zeros(w + 1)
        h[-1] = self.val
        return [(side, idx, G, h, SECOND_ORDER_CONE, False, False)]


class NormConstraint2(StageConstraint):
    """||A y|| <= c'y with y = z[inds]; SOC value [A y; c'y] (new_constraints.jl:72-116,
    FrictionConstraint.jl:1-39).  `compact=True` drops all-zero rows of A (they change neither the
    norm nor the projection) and all-zero columns of [A; c']."""

    def __init__(self, n, m, A, c, sense=SecondOrderCone, inds=":control", compact=True):
        assert _sense_code(sense) == SECOND_ORDER_CONE
        self.A, self.c, self.inds, self.compact = _f64(A), _f64(c), inds, compact

    def lower(self, n, m):
        side, idx = _side_inds(n, m, self.inds)
        G = np.vstack([self.A, self.c[None, :]])
        if self.compact:
            keep_r = [r for r in range(self.A.shape[0]) if np.any(self.A[r] != 0.0)] + [self.A.shape[0]]
            G = G[keep_r]
            keep_c = [j for j in range(G.shape[1]) if np.any(G[:, j] != 0.0)]
            G, idx = G[:, keep_c], idx[keep_c]
        return [(side, idx, G, np.zeros(G.shape[0]), SECOND_ORDER_CONE, False, False)]


class LinearConstraint(StageConstraint):
    """A y - b (=|<=) 0 or (A y - b) in SOC, y = z[inds].  A may carry leading (nk,) and/or (B,)
    dimensions for per-knot / per-instance data (grasp_problem.jl:35-67, grasp_mpc_helpers.jl:46-55)."""

    def __init__(self, n, m, A, b, sense, inds=":control", per_knot=False, per_instance=False):
        self.A, self.b, self.sense, self.inds = _f64(A), _f64(b), _sense_code(sense), inds
        self.per_knot, self.per_instance = per_knot, per_instance

    def lower(self, n, m):
        side, idx = _side_inds(n, m, self.inds)
        return [(side, idx, self.A, -self.b, self.sense, self.per_knot, self.per_instance)]


class TrackConstraint(StageConstraint):
    """A y - b (=|<=|in SOC) with time-varying data given once along a timeline: A (Nt,p,w), b (Nt,p).  Knot k of an
    instance whose timeline position is kidx uses row kidx + k -- what grasp_mpc_helpers.jl:26-55 achieves by
    rewriting cons[i].A / .b / .c in place before every solve."""

    def __init__(self, n, m, A, b, sense, inds=":control"):
        self.A, self.b, self.sense, self.inds = _f64(A), _f64(b), _sense_code(sense), inds

    def lower(self, n, m):
        side, idx = _side_inds(n, m, self.inds)
        return [(side, idx, self.A, -self.b, self.sense, "track", False)]


class ConstraintList:
    """TO.ConstraintList(n,m,N) with add_constraint!(cons, con, knots)."""

    def __init__(self, n: int, m: int, N: int):
        self.n, self.m, self.N = n, m, N
        self.flat: List[FlatConstraint] = []
        self.source = []  # (constraint object, knot range) as added, for gen_tracking_problem

    def add_constraint(self, con: StageConstraint, knots, name: str = "") -> None:
        """knots: range / (start, stop) 0-based half-open / single int."""
        if isinstance(knots, int):
            k0, k1 = knots, knots + 1
        elif isinstance(knots, range):
            assert knots.step == 1
            k0, k1 = knots.start, knots.stop
        else:
            k0, k1 = knots
        assert 0 <= k0 <= k1 <= self.N
        self.source.append((con, (k0, k1), name))
        for side, idx, G, h, sense, per_knot, per_instance in con.lower(self.n, self.m):
            kk1 = min(k1, self.N - 1) if side == CONTROL else k1  # u_N is not a decision variable
            if kk1 <= k0:
                continue
            track = per_knot == "track"
            per_knot = False if track else per_knot
            nk_full, nk = k1 - k0, kk1 - k0
            if per_knot and nk != nk_full:  # drop per-knot data of the clipped terminal knot
                G = G[..., :nk, :, :]
                h = h[..., :nk, :]
            self.flat.append(
                FlatConstraint(sense, side, k0, kk1, np.asarray(idx, np.int32), _f64(G), _f64(h), per_knot,
                               per_instance, name or type(con).__name__, track)
            )

    def __len__(self):
        return len(self.flat)

    def dual_len(self) -> int:
        return sum((c.k1 - c.k0) * c.p for c in self.flat)

    def dual_offsets(self) -> List[int]:
        off, o = [], 0
        for c in self.flat:
            off.append(o)
            o += (c.k1 - c.k0) * c.p
        return off


# --------------------------------------------------------------------------- problem


class Problem:
    """TO.Problem for a batch of B instances.

    x0: (n,) or (B,n).  X0/U0: warm start (N,n)/(B,N,n), (N-1,m)/(B,N-1,m).
    Mutators mirror the in-place updates the reference's MPC loops perform between solves; each
    marks a dirty flag that ALTROSolver.solve() turns into a C-ABI upload.
    """

    def __init__(self, model: LinearModel, obj: Objective, N: int, x0, constraints: Optional[ConstraintList] = None,
                 batch: int = 1, X0=None, U0=None, tf: Optional[float] = None):
        n, m = model.n, model.m
        self.model, self.obj, self.N, self.n, self.m, self.B = model, obj, int(N), n, m, int(batch)
        self.dt = model.dt if tf is None else tf / (N - 1)
        self.constraints = constraints if constraints is not None else ConstraintList(n, m, N)
        B = self.B
        self.x0 = _bcast(x0, (B, n))
        self.Xref = _bcast(obj.Xref, (B, N, n))
        self.Uref = _bcast(obj.Uref, (B, N - 1, m))
        self.X = np.zeros((B, N, n)) if X0 is None else _bcast(X0, (B, N, n))
        self.U = np.zeros((B, N - 1, m)) if U0 is None else _bcast(U0, (B, N - 1, m))
        self.kidx = np.zeros(B, dtype=np.int32)  # position of every instance on the shared timelines
        if model.per_instance:
            assert model.A.shape[0] == B
        self.dirty = {"x0": True, "ref": True, "dyn": True, "traj": True, "con": set(range(len(self.constraints)))}

    def size(self):
        return self.n, self.m, self.N

    def slice(self, i0: int, i1: int) -> "Problem":
        """Sub-batch [i0, i1) as an independent Problem (the shard one GPU owns, sharding.shard_range)."""
        mdl = self.model
        if mdl.per_instance:
            model = LinearModel(mdl.A[i0:i1].copy(), mdl.B[i0:i1].copy(), mdl.d[i0:i1].copy(), dt=mdl.dt,
                                per_instance=True, sched=None if mdl.sched is None else mdl.sched[i0:i1].copy())
        else:
            model = LinearModel(mdl.A.copy(), mdl.B.copy(), mdl.d.copy(), dt=mdl.dt, per_instance=False)
        cons = ConstraintList(self.n, self.m, self.N)
        cons.source = list(self.constraints.source)
        for c in self.constraints.flat:
            G = c.G[i0:i1].copy() if c.per_instance else c.G.copy()
            h = c.h[i0:i1].copy() if c.per_instance else c.h.copy()
            cons.flat.append(FlatConstraint(c.sense, c.side, c.k0, c.k1, c.inds.copy(), G, h, c.per_knot,
                                            c.per_instance, c.name, c.track))
        obj = Objective(self.obj.Q, self.obj.R, self.obj.Qf, self.Xref[i0:i1], self.Uref[i0:i1])
        p = Problem(model, obj, self.N, self.x0[i0:i1], cons, batch=i1 - i0, X0=self.X[i0:i1], U0=self.U[i0:i1])
        p.dt = self.dt
        p.kidx[...] = self.kidx[i0:i1]
        return p

    # --- mutators (TO.set_initial_state!, TO.update_trajectory!, initial_controls!, model.A[i] = ...)
    def set_initial_state(self, x0) -> None:
        self.x0[...] = np.broadcast_to(_f64(x0), self.x0.shape)
        self.dirty["x0"] = True

    def update_trajectory(self, Xref, Uref) -> None:
        self.Xref[...] = np.broadcast_to(_f64(Xref), self.Xref.shape)
        self.Uref[...] = np.broadcast_to(_f64(Uref), self.Uref.shape)
        self.dirty["ref"] = True

    def initial_controls(self, U0) -> None:
        self.U[...] = np.broadcast_to(_f64(U0), self.U.shape)
        self.dirty["traj"] = True

    def initial_states(self, X0) -> None:
        self.X[...] = np.broadcast_to(_f64(X0), self.X.shape)
        self.dirty["traj"] = True

    def set_dynamics(self, A=None, B=None, d=None) -> None:
        if A is not None:
            self.model.A[...] = A
        if B is not None:
            self.model.B[...] = B
        if d is not None:
            self.model.d[...] = d
        self.dirty["dyn"] = True

    def set_constraint_data(self, con_id: int, G=None, h=None) -> None:
        c = self.constraints.flat[con_id]
        if G is not None:
            c.G[...] = G
        if h is not None:
            c.h[...] = h
        self.dirty["con"].add(con_id)


@dataclass
class SolverOptions:
    """Altro.SolverOptions (defaults per SURVEY.md Appendix A.1); field names as the reference sets them
    (run_random_linear.jl:41-49, run_simple_rocket.jl:121-129, ALTROParams.jl:86-95,
    grasp_benchmark.jl:26-34, flexible_sat_mpc.jl:250-257)."""

    constraint_tolerance: float = 1e-6
    cost_tolerance: float = 1e-4
    cost_tolerance_intermediate: float = 1e-4
    gradient_tolerance: float = 10.0
    gradient_tolerance_intermediate: float = 1.0
    penalty_initial: float = 1.0
    penalty_scaling: float = 10.0
    penalty_max: float = 1e8
    dual_max: float = 1e8
    line_search_lower_bound: float = 1e-8
    line_search_upper_bound: float = 10.0
    max_cost_value: float = 1e8
    max_state_value: float = 1e8
    bp_reg_initial: float = 0.0
    bp_reg_increase_factor: float = 1.6
    bp_reg_max: float = 1e8
    bp_reg_min: float = 1e-8
    bp_reg_fp: float = 10.0
    iterations: int = 1000
    iterations_inner: int = 300
    iterations_outer: int = 30
    iterations_linesearch: int = 20
    dJ_counter_limit: int = 10
    reset_duals: bool = True
    reset_penalties: bool = True
    kickout_max_penalty: bool = False
    # switches for recollection-uncertain details
    dj_zero_converges: bool = True
    soc_hess_exact: bool = True
    soc_viol_proj: bool = True
    # the initial rollout's cost is not a line-search reference: the first forward pass of every iLQR solve takes
    # its full step (J_prev = +inf).  Matches the saved grasp statistics of the reference
    first_step_unconditional: bool = True
    # accepted for API compatibility (static variant, logging); projected_newton = True is refused at upload:
    # the polish step is not built and every benchmark of the reference switches it off
    projected_newton: bool = False
    static_bp: bool = True
    verbose: int = 0
    show_summary: bool = False

    def copy(self) -> "SolverOptions":
        return SolverOptions(**self.__dict__)


# Altro.TerminationStatus
UNSOLVED, SOLVE_SUCCEEDED, MAX_ITERATIONS, MAX_ITERATIONS_OUTER, MAXIMUM_COST, STATE_LIMIT, CONTROL_LIMIT, \
    NO_PROGRESS, COST_INCREASE, NOT_PD = range(10)
STATUS_NAMES = ["UNSOLVED", "SOLVE_SUCCEEDED", "MAX_ITERATIONS", "MAX_ITERATIONS_OUTER", "MAXIMUM_COST",
                "STATE_LIMIT", "CONTROL_LIMIT", "NO_PROGRESS", "COST_INCREASE", "NOT_PD"]
