"""MPC helpers: the host side of the reference's warm-started MPC shift loops.

gen_tracking_problem mirrors benchmarks/mpc.jl:11-47; MPCLoop mirrors the per-step procedure of
run_MPC (random_linear_problem.jl:121-139), mpc_update (rocket_landing/simple_rocket.jl:59-82) and
run_flexsat_mpc (flexible_sat_mpc.jl:264-272): advance the plant with the first control, add noise,
set_initial_state!, update_trajectory!, RD.shift_fill!, Altro.shift_fill!, then solve!.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

from ..problem import (ConstraintList, GoalConstraint, LinearModel, Problem, TrackingObjective)


def rng_for(seed: int, stream: int = 0) -> np.random.Generator:
    """Counter-based, platform-independent stream (seed = 0xA1720 + config id)."""
    return np.random.Generator(np.random.Philox(key=[int(seed), int(stream)]))


def gen_tracking_problem(prob: Problem, X_track: np.ndarray, U_track: np.ndarray, N: int, Qk: float = 10.0,
                         Rk: float = 0.1, Qfk: Optional[float] = None, batch: int = 1,
                         k_start: Optional[np.ndarray] = None) -> Problem:
    """mpc.jl:11-47.  Tracks N knots of (X_track, U_track) (shapes (Nl,n), (Nl-1,m)) starting at
    k_start[i] for instance i; same constraints minus GoalConstraint, ranges re-indexed to N."""
    n, m, Nl = prob.n, prob.m, prob.N
    Qfk = Qk if Qfk is None else Qfk
    k_start = np.zeros(batch, dtype=np.int64) if k_start is None else np.asarray(k_start, dtype=np.int64)
    Xref, Uref = window_reference(X_track, U_track, k_start, N)
    obj = TrackingObjective(np.full(n, Qk), np.full(m, Rk), Xref, Uref, Qf=np.full(n, Qfk))
    cons = ConstraintList(n, m, N)
    for con, (k0, k1), name in prob.constraints.source:
        if isinstance(con, GoalConstraint):
            continue
        if k1 > N:  # inds.start : N - (prob.N - inds.stop), mpc.jl:35-37
            k1 = N - (Nl - k1)
        if k1 > k0:
            cons.add_constraint(con, (k0, k1), name)
    mdl = prob.model
    model = LinearModel(mdl.A, mdl.B, mdl.d, dt=mdl.dt, per_instance=mdl.per_instance)
    p = Problem(model, obj, N, x0=Xref[:, 0, :], constraints=cons, batch=batch, X0=Xref, U0=Uref)
    p.kidx[...] = k_start  # position on the shared timelines (track constraints follow it)
    return p


def window_reference(X_track, U_track, k_start, N):
    """Z_track[k : k+N] per instance; indices past the end of the track repeat its last knot."""
    Nl = X_track.shape[0]
    kx = np.minimum(k_start[:, None] + np.arange(N)[None, :], Nl - 1)
    ku = np.minimum(k_start[:, None] + np.arange(N - 1)[None, :], Nl - 2)
    return np.ascontiguousarray(X_track[kx]), np.ascontiguousarray(U_track[ku])


class MPCLoop:
    """Warm-started MPC shift loop over a batch.  `solver` is anything with the ALTROSolver surface
    (solve(), shift_fill(), prob) -- the product solver or, in tests, the oracle adapter."""

    def __init__(self, solver, X_track=None, U_track=None, k_start=None,
                 noise: Optional[Callable[[np.ndarray, np.random.Generator], np.ndarray]] = None,
                 shift: bool = True, seed: int = 0xA1720):
        self.solver, self.prob = solver, solver.prob
        self.X_track, self.U_track = X_track, U_track
        self.k = np.zeros(self.prob.B, dtype=np.int64) if k_start is None else np.asarray(k_start, np.int64).copy()
        self.noise, self.shift = noise, shift
        self.rng = rng_for(seed, 1)

    def plant_step(self) -> np.ndarray:
        """x0+ = A_0 x0 + B_0 u_0 + d_0 with the first control of the last solution (+ noise)."""
        p, mdl = self.prob, self.prob.model
        A, Bm, d = mdl.A, mdl.B, mdl.d
        if mdl.per_knot:
            A, Bm, d = A[..., 0, :, :], Bm[..., 0, :, :], d[..., 0, :]
        x, u = p.X[:, 0, :], p.U[:, 0, :]
        xn = np.einsum("...ij,bj->bi", A, x) + np.einsum("...ij,bj->bi", Bm, u) + d
        if self.noise is not None:
            xn = xn + self.noise(xn, self.rng)
        return xn

    def advance(self) -> None:
        """Everything the reference does between two solve! calls."""
        p = self.prob
        x0 = self.plant_step()
        self.k += 1
        p.kidx += 1
        if hasattr(self.solver, "set_track_index"):
            self.solver.set_track_index(p.kidx)
        p.set_initial_state(x0)
        if self.X_track is not None:
            p.update_trajectory(*window_reference(self.X_track, self.U_track, self.k, p.N))
        if self.shift:
            self.solver.shift_fill(primal=True, dual=True)

    def step(self):
        self.advance()
        return self.solver.solve()
