"""ctypes front-end of the CPU oracle (oracle/altro_oracle.c).

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (see oracle/altro_oracle.h).  May be imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libaltro_oracle.so")
ORC_MAX_W = 64


class _Con(C.Structure):
    _fields_ = [("sense", C.c_int), ("side", C.c_int), ("k0", C.c_int), ("k1", C.c_int), ("p", C.c_int),
                ("w", C.c_int), ("inds", C.c_int * ORC_MAX_W), ("per_knot", C.c_int), ("per_instance", C.c_int),
                ("track", C.c_int), ("G", C.c_void_p), ("h", C.c_void_p)]


class _Problem(C.Structure):
    _fields_ = [("n", C.c_int), ("m", C.c_int), ("N", C.c_int), ("B", C.c_int), ("dt", C.c_double),
                ("dyn_per_knot", C.c_int), ("dyn_per_instance", C.c_int),
                ("A", C.c_void_p), ("Bm", C.c_void_p), ("d", C.c_void_p),
                ("dyn_slots", C.c_int), ("sched_len", C.c_int), ("step0", C.c_int), ("sched", C.c_void_p),
                ("kidx", C.c_void_p),
                ("Q", C.c_void_p), ("R", C.c_void_p), ("Qf", C.c_void_p),
                ("xref", C.c_void_p), ("uref", C.c_void_p), ("x0", C.c_void_p),
                ("ncon", C.c_int), ("con", C.POINTER(_Con))]


class _Opts(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        "constraint_tolerance", "cost_tolerance", "cost_tolerance_intermediate", "gradient_tolerance",
        "gradient_tolerance_intermediate", "penalty_initial", "penalty_scaling", "penalty_max", "dual_max",
        "line_search_lower_bound", "line_search_upper_bound", "max_cost_value", "max_state_value",
        "bp_reg_initial", "bp_reg_increase_factor", "bp_reg_max", "bp_reg_min", "bp_reg_fp")] + [
        (k, C.c_int) for k in (
            "iterations", "iterations_inner", "iterations_outer", "iterations_linesearch", "dJ_counter_limit",
            "reset_duals", "reset_penalties", "kickout_max_penalty", "dj_zero_converges", "soc_hess_exact",
            "soc_viol_proj", "first_step_unconditional")]


class _Run(C.Structure):
    _fields_ = [("steps", C.c_int), ("shift", C.c_int), ("noise_mode", C.c_int), ("Nt", C.c_int),
                ("w1", C.c_double), ("w2", C.c_double), ("noise", C.c_void_p), ("trackX", C.c_void_p),
                ("trackU", C.c_void_p), ("kidx", C.c_void_p)]


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("altro_oracle.c", "altro_oracle.h", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_dual_len.restype = C.c_int
        _lib.orc_solve_batch.restype = C.c_int
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _opts_struct(opts) -> _Opts:
    o = _Opts()
    for name, _ in _Opts._fields_:
        setattr(o, name, getattr(opts, name))
    return o


@dataclass
class OracleResult:
    X: np.ndarray
    U: np.ndarray
    lam: np.ndarray
    iterations: np.ndarray
    iterations_outer: np.ndarray
    status: np.ndarray
    ls_trials: np.ndarray
    cost: np.ndarray
    cost_al: np.ndarray
    c_max: np.ndarray
    penalty_max: np.ndarray


class OracleProblem:
    """Flattens an altro_mpc_icra2021_b200.problem.Problem into the oracle's C structs (keeps the arrays alive)."""

    def __init__(self, prob):
        self.prob = prob
        self._keep = []
        cons = prob.constraints.flat
        self.cons = (_Con * max(1, len(cons)))()
        for i, c in enumerate(cons):
            s = self.cons[i]
            s.sense, s.side, s.k0, s.k1, s.p, s.w = c.sense, c.side, c.k0, c.k1, c.p, c.w
            assert c.w <= ORC_MAX_W and c.p <= ORC_MAX_W
            for j, v in enumerate(c.inds):
                s.inds[j] = int(v)
            s.per_knot, s.per_instance = int(c.per_knot), int(c.per_instance)
            s.track = c.G.shape[0] if getattr(c, "track", False) else 0
            s.G, s.h = _ptr(c.G), _ptr(c.h)  # live views: in-place constraint-data updates are seen
        mdl = prob.model
        p = _Problem()
        p.n, p.m, p.N, p.B, p.dt = prob.n, prob.m, prob.N, prob.B, prob.dt
        p.dyn_per_knot, p.dyn_per_instance = int(mdl.per_knot), int(mdl.per_instance)
        p.A, p.Bm, p.d = _ptr(mdl.A), _ptr(mdl.B), _ptr(mdl.d)
        if getattr(mdl, "sched", None) is not None:
            p.dyn_slots, p.sched_len, p.step0, p.sched = mdl.A.shape[1], mdl.sched.shape[1], 0, _ptr(mdl.sched)
        p.Q, p.R, p.Qf = _ptr(prob.obj.Q), _ptr(prob.obj.R), _ptr(prob.obj.Qf)
        p.xref, p.uref, p.x0 = _ptr(prob.Xref), _ptr(prob.Uref), _ptr(prob.x0)
        p.kidx = _ptr(prob.kidx)  # live view: the host loop advances prob.kidx in place
        p.ncon = len(cons)
        p.con = C.cast(self.cons, C.POINTER(_Con))
        self.c = p
        self.P = lib().orc_dual_len(C.byref(p))
        self.lam = np.zeros((prob.B, self.P))

    def solve(self, opts, nthreads: int = 1, i0: int = 0, i1: int | None = None) -> OracleResult:
        """Solves in place on prob.X / prob.U (warm start -> solution) and self.lam."""
        pr = self.prob
        B = pr.B
        i1 = B if i1 is None else i1
        it, ito, st, ls = (np.zeros(B, np.int32) for _ in range(4))
        cost, cal, cmax, pmax = (np.zeros(B) for _ in range(4))
        o = _opts_struct(opts)
        self.c.step0 = getattr(self, "step_abs", 0)
        rc = lib().orc_solve_batch(C.byref(self.c), C.byref(o), i0, i1, nthreads, _ptr(pr.X), _ptr(pr.U),
                                   _ptr(self.lam), _ptr(it), _ptr(ito), _ptr(st), _ptr(ls), _ptr(cost),
                                   _ptr(cal), _ptr(cmax), _ptr(pmax))
        if rc != 0:
            raise RuntimeError(f"orc_solve_batch failed: {rc}")
        return OracleResult(pr.X.copy(), pr.U.copy(), self.lam.copy(), it, ito, st, ls, cost, cal, cmax, pmax)

    def mpc_run(self, opts, steps, noise=None, noise_model=(0, 1.0, 1.0), track=None, kidx=None, shift=True,
                nthreads: int = 1) -> dict:
        """Closed-loop run mirroring the product's altro_mpc_run; per-step results [steps][B]."""
        pr, B = self.prob, self.prob.B
        run = _Run()
        run.steps, run.shift, run.noise_mode = steps, int(shift), int(noise_model[0])
        run.w1, run.w2 = float(noise_model[1]), float(noise_model[2])
        keep = []
        if noise is not None:
            nz = np.ascontiguousarray(noise, dtype=np.float64)
            assert nz.shape == (steps, B, pr.n)
            keep.append(nz)
            run.noise = _ptr(nz)
        ki = np.ascontiguousarray(pr.kidx if kidx is None else kidx, dtype=np.int32).copy()
        keep.append(ki)
        run.kidx = _ptr(ki)
        if track is not None:
            Xt = np.ascontiguousarray(track[0], dtype=np.float64)
            Ut = np.ascontiguousarray(track[1], dtype=np.float64)
            keep += [Xt, Ut]
            run.trackX, run.trackU, run.Nt = _ptr(Xt), _ptr(Ut), Xt.shape[0]
        it, ito, st, ls = (np.zeros((steps, B), np.int32) for _ in range(4))
        cost, cmax = np.zeros((steps, B)), np.zeros((steps, B))
        x0l, u0l = np.zeros((steps, B, pr.n)), np.zeros((steps, B, pr.m))
        o = _opts_struct(opts)
        self.c.step0 = getattr(self, "step_abs", 0)
        self.step_abs = self.c.step0 + steps
        rc = lib().orc_mpc_run(C.byref(self.c), C.byref(o), C.byref(run), nthreads, _ptr(pr.X), _ptr(pr.U),
                               _ptr(self.lam), _ptr(it), _ptr(ito), _ptr(st), _ptr(ls), _ptr(cost), _ptr(cmax),
                               _ptr(x0l), _ptr(u0l))
        if rc != 0:
            raise RuntimeError(f"orc_mpc_run failed: {rc}")
        pr.kidx[...] = ki + steps
        return {"iterations": it, "iterations_outer": ito, "status": st, "ls_trials": ls, "cost": cost, "c_max": cmax,
                "x0": x0l, "u0": u0l}

    def set_trace(self, rows: int):
        self.trace = np.zeros((self.prob.B, rows, 10)) if rows else None
        lib().orc_set_trace(_ptr(self.trace) if rows else None, rows)
        return self.trace

    def shift_fill(self, primal=True, dual=True):
        pr = self.prob
        lib().orc_shift_fill(C.byref(self.c), int(primal), int(dual), _ptr(pr.X), _ptr(pr.U), _ptr(self.lam))

    def evaluate(self, opts, X=None, U=None):
        pr = self.prob
        X = pr.X if X is None else np.ascontiguousarray(X, dtype=np.float64)
        U = pr.U if U is None else np.ascontiguousarray(U, dtype=np.float64)
        cost, cmax = np.zeros(pr.B), np.zeros(pr.B)
        o = _opts_struct(opts)
        lib().orc_evaluate(C.byref(self.c), C.byref(o), _ptr(X), _ptr(U), _ptr(cost), _ptr(cmax))
        return cost, cmax


def soc_project(v):
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.zeros_like(v)
    lib().orc_soc_project(v.size, _ptr(v), _ptr(out))
    return out


def soc_project_jac(v):
    v = np.ascontiguousarray(v, dtype=np.float64)
    J = np.zeros((v.size, v.size))
    lib().orc_soc_project_jac(v.size, _ptr(v), _ptr(J))
    return J
