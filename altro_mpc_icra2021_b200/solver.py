"""ALTROSolver: the host-side mirror of Altro.ALTROSolver / solve! over the C ABI (include/altro_b200.h).

Reference surface mirrored (paths relative to /root/reference/benchmarks):
  ALTROSolver(prob, opts)          random_linear_mpc/random_linear_problem.jl:87, rocket_landing/simple_rocket.jl:128
  solve!(altro)                    random_linear_problem.jl:113, simple_rocket.jl:174, quadruped/.../altro_solver.jl:72
  benchmark_solve!(altro, ...)     random_linear_problem.jl:161
  set_options!(solver; ...)        flexible_satellite/flexible_sat_mpc.jl:163,250-257
  iterations / status / states / controls / cost / max_violation / stats.tsolve
                                   random_linear_problem.jl:166-178, simple_rocket.jl:178-198, altro_solver.jl:75-79
  Altro.shift_fill!(conSet) + RD.shift_fill!(Z)   random_linear_problem.jl:136,139, altro_solver.jl:65,68
The Problem object is shared by reference with the solver and may be mutated between solves; solve()
uploads whatever the mutators marked dirty, exactly like the Julia solver re-reads prob at every solve!.

There is no CPU fallback: importing works anywhere, but constructing a solver without the compiled
CUDA library or without a GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

from .problem import Problem, SolverOptions, STATUS_NAMES, SOLVE_SUCCEEDED

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ALTRO_B200_LIB") or os.path.join(_HERE, "libaltro_b200.so")  # (env override: development builds)

ABI_SYMBOLS = [
    "altro_default_options", "altro_create", "altro_destroy", "altro_last_error", "altro_set_stream",
    "altro_set_options", "altro_set_dynamics", "altro_set_dynamics_slots", "altro_set_cost_diag", "altro_set_reference", "altro_add_constraint", "altro_add_track_constraint", "altro_set_track_index",
    "altro_update_constraint_data", "altro_set_x0", "altro_set_trajectory", "altro_get_trajectory", "altro_dual_len",
    "altro_set_duals", "altro_get_duals", "altro_shift_fill", "altro_solve", "altro_sync", "altro_get_stats",
    "altro_get_timing", "altro_get_phase_cycles", "altro_set_trace", "altro_get_trace", "altro_snapshot", "altro_restore", "altro_set_track", "altro_mpc_transition", "altro_set_noise_bank", "altro_set_noise_model", "altro_get_x0", "altro_mpc_run",
    "altro_get_run_results",
    "altro_host_register", "altro_host_unregister", "altro_set_launch_config", "altro_get_launch_info",
    "altro_set_line_search_mode", "altro_get_line_search_mode", "altro_reserve_steps",
    "altro_set_kernel_mode", "altro_get_kernel_mode", "altro_set_run_queue",
    "altro_admm_solve", "altro_quadruped_linearize", "altro_quadruped_tick", "altro_quadruped_get_schedule", "altro_get_dynamics",
    "altro_measure_peaks",
]


class AltroOpts(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        "constraint_tolerance", "cost_tolerance", "cost_tolerance_intermediate", "gradient_tolerance",
        "gradient_tolerance_intermediate", "penalty_initial", "penalty_scaling", "penalty_max", "dual_max",
        "line_search_lower_bound", "line_search_upper_bound", "max_cost_value", "max_state_value",
        "bp_reg_initial", "bp_reg_increase_factor", "bp_reg_max", "bp_reg_min", "bp_reg_fp")] + [
        (k, C.c_int) for k in (
            "iterations", "iterations_inner", "iterations_outer", "iterations_linesearch", "dJ_counter_limit",
            "reset_duals", "reset_penalties", "kickout_max_penalty", "dj_zero_converges", "soc_hess_exact",
            "soc_viol_proj", "first_step_unconditional")]


class AltroError(RuntimeError):
    pass


_lib = None


def load_library() -> C.CDLL:
    """Loads libaltro_b200.so (built in-tree by __graft_entry__.build / csrc/Makefile). Fails loudly."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AltroError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                             f"g.build()'` (make -C altro_mpc_icra2021_b200/csrc). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        lib.altro_last_error.restype = C.c_char_p
        lib.altro_last_error.argtypes = [C.c_void_p]
        for name in ABI_SYMBOLS:
            fn = getattr(lib, name)
            if name != "altro_last_error":
                fn.restype = C.c_int
        lib.altro_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]
        _lib = lib
    return _lib


def _p(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"] and a.dtype in (np.float64, np.int32, np.int64), (a.dtype, a.flags)
    return a.ctypes.data_as(C.c_void_p)


@dataclass
class SolverStats:
    """solver.stats of the reference, per instance."""

    iterations: np.ndarray
    iterations_outer: np.ndarray
    status: np.ndarray
    ls_trials: np.ndarray
    cost: np.ndarray
    cost_al: np.ndarray
    c_max: np.ndarray
    penalty_max: np.ndarray
    tsolve: float  # ms, device time of the batched solve
    t_instance_us: np.ndarray  # per-instance solve time on the device, microseconds

    @property
    def status_names(self):
        return [STATUS_NAMES[s] for s in self.status]


class ALTROSolver:
    def __init__(self, prob: Problem, opts: Optional[SolverOptions] = None, device: int = 0,
                 threads_per_instance: int = 0, stream: Optional[int] = None, pin: bool = False,
                 speculative_line_search: Optional[bool] = None, kernel: str = "auto", **kwargs):
        self.lib = load_library()
        self.prob = prob
        self.opts = (opts or SolverOptions()).copy()
        for k, v in kwargs.items():  # ALTROSolver(prob, opts; kwargs...) overrides (run_simple_rocket.jl:66)
            setattr(self.opts, k, v)
        self.h = C.c_void_p()
        self._pinned = []
        rc = self.lib.altro_create(C.byref(self.h), device, prob.n, prob.m, prob.N, prob.B, prob.dt)
        if rc != 0:
            raise AltroError(f"altro_create failed ({rc}): {self.lib.altro_last_error(None).decode()}")
        if stream is not None:
            self._ck(self.lib.altro_set_stream(self.h, C.c_void_p(stream)))
        if threads_per_instance:
            self._ck(self.lib.altro_set_launch_config(self.h, threads_per_instance))
        if speculative_line_search is not None:
            self._ck(self.lib.altro_set_line_search_mode(self.h, int(bool(speculative_line_search))))
        if kernel != "auto":  # "cta": one CTA per instance, "lane": one thread per instance (small dimensions)
            self._ck(self.lib.altro_set_kernel_mode(self.h, {"cta": 1, "lane": 2}[kernel]))
        self._ck(self.lib.altro_set_cost_diag(self.h, _p(prob.obj.Q), _p(prob.obj.R), _p(prob.obj.Qf)))
        for c in prob.constraints.flat:
            cid = C.c_int()
            inds = np.ascontiguousarray(c.inds, dtype=np.int32)
            if c.track:
                self._ck(self.lib.altro_add_track_constraint(self.h, c.sense, c.side, c.k0, c.k1, c.p, c.w, _p(inds),
                                                             _p(c.G), _p(c.h), c.G.shape[0], C.byref(cid)))
            else:
                self._ck(self.lib.altro_add_constraint(self.h, c.sense, c.side, c.k0, c.k1, c.p, c.w, _p(inds),
                                                       int(c.per_knot), int(c.per_instance), _p(c.G), _p(c.h),
                                                       C.byref(cid)))
        prob.dirty["con"] = set()
        if any(c.track for c in prob.constraints.flat):
            self.set_track_index(prob.kidx)
        self.P = prob.constraints.dual_len()
        if pin:
            for a in (prob.x0, prob.Xref, prob.Uref, prob.X, prob.U, prob.model.A, prob.model.B, prob.model.d):
                if self.lib.altro_host_register(_p(a), C.c_size_t(a.nbytes)) == 0:
                    self._pinned.append(a)
        self._opts_dirty = True
        self.stats: Optional[SolverStats] = None
        self._results_stale = False

    # ------------------------------------------------------------------ plumbing
    def _ck(self, rc: int) -> None:
        if rc != 0:
            raise AltroError(f"altro error {rc}: {self.lib.altro_last_error(self.h).decode()}")

    def close(self) -> None:
        if getattr(self, "h", None) and self.h.value:
            for a in self._pinned + getattr(self, "_runbuf", []):
                self.lib.altro_host_unregister(_p(a))
            self._runbuf, self._runbuf_steps = [], 0
            self.lib.altro_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_options(self, **kwargs) -> None:
        """set_options!(solver; ...)."""
        for k, v in kwargs.items():
            if not hasattr(self.opts, k):
                raise AttributeError(k)
            setattr(self.opts, k, v)
        self._opts_dirty = True

    def _push_options(self) -> None:
        if self.opts.projected_newton:
            # every benchmark of the reference sets projected_newton = false; the polish step is not built and is not
            # silently skipped either
            raise AltroError("projected_newton = true is not supported by this solve path (DESIGN.md section 7)")
        o = AltroOpts()
        for name, _ in AltroOpts._fields_:
            setattr(o, name, getattr(self.opts, name))
        self._ck(self.lib.altro_set_options(self.h, C.byref(o)))
        self._opts_dirty = False

    def upload(self) -> None:
        """Pushes every dirty piece of the shared Problem to the device (async on the solver's stream)."""
        p, d = self.prob, self.prob.dirty
        if self._opts_dirty:
            self._push_options()
        if d["dyn"]:
            mdl = p.model
            if mdl.sched is not None:
                self._ck(self.lib.altro_set_dynamics_slots(self.h, mdl.A.shape[1], _p(mdl.A), _p(mdl.B), _p(mdl.d),
                                                           _p(mdl.sched), mdl.sched.shape[1]))
            else:
                self._ck(self.lib.altro_set_dynamics(self.h, int(mdl.per_knot), int(mdl.per_instance), _p(mdl.A),
                                                     _p(mdl.B), _p(mdl.d)))
            d["dyn"] = False
        if d["ref"]:
            self._ck(self.lib.altro_set_reference(self.h, _p(p.Xref), _p(p.Uref)))
            d["ref"] = False
        if d["x0"]:
            self._ck(self.lib.altro_set_x0(self.h, _p(p.x0)))
            d["x0"] = False
        if d["traj"]:
            self._ck(self.lib.altro_set_trajectory(self.h, _p(p.X), _p(p.U)))
            d["traj"] = False
        for cid in sorted(d["con"]):
            c = p.constraints.flat[cid]
            self._ck(self.lib.altro_update_constraint_data(self.h, cid, _p(c.G), _p(c.h)))
        d["con"] = set()

    # ------------------------------------------------------------------ solve!
    def solve(self, fetch: bool = True) -> "ALTROSolver":
        """solve!(solver): upload dirty problem data, run the batched AL-iLQR solve, and (fetch=True) bring
        the trajectories and statistics back into prob.X / prob.U / self.stats."""
        self.upload()
        self._ck(self.lib.altro_solve(self.h))
        self._results_stale = True
        if fetch:
            self.fetch()
        return self

    def fetch(self) -> SolverStats:
        p, B = self.prob, self.prob.B
        self._ck(self.lib.altro_get_trajectory(self.h, _p(p.X), _p(p.U)))
        it, ito, st, ls = (np.zeros(B, np.int32) for _ in range(4))
        cost, cal, cmax, pmax = (np.zeros(B) for _ in range(4))
        self._ck(self.lib.altro_get_stats(self.h, _p(it), _p(ito), _p(st), _p(ls), _p(cost), _p(cal), _p(cmax),
                                          _p(pmax)))
        ms = C.c_double()
        tns = np.zeros(B, np.int64)
        self._ck(self.lib.altro_get_timing(self.h, C.byref(ms), _p(tns)))
        self.stats = SolverStats(it, ito, st, ls, cost, cal, cmax, pmax, ms.value, tns / 1e3)
        self._results_stale = False
        return self.stats

    def snapshot(self) -> None:
        self.upload()
        self._ck(self.lib.altro_snapshot(self.h))
        self._kidx_snap = self.prob.kidx.copy()

    def restore(self) -> None:
        self._ck(self.lib.altro_restore(self.h))
        self.prob.kidx[...] = self._kidx_snap

    def sync(self) -> None:
        self._ck(self.lib.altro_sync(self.h))

    def device_ms(self) -> float:
        ms = C.c_double()
        self._ck(self.lib.altro_get_timing(self.h, C.byref(ms), None))
        return ms.value

    def benchmark_solve(self, samples: int = 10, evals: int = 10):
        """benchmark_solve!(solver; samples, evals): snapshot the warm start, then repeat {restore; solve!}.
        Duals are restored together with the primal trajectory.  Returns device times in ms."""
        self.upload()
        self._ck(self.lib.altro_snapshot(self.h))
        times = []
        for _ in range(samples * evals):
            self._ck(self.lib.altro_restore(self.h))
            self._ck(self.lib.altro_solve(self.h))
            times.append(self.device_ms())
        self.fetch()
        return np.array(times)

    # ------------------------------------------------------------------ warm-start shifts
    def shift_fill(self, primal: bool = True, dual: bool = True) -> None:
        """RD.shift_fill!(prob.Z) and Altro.shift_fill!(get_constraints(solver)), on the device."""
        self.upload()
        self._ck(self.lib.altro_shift_fill(self.h, int(primal), int(dual)))
        if primal:  # keep the host mirror of the warm start consistent
            p = self.prob
            p.X[:, :-1] = p.X[:, 1:].copy()
            if p.N > 2:
                p.U[:, :-1] = p.U[:, 1:].copy()

    def set_track_index(self, kidx) -> None:
        """Position of every instance on the shared timelines (track constraints, reference track)."""
        ki = np.ascontiguousarray(kidx, dtype=np.int32)
        self._ck(self.lib.altro_set_track_index(self.h, _p(ki)))

    def set_track(self, X_track, U_track, k_start) -> None:
        Xt = np.ascontiguousarray(X_track, dtype=np.float64)
        Ut = np.ascontiguousarray(U_track, dtype=np.float64)
        ks = np.ascontiguousarray(k_start, dtype=np.int32)
        self._ck(self.lib.altro_set_track(self.h, _p(Xt), _p(Ut), Xt.shape[0], _p(ks)))

    def set_noise_model(self, mode: int, w1: float = 1.0, w2: float = 1.0) -> None:
        """0 additive, 1 random-linear (|x|_inf relative), 2 rocket (position / velocity 2-norm relative)."""
        self._ck(self.lib.altro_set_noise_model(self.h, int(mode), C.c_double(w1), C.c_double(w2)))

    def get_x0(self) -> np.ndarray:
        x0 = np.zeros((self.prob.B, self.prob.n))
        self._ck(self.lib.altro_get_x0(self.h, _p(x0)))
        return x0

    def set_noise_bank(self, noise: Optional[np.ndarray]) -> None:
        """Device-resident process noise [steps][B][n], consumed in turn by mpc_transition(noise=None)."""
        if noise is None:
            self._ck(self.lib.altro_set_noise_bank(self.h, None, 0))
            return
        nz = np.ascontiguousarray(noise, dtype=np.float64)
        assert nz.shape[1:] == (self.prob.B, self.prob.n)
        self._ck(self.lib.altro_set_noise_bank(self.h, _p(nz), nz.shape[0]))

    def mpc_transition(self, noise: Optional[np.ndarray] = None, shift: bool = True) -> None:
        """Device-side MPC step: plant step with the first control (+ noise), reference window advance
        along the registered track, primal + dual shift.  No host traffic except the optional noise."""
        self.upload()
        nz = None if noise is None else np.ascontiguousarray(noise, dtype=np.float64)
        self._ck(self.lib.altro_mpc_transition(self.h, _p(nz), int(shift)))
        self.prob.kidx += 1

    def mpc_run(self, steps: int, shift: bool = True, fetch: bool = True, reuse_buffers: bool = False):
        """Closed-loop MPC run on the device: `steps` x {transition; solve!} per instance in one launch.
        Returns a dict of per-step results (see altro_mpc_run) when fetch=True."""
        self.upload()
        self._ck(self.lib.altro_mpc_run(self.h, int(steps), int(shift)))
        self.prob.kidx += int(steps)  # host mirror of the timeline positions the device just advanced
        self._results_stale = True
        return self.run_results(steps, reuse_buffers) if fetch else None

    def _run_buffers(self, steps: int, reuse: bool):
        """Host arrays the run results are copied into.  reuse=True hands out views of one page-locked set kept by
        the solver (valid until the next call with reuse=True): what a caller that reads results every control tick
        does -- fresh pageable arrays cost page faults and a staged copy on every call."""
        p, B = self.prob, self.prob.B
        shapes = [((B,), np.int32)] * 4 + [((B,), np.float64)] * 2 + [((B, p.n), np.float64), ((B, p.m), np.float64),
                                                                      ((B,), np.int64)]
        if not reuse:
            return [np.zeros((steps,) + sh, dt) for sh, dt in shapes]
        if getattr(self, "_runbuf_steps", 0) < steps:
            for a in getattr(self, "_runbuf", []):
                self.lib.altro_host_unregister(_p(a))
            self._runbuf = [np.zeros((steps,) + sh, dt) for sh, dt in shapes]
            for a in self._runbuf:
                self.lib.altro_host_register(_p(a), C.c_size_t(a.nbytes))  # best effort: pageable still works
            self._runbuf_steps = steps
        return [a[:steps] for a in self._runbuf]

    def reserve_host_results(self, steps: int) -> None:
        """Allocates and page-locks the reusable result buffers of run_results(..., reuse_buffers=True) ahead of time."""
        self._run_buffers(steps, True)

    def run_results(self, steps: int, reuse_buffers: bool = False) -> dict:
        p = self.prob
        it, ito, st, ls, cost, cmax, x0l, u0l, tns = self._run_buffers(steps, reuse_buffers)
        self._ck(self.lib.altro_get_run_results(self.h, steps, _p(it), _p(ito), _p(st), _p(ls), _p(cost), _p(cmax),
                                                _p(x0l), _p(u0l), _p(tns)))
        self._ck(self.lib.altro_get_trajectory(self.h, _p(p.X), _p(p.U)))
        p.x0[...] = x0l[-1]
        return {"iterations": it, "iterations_outer": ito, "status": st, "ls_trials": ls, "cost": cost, "c_max": cmax,
                "x0": x0l, "u0": u0l, "t_us": tns / 1e3, "device_ms": self.device_ms()}

    # ------------------------------------------------------------------ queries
    def states(self) -> np.ndarray:
        return self.prob.X

    def controls(self) -> np.ndarray:
        return self.prob.U

    def iterations(self) -> np.ndarray:
        return self.stats.iterations

    def status(self) -> np.ndarray:
        return self.stats.status

    def cost(self) -> np.ndarray:
        return self.stats.cost

    def max_violation(self) -> np.ndarray:
        return self.stats.c_max

    def get_duals(self) -> np.ndarray:
        lam = np.zeros((self.prob.B, max(self.P, 1)))
        if self.P:
            self._ck(self.lib.altro_get_duals(self.h, _p(lam)))
        return lam[:, :self.P]

    def set_duals(self, lam: np.ndarray) -> None:
        if self.P:
            self._ck(self.lib.altro_set_duals(self.h, _p(np.ascontiguousarray(lam, dtype=np.float64))))

    def set_trace(self, max_rows: int) -> None:
        """verbose mode: keep a per-iteration log of every instance (see altro_set_trace)."""
        self._ck(self.lib.altro_set_trace(self.h, int(max_rows)))
        self._trace_rows = int(max_rows)

    def get_trace(self) -> np.ndarray:
        out = np.zeros((self.prob.B, self._trace_rows, 10))
        self._ck(self.lib.altro_get_trace(self.h, _p(out)))
        return out

    def phase_cycles(self, enable: bool = True) -> np.ndarray:
        """Per-instance cycle counters per solver phase since the last call (see altro_get_phase_cycles)."""
        out = np.zeros((self.prob.B, 8), np.int64)
        self._ck(self.lib.altro_get_phase_cycles(self.h, int(enable), _p(out)))
        return out

    # ------------------------------------------------------------------ independent convex cross-check
    def admm_solve(self, rho: float = 10.0, eps: float = 1e-6, max_iter: int = 4000) -> dict:
        """Solves the current problems with the batched operator-splitting solver (csrc/admm.cu), the stand-in for the
        reference's OSQP / ECOS cross-check.  Does not touch the solver's own trajectories."""
        p, B = self.prob, self.prob.B
        self.upload()
        X, U = np.zeros((B, p.N, p.n)), np.zeros((B, p.N - 1, p.m))
        it, rp, rd = np.zeros(B, np.int32), np.zeros(B), np.zeros(B)
        self._ck(self.lib.altro_admm_solve(self.h, C.c_double(rho), C.c_double(eps), int(max_iter), _p(X), _p(U), _p(it),
                                           _p(rp), _p(rd)))
        return {"X": X, "U": U, "iterations": it, "r_prim": rp, "r_dual": rd}

    # ------------------------------------------------------------------ quadruped pre-solve kernels
    def quadruped_linearize(self, x_ref, foot, contacts, J, mass, u_ref=None) -> None:
        """update_dynamics_matrices! (altro_solver.jl:5-42) on the device: writes A_k, B_k, d_k of every instance and
        knot into the solver's model.  x_ref (B,12) or (B,N-1,12); foot (B,N-1,4,3) world; contacts (B,N-1,4)."""
        x = np.ascontiguousarray(x_ref, dtype=np.float64)
        u = None if u_ref is None else np.ascontiguousarray(u_ref, dtype=np.float64)
        f, c = np.ascontiguousarray(foot, dtype=np.float64), np.ascontiguousarray(contacts, dtype=np.float64)
        Jm = np.ascontiguousarray(J, dtype=np.float64)
        self.upload()
        self._ck(self.lib.altro_quadruped_linearize(self.h, _p(x), int(x.ndim == 3), _p(u), int(u is not None and u.ndim == 3),
                                                    _p(f), _p(c), _p(Jm), C.c_double(mass)))
        self.prob.dirty["dyn"] = False
        self._ck(self.lib.altro_sync(self.h))

    def quadruped_tick(self, t, x_ref, cur_foot, contact_phases, phase_times, nom_foot, J, mass, alpha=0.5,
                       foot_radius=0.02) -> None:
        """foot_history! + update_dynamics_matrices! of one control tick on the device (footsteps.jl:29-84,
        altro_solver.jl:5-42).  t (B,), x_ref (B,12) or (B,N-1,12), cur_foot (B,4,3) body frame,
        contact_phases (num_phases,4), phase_times (num_phases,)."""
        tt = np.ascontiguousarray(t, dtype=np.float64)
        x = np.ascontiguousarray(x_ref, dtype=np.float64)
        cf = np.ascontiguousarray(cur_foot, dtype=np.float64)
        cp = np.ascontiguousarray(contact_phases, dtype=np.float64)
        pt = np.ascontiguousarray(phase_times, dtype=np.float64)
        nf, Jm = np.ascontiguousarray(nom_foot, dtype=np.float64), np.ascontiguousarray(J, dtype=np.float64)
        self.upload()
        self._ck(self.lib.altro_quadruped_tick(self.h, _p(tt), _p(x), int(x.ndim == 3), _p(cf), int(cp.shape[0]), _p(cp),
                                               _p(pt), C.c_double(alpha), C.c_double(foot_radius), _p(nf), _p(Jm),
                                               C.c_double(mass)))
        self.prob.dirty["dyn"] = False
        self._ck(self.lib.altro_sync(self.h))

    def quadruped_schedule(self):
        B, K = self.prob.B, self.prob.N - 1
        c, f = np.zeros((B, K, 4)), np.zeros((B, K, 4, 3))
        self._ck(self.lib.altro_quadruped_get_schedule(self.h, _p(c), _p(f)))
        return c, f

    def get_dynamics(self):
        """The model as it is on the device (A, B, d in the layout of prob.model)."""
        mdl = self.prob.model
        A, Bm, d = np.zeros_like(mdl.A), np.zeros_like(mdl.B), np.zeros_like(mdl.d)
        self._ck(self.lib.altro_get_dynamics(self.h, _p(A), _p(Bm), _p(d)))
        return A, Bm, d

    def set_run_queue(self, steps_per_item: int) -> None:
        """Scheduling of mpc_run: >= 1 = persistent grid + work queue of (instance, steps_per_item steps) items
        (default 1), 0 = one CTA per instance for the whole run."""
        self._ck(self.lib.altro_set_run_queue(self.h, int(steps_per_item)))

    def reserve_steps(self, steps: int):
        """Pre-sizes the per-step statistics / log buffers (otherwise the first longer mpc_run reallocates them)."""
        self.upload()
        self._ck(self.lib.altro_reserve_steps(self.h, int(steps)))

    def launch_info(self) -> dict:
        self.upload()
        v = [C.c_int() for _ in range(5)]
        self._ck(self.lib.altro_get_launch_info(self.h, *[C.byref(x) for x in v]))
        keys = ("threads_per_instance", "smem_bytes", "regs_per_thread", "ctas_per_sm", "num_sms")
        info = {k: x.value for k, x in zip(keys, v)}
        spec = C.c_int()
        self._ck(self.lib.altro_get_line_search_mode(self.h, C.byref(spec)))
        info["speculative_line_search"] = bool(spec.value)
        km, lr, lsm, lpw = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._ck(self.lib.altro_get_kernel_mode(self.h, C.byref(km), C.byref(lr), C.byref(lsm), C.byref(lpw)))
        info["kernel"] = "lane" if km.value == 2 else "cta"
        if km.value == 2:
            info.update(lane_regs_per_thread=lr.value, lane_smem_bytes=lsm.value, instances_per_warp=lpw.value)
        return info

    def all_succeeded(self) -> bool:
        return bool(np.all(self.stats.status == SOLVE_SUCCEEDED))


def solve(solver: ALTROSolver) -> ALTROSolver:
    """solve!(solver) spelled as a function, like the reference's call sites."""
    return solver.solve()


def measure_peaks(device: int = 0) -> dict:
    lib = load_library()
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    rc = lib.altro_measure_peaks(device, C.byref(a), C.byref(b), C.byref(c))
    if rc != 0:
        raise AltroError(f"altro_measure_peaks failed: {lib.altro_last_error(None).decode()}")
    return {"dfma_tflops": a.value, "dmma_tflops": b.value, "copy_gbs": c.value}
