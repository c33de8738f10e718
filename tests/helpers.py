"""Test helpers: an ALTROSolver-shaped adapter around the CPU oracle, and small problem builders."""
import copy

import numpy as np

from altro_mpc_icra2021_b200.problem import (BoundConstraint, ConstraintList, LinearModel, LQRObjective, Problem,
                                             SolverOptions)
from oracle.oracle import OracleProblem


class OracleSolver:
    """Same surface as altro_mpc_icra2021_b200.solver.ALTROSolver, computed by the oracle (tests only)."""

    def __init__(self, prob, opts, nthreads=4):
        self.prob, self.opts, self.nthreads = prob, opts, nthreads
        self.op = OracleProblem(prob)
        self.stats = None

    def solve(self):
        self.stats = self.op.solve(self.opts, nthreads=self.nthreads)
        return self

    def shift_fill(self, primal=True, dual=True):
        self.op.shift_fill(primal, dual)

    def get_duals(self):
        return self.op.lam.copy()

    def set_track_index(self, kidx):
        pass  # the oracle reads prob.kidx in place


def random_lti(n, m, rng, stable=0.95):
    A = rng.standard_normal((n, n))
    A *= stable / max(abs(np.linalg.eigvals(A)))
    return A, rng.standard_normal((n, m))


def lqr_problem(n=4, m=2, N=15, batch=3, seed=0, u_bnd=None, dt=0.1):
    rng = np.random.default_rng(seed)
    A, B = random_lti(n, m, rng)
    model = LinearModel(A, B, d=0.01 * rng.standard_normal(n), dt=dt)
    obj = LQRObjective(1.0 + rng.random(n), 0.1 + rng.random(m), 10.0 * (1.0 + rng.random(n)), np.zeros(n), N)
    cons = ConstraintList(n, m, N)
    if u_bnd is not None:
        cons.add_constraint(BoundConstraint(n, m, u_min=-u_bnd, u_max=u_bnd), (0, N - 1))
    x0 = rng.standard_normal((batch, n))
    return Problem(model, obj, N, x0=x0, constraints=cons, batch=batch)


def assert_bit_identical(gpu_prob, gpu_stats, gpu_duals, ref, what=""):
    """The parity bar of this repo: bit-identical FP64 trajectories, duals, costs and integer statistics."""
    assert np.array_equal(gpu_prob.X, ref.X), f"{what}: X differs (max {np.abs(gpu_prob.X - ref.X).max():.3e})"
    assert np.array_equal(gpu_prob.U, ref.U), f"{what}: U differs (max {np.abs(gpu_prob.U - ref.U).max():.3e})"
    if ref.lam.size:
        assert np.array_equal(gpu_duals, ref.lam), f"{what}: duals differ"
    assert np.array_equal(gpu_stats.iterations, ref.iterations), f"{what}: iteration counts differ"
    assert np.array_equal(gpu_stats.iterations_outer, ref.iterations_outer), f"{what}: outer iterations differ"
    assert np.array_equal(gpu_stats.status, ref.status), f"{what}: status differs"
    assert np.array_equal(gpu_stats.ls_trials, ref.ls_trials), f"{what}: line-search trials differ"
    assert np.array_equal(gpu_stats.cost, ref.cost), f"{what}: cost differs"
    assert np.array_equal(gpu_stats.cost_al, ref.cost_al), f"{what}: AL cost differs"
    assert np.array_equal(gpu_stats.c_max, ref.c_max), f"{what}: c_max differs"
