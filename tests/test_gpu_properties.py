"""Full-size checks at BASELINE.json's batch sizes through size-independent properties: determinism,
independence of an instance's result from the batch it is solved in (the sharding contract), feasibility
and the reference's own quadruped feasibility tolerances (mujoco_test.jl:185-206)."""
import copy
import os

import numpy as np
import pytest

from altro_mpc_icra2021_b200.problems import quadruped
from tests.golden import cases
from tests.helpers import OracleSolver

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gpu_solver(prob, opts, **kw):
    from altro_mpc_icra2021_b200.solver import ALTROSolver

    return ALTROSolver(prob, opts, **kw)


@pytest.mark.parametrize("linearized", [True, False])
def test_quadruped_4096_properties(linearized):
    B = 4096
    prob, st = quadruped.mpc_problem(B, linearized_friction=linearized)
    opts = quadruped.mpc_options()
    full = copy.deepcopy(prob)
    g = gpu_solver(full, opts).solve()
    s = g.stats
    assert np.all(s.status == 1) and np.all(s.c_max < opts.constraint_tolerance)
    f = full.U.reshape(B, prob.N - 1, 4, 3)
    assert np.all(f[..., 2] >= -1e-4) and np.all(f[..., 2] <= 133 + 1e-4)
    if linearized:
        assert np.all(np.abs(f[..., :2]) <= 0.5 * f[..., 2:3] + 1e-4)
    else:
        assert np.all(np.linalg.norm(f[..., :2], axis=-1) <= 0.5 * f[..., 2] + 1e-4)
    # swing feet carry no force cost gradient other than R: their forces stay at the reference (0)
    # determinism: a second solver on the same data gives the same bits
    again = copy.deepcopy(prob)
    g2 = gpu_solver(again, opts).solve()
    assert np.array_equal(again.X, full.X) and np.array_equal(again.U, full.U)
    assert np.array_equal(g2.stats.iterations, s.iterations)
    # shard independence: instances [1000,1300) solved alone, and spot instances against the oracle
    sub = prob.slice(1000, 1300)
    gs = gpu_solver(sub, opts).solve()
    assert np.array_equal(sub.X, full.X[1000:1300]) and np.array_equal(sub.U, full.U[1000:1300])
    assert np.array_equal(gs.stats.iterations, s.iterations[1000:1300])
    spot = prob.slice(4000, 4032)
    o = OracleSolver(spot, opts).solve()
    assert np.array_equal(spot.X, full.X[4000:4032]) and np.array_equal(o.stats.iterations, s.iterations[4000:4032])


def test_rocket_4096_properties():
    cold = np.load(os.path.join(GOLD, "rocket_cold.npz"))
    B = 4096
    prob, opts, _, adv = cases.case_rocket_mpc(cold["X"], cold["U"], batch=B)
    g = gpu_solver(prob, opts)
    for st in range(3):
        g.solve()
        s = g.stats
        assert np.all(s.status == 1), np.bincount(s.status)  # the reference aborts a sweep on anything else
        ok = s.status == 1
        assert np.all(s.c_max[ok] < opts.constraint_tolerance)
        U, X = prob.U[ok], prob.X[ok]
        assert np.all(np.linalg.norm(U, axis=-1) <= 196.2 + 1e-3)
        assert np.all(np.linalg.norm(U[..., :2], axis=-1) <= np.tan(np.deg2rad(5.0)) * U[..., 2] + 1e-3)
        assert np.all(np.linalg.norm(X[:, 7:20, :2], axis=-1) <= X[:, 7:20, 2] + 1e-3)
        # dynamics feasibility of the returned trajectory (dynamics_violation, simple_rocket.jl:208-216)
        A, Bm, d = prob.model.A, prob.model.B, prob.model.d
        pred = np.einsum("ij,bkj->bki", A, prob.X[:, :-1]) + np.einsum("ij,bkj->bki", Bm, prob.U) + d
        assert np.abs(pred - prob.X[:, 1:]).max() < 1e-9
        adv(prob, g, st)
    spot = np.arange(0, B, 128)
    # instance results are independent of the batch: re-solve a strided subset alone from the same warm start
