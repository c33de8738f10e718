"""Host-side mirror of the reference API: constraint lowering, dirty tracking, problem slicing, sharding."""
import numpy as np
import pytest

from altro_mpc_icra2021_b200 import sharding
from altro_mpc_icra2021_b200.problem import (BoundConstraint, ConstraintList, GoalConstraint, Inequality,
                                             LinearConstraint, LinearModel, LQRObjective, NormConstraint,
                                             NormConstraint2, Problem, SecondOrderCone, SolverOptions, CONTROL, EQUALITY,
                                             INEQUALITY, SECOND_ORDER_CONE, STATE)
from altro_mpc_icra2021_b200.problems import quadruped
from tests.helpers import lqr_problem


def test_bound_constraint_rows_follow_trajopt_order():
    cons = ConstraintList(3, 2, 5)
    cons.add_constraint(BoundConstraint(3, 2, x_max=[1.0, np.inf, 2.0], u_min=-0.5, u_max=[0.5, np.inf]), (0, 5))
    sx, su = cons.flat
    assert sx.side == STATE and sx.sense == INEQUALITY and list(sx.inds) == [0, 2] and (sx.k0, sx.k1) == (0, 5)
    assert np.array_equal(sx.G, [[1, 0], [0, 1]]) and np.array_equal(sx.h, [-1, -2])
    # rows [u - u_max (finite); u_min - u], control blocks clipped to N-1 (u_N is not a decision variable)
    assert su.side == CONTROL and (su.k0, su.k1) == (0, 4) and list(su.inds) == [0, 1]
    assert np.array_equal(su.G, [[1, 0], [-1, 0], [0, -1]]) and np.array_equal(su.h, [-0.5, -0.5, -0.5])
    assert cons.dual_len() == 5 * 2 + 4 * 3 and cons.dual_offsets() == [0, 10]


def test_cone_constraints_lower_to_affine_blocks():
    cons = ConstraintList(6, 3, 10)
    cons.add_constraint(NormConstraint(6, 3, 7.0, SecondOrderCone, ":control"), (0, 9))
    A = np.zeros((6, 6))
    A[0, 0] = A[1, 1] = 1
    c = np.zeros(6)
    c[2] = 0.5
    cons.add_constraint(NormConstraint2(6, 3, A, c, SecondOrderCone, ":state"), (2, 9))
    cons.add_constraint(NormConstraint2(6, 3, A, c, SecondOrderCone, ":state", compact=False), (2, 9))
    cons.add_constraint(GoalConstraint(np.arange(6.0)), 9)
    nrm, compact, full, goal = cons.flat
    assert nrm.sense == SECOND_ORDER_CONE and nrm.p == 4 and nrm.h[-1] == 7.0 and np.array_equal(nrm.G[:3], np.eye(3))
    assert compact.p == 3 and list(compact.inds) == [0, 1, 2] and full.p == 7 and full.w == 6
    z = np.random.default_rng(0).standard_normal(6)
    v1, v2 = compact.G @ z[compact.inds] + compact.h, full.G @ z[full.inds] + full.h
    assert np.isclose(np.linalg.norm(v1[:-1]), np.linalg.norm(v2[:-1])) and v1[-1] == v2[-1]
    assert goal.sense == EQUALITY and (goal.k0, goal.k1) == (9, 10) and np.array_equal(goal.h, -np.arange(6.0))


def test_per_knot_constraint_data_is_clipped_with_the_range():
    n, m, N = 2, 2, 6
    A = np.arange(N * 3 * 2, dtype=float).reshape(N, 3, 2)
    b = np.arange(N * 3, dtype=float).reshape(N, 3)
    cons = ConstraintList(n, m, N)
    cons.add_constraint(LinearConstraint(n, m, A, b, Inequality, ":control", per_knot=True), (0, N))
    c = cons.flat[0]
    assert (c.k0, c.k1) == (0, N - 1) and c.G.shape == (N - 1, 3, 2) and np.array_equal(c.h, -b[:N - 1])


def test_mutators_mark_dirty_and_keep_buffers_in_place():
    prob = lqr_problem(batch=3, u_bnd=1.0)
    for k in ("x0", "ref", "dyn", "traj"):
        prob.dirty[k] = False
    prob.dirty["con"] = set()
    x0_buf = prob.x0
    prob.set_initial_state(np.ones(prob.n))
    assert prob.x0 is x0_buf and np.all(prob.x0 == 1.0) and prob.dirty["x0"] and not prob.dirty["ref"]
    prob.update_trajectory(np.zeros((prob.N, prob.n)), np.ones((prob.N - 1, prob.m)))
    assert prob.dirty["ref"] and np.all(prob.Uref == 1.0)
    prob.set_constraint_data(0, h=prob.constraints.flat[0].h * 2)
    assert prob.dirty["con"] == {0}
    prob.initial_controls(np.zeros(prob.m))
    assert prob.dirty["traj"]


def test_options_copy_is_independent():
    a = SolverOptions(cost_tolerance=1e-3)
    b = a.copy()
    b.cost_tolerance = 1.0
    assert a.cost_tolerance == 1e-3 and a.penalty_scaling == 10.0 and a.iterations_outer == 30


def test_projected_newton_is_refused_not_ignored():
    """The polish step is not built (every benchmark of the reference sets projected_newton = false): asking for it is
    an error at option upload, before any device call."""
    from altro_mpc_icra2021_b200.solver import ALTROSolver, AltroError

    sv = ALTROSolver.__new__(ALTROSolver)  # no library, no device: _push_options must fail before touching either
    sv.opts = SolverOptions(projected_newton=True)
    with pytest.raises(AltroError, match="projected_newton"):
        sv._push_options()
    sv.h = None


@pytest.mark.parametrize("total,world", [(4096, 8), (10, 3), (5, 8), (1, 1)])
def test_shard_ranges_partition_the_batch(total, world):
    edges = [sharding.shard_range(total, r, world) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == total
    assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
    sizes = [b - a for a, b in edges]
    assert max(sizes) - min(sizes) <= 1


def test_problem_slice_carries_per_instance_data():
    prob, _ = quadruped.mpc_problem(7)
    sub = prob.slice(2, 5)
    assert sub.B == 3 and np.array_equal(sub.model.B, prob.model.B[2:5]) and np.array_equal(sub.x0, prob.x0[2:5])
    assert len(sub.constraints) == len(prob.constraints) and sub.dt == prob.dt
    sub.x0[...] = 0
    assert np.any(prob.x0[2:5] != 0)  # a copy, not a view
