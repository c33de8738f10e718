"""Unit pieces of the oracle: second-order-cone projection, its Jacobian, the closed-form AL Hessian
the solver uses against the dense mu*G'*dPi*G form, and the warm-start shift."""
import numpy as np
import pytest

from altro_mpc_icra2021_b200.problem import (ConstraintList, LinearConstraint, LinearModel, LQRObjective, Problem,
                                             SecondOrderCone, SolverOptions, CONTROL)
from oracle import oracle as orc


def proj_ref(v):
    a, t = np.linalg.norm(v[:-1]), v[-1]
    if a <= -t:
        return np.zeros_like(v)
    if a <= t:
        return v.copy()
    c = 0.5 * (1 + t / a)
    return np.concatenate([c * v[:-1], [c * a]])


@pytest.mark.parametrize("p", [2, 3, 4, 7])
def test_soc_projection_identities(p):
    rng = np.random.default_rng(p)
    for _ in range(200):
        v = rng.standard_normal(p) * 10 ** rng.uniform(-3, 3)
        pv = orc.soc_project(v)
        assert np.allclose(pv, proj_ref(v), rtol=1e-14, atol=0)
        assert np.linalg.norm(pv[:-1]) <= pv[-1] * (1 + 1e-12) + 1e-300  # in the cone
        assert np.allclose(orc.soc_project(pv), pv, rtol=1e-12, atol=1e-300)  # idempotent
        # Moreau: v = Pi_K(v) - Pi_K(-v) and the two parts are orthogonal
        pm = orc.soc_project(-v)
        assert np.allclose(pv - pm, v, rtol=1e-12, atol=1e-12 * np.abs(v).max())
        assert abs(pv @ pm) <= 1e-10 * max(1.0, v @ v)
        assert np.allclose(orc.soc_project(3.7 * v), 3.7 * pv, rtol=1e-13)  # positively homogeneous


def test_soc_projection_edge_cases():
    assert np.array_equal(orc.soc_project(np.array([0.0, 0.0, 0.0])), [0, 0, 0])
    assert np.array_equal(orc.soc_project(np.array([0.0, 0.0, 2.0])), [0, 0, 2.0])  # apex direction
    assert np.array_equal(orc.soc_project(np.array([0.0, 0.0, -2.0])), [0, 0, 0])
    assert np.array_equal(orc.soc_project(np.array([3.0, 4.0, 5.0])), [3, 4, 5])  # on the boundary: inside branch
    assert np.array_equal(orc.soc_project(np.array([3.0, 4.0, -5.0])), [0, 0, 0])  # on the polar boundary


@pytest.mark.parametrize("p", [2, 3, 4, 6])
def test_soc_jacobian_matches_finite_differences(p):
    rng = np.random.default_rng(10 + p)
    for _ in range(50):
        v = rng.standard_normal(p)
        a, t = np.linalg.norm(v[:-1]), v[-1]
        if min(abs(a - t), abs(a + t)) < 1e-3:
            continue
        J = orc.soc_project_jac(v)
        Jfd = np.zeros((p, p))
        for j in range(p):
            e = np.zeros(p)
            e[j] = 1e-6
            Jfd[:, j] = (proj_ref(v + e) - proj_ref(v - e)) / 2e-6
        assert np.allclose(J, Jfd, atol=1e-7)
        assert np.allclose(J, J.T)
        assert np.allclose(J @ v, proj_ref(v), atol=1e-12)  # homogeneity: dPi(v) v = Pi(v)


def _one_knot_soc_problem(G, h, u0, lam0, mu):
    """A 2-knot problem whose only AL term is one SOC block on u_0, to probe the solver's expansion."""
    n, m = 1, G.shape[1]
    model = LinearModel(np.zeros((1, 1)), np.zeros((1, m)), dt=1.0)
    obj = LQRObjective(np.ones(n), np.ones(m), np.ones(n), np.zeros(n), 2)
    cons = ConstraintList(n, m, 2)
    cons.add_constraint(LinearConstraint(n, m, G, -h, SecondOrderCone, (CONTROL, np.arange(m))), (0, 1))
    prob = Problem(model, obj, 2, x0=np.zeros(n), constraints=cons, U0=u0[None, :])
    return prob


@pytest.mark.parametrize("exact", [True, False])
def test_al_minimiser_matches_dense_newton(exact):
    """One iLQR iteration from u0 on a single-knot problem is a Newton step of the AL function; compare with a
    numpy Newton step built from the dense projection Jacobian (exact: mu G'dPi G, Gauss-Newton: mu G'dPi'dPi G)."""
    rng = np.random.default_rng(3)
    for trial in range(20):
        m, p = 3, 4
        G = rng.standard_normal((p, m))
        h = rng.standard_normal(p)
        u0 = rng.standard_normal(m)
        mu = 10.0 ** rng.uniform(0, 3)
        prob = _one_knot_soc_problem(G, h, u0, None, mu)
        opts = SolverOptions(penalty_initial=mu, iterations_outer=1, iterations_inner=1, soc_hess_exact=exact)
        op = orc.OracleProblem(prob)
        tr = op.set_trace(4)
        op.solve(opts)
        op.set_trace(0)
        lb = -mu * (G @ u0 + h)
        a, t = np.linalg.norm(lb[:-1]), lb[-1]
        if min(abs(a - t), abs(a + t)) < 1e-6 * max(a, abs(t)):
            continue
        Pv, J = proj_ref(lb), orc.soc_project_jac(lb)
        grad = u0 - G.T @ Pv  # R = I, dt = 1, uref = 0
        H = np.eye(m) + mu * G.T @ (J if exact else J.T @ J) @ G
        d = -np.linalg.solve(H, grad)
        # expected decrease reported by the backward pass: dV1 = d'Qu, dV2 = 1/2 d'Quu d
        assert np.isclose(tr[0, 0, 6], d @ grad, rtol=1e-9, atol=1e-12)
        assert np.isclose(tr[0, 0, 7], 0.5 * d @ H @ d, rtol=1e-9, atol=1e-12)


def test_shift_fill_semantics():
    from tests.helpers import lqr_problem

    prob = lqr_problem(n=3, m=2, N=6, batch=2, u_bnd=0.5)
    prob.X[...] = np.arange(prob.X.size).reshape(prob.X.shape)
    prob.U[...] = np.arange(prob.U.size).reshape(prob.U.shape)
    X0, U0 = prob.X.copy(), prob.U.copy()
    op = orc.OracleProblem(prob)
    op.lam[...] = np.arange(op.lam.size).reshape(op.lam.shape)
    L0 = op.lam.copy()
    op.shift_fill(True, True)
    assert np.array_equal(prob.X[:, :-1], X0[:, 1:]) and np.array_equal(prob.X[:, -1], X0[:, -1])
    assert np.array_equal(prob.U[:, :-1], U0[:, 1:]) and np.array_equal(prob.U[:, -1], U0[:, -1])
    p = prob.constraints.flat[0].p
    L = op.lam.reshape(2, 5, p)
    assert np.array_equal(L[:, :-1], L0.reshape(2, 5, p)[:, 1:]) and np.array_equal(L[:, -1], L0.reshape(2, 5, p)[:, -1])
