"""The quadruped's pre-solve step: linearisation and gait / footstep history.  CPU part: the numpy
restatement of linearized_dynamics.jl:1-66 against the closed form at the benchmark's linearisation point and against
finite differences.  GPU part: the device kernels (csrc/quadruped.cu) against the restatement at random orientations."""
import copy

import numpy as np
import pytest

from altro_mpc_icra2021_b200.problems import mpc, quadruped as Q


def random_case(rng, general=True):
    x = Q.X_DES + rng.standard_normal(12) * np.array([0.05] * 3 + ([0.3] * 3 if general else [0.0] * 3) + [0.3] * 3 + ([0.5] * 3 if general else [0.0] * 3))
    u = rng.standard_normal(12) * 5.0 + Q.U_HOVER if general else np.zeros(12)
    foot = x[0:3] + Q.NOM_FOOT + 0.02 * rng.standard_normal((4, 3))
    contacts = (rng.random(4) < 0.7).astype(float)
    return x, u, foot, contacts


def test_restatement_equals_closed_form_at_the_benchmark_linearisation_point():
    rng = np.random.default_rng(0)
    x, u, foot, contacts = random_case(rng, general=False)
    x = Q.X_DES.copy()
    A, B, d = Q.linearize_reference(x, u, foot, contacts)
    A0, B0, d0 = Q.linearized_dynamics(contacts[None, None, :], (foot - x[0:3])[None])
    assert np.allclose(A, A0[0, 0], atol=1e-14) and np.allclose(B, B0[0, 0], atol=1e-14) and np.allclose(d, d0[0, 0], atol=1e-13)


def test_restatement_jacobians_against_finite_differences_at_a_general_point():
    rng = np.random.default_rng(1)
    x, u, foot, contacts = random_case(rng)
    A, B, d = Q.linearize_reference(x, u, foot, contacts)
    f = lambda xx, uu: Q.nonlinear_dynamics(xx, uu, foot, contacts)
    h = 1e-6
    for j in range(12):
        e = np.zeros(12)
        e[j] = h
        assert np.allclose((f(x + e, u) - f(x - e, u)) / (2 * h) * Q.DT + (np.arange(12) == j), A[:, j], atol=1e-7)
        assert np.allclose((f(x, u + e) - f(x, u - e)) / (2 * h) * Q.DT, B[:, j], atol=1e-7)
    assert np.allclose(A @ x + B @ u + d, x + Q.DT * f(x, u), atol=1e-12)  # the affine model is exact at (x_ref, u_ref)
    R = Q.mrp_rotation(x[3:6])
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-14) and np.isclose(np.linalg.det(R), 1.0)


def test_foot_history_contacts_equal_the_builder_schedule():
    t = 0.37
    c, foot, plan = Q.foot_history_reference(t, Q.X_DES, Q.NOM_FOOT, Q.NOM_FOOT, Q.N_HORIZON - 1)
    assert np.array_equal(c, Q.contact_schedule(np.array([t]), Q.N_HORIZON)[0])
    assert np.allclose(foot[0], Q.X_DES[0:3] + Q.NOM_FOOT)


@pytest.mark.gpu
def test_device_linearisation_matches_the_restatement():
    from altro_mpc_icra2021_b200.solver import ALTROSolver
    B = 24
    prob, _ = Q.mpc_problem(B, linearized_friction=True, seed=5)
    K = prob.N - 1
    rng = np.random.default_rng(3)
    xs, us, feet, cons = (np.zeros((B, K, 12)), np.zeros((B, K, 12)), np.zeros((B, K, 4, 3)), np.zeros((B, K, 4)))
    for b in range(B):
        for k in range(K):
            xs[b, k], us[b, k], feet[b, k], cons[b, k] = random_case(rng)
    g = ALTROSolver(prob, Q.mpc_options())
    g.quadruped_linearize(xs, feet, cons, Q.J_BODY, Q.MASS, u_ref=us)
    A, Bm, d = g.get_dynamics()
    for b in range(0, B, 5):
        for k in range(0, K, 3):
            A0, B0, d0 = Q.linearize_reference(xs[b, k], us[b, k], feet[b, k], cons[b, k])
            assert np.allclose(A[b, k], A0, rtol=1e-12, atol=1e-13), (b, k, np.abs(A[b, k] - A0).max())
            assert np.allclose(Bm[b, k], B0, rtol=1e-12, atol=1e-13) and np.allclose(d[b, k], d0, rtol=1e-11, atol=1e-12)
    # at the benchmark's point (x_des, u_ref = 0) the device model reproduces the host builder's, and the solve agrees
    p2, st = Q.mpc_problem(B, linearized_friction=True, seed=5)
    pd = copy.deepcopy(p2)
    gh = ALTROSolver(p2, Q.mpc_options()).solve()
    gd = ALTROSolver(pd, Q.mpc_options())
    contacts = Q.contact_schedule(st["t0"], p2.N)
    foot = (Q.X_DES[0:3] + st["foot_rel"])[:, None].repeat(K, 1)
    gd.quadruped_linearize(np.broadcast_to(Q.X_DES, (B, 12)), foot, contacts, Q.J_BODY, Q.MASS)
    A, Bm, d = gd.get_dynamics()
    assert np.allclose(A, p2.model.A, atol=1e-15) and np.allclose(Bm, p2.model.B, atol=1e-15) and np.allclose(d, p2.model.d, atol=1e-14)
    gd.solve()
    assert np.array_equal(gd.stats.iterations, gh.stats.iterations) and np.allclose(pd.U, p2.U, rtol=1e-9, atol=1e-9)


@pytest.mark.gpu
def test_device_gait_and_footstep_history_match_the_restatement():
    from altro_mpc_icra2021_b200.solver import ALTROSolver
    B = 40
    prob, _ = Q.mpc_problem(B, linearized_friction=True, seed=7)
    K = prob.N - 1
    rng = np.random.default_rng(9)
    g = ALTROSolver(prob, Q.mpc_options())
    t = rng.random(B) * 2.0
    cur = Q.NOM_FOOT[None] + 0.02 * rng.standard_normal((B, 4, 3))
    xref = np.broadcast_to(Q.X_DES, (B, K, 12)).copy()
    xref[:, :, 0:2] += 0.05 * rng.standard_normal((B, 1, 2))
    xref[:, :, 3:6] += 0.1 * rng.standard_normal((B, 1, 3))
    xref[:, :, 6:8] = 0.3 * rng.standard_normal((B, 1, 2))
    planner = cur.copy()
    for tick in range(3):  # the planner state carries over from tick to tick
        g.quadruped_tick(t, xref, cur, Q.TROT.T, Q.PHASE_T, Q.NOM_FOOT, Q.J_BODY, Q.MASS)
        c, f = g.quadruped_schedule()
        A, Bm, d = g.get_dynamics()
        for b in range(B):
            c0, f0, planner[b] = Q.foot_history_reference(t[b], xref[b], cur[b], planner[b], K)
            assert np.array_equal(c[b], c0), (tick, b)
            assert np.allclose(f[b], f0, rtol=1e-13, atol=1e-14), (tick, b)
            if b % 9 == 0:
                A0, B0, d0 = Q.linearize_reference(xref[b, 4], np.zeros(12), f0[4], c0[4])
                assert np.allclose(A[b, 4], A0, atol=1e-13) and np.allclose(Bm[b, 4], B0, atol=1e-13)
        t = t + Q.DT
    prob.set_initial_state(xref[:, 0])
    g.solve()  # the solver runs on the model the kernels just wrote
    assert np.mean(g.stats.status == 1) > 0.9
