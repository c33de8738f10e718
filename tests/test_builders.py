"""Problem builders against known-answer values derived from the reference's own formulas
(flexible_sat_mpc.jl:72-130, rocket_landing_problem.jl:19-39,
MPC.yaml, Woofer.yaml, mpc.jl:33-40)."""
import numpy as np

from altro_mpc_icra2021_b200.problem import CONTROL, INEQUALITY, SECOND_ORDER_CONE, STATE
from altro_mpc_icra2021_b200.problems import flexsat, mpc, quadruped, random_linear, rocket


def test_flexsat_zoh_known_answers():
    Ad, Bd = flexsat.generate_AB()
    assert np.isclose(np.linalg.norm(Ad), 3.726671412481, rtol=1e-12)
    assert np.isclose(np.linalg.norm(Bd), 1.419673625114, rtol=1e-12)
    assert np.isclose(np.trace(Ad), 10.910525829915, rtol=1e-12)
    assert np.isclose(Ad[0, 3], 0.125) and np.isclose(Ad[3, 6], -0.003281905011296, rtol=1e-10)
    assert np.isclose(Ad[9, 6], -0.073781799799133, rtol=1e-10) and np.isclose(Ad[11, 11], 0.846910980397957, rtol=1e-12)
    assert np.allclose(Bd[3], [-0.968647749863, 0.062514115021, 0.033400776196], rtol=1e-9)
    assert np.allclose(Bd[9], [-0.033252650035, 0.004434198057, 0.250772102422], rtol=1e-9)
    assert np.isclose(max(abs(np.linalg.eigvals(Ad))), 1.0, atol=1e-9)


def test_rocket_model_and_constraints():
    p = rocket.cold_problem()
    A, B, d = p.model.A, p.model.B, p.model.d
    assert np.allclose(A, np.block([[np.eye(3), 0.05 * np.eye(3)], [np.zeros((3, 3)), np.eye(3)]]))
    assert np.allclose(B, np.vstack([1.25e-4 * np.eye(3), 5e-3 * np.eye(3)]))
    assert np.allclose(d, [0, 0, -0.0122625, 0, 0, -0.4905])
    names = {c.name: c for c in p.constraints.flat}
    assert (names["GoalConstraint"].k0, names["GoalConstraint"].k1) == (300, 301)
    th = names["max_thrust"]
    assert th.sense == SECOND_ORDER_CONE and th.side == CONTROL and (th.k0, th.k1) == (0, 300) and np.isclose(th.h[-1], 196.2)
    ang = names["thrust_angle"]
    assert np.isclose(ang.G[-1, 2], 0.087488663526, rtol=1e-10)
    gl = names["glideslope"]
    assert gl.side == STATE and (gl.k0, gl.k1) == (7, 300) and np.isclose(gl.G[-1, -1], 1.0)
    assert np.allclose(p.U[0, 0], [0, 0, 98.1])


def test_tracking_problem_reindexes_constraints():
    cold = rocket.cold_problem()
    X = np.zeros((301, 6))
    U = np.zeros((300, 3))
    pm, ks = rocket.mpc_problem(cold, X, U, 21, batch=5)
    rng = {c.name: (c.k0, c.k1) for c in pm.constraints.flat}
    # mpc.jl:33-40: goal dropped, thrust cones 1..20, glideslope 8..20 (1-based)
    assert "GoalConstraint" not in rng
    assert rng == {"max_thrust": (0, 20), "thrust_angle": (0, 20), "glideslope": (7, 20)}
    assert np.allclose(pm.obj.Q, 10.0) and np.allclose(pm.obj.R, 0.1) and np.allclose(pm.obj.Qf, 10.0)
    assert pm.B == 5 and pm.Xref.shape == (5, 21, 6)


def test_quadruped_batch_structure():
    assert np.isclose(quadruped.MASS, 7.692) and np.isclose(quadruped.U_HOVER[2], 18.86463, rtol=1e-6)
    assert np.isclose(quadruped.NOM_FOOT[0, 2], -0.264575131106, rtol=1e-10)
    p, st = quadruped.mpc_problem(6)
    A, B, d = p.model.A, p.model.B, p.model.d
    assert A.shape == (6, 14, 12, 12) and np.allclose(A[:, :, :3, 6:9], 0.03 * np.eye(3))
    assert np.allclose(A[:, :, 3:6, 9:12], 0.25 * 0.03 * np.eye(3)) and np.allclose(d[:, :, 8], -9.81 * 0.03)
    c = quadruped.contact_schedule(np.array([0.0, 0.21, 0.41, 0.61]))
    assert np.array_equal(c[:, 0], [[1, 1, 1, 1], [1, 0, 0, 1], [1, 1, 1, 1], [0, 1, 1, 0]])
    # swing feet have zero columns in B_k
    con = quadruped.contact_schedule(st["t0"])
    for i in range(4):
        assert np.all((B[:, :, :, 3 * i:3 * i + 3] != 0).any(axis=(2, 3)) == (con[:, :, i] > 0))
    kinds = [(c.name, c.sense, c.p, c.w) for c in p.constraints.flat]
    assert kinds[:4] == [(f"friction{i}", INEQUALITY, 4, 3) for i in range(4)] and kinds[4][2:] == (8, 4)
    ps, _ = quadruped.mpc_problem(2, linearized_friction=False)
    assert [(c.sense, c.p, c.w) for c in ps.constraints.flat[:4]] == [(SECOND_ORDER_CONE, 4, 3)] * 4


def test_random_linear_generator_is_stable_and_deterministic():
    a = random_linear.mpc_problem(12, 6, 21, batch=4)
    b = random_linear.mpc_problem(12, 6, 21, batch=4)
    assert np.array_equal(a[0].model.A, b[0].model.A) and np.array_equal(a[0].Xref, b[0].Xref)
    assert max(abs(np.linalg.eigvals(a[0].model.A))) <= 1.0
    bound = a[0].constraints.flat[0]
    assert bound.p == 12 and bound.w == 6 and (bound.k0, bound.k1) == (0, 20) and np.allclose(bound.h, -3.0)


def test_window_reference_clamps_at_track_end():
    X = np.arange(10.0)[:, None] * np.ones((10, 2))
    U = np.arange(9.0)[:, None] * np.ones((9, 1))
    Xr, Ur = mpc.window_reference(X, U, np.array([0, 7]), 4)
    assert np.array_equal(Xr[1, :, 0], [7, 8, 9, 9]) and np.array_equal(Ur[1, :, 0], [7, 8, 8])
