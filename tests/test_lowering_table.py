"""One lowering table for both host mirrors (altro_mpc_icra2021_b200/lowering.tbl): the Python classes must produce
exactly what the table's rules say, and the Julia shim (julia/AltroB200TO.jl, not executable here) must read the same
file and implement every rule the table names."""
import os
import re

import numpy as np

from altro_mpc_icra2021_b200 import problem as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def table():
    rows = {}
    for line in open(os.path.join(ROOT, "altro_mpc_icra2021_b200", "lowering.tbl")):
        if line.strip() and not line.startswith("#"):
            f = [x.strip() for x in line.split("|")]
            rows[f[0]] = dict(sense=f[1], side=f[2], G=f[3], h=f[4], layout=f[5], ref=f[6])
    return rows


def test_julia_shim_reads_the_table_and_implements_every_rule():
    jl = open(os.path.join(ROOT, "julia", "AltroB200TO.jl")).read()
    assert "lowering.tbl" in jl and "read_lowering_table" in jl
    t = table()
    for kind, key in (("G_rule", "G"), ("h_rule", "h"), ("side_inds", "side")):
        have = set(re.findall(kind + r"\(::Val\{:(\w+)\}", jl))
        need = {r[key] for r in t.values()}
        assert need <= have, (kind, need - have)
    for sense in {r["sense"] for r in t.values()}:
        assert sense == "from_con" or f":{sense}" in jl or f"TO.{sense}" in jl
    # every constraint type the reference's benchmark scripts add is in the table
    for name in ("BoundConstraint", "GoalConstraint", "NormConstraint", "NormConstraint2", "AffineSOCTraj",
                 "LinearConstraintTraj", "LinearizedFrictionConstraint"):
        assert name in t


def rule_G(rule, w, data):
    if rule == "identity":
        return np.eye(w)
    if rule == "identity_0":
        return np.vstack([np.eye(w), np.zeros((1, w))])
    if rule == "stack_A_c":
        return np.vstack([data["A"], data["c"][None, :]])
    if rule == "A":
        return data["A"]
    if rule == "friction_pyramid":
        mu = data["mu"]
        return np.array([[1, 0, -mu], [-1, 0, -mu], [0, 1, -mu], [0, -1, -mu]], float)
    raise KeyError(rule)


def rule_h(rule, p, data):
    return {"zero": lambda: np.zeros(p), "minus_b": lambda: -data["b"], "minus_xf": lambda: -data["xf"],
            "val_last": lambda: np.r_[np.zeros(p - 1), data["val"]]}[rule]()


def test_python_classes_follow_the_table():
    t = table()
    rng = np.random.default_rng(0)
    n, m = 6, 3
    SENSE = {"Equality": P.EQUALITY, "Inequality": P.INEQUALITY, "SecondOrderCone": P.SECOND_ORDER_CONE}
    # GoalConstraint
    xf = rng.standard_normal(n)
    (side, idx, G, h, sense, pk, pi), = P.GoalConstraint(xf).lower(n, m)
    r = t["GoalConstraint"]
    assert side == P.STATE and sense == SENSE[r["sense"]] and np.array_equal(G, rule_G(r["G"], n, {}))
    assert np.array_equal(h, rule_h(r["h"], n, {"xf": xf}))
    # NormConstraint
    (side, idx, G, h, sense, pk, pi), = P.NormConstraint(n, m, 7.5, P.SecondOrderCone, ":control").lower(n, m)
    r = t["NormConstraint"]
    assert side == P.CONTROL and sense == SENSE[r["sense"]] and np.array_equal(G, rule_G(r["G"], m, {}))
    assert np.array_equal(h, rule_h(r["h"], m + 1, {"val": 7.5}))
    # NormConstraint2 / FrictionConstraint / AffineSOCTraj: [A; c'] (compact=False keeps the table's literal form)
    A, c = rng.standard_normal((3, 3)), rng.standard_normal(3)
    (side, idx, G, h, sense, pk, pi), = P.NormConstraint2(n, m, A, c, P.SecondOrderCone, ":control", compact=False).lower(n, m)
    for name in ("NormConstraint2", "FrictionConstraint", "AffineSOCTraj"):
        r = t[name]
        assert sense == SENSE[r["sense"]] and np.array_equal(G, rule_G(r["G"], 3, {"A": A, "c": c}))
        assert np.array_equal(h, rule_h(r["h"], 4, {}))
    # LinearConstraint family: A y - b with the constraint's own sense
    A, b = rng.standard_normal((2, 3)), rng.standard_normal(2)
    for name in ("LinearConstraint", "LinearConstraint2", "LinearConstraintTraj"):
        r = t[name]
        (side, idx, G, h, sense, pk, pi), = P.LinearConstraint(n, m, A, b, P.Equality, ":control").lower(n, m)
        assert r["sense"] == "from_con" and sense == P.EQUALITY
        assert np.array_equal(G, rule_G(r["G"], 3, {"A": A})) and np.array_equal(h, rule_h(r["h"], 2, {"b": b}))
    # LinearizedFrictionConstraint: the quadruped builder's friction rows
    from altro_mpc_icra2021_b200.problems import quadruped
    r = t["LinearizedFrictionConstraint"]
    assert np.array_equal(quadruped.friction_rows(0.5), rule_G(r["G"], 3, {"mu": 0.5})) and r["sense"] == "Inequality"
    # BoundConstraint: +e_j rows for finite upper bounds, then -e_j rows for finite lower bounds, per side
    u_min, u_max = np.array([-np.inf, -1.0, 0.0]), np.array([2.0, np.inf, 133.0])
    blocks = P.BoundConstraint(n, m, u_min=u_min, u_max=u_max).lower(n, m)
    assert len(blocks) == 1 and t["BoundConstraint"]["G"] == "bound_rows" and t["BoundConstraint"]["h"] == "bound_rhs"
    side, idx, G, h, sense, pk, pi = blocks[0]
    assert side == P.CONTROL and sense == P.INEQUALITY and list(idx) == [0, 1, 2]
    assert np.array_equal(G, np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0], [0, 0, -1]], float))
    assert np.array_equal(h, np.array([-2.0, -133.0, -1.0, 0.0]))
