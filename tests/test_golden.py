"""The oracle reproduces the committed golden vectors bit for bit (tests/golden/make_golden.py).
PARITY UNPINNED against Julia Altro.jl (not runnable here); these pin the oracle against drift."""
import os

import numpy as np
import pytest

from tests.golden import cases
from tests.helpers import OracleSolver

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def check(out, gold):
    assert set(out) == set(gold.files)
    for k in gold.files:
        assert np.array_equal(out[k], gold[k]), k


def test_rocket_cold_and_mpc_golden():
    gold = np.load(os.path.join(GOLD, "rocket_cold.npz"))
    prob, opts = cases.rocket_track()
    s = OracleSolver(prob, opts, nthreads=1).solve()
    assert np.array_equal(prob.X[0], gold["X"]) and np.array_equal(prob.U[0], gold["U"])
    assert np.array_equal(s.stats.iterations, gold["iters"]) and s.stats.status[0] == 1
    out = cases.run_case(lambda p, o: OracleSolver(p, o, nthreads=2), *cases.case_rocket_mpc(gold["X"], gold["U"]))
    check(out, np.load(os.path.join(GOLD, "rocket_mpc.npz")))


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_family_golden(name):
    out = cases.run_case(lambda p, o: OracleSolver(p, o, nthreads=2), *cases.CASES[name]())
    check(out, np.load(os.path.join(GOLD, f"{name}.npz")))
    last = max(int(k[6:]) for k in out if k.startswith("status"))
    assert np.all(out[f"status{last}"] == 1)


def test_thread_count_does_not_change_results():
    a = cases.run_case(lambda p, o: OracleSolver(p, o, nthreads=1), *cases.case_random_linear(batch=9))
    b = cases.run_case(lambda p, o: OracleSolver(p, o, nthreads=5), *cases.case_random_linear(batch=9))
    for k in a:
        assert np.array_equal(a[k], b[k]), k
