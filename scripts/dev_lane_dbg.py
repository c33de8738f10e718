"""Dev: lane kernel vs oracle on a small rocket batch, prints mismatch statistics."""
import copy, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.golden import cases
from tests.helpers import OracleSolver
from altro_mpc_icra2021_b200.solver import ALTROSolver
from altro_mpc_icra2021_b200.problems import mpc
cold = np.load("tests/golden/rocket_cold.npz")
prob, opts, _, _ = cases.case_rocket_mpc(cold["X"], cold["U"], batch=int(os.environ.get("B", 64)))
prob.set_initial_state(prob.x0 + 0.05 * mpc.rng_for(2, 3).standard_normal(prob.x0.shape))
pg = copy.deepcopy(prob)
o = OracleSolver(prob, opts).solve()
g = ALTROSolver(pg, opts, kernel="lane").solve()
print(os.environ.get("ALTRO_B200_LIB"), g.launch_info().get("instances_per_warp"), "X equal", np.array_equal(pg.X, prob.X), "maxdiff %.3e" % np.abs(pg.X - prob.X).max(),
      "iters equal", np.array_equal(g.stats.iterations, o.stats.iterations), g.stats.iterations[:12], o.stats.iterations[:12], "tsolve ms", g.stats.tsolve)
