"""Development: find a rocket MPC instance where GPU and oracle diverge and print both traces."""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
np.set_printoptions(linewidth=200, precision=6)
from altro_mpc_icra2021_b200 import solver as S
from altro_mpc_icra2021_b200.problems import rocket, mpc
from oracle.oracle import OracleProblem

B = 256
cold = rocket.cold_problem()
oc = OracleProblem(cold); rc = oc.solve(rocket.cold_options())
Xt, Ut = rc.X[0], rc.U[0]
pm, ks = rocket.mpc_problem(cold, Xt, Ut, 21, batch=B)
pg = copy.deepcopy(pm)
opts = rocket.mpc_options()
op = OracleProblem(pm); sg = S.ALTROSolver(pg, opts)
ro = op.solve(opts, 8); sg.solve()
rng = mpc.rng_for(7, 7)
class _S:
    prob = pm
lo = mpc.MPCLoop(_S(), Xt, Ut, ks, noise=None)
x0 = lo.plant_step(); x0 = x0 + rocket.noise(x0, rng)
k = ks + 1
Xr, Ur = mpc.window_reference(Xt, Ut, k, pm.N)
for p_ in (pm, pg):
    p_.set_initial_state(x0); p_.update_trajectory(Xr, Ur)
op.shift_fill(True, True); sg.shift_fill(True, True)
print("warm start equal:", np.abs(pm.U - pg.U).max(), np.abs(op.lam - sg.get_duals()).max())
tro = op.set_trace(64); sg.set_trace(64)
ro = op.solve(opts, 1); sg.solve(); g = sg.stats; trg = sg.get_trace()
bad = np.flatnonzero((g.iterations != ro.iterations) | (np.abs(pg.X - ro.X).max(axis=(1, 2)) > 1e-6))
print("bad instances", bad[:20], len(bad))
for i in bad[:3]:
    print("=== instance", i, "iters g/o", g.iterations[i], ro.iterations[i], "outer", g.iterations_outer[i], ro.iterations_outer[i])
    nrow = max(g.iterations[i], ro.iterations[i]) + 1
    for r in range(min(nrow, 12)):
        print(" o", tro[i, r]); print(" g", trg[i, r])
