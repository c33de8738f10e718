"""Dev: a few tiny solves (small enough for a sanitizer where one is available; it is closed on the build pool): staged-panel path at n = 64, the queued
closed-loop run on rocket, the quadruped kernel."""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from altro_mpc_icra2021_b200 import solver as S
from altro_mpc_icra2021_b200.problem import SolverOptions
from altro_mpc_icra2021_b200.problems import mpc, quadruped, random_linear
from tests.helpers import lqr_problem
os.environ["ALTRO_B200_TMA"] = "1"
p = lqr_problem(n=64, m=16, N=5, batch=2, seed=3, u_bnd=0.4)
sv = S.ALTROSolver(p, SolverOptions(constraint_tolerance=1e-6)); sv.solve(); print("staged n=64", sv.stats.iterations, sv.launch_info()["smem_bytes"]); sv.close()
prob, Xt, Ut, ks = random_linear.mpc_problem(12, 6, 11, batch=6, seed=5)
sv = S.ALTROSolver(prob, random_linear.mpc_options()); sv.set_track(Xt, Ut, ks)
sv.set_noise_model(1, 0.01, 0.0); sv.set_noise_bank(mpc.rng_for(1, 2).standard_normal((3, 6, 12))); sv.solve()
r = sv.mpc_run(3); print("queued run", r["iterations"].ravel()); sv.close()
pq, _ = quadruped.mpc_problem(3); oq = quadruped.mpc_options()
sv = S.ALTROSolver(pq, oq); sv.solve(); print("quadruped", sv.stats.iterations); sv.close()
