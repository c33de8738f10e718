"""Dev: one large-state random-linear batch (n=200, m=25) through a short closed-loop run; prints throughput."""
import os, sys, copy, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from altro_mpc_icra2021_b200 import solver as S
from altro_mpc_icra2021_b200.problems import mpc, random_linear
n, m = int(os.environ.get("NN", 200)), int(os.environ.get("MM", 25))
B, K = int(os.environ.get("B", 148)), int(os.environ.get("K", 2))
prob, Xt, Ut, ks = random_linear.mpc_problem(n, m, 21, batch=B, seed=500 + n + m)
sv = S.ALTROSolver(prob, random_linear.mpc_options())
sv.set_track(Xt, Ut, ks)
noise = mpc.rng_for(n, m).standard_normal((K, B, n))
sv.set_noise_model(1, 0.01, 0.0); sv.set_noise_bank(noise)
sv.solve()
rg = sv.mpc_run(K)
print(sv.launch_info(), "solves/s %.0f" % (B * K / (rg["device_ms"] * 1e-3)), "device_ms %.1f" % rg["device_ms"], "iters %.2f" % rg["iterations"].mean(), "p50 us %.0f" % np.median(rg["t_us"]))
