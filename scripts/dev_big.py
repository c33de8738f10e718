"""Development: phase cycle breakdown of the large-dimension path (run on a GPU box)."""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from altro_mpc_icra2021_b200 import solver as S
from altro_mpc_icra2021_b200.problems import mpc, random_linear
n, m = int(sys.argv[1]), int(sys.argv[2]); B = int(os.environ.get("B", "296")); K = 4
prob, Xt, Ut, ks = random_linear.mpc_problem(n, m, 21, batch=B, seed=500 + n + m)
opts = random_linear.mpc_options()
sv = S.ALTROSolver(prob, opts, threads_per_instance=int(os.environ.get("T", "0")))
sv.set_track(Xt, Ut, ks); sv.set_noise_model(1, 0.01, 0.0); sv.set_noise_bank(mpc.rng_for(n, m).standard_normal((K, B, n)))
print(sv.launch_info())
sv.solve(); sv.phase_cycles(True)
r = sv.mpc_run(K); ph = sv.phase_cycles(False).astype(float)
it, ls = r["iterations"], r["ls_trials"]
F = lambda n, m: 4 * n**3 + 8 * n * n * m + 6 * n * m * m + m**3 / 3
flops = it.sum() * 20 * F(n, m)
print(f"{n}x{m}: {r['device_ms']:.1f} ms, {B*K/r['device_ms']*1e3:.0f} solves/s, iters {it.mean():.2f} ls {ls.mean():.2f}, bp flops {flops/r['device_ms']/1e9:.2f} TFLOP/s")
tot = ph[:, 3].sum()
names = ["init rollout+cost", "backward(+expand)", "forward", "whole", "expand", "ls rollouts", "ls costs"]
print("cycle shares:", {nm: round(ph[:, i].sum() / tot, 3) for i, nm in enumerate(names)})
print("cycles per: bp %.0f  expand %.0f  ls-rollout %.0f  ls-cost %.0f" % ((ph[:, 1].sum() - ph[:, 4].sum()) / it.sum(), ph[:, 4].sum() / it.sum(), ph[:, 5].sum() / ls.sum(), ph[:, 6].sum() / ls.sum()))
