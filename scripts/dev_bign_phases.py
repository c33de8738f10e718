"""Dev: where the cycles of a large-state solve go (phase counters of altro_get_phase_cycles; with a library built
with -DALTRO_PHASE_TIMERS and ALTRO_B200_PHASE_DETAIL=1 the seven backward-pass slots per knot)."""
import os, sys
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
sv.phase_cycles(True)
rg = sv.mpc_run(K)
ph = sv.phase_cycles(False).astype(np.float64)
it = rg["iterations"].sum()
print(sv.launch_info()["smem_bytes"], "solves/s %.0f" % (B * K / (rg["device_ms"] * 1e-3)), "p50 us %.0f" % np.median(rg["t_us"]), "iters/solve %.2f" % rg["iterations"].mean())
if os.environ.get("ALTRO_B200_PHASE_DETAIL"):
    names = ["P1 SA,SB", "P3 factor(w0)", "P2 A'SA..", "prep/wait", "P4 gains", "P5 T1", "P6 S"]
    tot = ph[:, :7].sum()
    for i, nm in enumerate(names):
        print("  %-14s %5.1f %%   %8.1f us per knot-step" % (nm, 100 * ph[:, i].sum() / tot, ph[:, i].sum() / (it * 20) / 1.9e3))
    print("  backward pass total %.1f us per knot-step" % (tot / (it * 20) / 1.9e3))
else:
    names = ["initial rollout+cost", "backward (incl. expansion)", "forward pass", "whole solve", "expansion", "ls rollouts", "ls costs"]
    for i, nm in enumerate(names):
        print("  %-28s %5.1f %% of solve   %9.1f us per iteration" % (nm, 100 * ph[:, i].sum() / ph[:, 3].sum(), ph[:, i].sum() / it / 1.9e3))
