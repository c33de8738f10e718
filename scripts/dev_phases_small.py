"""Dev: coarse phase split (altro_get_phase_cycles) of a closed-loop run of one bench workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from altro_mpc_icra2021_b200 import solver as S
name = sys.argv[1] if len(sys.argv) > 1 else "rocket"
K, W = int(os.environ.get("K", 20)), 3
mk = lambda prob, opts, **kw: S.ALTROSolver(prob, opts, **{k: v for k, v in kw.items() if k != "pin"})
wl = bench.Workload(name, int(os.environ.get("B", 4096)), 78, mk)
sv = mk(wl.prob, wl.opts)
if wl.track is not None:
    sv.set_track(wl.track[0], wl.track[1], wl.k)
sv.set_noise_model(*wl.noise_model)
sv.set_noise_bank(wl.noise_samples(W + K))
sv.solve()
sv.mpc_run(W, shift=wl.shift)
sv.phase_cycles(True)
rr = sv.mpc_run(K, shift=wl.shift)
ph = sv.phase_cycles(False).astype(np.float64)
it, tr = rr["iterations"].sum(), rr["ls_trials"].sum()
print(name, sv.launch_info())
print("solves/s %.0f  iters/solve %.2f  trials/solve %.2f  cycles per solve %.0f" % (wl.batch * K / (rr["device_ms"] * 1e-3), it / (wl.batch * K), tr / (wl.batch * K), ph[:, 3].sum() / (wl.batch * K)))
if os.environ.get("ALTRO_B200_PHASE_DETAIL"):
    nk = wl.prob.N - 1
    tot = ph[:, :7].sum()
    for i, nm in enumerate(["P1 SA,SB", "P3 factor (warp 0)", "P2 A'SA..", "prep / wait", "P4 gains", "P5 T1", "P6 S"]):
        print("  %-20s %5.1f %%   %7.0f cycles per knot" % (nm, 100 * ph[:, i].sum() / tot, ph[:, i].sum() / (it * nk)))
    print("  backward knots total %.0f cycles per knot" % (tot / (it * nk)))
    sys.exit(0)
names = ["initial rollout+cost", "backward (incl. expansion)", "forward pass", "whole solve", "expansion", "ls rollouts", "ls costs"]
for i, nm in enumerate(names):
    print("  %-28s %5.1f %% of solve   %8.0f cycles per iteration" % (nm, 100 * ph[:, i].sum() / ph[:, 3].sum(), ph[:, i].sum() / it))
