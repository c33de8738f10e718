"""Development: run-to-run variation of the same fused run replayed from one snapshot (run on a GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from altro_mpc_icra2021_b200 import solver as S
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "rocket"
B = int(os.environ.get("B", "4096")); K = int(os.environ.get("K", "100")); T = int(os.environ.get("T", "0"))
idle = float(os.environ.get("IDLE", "0"))
wl = bench.Workload(name, B, 0xA1722, lambda p, o: S.ALTROSolver(p, o))
sv = S.ALTROSolver(wl.prob, wl.opts, threads_per_instance=T)
if wl.track is not None: sv.set_track(wl.track[0], wl.track[1], wl.k)
sv.set_noise_model(*wl.noise_model); sv.set_noise_bank(wl.noise_samples(K + 3))
sv.solve(); sv.mpc_run(3, shift=wl.shift); sv.snapshot()
out = []
for rep in range(6):
    sv.restore()
    if idle and rep in (3, 4): time.sleep(idle)
    r = sv.mpc_run(K, shift=wl.shift)
    tot = r["t_us"].sum(axis=0)
    out.append((round(r["device_ms"], 1), round(tot.max() / 1e3, 1), int(np.argmax(tot)), round(tot.mean() / 1e3, 2)))
print(name, "device_ms / chain max ms / argmax / mean ms per replay:", out)
