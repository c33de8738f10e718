"""Development: a lone CTA per SM (B = 148) fused run, for latency profiling under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from altro_mpc_icra2021_b200 import solver as S
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "rocket"
B = int(os.environ.get("B", "148")); K = int(os.environ.get("K", "10"))
wl = bench.Workload(name, B, 0xA1722, lambda p, o: S.ALTROSolver(p, o))
sv = S.ALTROSolver(wl.prob, wl.opts)
if wl.track is not None: sv.set_track(wl.track[0], wl.track[1], wl.k)
sv.set_noise_model(*wl.noise_model); sv.set_noise_bank(wl.noise_samples(2 * K + 3))
sv.solve()
r = sv.mpc_run(K, shift=wl.shift)
print(name, sv.launch_info(), r["device_ms"])
