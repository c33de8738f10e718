"""Development: which instances make the tail of a fused run (run on a GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from altro_mpc_icra2021_b200 import solver as S
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "rocket"
B = int(os.environ.get("B", "4096")); K = int(os.environ.get("K", "100"))
wl = bench.Workload(name, B, 0xA1722, lambda p, o: S.ALTROSolver(p, o))
sv = S.ALTROSolver(wl.prob, wl.opts)
if wl.track is not None: sv.set_track(wl.track[0], wl.track[1], wl.k)
sv.set_noise_model(*wl.noise_model); sv.set_noise_bank(wl.noise_samples(K + 6))
sv.solve()
print("cold: iters", np.bincount(sv.stats.iterations)[:60].nonzero()[0][[0, -1]], "status", np.bincount(sv.stats.status))
sv.mpc_run(3, shift=wl.shift)
r = sv.mpc_run(K, shift=wl.shift)
it, st, t, ls, cm = r["iterations"], r["status"], r["t_us"], r["ls_trials"], r["c_max"]
print("opts: iterations", wl.opts.iterations, "outer", wl.opts.iterations_outer, "inner", wl.opts.iterations_inner, "ctol", wl.opts.constraint_tolerance)
print("device ms", r["device_ms"], "solves/s", B * K / r["device_ms"] * 1e3)
print("status histogram", np.bincount(st.ravel()))
q = [50, 90, 99, 99.9, 100]
print("iters pct", np.percentile(it, q), "ls pct", np.percentile(ls, q), "t_us pct", np.percentile(t, q))
tot = t.sum(axis=0)
order = np.argsort(-tot)
print("per-instance total ms pct", np.percentile(tot, q) / 1e3, "sum/slots(1184) ms", tot.sum() / 1184 / 1e3)
for i in order[:8]:
    bad = np.argsort(-t[:, i])[:5]
    print(f"inst {i}: total {tot[i]/1e3:.1f} ms; iters sum {it[:, i].sum()} max {it[:, i].max()}; statuses {np.bincount(st[:, i])}; worst steps {bad.tolist()} t {t[bad, i].astype(int).tolist()} it {it[bad, i].tolist()} ls {ls[bad, i].tolist()} cmax {cm[bad, i]}")
# share of total time spent in non-converged solves
ok = st == st.ravel()[np.argmin(t.ravel())]
print("share of CTA time in solves whose status differs from the fastest solve's:", t[~ok].sum() / t.sum(), "count", (~ok).sum())
print("time share by iteration count bucket:", {b: round(float(t[(it >= lo) & (it < hi)].sum() / t.sum()), 3) for b, (lo, hi) in {"<=3": (0, 4), "4-10": (4, 11), "11-50": (11, 51), ">50": (51, 10**9)}.items()})
np.savez_compressed(os.path.join("gpurun_out", f"tail_{name}.npz"), t_us=t, it=it)
