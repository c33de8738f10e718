"""Dev: per-family iteration / line-search statistics of the CPU oracle on the bench workloads."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oracle.oracle import OracleProblem

class _OS:
    def __init__(self, prob, opts):
        self.prob, self.opts, self.op = prob, opts, OracleProblem(prob)
    def solve(self):
        self.stats = self.op.solve(self.opts, nthreads=8)
        return self

def main():
    names = sys.argv[1].split(",")
    mods = {}
    for a in sys.argv[2:]:
        k, v = a.split("=")
        mods[k] = float(v) if "." in v or "e" in v else int(v)
    B, steps = int(os.environ.get("B", 256)), int(os.environ.get("STEPS", 50))
    for name in names:
        wl = bench.Workload(name, B, 0xA1720 + 2, _OS)
        for k, v in mods.items(): setattr(wl.opts, k, v)
        op = OracleProblem(wl.prob)
        op.solve(wl.opts, nthreads=8)
        t0 = time.perf_counter()
        r = op.mpc_run(wl.opts, steps, wl.noise_samples(steps), wl.noise_model, wl.track, wl.k.copy(), wl.shift, 8)
        dt = time.perf_counter() - t0
        it, ls, st, ou = r["iterations"].ravel(), r["ls_trials"].ravel(), r["status"].ravel(), r["iterations_outer"].ravel()
        print("%-14s iters mean %.2f med %d p99 %d max %d | outer %.2f | ls mean %.2f (%.2f/iter) max %d | fail %.4f%% %s | %.0f solves/s" % (
            name, it.mean(), np.median(it), np.quantile(it, .99), it.max(), ou.mean(), ls.mean(), ls.sum() / it.sum(), ls.max(),
            100 * np.mean(st != 1), dict(zip(*np.unique(st[st != 1], return_counts=True))), B * steps / dt), np.bincount(it)[:9])
        if os.environ.get("HIST"):
            for k in np.unique(it)[:5]:
                v = ls[it == k]
                print("    iters", k, "n", len(v), "trials", {int(a): int(b) for a, b in zip(*np.unique(v, return_counts=True))})

if __name__ == "__main__":
    main()
