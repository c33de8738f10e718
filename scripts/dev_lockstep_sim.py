"""Dev: list-scheduling simulation of the lock-step API (one launch per MPC step, grid order) on the iteration counts
the CPU oracle produces for a workload: makespan in index order, longest-first by the PREVIOUS step's iteration count,
longest-first with perfect knowledge, and the lower bound max(longest, sum / slots).  profiles/r2_summary.md quotes it."""
import sys, heapq, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oracle.oracle import OracleProblem
class _OS:
    def __init__(self, prob, opts):
        self.prob, self.opts, self.op = prob, opts, OracleProblem(prob)
    def solve(self):
        self.stats = self.op.solve(self.opts, nthreads=16); return self
wl = bench.Workload(sys.argv[1] if len(sys.argv) > 1 else "rocket", 2048, 78, _OS)
K = 20
op = OracleProblem(wl.prob); op.solve(wl.opts, 16)
k = wl.k.copy()
ro = op.mpc_run(wl.opts, K, wl.noise_samples(K), wl.noise_model, wl.track, k, wl.shift, 16)
it = ro["iterations"]  # (K,B)
print("iters mean %.2f max %d" % (it.mean(), it.max()))
def makespan(dur, order, slots):
    h = [0.0] * slots; heapq.heapify(h)
    end = 0.0
    for i in order:
        t = heapq.heappop(h); t2 = t + dur[i]; end = max(end, t2); heapq.heappush(h, t2)
    return end
slots = 592  # 2048 instances on half a GPU's worth of slots: same 3.46 waves as 4096 on 1184
res = {"index": [], "lpt_prev": [], "lpt_oracle": [], "bound": []}
for s in range(1, K):
    dur = it[s].astype(float) + 0.6   # fixed part of a solve ~ 0.6 iterations' worth
    B = len(dur)
    res["index"].append(makespan(dur, range(B), slots))
    res["lpt_prev"].append(makespan(dur, np.argsort(-it[s - 1], kind="stable"), slots))
    res["lpt_oracle"].append(makespan(dur, np.argsort(-dur, kind="stable"), slots))
    res["bound"].append(max(dur.max(), dur.sum() / slots))
for k, v in res.items(): print(k, "mean makespan %.2f" % np.mean(v))
print("corr(prev, cur) %.2f" % np.corrcoef(it[:-1].ravel(), it[1:].ravel())[0, 1])
