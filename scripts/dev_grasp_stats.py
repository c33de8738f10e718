"""Dev: grasp MPC iteration statistics of the oracle vs the reference's saved data (grasp_benchmark_data.jld2)."""
import sys, time
import numpy as np
from altro_mpc_icra2021_b200.problems import grasp, mpc
from oracle.oracle import OracleProblem

def run(Nm, B=15, opts_mod=None, trace=False, steps=None, seed=11):
    cold = grasp.cold_problem()
    rc = OracleProblem(cold).solve(grasp.cold_options())
    Xt, Ut = rc.X[0], rc.U[0]
    steps = (251 - Nm) if steps is None else steps
    pm = mpc.gen_tracking_problem(cold, Xt, Ut, Nm, Qk=1e3, Rk=1.0, Qfk=10.0, batch=B, k_start=np.zeros(B, np.int64))
    opts = grasp.mpc_options()
    if opts_mod:
        for k, v in opts_mod.items(): setattr(opts, k, v)
    op = OracleProblem(pm)
    r0 = op.solve(opts, nthreads=8)
    noise = mpc.rng_for(seed, 7).standard_normal((steps, B, 6))
    tr = op.set_trace(64) if trace else None
    out = op.mpc_run(opts, steps, noise, (1, 0.01, 0.0), (Xt, Ut), None, True, nthreads=1 if trace else 8)
    return out, tr

if __name__ == "__main__":
    mods = {}
    for a in sys.argv[1:]:
        k, v = a.split("=")
        mods[k] = float(v) if "." in v or "e" in v else int(v)
    for Nm in (11, 21, 31, 41, 51):
        out, _ = run(Nm, opts_mod=mods)
        it = out["iterations"]
        print(Nm, "mean %.2f median %.1f min %d max %d | outer mean %.2f | ls %.2f | fail %d | per-run means %s" % (
            it.mean(), np.median(it), it.min(), it.max(), out["iterations_outer"].mean(), out["ls_trials"].mean(),
            int((out["status"] != 1).sum()), np.round(it.mean(axis=0)[:5], 2)), np.bincount(it.ravel())[:12])
