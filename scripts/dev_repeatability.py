"""Development: repeatability of the large-dimension path (GPU run vs its own repeats and vs the oracle)."""
import sys, copy, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from altro_mpc_icra2021_b200 import solver as S
from altro_mpc_icra2021_b200.problems import mpc, random_linear
from oracle.oracle import OracleProblem
B, K, R = int(os.environ.get("B", "296")), 4, int(os.environ.get("R", "8"))
for n, m in [tuple(int(v) for v in p.split("x")) for p in os.environ.get("POINTS", "200x25,128x32").split(",")]:
    prob, Xt, Ut, ks = random_linear.mpc_problem(n, m, 21, batch=B, seed=500 + n + m)
    opts = random_linear.mpc_options()
    noise = mpc.rng_for(n, m).standard_normal((K, B, n))
    runs = []
    for r in range(R):
        pg = copy.deepcopy(prob)
        sv = S.ALTROSolver(pg, opts); sv.set_track(Xt, Ut, ks.copy()); sv.set_noise_model(1, 0.01, 0.0); sv.set_noise_bank(noise)
        sv.solve(); rg = sv.mpc_run(K); rg["X"] = pg.X.copy(); runs.append(rg); sv.close()
    keys = ("iterations", "ls_trials", "status", "cost", "c_max", "x0", "u0", "X")
    for r in range(1, R):
        diff = [k for k in keys if not np.array_equal(runs[0][k], runs[r][k])]
        if diff:
            d = np.argwhere(runs[0]["iterations"] != runs[r]["iterations"])
            print(n, m, "GPU repeat", r, "differs in", diff, "iteration diffs at (step, inst):", d[:5].tolist())
            for k in diff:
                w = np.argwhere(runs[0][k] != runs[r][k])
                print("   ", k, "count", len(w), "first", w[:6].tolist(), [(float(runs[0][k][tuple(x)]), float(runs[r][k][tuple(x)])) for x in w[:4]])
    if os.environ.get("CPU"):
        pc = copy.deepcopy(prob); op = OracleProblem(pc); op.solve(opts, 16)
        ro = op.mpc_run(opts, K, noise, (1, 0.01, 0.0), (Xt, Ut), ks.copy(), True, 16)
        print(n, m, "oracle == GPU run 0:", all(np.array_equal(runs[0][k], ro[k]) for k in ro))
    print(n, m, "done", R, "GPU repeats", flush=True)
