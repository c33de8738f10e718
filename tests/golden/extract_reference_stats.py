"""Recovers the iteration counts and solve times the reference SAVED from its own runs of the real Altro.jl, and
writes them to tests/golden/reference_stats.json (run here, in the build container: /root/reference does not exist
on the GPU box).

    python -m tests.golden.extract_reference_stats

Sources (written by the reference's benchmark scripts, read-only):
  benchmarks/grasp_optimization/grasp_benchmark_data.jld2   grasp_benchmark.jl:88  <- run_grasp_mpc (grasp_mpc.jl:93-94):
        per MPC step `altro.stats.tsolve` (ms) and `iterations(altro)`; 3 comparison solvers x Ns = 11,21,31,41,51
  horizon_comp.jld2 / state_dim_comp.jld2 / control_dim_comp.jld2   run_random_linear.jl:125,139,153 <- run_MPC
        (random_linear_problem.jl:171-173): `iterations(altro)` and the median of benchmark_solve!, 100 steps each

JLD2 is an HDF5 dialect and no HDF5 reader is installed, so the arrays are found by a raw scan: a run of >= 90
little-endian Float64 (or Int64) values that are all small integers is an iteration array; the Float64 run of twice
that length behind it that holds plausible milliseconds is the [altro | other solver] time matrix (column-major).
These are the only numbers about the solve path that the reference holds; they pin STATISTICS, not trajectories.
"""
import json
import os

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def runs(buf, dtype, pred, minlen):
    out = []
    for off in range(8):
        bb = buf[off:]
        a = np.frombuffer(bb[:len(bb) // 8 * 8], dtype=dtype)
        with np.errstate(all="ignore"):
            idx = np.flatnonzero(pred(a))
        if idx.size == 0:
            continue
        for s in np.split(idx, np.flatnonzero(np.diff(idx) != 1) + 1):
            if len(s) >= minlen:
                out.append((off + int(s[0]) * 8, a[s].copy()))
    return sorted(out, key=lambda t: t[0])


def grasp():
    buf = open(os.path.join(REF, "benchmarks/grasp_optimization/grasp_benchmark_data.jld2"), "rb").read()
    its = runs(buf, "<f8", lambda a: (a == np.round(a)) & (a >= 1) & (a <= 300), 150)
    tms = runs(buf, "<f8", lambda a: (a > 0.01) & (a < 500) & (a != np.round(a)), 150)
    out = []
    for k, (off, v) in enumerate(its):
        L = len(v)
        nxt = its[k + 1][0] if k + 1 < len(its) else len(buf)
        # the time matrix is the last 2L-long float run before the next result dict
        cand = [(o, t) for o, t in tms if off < o < nxt and len(t) == 2 * L]
        t = cand[-1][1]
        out.append({"solver_index": k // 5, "N_mpc": 251 - L, "iterations": v.astype(int).tolist(),
                    "altro_ms": [round(float(x), 6) for x in t[:L]]})
    assert len(out) == 15 and sorted({r["N_mpc"] for r in out}) == [11, 21, 31, 41, 51]
    return out


def random_linear():
    out = {}
    for name in ("horizon_comp", "state_dim_comp", "control_dim_comp"):
        buf = open(os.path.join(REF, name + ".jld2"), "rb").read()
        rows = []
        for off, v in runs(buf, "<i8", lambda a: (a >= 1) & (a <= 1000), 100):
            v = v[:100]
            if v.max() <= 30:  # ALTRO iterations of the 100 MPC steps (the next 100 are OSQP's: 25/50/75/100)
                rows.append({"offset": off, "iterations": v.tolist()})
        out[name] = rows
    return out


def main():
    data = {"source": "raw scan of the reference's saved JLD2 results, see extract_reference_stats.py",
            "grasp": grasp(), "random_linear": random_linear()}
    with open(os.path.join(HERE, "reference_stats.json"), "w") as f:
        json.dump(data, f, separators=(",", ":"))
    g = np.concatenate([r["iterations"] for r in data["grasp"]])
    print("grasp:", len(data["grasp"]), "runs,", g.size, "solves, mean %.3f median %d min %d max %d" % (
        g.mean(), np.median(g), g.min(), g.max()), np.bincount(g)[:12])
    for k, rows in data["random_linear"].items():
        print(k, [np.bincount(r["iterations"])[:7].tolist() for r in rows])


if __name__ == "__main__":
    main()
