"""Random-linear state / control dimension sweep (run_random_linear.jl:128-152): GPU closed-loop run vs the CPU
oracle on the same batch, with a bit-parity check on every point.  Prints a markdown table."""
import os, sys, copy, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from altro_mpc_icra2021_b200 import solver as S
from altro_mpc_icra2021_b200.problems import mpc, random_linear
from oracle.oracle import OracleProblem

points = [(2, 2), (15, 2), (25, 2), (35, 2), (45, 2), (55, 2), (30, 2), (30, 6), (30, 10), (30, 15), (30, 20), (30, 25),
          (64, 16), (100, 25), (128, 32), (200, 25)]
if os.environ.get("POINTS"):
    points = [tuple(int(v) for v in p.split("x")) for p in os.environ["POINTS"].split(",")]
B0 = int(os.environ.get("B", "1024")); K0 = int(os.environ.get("K", "10"))
nthreads = len(os.sched_getaffinity(0))
print("| n | m | N | batch | threads/inst | smem/inst KB | GPU solves/s | p50 us | CPU solves/s (%d cores) | speed-up | iters | bit-identical |" % nthreads)
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for n, m in points:
    B, K = (B0, K0) if n < 64 else (min(B0, 296), min(K0, 4))  # large points: the CPU side is O(n^3) per knot
    try:
        prob, Xt, Ut, ks = random_linear.mpc_problem(n, m, 21, batch=B, seed=500 + n + m)
        opts = random_linear.mpc_options()
        pg = copy.deepcopy(prob)
        sv = S.ALTROSolver(pg, opts)
        sv.set_track(Xt, Ut, ks); info = sv.launch_info()
        noise = mpc.rng_for(n, m).standard_normal((K, B, n))
        sv.set_noise_model(1, 0.01, 0.0); sv.set_noise_bank(noise)
        sv.solve(); rg = sv.mpc_run(K)
        op = OracleProblem(prob); op.solve(opts, nthreads)
        t0 = time.perf_counter(); ro = op.mpc_run(opts, K, noise, (1, 0.01, 0.0), (Xt, Ut), ks, True, nthreads); tc = time.perf_counter() - t0
        same = all(np.array_equal(rg[k], ro[k]) for k in ro) and np.array_equal(pg.X, prob.X)
        gps, cps = B * K / (rg["device_ms"] * 1e-3), B * K / tc
        print(f"| {n} | {m} | 21 | {B} | {info['threads_per_instance']} | {info['smem_bytes']/1024:.1f} | {gps:.0f} | {np.median(rg['t_us']):.0f} | {cps:.0f} | {gps/cps:.1f}x | {rg['iterations'].mean():.2f} | {same} |", flush=True)
        sv.close()
    except Exception as e:
        print(f"| {n} | {m} | 21 | {B} | - | - | unsupported: {str(e)[:90]} | | | | | |", flush=True)
