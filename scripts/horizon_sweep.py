"""Grasp and flexible-satellite horizon sweeps at 8192 instances (BASELINE.json configs[3]; grasp_benchmark.jl sweeps
N = 11..51, flexible_sat_mpc.jl uses N = 80): GPU closed-loop run vs the CPU oracle on the same batch, with a
bit-parity check on every point.  Prints a markdown table."""
import os, sys, copy, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from altro_mpc_icra2021_b200 import solver as S
from altro_mpc_icra2021_b200.problems import mpc, grasp, flexsat
from oracle.oracle import OracleProblem

B = int(os.environ.get("B", "8192")); K = int(os.environ.get("K", "20"))
nthreads = len(os.sched_getaffinity(0))
print("| family | N | batch | threads/inst | smem/inst KB | CTAs/SM | GPU solves/s | p50 us | CPU solves/s (%d cores) | speed-up | iters | bit-identical |" % nthreads)
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
cold = grasp.cold_problem()
cs = S.ALTROSolver(cold, grasp.cold_options()); cs.solve()
assert cs.stats.status[0] == 1
Xt, Ut = cold.X[0].copy(), cold.U[0].copy()
points = [("grasp", N) for N in (11, 21, 31, 41, 51)] + [("flexsat", N) for N in (40, 80, 120)]
for fam, N in points:
    if fam == "grasp":
        prob, ks = grasp.mpc_problem(cold, Xt, Ut, N, batch=B, seed=77)
        opts, track, model, shift = grasp.mpc_options(), (Xt, Ut), (1, 0.01, 0.0), True
    else:
        prob, ks = flexsat.mpc_problem(N, batch=B, seed=78), np.zeros(B, dtype=np.int64)
        opts, track, model, shift = flexsat.mpc_options(), None, (0, 2e-4, 0.0), False
    steps = K if fam == "grasp" else max(4, K // 4)  # the satellite's CPU side is 10x slower per solve
    pg = copy.deepcopy(prob)
    sv = S.ALTROSolver(pg, opts)
    if track is not None: sv.set_track(track[0], track[1], ks)
    info = sv.launch_info()
    noise = mpc.rng_for(N, 9).standard_normal((steps, B, prob.n))
    sv.set_noise_model(*model); sv.set_noise_bank(noise)
    sv.solve(); rg = sv.mpc_run(steps, shift=shift)
    op = OracleProblem(prob); op.solve(opts, nthreads)
    t0 = time.perf_counter(); ro = op.mpc_run(opts, steps, noise, model, track, ks, shift, nthreads); tc = time.perf_counter() - t0
    same = all(np.array_equal(rg[k], ro[k]) for k in ro) and np.array_equal(pg.X, prob.X)
    gps, cps = B * steps / (rg["device_ms"] * 1e-3), B * steps / tc
    print(f"| {fam} | {N} | {B} | {info['threads_per_instance']} | {info['smem_bytes']/1024:.1f} | {info['ctas_per_sm']} | {gps:.0f} | {np.median(rg['t_us']):.0f} | {cps:.0f} | {gps/cps:.1f}x | {rg['iterations'].mean():.2f} | {same} |", flush=True)
    sv.close()
