"""Development: fused-run timing and per-phase cycle breakdown (run on a GPU box)."""
import os, sys, copy, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from altro_mpc_icra2021_b200 import solver as S
from altro_mpc_icra2021_b200.problems import mpc
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "rocket"
B = int(os.environ.get("B", "4096")); K = int(os.environ.get("K", "20")); T = int(os.environ.get("T", "0"))
wl = bench.Workload(name, B, 0xA1722, lambda p, o: S.ALTROSolver(p, o))
sv = S.ALTROSolver(wl.prob, wl.opts, threads_per_instance=T)
if wl.track is not None: sv.set_track(wl.track[0], wl.track[1], wl.k)
print(name, sv.launch_info())
sv.set_noise_model(*wl.noise_model); sv.set_noise_bank(wl.noise_samples(3 * K + 3))
sv.solve()
if wl.qstate is None:
    sv.mpc_run(3, shift=wl.shift)
    sv.phase_cycles(True)
    r = sv.mpc_run(K, shift=wl.shift)
    ph = sv.phase_cycles(False).astype(float)
    ms = r["device_ms"]
    it, ls = r["iterations"], r["ls_trials"]
    tot_i = r["t_us"].sum(axis=0)
    print(f"fused {K} steps: {ms:.2f} ms -> {B*K/ms*1e3:.0f} solves/s; per-solve us p50 {np.median(r['t_us']):.0f} mean {r['t_us'].mean():.0f} max {r['t_us'].max():.0f}; "
          f"per-instance total ms mean {tot_i.mean()/1e3:.2f} max {tot_i.max()/1e3:.2f}; iters {it.mean():.2f} ls {ls.mean():.2f}")
    if os.environ.get("ALTRO_B200_PHASE_DETAIL"):
        nit = ph[:, 7].sum()
        print("bp sub-phase cycles per iteration:", {nm: int(ph[:, i].sum() / nit) for i, nm in enumerate(["P1 SA/SB", "P3 warp0 LDL+solve", "P2 Q", "P3 wait prep", "P4 solve", "P5 T1", "P6 S"])})
        sys.exit(0)
    tot = ph[:, 3].sum()
    names = ["init rollout+cost", "backward(+expand)", "forward", "whole", "expand", "ls rollouts", "ls costs"]
    print("cycle shares:", {nm: round(ph[:, i].sum() / tot, 3) for i, nm in enumerate(names)})
    print("cycles per: bp %.0f  expand %.0f  ls-rollout %.0f  ls-cost %.0f" % (
        (ph[:, 1].sum() - ph[:, 4].sum()) / it.sum(), ph[:, 4].sum() / it.sum(), ph[:, 5].sum() / ls.sum(), ph[:, 6].sum() / ls.sum()))
else:
    sv.phase_cycles(True)
    sv.solve(); st = sv.stats
    ph = sv.phase_cycles(False).astype(float)
    print(f"one solve: {st.tsolve:.2f} ms -> {B/st.tsolve*1e3:.0f} solves/s; us p50 {np.median(st.t_instance_us):.0f} max {st.t_instance_us.max():.0f}; iters {st.iterations.mean():.2f} ls {st.ls_trials.mean():.2f}")
    tot = ph[:, 3].sum(); it, ls = st.iterations, st.ls_trials
    print("cycles per: bp %.0f  expand %.0f  ls-rollout %.0f  ls-cost %.0f" % (
        (ph[:, 1].sum() - ph[:, 4].sum()) / it.sum(), ph[:, 4].sum() / it.sum(), ph[:, 5].sum() / ls.sum(), ph[:, 6].sum() / ls.sum()))
