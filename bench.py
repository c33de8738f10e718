#!/usr/bin/env python
"""Batched-MPC benchmark of the ALTRO hot path (BASELINE.json: batched MPC solves/sec, device-timed).

One *step* = one pass of the hot path over one batch: the warm-started MPC transition of every instance
(plant step with the first control + noise, tracking-reference window, primal + dual shift_fill) followed by
the batched AL-iLQR solve!  -- the body of the reference's MPC loops (random_linear_problem.jl:121-161,
simple_rocket.jl:59-82,163-174, altro_solver.jl:44-72).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload rocket|quadruped|quadruped_soc|random_linear|grasp|flexsat]
  torchrun ... bench.py --gpus N ...      one rank per GPU, 4096 instances per GPU (weak scaling), no collective on
                                          the solve path; NCCL only gathers the per-instance statistics at the end
  python bench.py --impl reference ...    the reference arm: the CPU oracle (a port -- Julia Altro.jl cannot run here)
                                          with all host threads on the same workloads

The headline keys describe configs[1] of BASELINE.json (rocket, 4096 instances per GPU); `workloads` carries the same
measurements for configs[2] (quadruped, 4096 instances per GPU) so that the 1/2/4/8-GPU scaling runs record both
families the target names.  --workload X makes X the headline and drops the secondary workloads.

`value`        : whole-job solves/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks
                 (closed-loop run: K x {transition; solve!} per instance in one launch).
`e2e`          : the same run through the public API with HOST buffers (pinned H2D of the disturbances, D2H of the
                 closed-loop states, controls, statistics, X, U inside the timed region).
`per_step_api` : the reference-shaped usage, one solve! call per MPC step: `value` with device-resident inputs (one
                 transition + one solve launch per step), `e2e` with x0 and the reference window uploaded from host
                 buffers and X, U and the statistics read back EVERY step.
`roofline`     : FP64 (DFMA) roofline of the solve kernel -- algorithmic flops from the per-instance iteration and
                 line-search counts (SURVEY.md 8d formulas) / CUDA-event duration / DFMA peak measured in this run.
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from altro_mpc_icra2021_b200.problem import STATUS_NAMES  # noqa: E402
from altro_mpc_icra2021_b200.problems import flexsat, grasp, mpc, quadruped, random_linear, rocket  # noqa: E402

METRIC = "batched MPC solves/sec (whole box, device-timed)"
UNIT = "solves/s"
L2_FLUSH_BYTES = 256 << 20


# ----------------------------------------------------------------------------- workloads


class Workload:
    """A batch of MPC instances plus everything the per-step update needs."""

    def __init__(self, name, batch, seed, make_solver):
        self.name, self.batch, self.seed = name, batch, seed
        self.track = None
        self.noise_model = (0, 0.0, 0.0)
        self.shift = True
        if name == "rocket":
            # configs[1]: 4096 instances, three second-order cones, warm-started shift MPC, N = 21
            cold = rocket.cold_problem()
            cs = make_solver(cold, rocket.cold_options())
            cs.solve()
            assert cs.stats.status[0] == 1, "cold solve failed"
            Xt, Ut = cold.X[0].copy(), cold.U[0].copy()
            self.prob, self.k = rocket.mpc_problem(cold, Xt, Ut, 21, batch=batch, seed=seed)
            self.opts, self.track, self.noise_model = rocket.mpc_options(), (Xt, Ut), (2, 1e-3, 1e-2)
            self.desc = "rocket_landing: n=6 m=3 N=21, thrust-norm + thrust-angle + glideslope SOC, tracking MPC"
        elif name == "grasp":
            # configs[3]: two-finger grasp, per-knot equality + inequality + 2 SOC, data along shared timelines
            cold = grasp.cold_problem()
            cs = make_solver(cold, grasp.cold_options())
            cs.solve()
            assert cs.stats.status[0] == 1, "cold solve failed"
            Xt, Ut = cold.X[0].copy(), cold.U[0].copy()
            self.prob, self.k = grasp.mpc_problem(cold, Xt, Ut, 21, batch=batch, seed=seed)
            self.opts, self.track, self.noise_model = grasp.mpc_options(), (Xt, Ut), (1, 0.01, 0.0)
            self.desc = "grasp_optimization: n=m=6 N=21, torque-balance eq + max-force ineq + 2 friction SOC, tracking MPC"
        elif name == "random_linear":
            self.prob, Xt, Ut, self.k = random_linear.mpc_problem(12, 6, 21, batch=batch, seed=seed)
            self.opts, self.track, self.noise_model = random_linear.mpc_options(), (Xt, Ut), (1, 0.01, 0.0)
            self.desc = "random_linear_mpc: n=12 m=6 N=21, control bounds, tracking MPC"
        elif name in ("quadruped", "quadruped_soc"):
            # configs[2]: dynamics stored once per gait phase + schedule, so the closed-loop run stays on the device
            self.prob, _ = quadruped.mpc_problem(batch, linearized_friction=(name == "quadruped"), seed=seed,
                                                 gait_slots=512)
            self.opts, self.noise_model = quadruped.mpc_options(), (0, 1e-3, 0.0)
            self.k = np.zeros(batch, dtype=np.int64)
            self.desc = ("quadruped: n=m=12 N=15, per-instance LTV dynamics, "
                         + ("linearised friction pyramid" if name == "quadruped" else "second-order friction cones")
                         + " + fz bounds")
        elif name == "flexsat":
            self.prob, self.opts, self.noise_model = flexsat.mpc_problem(80, batch=batch, seed=seed), flexsat.mpc_options(), (0, 2e-4, 0.0)
            self.k = np.zeros(batch, dtype=np.int64)
            self.shift = False
            self.desc = "flexible_satellite: n=12 m=3 N=80, control bounds, no shift"
        else:
            raise SystemExit(f"unknown workload {name}")
        self.rng = mpc.rng_for(seed, 1234)

    def noise_samples(self, steps):
        return self.rng.standard_normal((steps, self.prob.B, self.prob.n))

    def apply_noise(self, x, z):
        mode, w1, w2 = self.noise_model
        if mode == 0:
            return x + w1 * z
        if mode == 1:
            return x + z * np.abs(x).max(axis=-1, keepdims=True) * w1
        h = x.shape[-1] // 2
        sp = np.linalg.norm(x[..., :h], axis=-1, keepdims=True) * w1
        sv = np.linalg.norm(x[..., h:], axis=-1, keepdims=True) * w2
        return x + z * np.concatenate([np.broadcast_to(sp, x[..., :h].shape), np.broadcast_to(sv, x[..., h:].shape)], -1)


def flops_model(prob, iters, outer, trials):
    """flops/solve = I (N-1)(F_bp + F_ex) + (T + rollouts) (N-1) F_fp, summed over instances."""
    n, m, N = prob.n, prob.m, prob.N
    f_bp = 4 * n ** 3 + 8 * n * n * m + 6 * n * m * m + m ** 3 / 3 + (2 * n * n + 6 * n * m + 4 * m * m)
    f_ex = 2.0 * (n + m) * (N - 1)
    f_fp_con = 0.0
    for c in prob.constraints.flat:
        nk = c.k1 - c.k0
        f_ex += nk * (2 * c.p * c.w * c.w + 2 * c.p * c.w)
        f_fp_con += nk * (2 * c.p * c.w + 3 * c.p)
    f_fp = (N - 1) * (2 * n * n + 4 * n * m + 2 * n + 4 * (n + m)) + f_fp_con
    I, O, T = (np.asarray(a, dtype=np.float64) for a in (iters, outer, trials))
    return float(np.sum(I * ((N - 1) * f_bp + f_ex) + (T + O) * f_fp))


def algorithmic_bytes(prob, P):
    """Compulsory HBM traffic per solve: problem in, solution out."""
    n, m, N = prob.n, prob.m, prob.N
    dyn = (n * n + n * m + n) * (N - 1) if prob.model.per_instance else 0
    cdat = sum((c.G.size + c.h.size) // prob.B for c in prob.constraints.flat if c.per_instance)
    rd = dyn + (n + m) * N + n + m * (N - 1) + P + cdat
    wr = n * N + m * (N - 1) + P + 8
    return 8.0 * (rd + wr)


# ----------------------------------------------------------------------------- helpers


class ClockSampler(threading.Thread):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx, self.rows, self.stop_flag = gpu_index, [], threading.Event()
        # NVML in-process (one handle, opened before the timed region): spawning nvidia-smi while the fused
        # launch runs stalled the device for tens of ms per call and made the timed run 20 % slower than its replays
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml, self.handle = pynvml, pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        except Exception:
            self.nvml = None

    def sample_nvml(self):
        nv, h = self.nvml, self.handle
        sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        try:
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        flags = ["Active" if mask & bits[k] else "Not Active"
                 for k in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")]
        try:
            pw = nv.nvmlDeviceGetPowerUsage(h) / 1e3
        except Exception:
            pw = 0.0
        self.rows.append([str(self.idx), str(sm), str(mx), f"{pw:.1f}"] + flags)

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nvml is not None:
                    self.sample_nvml()
                else:
                    out = subprocess.run(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    for line in out.strip().splitlines():
                        self.rows.append([x.strip() for x in line.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.05 if self.nvml is not None else 0.2)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def oracle_arm(wl: Workload, steps, warmup, nthreads):
    """Times the CPU oracle (port of the reference algorithm) on the same workload with `nthreads` host threads.
    The one place besides tests/ and smoke() that executes oracle/ -- as the measured CPU baseline only.
    Like the GPU arm, every instance runs its own closed-loop MPC loop (no lock-step between steps); instances are
    handed to the threads dynamically (the analogue of Threads.@threads over the batch)."""
    from oracle.oracle import OracleProblem

    prob = copy.deepcopy(wl.prob)
    op = OracleProblem(prob)
    op.solve(wl.opts, nthreads=nthreads)  # initial solve (random_linear_problem.jl:113), not timed
    k = wl.k.copy()
    if warmup:
        op.mpc_run(wl.opts, warmup, wl.noise_samples(warmup), wl.noise_model, wl.track, k, wl.shift, nthreads)
        k = k + warmup
    zs = wl.noise_samples(steps)
    t0 = time.perf_counter()
    r = op.mpc_run(wl.opts, steps, zs, wl.noise_model, wl.track, k, wl.shift, nthreads)
    total = time.perf_counter() - t0
    return {"value": prob.B * steps / total, "seconds": total, "iters_mean": float(r["iterations"].mean()),
            "success": float(np.mean(r["status"] == 1))}


# ----------------------------------------------------------------------------- one workload on the GPU


def bench_gpu_workload(name, args, K, W, rank, world, local, stream, dev, flush, peaks, make_solver, barrier, S, sharding):
    """All GPU measurements of one workload; returns the dict that becomes the headline line or a `workloads` entry."""
    import torch

    seed = 0xA1720 + 2 + 1000 * rank
    wl = Workload(name, args.batch, seed, make_solver)
    prob, B = wl.prob, wl.batch
    sv = make_solver(prob, wl.opts, threads_per_instance=args.threads_per_instance, pin=True)
    if wl.track is not None:
        sv.set_track(wl.track[0], wl.track[1], wl.k)  # before the first launch: the window is then read from the track
    info = sv.launch_info()
    sv.set_noise_model(*wl.noise_model)
    zs_all = wl.noise_samples(max(W, 1) + K)
    sv.set_noise_bank(zs_all)

    # ---- initial solve + warm-up (untimed)
    sv.solve()
    init_ok = float(np.mean(sv.stats.status == 1))
    if W:
        sv.mpc_run(W, shift=wl.shift)
    sv.reserve_steps(K)  # log buffers sized before the timed region (no cudaFree/cudaMalloc inside it)
    sv.snapshot()  # the per-step and e2e phases replay the same K steps from this state
    k_snap = wl.k.copy() + W
    barrier()
    # ---- timed region: K steps in one closed-loop launch, CUDA events on the launching stream
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    flush.zero_()  # L2 flush before the timed launch (256 MB write, outside the event pair)
    e0.record(stream)
    sv.mpc_run(K, shift=wl.shift, fetch=False)
    e1.record(stream)
    rr = sv.run_results(K)  # D2H after the event pair: needed for the flop model, not timed
    barrier()
    wall = time.perf_counter() - t_wall
    sampler.stop_flag.set()
    sampler.join()
    kern_s = rr["device_ms"] * 1e-3
    step_s = e0.elapsed_time(e1) * 1e-3
    total_s = sharding.max_over_ranks(step_s)
    flops = flops_model(prob, rr["iterations"], rr["iterations_outer"], rr["ls_trials"])
    status_counts = {STATUS_NAMES[int(k)]: int(v) for k, v in zip(*np.unique(rr["status"], return_counts=True))}
    gathered = sharding.gather_stats({"iterations": rr["iterations"][-1], "status": rr["status"][-1]})

    # ---- per-step API, device-resident inputs: one transition + one solve launch per MPC step
    sv.restore()
    evl = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for st in range(K):
        flush.zero_()
        evl[st][0].record(stream)
        sv.mpc_transition(None, shift=wl.shift)
        sv.solve(fetch=False)
        evl[st][1].record(stream)
    barrier()
    t_lock = sharding.max_over_ranks(sum(a.elapsed_time(b) for a, b in evl) * 1e-3)
    sv.fetch()
    per_step = {"value": B * world * K / t_lock, "unit": UNIT, "ms_per_step": 1e3 * t_lock / K,
                "note": "one altro_mpc_transition + one altro_solve launch per MPC step (solve! once per step, as the "
                        "reference's loops do); every step waits for its slowest instance"}

    e2e = None
    if not args.no_e2e:
        # ---- e2e of the closed-loop run: pinned H2D of the K x B x n disturbances, the run, D2H of everything
        barrier()
        sv.restore()
        zs = np.ascontiguousarray(zs_all[W:W + K])  # the same disturbances the device-timed run consumed
        sv.reserve_host_results(K)  # result buffers of a caller that reads results every run: allocated and page-locked once
        sv.lib.altro_host_register(S._p(zs), zs.nbytes)
        h2d = zs.nbytes / K
        d2h = (K * B * (prob.n + prob.m) * 8 + K * B * (4 * 4 + 2 * 8 + 8) + prob.X.nbytes + prob.U.nbytes) / K
        t0 = time.perf_counter()
        sv.set_noise_bank(zs)  # H2D of this run's inputs
        sv.mpc_run(K, shift=wl.shift, fetch=True, reuse_buffers=True)  # run + D2H of closed-loop states, controls, statistics, X, U (page-locked result buffers kept by the solver)
        t_e2e = time.perf_counter() - t0
        sv.lib.altro_host_unregister(S._p(zs))
        barrier()
        t_e2e = sharding.max_over_ranks(t_e2e)
        e2e = {"value": B * world * K / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * t_e2e / K}
        # ---- e2e of the per-step API: every step the caller uploads its inputs from host buffers and reads the
        #      solution and the statistics back (set_initial_state! / update_trajectory! / solve! / states / stats)
        sv.restore()
        sv.fetch()  # host mirror of X, U at the snapshot
        wl.k = k_snap.copy()
        prob.kidx[...] = wl.k if wl.track is not None or name == "grasp" else prob.kidx
        t_api, h2d_s, d2h_s = 0.0, 0, prob.X.nbytes + prob.U.nbytes + B * (4 * 4 + 4 * 8 + 8)
        for st in range(K):
            z = zs_all[W + st]
            if wl.track is not None and prob.model.sched is None:
                # host side of the reference's loop (plant step + noise, window of the tracked trajectory): the caller's work
                x0 = wl.apply_noise(prob.X[:, 1, :], z)
                wl.k = wl.k + 1
                Xr, Ur = mpc.window_reference(wl.track[0], wl.track[1], wl.k, prob.N)
                t0 = time.perf_counter()
                prob.kidx += 1
                if name == "grasp":
                    sv.set_track_index(prob.kidx)
                prob.set_initial_state(x0)      # H2D x0
                prob.update_trajectory(Xr, Ur)   # H2D reference window
                if wl.shift:
                    sv.shift_fill(True, True)
                sv.solve(fetch=True)             # solve! + D2H X, U, statistics
                t_api += time.perf_counter() - t0
                h2d_s = x0.nbytes + Xr.nbytes + Ur.nbytes
            else:
                # gait-scheduled / untracked workloads: the step's disturbance goes up, the device does the plant step
                zz = np.ascontiguousarray(z)
                t0 = time.perf_counter()
                sv.mpc_transition(zz, shift=wl.shift)  # H2D disturbance
                sv.solve(fetch=True)                   # solve! + D2H X, U, statistics
                t_api += time.perf_counter() - t0
                h2d_s = zz.nbytes
        barrier()
        t_api = sharding.max_over_ranks(t_api)
        per_step["e2e"] = {"value": B * world * K / t_api, "unit": UNIT, "h2d_bytes_per_step": int(h2d_s),
                           "d2h_bytes_per_step": int(d2h_s), "ms_per_step": 1e3 * t_api / K}

    achieved = flops / kern_s / 1e12
    alg_bytes = algorithmic_bytes(prob, sv.P)
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        hbm_src = "MEASURED_PEAKS.json"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic = None
    tnote = "no ncu capture for this workload / batch"
    tfile = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(tfile):  # DRAM bytes per launch from the committed ncu --set full capture of this kernel
        tj = json.load(open(tfile)).get(name)
        if tj and tj.get("batch") == B and tj.get("kernel") == info.get("kernel"):
            traffic = tj["per_launch_fixed"] + tj["per_step"] * K
            tnote = tj["source"]
    res = {
        "value": B * world * K / total_s, "unit": UNIT, "ms_per_step": 1e3 * total_s / K,
        "config": {"workload": wl.desc, "instances_per_gpu": B, "global_batch": B * world,
                   "parallelism": f"instance-sharded x{world}, no collective on the solve path",
                   "l2": f"L2 flushed before every timed launch ({L2_FLUSH_BYTES >> 20} MB write)",
                   "launch": info,
                   "step": "closed-loop MPC run: per instance K x {transition + shift_fill + AL-iLQR solve} in one "
                           "launch, instances advance independently"},
        "p50_solve_us": float(np.median(rr["t_us"])), "max_solve_us": float(rr["t_us"].max()),
        "iters_mean": float(rr["iterations"].mean()), "iters_max": int(rr["iterations"].max()),
        "ls_trials_mean": float(rr["ls_trials"].mean()),
        "success": float(np.mean(rr["status"] == 1)), "status_counts": status_counts,
        "success_all_ranks_last_step": float(np.mean(gathered["status"] == 1)),
        "iters_all_ranks_last_step": float(np.mean(gathered["iterations"])),
        "init_success": init_ok,
        "gpu_launches": 2,  # altro_{lane,solve}_kernel + advance_kidx_kernel inside the timed region
        "steps_per_launch": K,
        "clocks": sampler.summary(),
        # "tensor" = the compute roofline of the two the contract names; here it is the FP64 one (DFMA, DMMA m8n8k4 for
        # the CTA kernels).  The kernels are latency / issue bound, not pipe bound: see `limiter`.
        "roofline": {"bound": "tensor", "pipe": "fp64", "achieved": achieved, "peak": peaks["dfma_tflops"],
                     "unit": "TFLOP/s", "frac": achieved / peaks["dfma_tflops"], "traffic": traffic,
                     "traffic_note": tnote,
                     "limiter": "dependent-instruction latency and issue slots (profiles/r2_summary.md), not the FP64 pipe",
                     "peak_source": "DFMA stream measured live by altro_measure_peaks in this run "
                                    "(MEASURED_PEAKS.json has no FP64 figure); DMMA m8n8k4 peak "
                                    f"{peaks['dmma_tflops']:.1f} TFLOP/s",
                     "kernel": "altro_lane_kernel" if info.get("kernel") == "lane" else "altro_solve_kernel",
                     "kernel_ms_per_step": 1e3 * kern_s / K,
                     "kernel_share_of_step": kern_s / step_s,
                     "flops_per_step": flops / K,
                     "hbm": {"achieved": alg_bytes * B * K / kern_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": alg_bytes * B * K / kern_s / 1e9 / hbm_peak, "peak_source": hbm_src,
                             "algorithmic_bytes_per_solve": alg_bytes}},
        "wall_s_timed_region": wall,
        "per_step_api": per_step,
        "lockstep": {k: per_step[k] for k in ("value", "unit", "ms_per_step", "note")},
    }
    if e2e:
        res["e2e"] = e2e
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        nt = host_threads()
        cpu_steps = args.cpu_steps
        if cpu_steps <= 0:  # probe 4 steps, then size the sample for about 20 core-seconds (bounded by K)
            probe = oracle_arm(Workload(name, args.batch, seed, make_solver), 4, 1, nt)
            cpu_steps = int(min(max(4, round(20.0 / nt * probe["value"] / B)), max(K, 4)))
        r = oracle_arm(Workload(name, args.batch, seed, make_solver), cpu_steps, 1, nt)
        res["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": nt, "kind": "port",
                               "sample": f"{cpu_steps} MPC steps x {B} instances of the same workload, CPU oracle "
                                         f"(oracle/altro_oracle.c, pthreads over instances), {r['seconds']:.2f} s",
                               "iters_mean": r["iters_mean"]}
    sv.close()
    return res


def reference_workload(name, args, K, W, seed):
    """The reference arm of one workload: the CPU oracle (port) on all host threads, whole batch per step."""
    from oracle.oracle import OracleProblem

    class _OS:
        def __init__(self, prob, opts):
            self.prob, self.opts, self.op = prob, opts, OracleProblem(prob)

        def solve(self):
            self.stats = self.op.solve(self.opts, nthreads=1)
            return self

    wl = Workload(name, args.batch, seed, _OS)
    nt = host_threads()
    r = oracle_arm(wl, K, W, nt)
    return {"value": r["value"], "unit": UNIT, "ms_per_step": 1e3 * r["seconds"] / K,
            "config": {"workload": wl.desc, "batch_per_step": wl.batch, "parallelism": f"{nt} host threads",
                       "note": "the CPU arm solves `batch_per_step` instances per step whatever --gpus says; the GPU arm "
                               "solves that many PER GPU (weak scaling): compare throughputs"},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": nt, "kind": "port",
                             "sample": f"{K} MPC steps x {wl.batch} instances (whole batch), CPU oracle "
                                       f"oracle/altro_oracle.c, pthreads over instances"},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "iters_mean": r["iters_mean"], "success": r["success"]}


# ----------------------------------------------------------------------------- main


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100, help="MPC steps timed (the reference's loops run 100-280)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, help="headline workload (default rocket, with quadruped as secondary)")
    ap.add_argument("--secondary", default=None, help="comma-separated secondary workloads (default: quadruped)")
    ap.add_argument("--batch", type=int, default=4096, help="instances per GPU")
    ap.add_argument("--threads-per-instance", type=int, default=0)
    ap.add_argument("--cpu-steps", type=int, default=0,
                    help="MPC steps of the CPU-baseline sample (0 = sized for about 20 core-seconds of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    K, W = args.steps, max(args.warmup, 0)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    headline = args.workload or os.environ.get("ALTRO_BENCH_WORKLOAD") or "rocket"
    if args.secondary is not None:
        secondary = [w for w in args.secondary.split(",") if w and w != "none"]
    else:
        secondary = ["quadruped"] if args.workload is None and "ALTRO_BENCH_WORKLOAD" not in os.environ else []
    base = {"metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic"}

    if args.impl == "reference":
        if rank != 0:
            return  # rank 0 alone runs the CPU arm
        seed = 0xA1720 + 2
        line = dict(base, impl="reference")
        line.update(reference_workload(headline, args, K, W, seed))
        if secondary:
            line["workloads"] = {w: reference_workload(w, args, K, W, seed) for w in secondary}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    from altro_mpc_icra2021_b200 import sharding
    from altro_mpc_icra2021_b200 import solver as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    sharding.init_distributed("nccl" if world > 1 else None)
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream()

    def make_solver(prob, opts, **kw):
        return S.ALTROSolver(prob, opts, device=local, stream=stream.cuda_stream, **kw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks = S.measure_peaks(local)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    results = {}
    for name in [headline] + secondary:
        results[name] = bench_gpu_workload(name, args, K, W, rank, world, local, stream, dev, flush, peaks, make_solver,
                                           barrier, S, sharding)
    if rank == 0:
        line = dict(base)
        line.update(results[headline])
        line["n_gpus"] = world
        if secondary:
            line["workloads"] = {w: results[w] for w in secondary}
            line["gpu_launches"] = sum(results[w]["gpu_launches"] for w in results)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
