#!/usr/bin/env python
"""Batched-MPC benchmark of the ALTRO hot path (BASELINE.json: batched MPC solves/sec, device-timed).

One *step* = one pass of the hot path over one batch: the warm-started MPC transition of every instance
(plant step with the first control + noise, tracking-reference window, primal + dual shift_fill) followed by
the batched AL-iLQR solve!  -- the body of the reference's MPC loops (random_linear_problem.jl:121-161,
simple_rocket.jl:59-82,163-174, altro_solver.jl:44-72).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload rocket|quadruped|quadruped_soc|random_linear|flexsat]
  torchrun ... bench.py --gpus N ...      one rank per GPU, 4096 instances per GPU (weak scaling), no collective on
                                          the solve path; NCCL only gathers the per-instance statistics at the end
  python bench.py --impl reference ...    the reference arm: the CPU oracle (a port -- Julia Altro.jl cannot run here)
                                          with all host threads on the same workload

`value`  : whole-job solves/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks.
`e2e`    : same metric through the public API with HOST buffers (pinned H2D of x0 + reference, D2H of X, U, stats).
`roofline`: FP64 (DFMA) roofline of the solve kernel -- algorithmic flops from the per-instance iteration and
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
        self.qstate = None
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

    def host_advance(self, prob, solver, z):
        """The reference's between-solve update done by the caller on HOST buffers (e2e and CPU arms)."""
        prob.kidx += 1
        if self.qstate is not None:  # quadruped control tick: new contact schedule -> new B_k, plant step + noise
            quadruped.advance(prob, self.qstate, self.rng)
            solver.shift_fill(True, True)
            return
        x0 = self.apply_noise(prob.X[:, 1, :], z)  # x_1 of the last solution = plant step with its first control
        prob.set_initial_state(x0)
        if self.track is not None:
            self.k = self.k + 1
            prob.update_trajectory(*mpc.window_reference(self.track[0], self.track[1], self.k, prob.N))
        if self.shift:
            solver.shift_fill(True, True)


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
    if wl.qstate is not None:  # quadruped: the contact schedule is rebuilt on the host every tick (lock-step)
        class _S:
            def shift_fill(self, primal=True, dual=True):
                op.shift_fill(primal, dual)

        total, iters, status = 0.0, [], []
        for st in range(steps + warmup):
            wl.host_advance(prob, _S(), None)
            t0 = time.perf_counter()
            r = op.solve(wl.opts, nthreads=nthreads)
            if st >= warmup:
                total += time.perf_counter() - t0
                iters.append(r.iterations.mean())
                status.append(np.mean(r.status == 1))
        return {"value": prob.B * steps / total, "seconds": total, "iters_mean": float(np.mean(iters)),
                "success": float(np.mean(status))}
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


# ----------------------------------------------------------------------------- main


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100, help="MPC steps timed (the reference's loops run 100-280)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rocket")
    ap.add_argument("--batch", type=int, default=4096, help="instances per GPU")
    ap.add_argument("--threads-per-instance", type=int, default=0)
    ap.add_argument("--cpu-steps", type=int, default=0,
                    help="MPC steps of the CPU-baseline sample (0 = sized for about 20 core-seconds of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    K, W = args.steps, max(args.warmup, 0)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    seed = 0xA1720 + 2 + 1000 * rank

    if args.impl == "reference":
        if rank != 0:
            return  # rank 0 alone runs the CPU arm
        wl_builder_solver = None
        try:
            import torch

            have_gpu = torch.cuda.is_available()
        except Exception:
            have_gpu = False
        # the rocket workload needs the cold-solved track; the CPU arm computes it with the oracle itself
        from oracle.oracle import OracleProblem

        class _OS:
            def __init__(self, prob, opts):
                self.prob, self.opts, self.op = prob, opts, OracleProblem(prob)

            def solve(self):
                self.stats = self.op.solve(self.opts, nthreads=1)
                return self

        wl = Workload(args.workload, args.batch, seed, _OS)
        nt = host_threads()
        r = oracle_arm(wl, K, W, nt)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": K, "warmup": W, "ms_per_step": 1e3 * r["seconds"] / K, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl.desc, "batch_per_step": wl.batch, "parallelism": f"{nt} host threads"},
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": nt, "kind": "port",
                                 "sample": f"{K} MPC steps x {wl.batch} instances (whole batch), CPU oracle "
                                           f"oracle/altro_oracle.c, pthreads over instances"},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "iters_mean": r["iters_mean"], "success": r["success"]}
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

    wl = Workload(args.workload, args.batch, seed, make_solver)
    prob, B = wl.prob, wl.batch
    sv = make_solver(prob, wl.opts, threads_per_instance=args.threads_per_instance, pin=True)
    if wl.track is not None:
        sv.set_track(wl.track[0], wl.track[1], wl.k)  # before the first launch: the window is then read from the track
    info = sv.launch_info()
    peaks = S.measure_peaks(local)
    sv.set_noise_model(*wl.noise_model)
    zs_all = wl.noise_samples(max(W, 1) + K)
    sv.set_noise_bank(zs_all)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    qrng = mpc.rng_for(seed, 77)
    fused = wl.qstate is None  # closed-loop run in one launch; quadruped rebuilds B_k on the host every tick
    S_launch = K if fused else 1

    def q_tick(fetch):
        quadruped.advance(prob, wl.qstate, qrng)
        sv.upload()
        sv.shift_fill(True, True)
        sv.solve(fetch=fetch)

    # ---- initial solve + warm-up (untimed)
    sv.solve()
    init_ok = float(np.mean(sv.stats.status == 1))
    if fused:
        if W:
            sv.mpc_run(W, shift=wl.shift)
    else:
        for _ in range(W):
            q_tick(True)
    if fused:
        sv.reserve_steps(K)  # log buffers sized before the timed region (no cudaFree/cudaMalloc inside it)
        sv.snapshot()  # the lock-step and e2e phases replay the same K steps from this state
    barrier()
    # ---- timed region: K steps, CUDA events on the launching stream around every launch
    sampler = ClockSampler(local)
    sampler.start()
    n_launch = K // S_launch
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_launch)]
    kern_ms, flops, per_step = [], 0.0, []
    t_wall = time.perf_counter()
    for li in range(n_launch):
        flush.zero_()  # L2 flush between timed launches (256 MB write, outside the event pair)
        if fused:
            ev[li][0].record(stream)
            sv.mpc_run(S_launch, shift=wl.shift, fetch=False)
            ev[li][1].record(stream)
            rr = sv.run_results(S_launch)  # D2H after the event pair: needed for the flop model, not timed
            kern_ms.append(rr["device_ms"])
            flops += flops_model(prob, rr["iterations"], rr["iterations_outer"], rr["ls_trials"])
            for st in range(S_launch):
                per_step.append((float(rr["iterations"][st].mean()), float(rr["ls_trials"][st].mean()),
                                 float(np.mean(rr["status"][st] == 1)), float(np.median(rr["t_us"][st])),
                                 float(rr["t_us"][st].max())))
            last_status, last_iters = rr["status"][-1], rr["iterations"][-1]
        else:
            quadruped.advance(prob, wl.qstate, qrng)
            sv.upload()
            ev[li][0].record(stream)
            sv.shift_fill(True, True)
            sv.solve(fetch=False)
            ev[li][1].record(stream)
            stt = sv.fetch()
            kern_ms.append(stt.tsolve)
            flops += flops_model(prob, stt.iterations, stt.iterations_outer, stt.ls_trials)
            per_step.append((float(stt.iterations.mean()), float(stt.ls_trials.mean()), float(np.mean(stt.status == 1)),
                             float(np.median(stt.t_instance_us)), float(stt.t_instance_us.max())))
            last_status, last_iters = stt.status, stt.iterations
    barrier()
    wall = time.perf_counter() - t_wall
    sampler.stop_flag.set()
    sampler.join()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_s = sharding.max_over_ranks(sum(step_ms) * 1e-3)
    kern_s = sum(kern_ms) * 1e-3
    gathered = sharding.gather_stats({"iterations": last_iters, "status": last_status})

    # ---- lock-step variant for the record: one transition + one solve launch per MPC step (every step waits
    #      for the slowest instance of the batch)
    lock = None
    if fused:
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
        lock = {"value": B * world * K / t_lock, "unit": UNIT, "ms_per_step": 1e3 * t_lock / K,
                "note": "one launch per MPC step: each step waits for the slowest instance"}
        sv.fetch()

    # ---- e2e: same metric through the public API with host buffers
    e2e = None
    if not args.no_e2e:
        barrier()
        if fused:
            sv.restore()
            zs = np.ascontiguousarray(zs_all[W:W + K])  # the same disturbances the device-timed run consumed
            sv.lib.altro_host_register(S._p(zs), zs.nbytes)
            h2d = zs.nbytes
            d2h = K * B * (prob.n + prob.m) * 8 + K * B * (4 * 4 + 2 * 8 + 8) + prob.X.nbytes + prob.U.nbytes
            t0 = time.perf_counter()
            sv.set_noise_bank(zs)  # H2D of this run's inputs
            sv.mpc_run(K, shift=wl.shift, fetch=True)  # run + D2H of closed-loop states, controls, statistics, X, U
            t_e2e = time.perf_counter() - t0
            sv.lib.altro_host_unregister(S._p(zs))
            h2d, d2h = h2d / K, d2h / K
        else:
            h2d = prob.x0.nbytes + prob.model.A.nbytes + prob.model.B.nbytes + prob.model.d.nbytes
            d2h = prob.X.nbytes + prob.U.nbytes + B * (4 * 4 + 4 * 8 + 8)
            t_e2e = 0.0
            for st in range(K):
                quadruped.advance(prob, wl.qstate, qrng)  # host-side linearisation (the caller's work, not timed)
                t0 = time.perf_counter()
                sv.shift_fill(True, True)  # pinned H2D of x0 and A_k, B_k, d_k, then the shifts on the device
                sv.solve(fetch=True)  # solve + D2H of X, U and statistics
                t_e2e += time.perf_counter() - t0
        barrier()
        t_e2e = sharding.max_over_ranks(t_e2e)
        e2e = {"value": B * world * K / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * t_e2e / K}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    value = B * world * K / total_s
    achieved = flops / kern_s / 1e12
    bytes_alg = algorithmic_bytes(prob, sv.P) * B * K
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        hbm_src = "MEASURED_PEAKS.json"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    ps = np.array(per_step)
    traffic = None
    if wl.name == "rocket" and B == 4096 and fused:
        # 23.66 MB for the 20-step launch (ncu, profiles/r1_solve_kernel_ncu_metrics.json): per launch the state of the
        # run, track and constraint data in and the solution out (18.7 MB), per step 4096 x 48 B of disturbances in
        # and the step's closed-loop log and statistics out (0.25 MB); scaled to this launch's step count
        traffic = 18.7e6 + 0.25e6 * S_launch
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * total_s / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl.desc, "instances_per_gpu": B, "global_batch": B * world,
                   "parallelism": f"instance-sharded x{world}, no collective on the solve path",
                   "l2": f"L2 flushed between timed steps ({L2_FLUSH_BYTES >> 20} MB write)",
                   "launch": info,
                   "step": ("closed-loop MPC run: per instance K x {transition + shift_fill + AL-iLQR solve} in one "
                            "launch, instances advance independently") if fused else
                           "host-built dynamics upload + device shift_fill + batched AL-iLQR solve, one launch per step"},
        "p50_solve_us": float(np.median(ps[:, 3])), "max_solve_us": float(ps[:, 4].max()),
        "iters_mean": float(ps[:, 0].mean()), "ls_trials_mean": float(ps[:, 1].mean()),
        "success": float(ps[:, 2].mean()), "success_all_ranks_last_step": float(np.mean(gathered["status"] == 1)),
        "iters_all_ranks_last_step": float(np.mean(gathered["iterations"])),
        "init_success": init_ok,
        "gpu_launches": int(2 * n_launch if fused else 2 * K),
        "steps_per_launch": S_launch,
        "clocks": sampler.summary(),
        # "tensor" = the compute roofline of the two the contract names; here it is the FP64 one (DFMA + DMMA m8n8k4)
        "roofline": {"bound": "tensor", "pipe": "fp64", "achieved": achieved, "peak": peaks["dfma_tflops"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["dfma_tflops"], "traffic": traffic,
                     "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of the "
                                     "20-step rocket launch (profiles/r1_summary.md), scaled to this launch's step count",
                     "peak_source": "DFMA stream measured live by altro_measure_peaks in this run "
                                    "(MEASURED_PEAKS.json has no FP64 figure); DMMA m8n8k4 peak "
                                    f"{peaks['dmma_tflops']:.1f} TFLOP/s",
                     "kernel": "altro_solve_kernel", "kernel_ms_per_step": 1e3 * kern_s / K,
                     "kernel_share_of_step": kern_s / (sum(step_ms) * 1e-3),
                     "flops_per_step": flops / K,
                     "hbm": {"achieved": bytes_alg / kern_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": bytes_alg / kern_s / 1e9 / hbm_peak, "peak_source": hbm_src,
                             "algorithmic_bytes_per_solve": algorithmic_bytes(prob, sv.P)}},
        "wall_s_timed_region": wall,
    }
    if e2e:
        line["e2e"] = e2e
    if lock:
        line["lockstep"] = lock
    if world == 1 and not args.no_cpu_baseline:
        nt = host_threads()
        cpu_steps = args.cpu_steps
        if cpu_steps <= 0:  # probe 4 steps, then size the sample for about 20 core-seconds (bounded by K)
            probe = oracle_arm(Workload(args.workload, args.batch, seed, make_solver), 4, 1, nt)
            cpu_steps = int(min(max(4, round(20.0 / nt * probe["value"] / B)), max(K, 4)))
        r = oracle_arm(Workload(args.workload, args.batch, seed, make_solver), cpu_steps, 1, nt)
        line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": nt, "kind": "port",
                                "sample": f"{cpu_steps} MPC steps x {B} instances of the same workload, CPU oracle "
                                          f"(oracle/altro_oracle.c, pthreads over instances), {r['seconds']:.2f} s",
                                "iters_mean": r["iters_mean"]}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
