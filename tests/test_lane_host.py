"""The lane-per-instance solver (csrc/altro_lane.cuh) compiled for the HOST and run against the CPU oracle: the
per-lane code is __host__ __device__, so the statements the GPU executes are checked here bit for bit without a GPU
(the GPU runs of the same code are covered by tests/test_gpu_parity.py)."""
import copy
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from altro_mpc_icra2021_b200.problems import grasp, mpc, rocket
from oracle import oracle as orc
from oracle.oracle import OracleProblem, _opts_struct, _ptr, _Run
from tests.helpers import lqr_problem

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")


@pytest.fixture(scope="module")
def lane_lib():
    out = os.path.join(HERE, "native", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "liblane_host.so")
    src = os.path.join(HERE, "native", "lane_host.cu")
    hdr = os.path.join(HERE, "..", "altro_mpc_icra2021_b200", "csrc", "altro_lane.cuh")
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(so):
        subprocess.check_call(["nvcc", "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a",
                               "-fmad=false", "-Xcompiler", "-fPIC", "-shared", "-o", so, src])
    lib = C.CDLL(so)
    lib.lane_host_run.restype = C.c_int
    return lib


def lane_run(lib, op: OracleProblem, opts, steps=0, noise=None, noise_model=(0, 1.0, 1.0), track=None, shift=True):
    """Same contract as OracleProblem.solve / mpc_run, computed by the host build of the lane solver."""
    pr, B = op.prob, op.prob.B
    slots = max(steps, 1)
    it, ito, st, ls = (np.zeros((slots, B), np.int32) for _ in range(4))
    cost, cal, cmax = (np.zeros((slots, B)) for _ in range(3))
    x0l, u0l = np.zeros((slots, B, pr.n)), np.zeros((slots, B, pr.m))
    run, keep = None, []
    if steps:
        run = _Run()
        run.steps, run.shift, run.noise_mode = steps, int(shift), int(noise_model[0])
        run.w1, run.w2 = float(noise_model[1]), float(noise_model[2])
        if noise is not None:
            nz = np.ascontiguousarray(noise, dtype=np.float64)
            keep.append(nz)
            run.noise = _ptr(nz)
        ki = np.ascontiguousarray(pr.kidx, dtype=np.int32).copy()
        keep.append(ki)
        run.kidx = _ptr(ki)
        if track is not None:
            Xt, Ut = (np.ascontiguousarray(a, dtype=np.float64) for a in track)
            keep += [Xt, Ut]
            run.trackX, run.trackU, run.Nt = _ptr(Xt), _ptr(Ut), Xt.shape[0]
    o = _opts_struct(opts)
    rc = lib.lane_host_run(C.byref(op.c), C.byref(o), C.byref(run) if run is not None else None, _ptr(pr.X),
                           _ptr(pr.U), _ptr(op.lam), _ptr(it), _ptr(ito), _ptr(st), _ptr(ls), _ptr(cost), _ptr(cal),
                           _ptr(cmax), _ptr(x0l), _ptr(u0l))
    assert rc == 0
    if steps:
        pr.kidx[...] = ki + steps
    return {"iterations": it, "iterations_outer": ito, "status": st, "ls_trials": ls, "cost": cost, "cost_al": cal,
            "c_max": cmax, "x0": x0l, "u0": u0l}


def same(a: dict, b: dict, keys):
    for k in keys:
        assert np.array_equal(np.asarray(a[k]).reshape(np.asarray(b[k]).shape), b[k]), k


def test_lane_plain_solve_rocket_and_unconstrained(lane_lib):
    cold = np.load(os.path.join(GOLD, "rocket_cold.npz"))
    pm, _ = rocket.mpc_problem(rocket.cold_problem(), cold["X"], cold["U"], 21, batch=24)
    pm.set_initial_state(pm.x0 + 0.05 * mpc.rng_for(1, 2).standard_normal(pm.x0.shape))
    for prob, opts in ((pm, rocket.mpc_options()), (lqr_problem(n=4, m=2, N=15, batch=5, seed=2, u_bnd=0.3), None)):
        from altro_mpc_icra2021_b200.problem import SolverOptions
        opts = opts or SolverOptions(penalty_initial=10.0)
        pa, pb = copy.deepcopy(prob), copy.deepcopy(prob)
        oa, ob = OracleProblem(pa), OracleProblem(pb)
        r = oa.solve(opts)
        g = lane_run(lane_lib, ob, opts)
        assert np.array_equal(pa.X, pb.X) and np.array_equal(pa.U, pb.U) and np.array_equal(oa.lam, ob.lam)
        ref = {"iterations": r.iterations, "iterations_outer": r.iterations_outer, "status": r.status,
               "ls_trials": r.ls_trials, "cost": r.cost, "cost_al": r.cost_al, "c_max": r.c_max}
        same(g, ref, ref.keys())
        assert np.all(r.status == 1)


@pytest.mark.parametrize("family", ["rocket", "grasp"])
def test_lane_closed_loop_run_matches_oracle(lane_lib, family):
    B, steps = 16, 12
    if family == "rocket":
        cold = np.load(os.path.join(GOLD, "rocket_cold.npz"))
        Xt, Ut = cold["X"], cold["U"]
        pm, ks = rocket.mpc_problem(rocket.cold_problem(), Xt, Ut, 21, batch=B)
        opts, nm = rocket.mpc_options(), (2, 1e-3, 1e-2)
    else:
        cp = grasp.cold_problem()
        rc = OracleProblem(cp).solve(grasp.cold_options())
        Xt, Ut = rc.X[0], rc.U[0]
        pm, ks = grasp.mpc_problem(cp, Xt, Ut, 21, batch=B, seed=3)
        opts, nm = grasp.mpc_options(), (1, 0.01, 0.0)
    pa, pb = copy.deepcopy(pm), copy.deepcopy(pm)
    oa, ob = OracleProblem(pa), OracleProblem(pb)
    oa.solve(opts, nthreads=4)
    lane_run(lane_lib, ob, opts)
    assert np.array_equal(pa.X, pb.X)
    noise = mpc.rng_for(9, 9).standard_normal((steps, B, pm.n))
    ra = oa.mpc_run(opts, steps, noise, nm, (Xt, Ut), None, True, nthreads=4)
    rb = lane_run(lane_lib, ob, opts, steps, noise, nm, (Xt, Ut), True)
    same(rb, ra, ["iterations", "iterations_outer", "status", "ls_trials", "cost", "c_max", "x0", "u0"])
    assert np.array_equal(pa.X, pb.X) and np.array_equal(pa.U, pb.U) and np.array_equal(oa.lam, ob.lam)
    assert np.array_equal(pa.x0, pb.x0) and np.array_equal(pa.Xref, pb.Xref) and np.array_equal(pa.Uref, pb.Uref)
    assert ra["iterations"].max() > 2 and np.all(ra["status"] == 1)
