#!/usr/bin/env python
"""Serialises a batch of MPC instances -- every input the solve path reads, per MPC step -- so that someone WITH Julia
can run the real Altro.jl on exactly these inputs (julia/run_reference_batch.jl) and close the parity loop that is
open here (no julia in the build environment).

    python scripts/dump_case.py rocket out_dir [--batch 8] [--steps 20]

Writes out_dir/problem.json (dimensions, options, cost weights, constraint blocks by the names of lowering.tbl) and,
per step s, out_dir/step_%04d.bin: little-endian float64 arrays in the order listed in problem.json["step_layout"]
(x0, Xref, Uref, U0 warm start, duals in, then the results of THIS repo's solver: X, U, duals out, iterations, cost,
c_max).  Plain binary + JSON so that Julia needs no extra package to read them.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle.oracle import OracleProblem  # noqa: E402  (the CPU restatement produces the recorded results)


class _OS:
    def __init__(self, prob, opts):
        self.prob, self.opts, self.op = prob, opts, OracleProblem(prob)

    def solve(self):
        self.stats = self.op.solve(self.opts, nthreads=4)
        return self


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("out")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    wl = bench.Workload(a.workload, a.batch, 0xA1720 + 2, _OS)
    p = wl.prob
    op = OracleProblem(p)
    op.solve(wl.opts, nthreads=4)
    cons = [dict(name=c.name, sense=int(c.sense), side=int(c.side), k0=int(c.k0), k1=int(c.k1), inds=c.inds.tolist(),
                 per_knot=bool(c.per_knot), per_instance=bool(c.per_instance), track=bool(c.track),
                 G=np.asarray(c.G).tolist(), h=np.asarray(c.h).tolist()) for c in p.constraints.flat]
    mdl = p.model
    meta = dict(workload=wl.desc, n=p.n, m=p.m, N=p.N, B=p.B, dt=p.dt, steps=a.steps,
                options={k: (float(v) if isinstance(v, float) else int(v)) for k, v in wl.opts.__dict__.items()},
                Q=p.obj.Q.tolist(), R=p.obj.R.tolist(), Qf=p.obj.Qf.tolist(),
                dynamics=dict(per_knot=bool(mdl.per_knot), per_instance=bool(mdl.per_instance),
                              A=mdl.A.tolist(), B=mdl.B.tolist(), d=mdl.d.tolist(),
                              sched=None if mdl.sched is None else mdl.sched.tolist()),
                constraints=cons, noise_model=list(wl.noise_model), shift=bool(wl.shift),
                kidx0=p.kidx.tolist(),
                step_layout=[["x0", [p.B, p.n]], ["Xref", [p.B, p.N, p.n]], ["Uref", [p.B, p.N - 1, p.m]],
                             ["U0", [p.B, p.N - 1, p.m]], ["lam_in", [p.B, op.P]], ["X", [p.B, p.N, p.n]],
                             ["U", [p.B, p.N - 1, p.m]], ["lam_out", [p.B, op.P]], ["iterations", [p.B]],
                             ["cost", [p.B]], ["c_max", [p.B]]],
                layout_note="row-major (C order); Julia: reshape(read, reverse(dims)...) gives the reversed dimension order")
    with open(os.path.join(a.out, "problem.json"), "w") as f:
        json.dump(meta, f)
    zs = wl.noise_samples(a.steps)
    k = wl.k.copy()
    for s in range(a.steps):
        # one step of the closed loop = transition (mirrored here to record its outputs) + solve
        lam_in_prev, U_prev = op.lam.copy(), p.U.copy()
        r = op.mpc_run(wl.opts, 1, zs[s:s + 1], wl.noise_model, wl.track, k, wl.shift, nthreads=4)
        k = k + 1
        U0 = U_prev.copy()
        lam_in = lam_in_prev.copy()
        if wl.shift:
            U0[:, :-1] = U_prev[:, 1:]
            off = 0
            for c in p.constraints.flat:
                nk, pp = c.k1 - c.k0, c.p
                blk = lam_in[:, off:off + nk * pp].reshape(p.B, nk, pp)
                blk[:, :-1] = blk[:, 1:].copy()
                off += nk * pp
        with open(os.path.join(a.out, "step_%04d.bin" % s), "wb") as f:
            for arr in (r["x0"][0], p.Xref, p.Uref, U0, lam_in, p.X, p.U, op.lam, r["iterations"][0].astype(np.float64),
                        r["cost"][0], r["c_max"][0]):
                f.write(np.ascontiguousarray(arr, dtype="<f8").tobytes())
    print("wrote", a.steps, "steps of", wl.desc, "to", a.out)


if __name__ == "__main__":
    main()
