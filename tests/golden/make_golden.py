"""Generates tests/golden/*.npz with the CPU oracle (run from the repo root: python -m tests.golden.make_golden).

PARITY UNPINNED: the reference (Julia Altro.jl) cannot run in this environment and ships no golden
trajectories, so these vectors pin the oracle's own semantics (SURVEY.md Appendix A with the switches of
Appendix D at their defaults) against drift, and give the CUDA path a fixed target that needs no oracle build."""
import os

import numpy as np

from tests.golden import cases
from tests.helpers import OracleSolver

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    prob, opts = cases.rocket_track()
    s = OracleSolver(prob, opts, nthreads=1).solve()
    Xt, Ut = prob.X[0].copy(), prob.U[0].copy()
    np.savez_compressed(os.path.join(HERE, "rocket_cold.npz"), X=Xt, U=Ut, iters=s.stats.iterations,
                        outer=s.stats.iterations_outer, cost=s.stats.cost, cmax=s.stats.c_max)
    out = cases.run_case(lambda p, o: OracleSolver(p, o, nthreads=1), *cases.case_rocket_mpc(Xt, Ut))
    np.savez_compressed(os.path.join(HERE, "rocket_mpc.npz"), **out)
    for name, make in cases.CASES.items():
        out = cases.run_case(lambda p, o: OracleSolver(p, o, nthreads=1), *make())
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, {k: v.shape for k, v in out.items() if k.startswith(("X", "iters"))})


if __name__ == "__main__":
    main()
