"""Development check: GPU solve vs CPU oracle on the benchmark families (run on a GPU box)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from altro_mpc_icra2021_b200 import solver as S
from altro_mpc_icra2021_b200.problems import random_linear, rocket, quadruped, flexsat, mpc
from oracle.oracle import OracleProblem


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def compare(name, prob, opts, steps=3, advance=None, T=0):
    import copy
    pg = copy.deepcopy(prob)
    op = OracleProblem(prob)
    sg = S.ALTROSolver(pg, opts, threads_per_instance=T)
    print(name, sg.launch_info(), flush=True)
    for st in range(steps):
        t = time.time(); ro = op.solve(opts, nthreads=8); tc = time.time() - t
        sg.solve(); g = sg.stats
        lam_g = sg.get_duals()
        print(f"  step {st}: bitX {np.mean(pg.X == ro.X):.4f} bitU {np.mean(pg.U == ro.U):.4f} bitlam {np.mean(lam_g == ro.lam) if lam_g.size else 1:.4f} bitJ {np.mean(g.cost_al == ro.cost_al):.4f} ls_eq {np.mean(g.ls_trials == ro.ls_trials):.4f} relX {rel(pg.X, ro.X):.2e} relU {rel(pg.U, ro.U):.2e} relJ {rel(g.cost, ro.cost):.2e} "
              f"cmax g/o {g.c_max.max():.2e}/{ro.c_max.max():.2e} dcmax {np.abs(g.c_max-ro.c_max).max():.2e} "
              f"rellam {rel(lam_g, ro.lam) if lam_g.size else 0:.2e} "
              f"iters_eq {np.mean(g.iterations == ro.iterations):.4f} outer_eq {np.mean(g.iterations_outer == ro.iterations_outer):.4f} "
              f"status_eq {np.mean(g.status == ro.status):.4f} ok {np.mean(g.status == 1):.3f} "
              f"it_mean {g.iterations.mean():.2f} ls_mean {g.ls_trials.mean():.2f} gpu_ms {g.tsolve:.3f} cpu8_ms {tc*1e3:.1f} "
              f"p50_us {np.median(g.t_instance_us):.1f}", flush=True)
        if advance is not None:
            advance(prob, op, pg, sg, st)
    sg.close()


def main():
    B = int(os.environ.get("B", "256"))
    print(json.dumps(S.measure_peaks()), flush=True)
    # rocket
    cold = rocket.cold_problem()
    import copy
    cold_g = copy.deepcopy(cold)
    oc = OracleProblem(cold); rc = oc.solve(rocket.cold_options())
    sc = S.ALTROSolver(cold_g, rocket.cold_options()); sc.solve()
    print("rocket cold: iters", sc.stats.iterations, rc.iterations, "relX", rel(cold_g.X, rc.X), "relU", rel(cold_g.U, rc.U),
          "cmax", sc.stats.c_max, rc.c_max, "ms", sc.stats.tsolve, sc.launch_info(), flush=True)
    Xt, Ut = rc.X[0], rc.U[0]
    pm, ks = rocket.mpc_problem(cold, Xt, Ut, 21, batch=B)
    rng = mpc.rng_for(7, 7)

    def adv_track(noise_fn):
        def adv(prob, op, pg, sg, st):
            # identical host-side MPC update applied to both copies (same noise)
            class _S:  # oracle adapter
                def __init__(s): s.prob = prob
                def shift_fill(s, primal=True, dual=True): op.shift_fill(primal, dual)
            lo = mpc.MPCLoop(_S(), Xt_, Ut_, adv.k, noise=None)
            x0 = lo.plant_step()
            x0 = x0 + noise_fn(x0, rng)
            adv.k = adv.k + 1
            Xr, Ur = mpc.window_reference(Xt_, Ut_, adv.k, prob.N)
            for p_, s_ in ((prob, None), (pg, sg)):
                p_.set_initial_state(x0); p_.update_trajectory(Xr, Ur)
            op.shift_fill(True, True); sg.shift_fill(True, True)
        return adv
    Xt_, Ut_ = Xt, Ut
    a = adv_track(rocket.noise); a.k = ks.copy()
    compare("rocket", pm, rocket.mpc_options(), steps=4, advance=a)
    # random linear
    pr, Xt_, Ut_, ks = random_linear.mpc_problem(12, 6, 21, batch=B)
    a = adv_track(random_linear.noise); a.k = ks.copy()
    compare("randlin", pr, random_linear.mpc_options(), steps=4, advance=a)
    # quadruped
    for lin in (True, False):
        pq, st = quadruped.mpc_problem(B, linearized_friction=lin)
        compare("quadruped lin=%s" % lin, pq, quadruped.mpc_options(), steps=2)
    pf = flexsat.mpc_problem(80, batch=B)
    compare("flexsat", pf, flexsat.mpc_options(), steps=2)


if __name__ == "__main__":
    main()
