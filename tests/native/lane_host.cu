// Host build of the lane-per-instance solver (altro_mpc_icra2021_b200/csrc/altro_lane.cuh): the per-lane code is
// __host__ __device__, so the exact statements the GPU executes can be run here, one instance after the other, on
// the oracle's problem structs and compared with the oracle bit for bit without a GPU (tests/test_lane_host.py).
// Test infrastructure: built by the test with `nvcc -shared`, never part of libaltro_b200.so.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../oracle/altro_oracle.h"
#include "../../altro_mpc_icra2021_b200/csrc/altro_lane.cuh"

using namespace altro;

template <int NX, int NU>
static void run_all(const Params &P, const LaneLayout &L, int B)
{
    std::vector<double> ws((size_t)L.total * B, 0.0), scr(LANE_SCRATCH + P.N * (1 + P.ncon), 0.0);
    LaneConst<NX, NU> C;
    for (int i = 0; i < NX * NX; ++i) C.A[i] = P.A[i];
    for (int i = 0; i < NX * NU; ++i) C.B[i] = P.Bm[i];
    for (int i = 0; i < NX; ++i) { C.d[i] = P.d[i]; C.Q[i] = P.Q[i]; C.Qf[i] = P.Qf[i]; }
    for (int i = 0; i < NU; ++i) C.R[i] = P.R[i];
    for (int i = 0; i < B; ++i) {
        Lane<NX, NU, 1> ln(P, C, L, ws.data() + i, (size_t)B, scr.data(), P.con, i);
        ln.load();
        while (ln.phase != LP_DONE) ln.step();
        ln.store();
    }
}

extern "C" int lane_host_run(const orc_problem_t *pb, const orc_opts_t *o, const orc_run_t *run, double *X, double *U,
                             double *lam, int *iters, int *outer, int *status, int *trials, double *cost,
                             double *cost_al, double *cmax, double *x0_log, double *u0_log)
{
    static_assert(sizeof(orc_opts_t) == sizeof(altro_opts_t), "option structs must mirror each other");
    const int n = pb->n, m = pb->m, N = pb->N, B = pb->B, steps = run ? run->steps : 0, slots = steps > 0 ? steps : 1;
    Params P;
    memset(&P, 0, sizeof(P));
    std::vector<ConDesc> cd(pb->ncon > 0 ? pb->ncon : 1);
    std::vector<std::vector<int>> cols(pb->ncon);
    std::vector<std::vector<double>> coefs(pb->ncon);
    int Pd = 0;
    for (int c = 0; c < pb->ncon; ++c) {
        const orc_con_t &s = pb->con[c];
        ConDesc &d = cd[c];
        memset(&d, 0, sizeof(d));
        d.sense = s.sense; d.side = s.side; d.k0 = s.k0; d.k1 = s.k1; d.p = s.p; d.w = s.w;
        d.per_knot = s.per_knot; d.per_instance = s.per_instance; d.track = s.track;
        d.dual_off = Pd;
        d.G = s.G; d.h = s.h;
        for (int j = 0; j < s.w; ++j) d.inds[j] = s.inds[j];
        bool rs = !s.per_knot && !s.per_instance && !s.track && s.sense != ORC_SOC;
        cols[c].assign(s.p, 0);
        coefs[c].assign(s.p, 0.0);
        for (int r = 0; r < s.p && rs; ++r) {
            int nz = 0;
            for (int j = 0; j < s.w; ++j)
                if (s.G[r * s.w + j] != 0.0) { ++nz; cols[c][r] = j; coefs[c][r] = s.G[r * s.w + j]; }
            rs = nz <= 1;
        }
        d.rowsparse = rs ? 1 : 0;
        d.rs_col = cols[c].data();
        d.rs_coef = coefs[c].data();
        Pd += (s.k1 - s.k0) * s.p;
    }
    P.n = n; P.m = m; P.N = N; P.B = B; P.P = Pd; P.ncon = pb->ncon; P.dt = pb->dt;
    P.dyn_per_knot = pb->dyn_per_knot; P.dyn_per_instance = pb->dyn_per_instance;
    P.A = pb->A; P.Bm = pb->Bm; P.d = pb->d;
    P.dyn_slots = pb->dyn_slots; P.sched_len = pb->sched_len; P.step0 = pb->step0; P.dyn_sched = pb->sched;
    P.Q = pb->Q; P.R = pb->R; P.Qf = pb->Qf;
    P.xref = const_cast<double *>(pb->xref); P.uref = const_cast<double *>(pb->uref); P.x0 = const_cast<double *>(pb->x0);
    P.X = X; P.U = U; P.lam = lam;
    std::vector<double> tmp_cal((size_t)slots * B), tmp_pen((size_t)slots * B);
    P.iters = iters; P.outer = outer; P.status = status; P.trials = trials;
    P.cost = cost; P.cost_al = cost_al ? cost_al : tmp_cal.data(); P.cmax = cmax; P.penmax = tmp_pen.data();
    P.con = cd.data();
    memcpy(&P.o, o, sizeof(altro_opts_t));
    std::vector<double> px, pu;
    P.kidx = pb->kidx;
    if (run) {
        P.steps = steps; P.shift = run->shift; P.noise_mode = run->noise_mode; P.Nt = run->Nt;
        P.noise_w1 = run->w1; P.noise_w2 = run->w2; P.noise = run->noise;
        P.kidx = run->kidx;
        P.x0_log = x0_log; P.u0_log = u0_log;
        if (run->trackX) {  // padded with N copies of the last knot, like altro_set_track
            const int Nt = run->Nt;
            px.resize((size_t)(Nt + N) * n);
            pu.resize((size_t)(Nt - 1 + N) * m);
            for (int k = 0; k < Nt + N; ++k) memcpy(&px[(size_t)k * n], run->trackX + (size_t)(k < Nt - 1 ? k : Nt - 1) * n, n * sizeof(double));
            for (int k = 0; k < Nt - 1 + N; ++k) memcpy(&pu[(size_t)k * m], run->trackU + (size_t)(k < Nt - 2 ? k : Nt - 2) * m, m * sizeof(double));
            P.trackX = px.data();
            P.trackU = pu.data();
        }
    }
    const LaneLayout L = make_lane_layout(n, m, N, Pd);
    if (n == 6 && m == 3) run_all<6, 3>(P, L, B);
    else if (n == 6 && m == 6) run_all<6, 6>(P, L, B);
    else if (n == 4 && m == 2) run_all<4, 2>(P, L, B);
    else return -1;
    return 0;
}
