// altro_abi.cu -- C ABI (include/altro_b200.h) over the sm_100a batched ALTRO kernels.
//
// Owns the device-resident problem batch (dynamics, weights, references, constraint data,
// trajectories, duals, per-instance statistics) of one handle and launches the solve /
// shift / MPC-transition kernels on the handle's stream.  No CPU fallback anywhere: every
// compute entry point fails with ALTRO_ERR_CUDA when no device is usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "altro_lane.cuh"
#include "altro_quadruped.cuh"
#include "altro_admm.cuh"

using namespace altro;

namespace {

thread_local std::string g_err;

struct HostCon {
    int sense, side, k0, k1, p, w, per_knot, per_instance, rowsparse, track = 0;
    std::vector<int> inds, rs_col;
    std::vector<double> rs_coef;
    double *G_dev = nullptr, *h_dev = nullptr, *rs_coef_dev = nullptr;
    int *rs_col_dev = nullptr;
    size_t g_count = 0, h_count = 0;
};

}  // namespace

struct altro_handle_s {
    int device = 0, n = 0, m = 0, N = 0, B = 0;
    double dt = 0.0;
    altro_opts_t opts;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    // dynamics
    int dyn_per_knot = 0, dyn_per_instance = 0;
    double *A = nullptr, *Bm = nullptr, *d = nullptr;
    size_t dyn_count = 0;
    int dyn_slots = 0, sched_len = 0, step_abs = 0, step_abs_snap = 0;  // gait-scheduled models (quadruped)
    int *sched = nullptr;
    // cost / reference / state
    double *Q = nullptr, *R = nullptr, *Qf = nullptr, *xref = nullptr, *uref = nullptr, *x0 = nullptr;
    double *X = nullptr, *U = nullptr, *lam = nullptr;
    double *X_snap = nullptr, *U_snap = nullptr, *lam_snap = nullptr, *x0_snap = nullptr, *xref_snap = nullptr,
           *uref_snap = nullptr;
    int *kidx_snap = nullptr;
    int bank_pos_snap = 0;
    // stats
    int *iters = nullptr, *outer = nullptr, *status = nullptr, *trials = nullptr;
    double *cost = nullptr, *cost_al = nullptr, *cmax = nullptr, *penmax = nullptr;
    long long *t_ns = nullptr;
    int stat_cap = 1;  // statistics arrays hold stat_cap x B entries (one slot per step of a closed-loop run)
    int last_run_steps = 0;       // steps of the last launch if it was a closed-loop run, else 0
    bool pending_transition = false;  // altro_mpc_transition done, its solve not yet launched
    double *x0_log = nullptr, *u0_log = nullptr;
    long long *phase = nullptr;
    double *trace = nullptr;
    int trace_rows = 0;
    // constraints
    std::vector<HostCon> cons;
    ConDesc *con_dev = nullptr;
    int *itab_dev = nullptr;
    double *ex_glob = nullptr;
    double *ws = nullptr;  // large state dimension: [B][lay.ws_doubles]
    int P = 0, EX = 0, ITAB = 0, NSRC = 0, NTL = 0;
    bool finalized = false, have_dyn = false, have_cost = false, have_ref = false, have_x0 = false;
    // MPC track
    double *trackX = nullptr, *trackU = nullptr, *noise = nullptr, *noise_bank = nullptr;
    int *kidx = nullptr;
    int Nt = 0, bank_steps = 0, bank_pos = 0, bank_cap = 0, noise_mode = 0;
    double noise_w1 = 1.0, noise_w2 = 1.0;
    // launch
    int threads_req = 0, threads = 0, smem = 0, regs = 0, ctas_per_sm = 0, num_sms = 0, dyn_in_smem = 0, ref_in_smem = 1;
    const void *kernel = nullptr;
    Layout lay;
    int spec = 0, spec_req = -1;  // speculative line search: in use / requested (-1 = automatic)
    // lane-per-instance kernel (altro_lane.cuh): small dimensions, one thread per instance
    // convex cross-check (admm.cu): workspace and outputs
    double *admm_ws = nullptr, *admm_X = nullptr, *admm_U = nullptr, *admm_rp = nullptr, *admm_rd = nullptr;
    int *admm_it = nullptr;
    // quadruped pre-solve kernels (quadruped.cu): device scratch, footstep-planner state
    double *q_xref = nullptr, *q_uref = nullptr, *q_foot = nullptr, *q_contacts = nullptr, *q_t = nullptr,
           *q_curfoot = nullptr, *q_planner = nullptr;
    bool q_planner_set = false;
    int run_chunk = 1;    // closed-loop runs: steps per work item of the persistent grid (0 = one CTA per instance)
    int *q_ctrl = nullptr;  // [2 + B]: queue head, error flag, per-instance completed steps
    int kernel_mode = 0;  // 0 automatic, 1 CTA per instance, 2 lane per instance
    const void *lane_kern = nullptr;
    LaneLayout lane_lay{};
    double *lane_ws = nullptr;
    size_t lane_stride = 0;
    int lane_smem = 0, lane_regs = 0, lane_lpw = 8, lane_scratch = 0;
    std::vector<double> hA, hB, hd, hQ, hR, hQf;  // host copies (shared LTI model, weights): lane-kernel parameters
};

namespace {

int fail(altro_handle_t h, int code, const std::string &msg)
{
    if (h) h->err = msg;
    g_err = msg;
    return code;
}

#define CK(h, call)                                                                                     \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(h, ALTRO_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
    } while (0)

template <typename Tp>
cudaError_t dalloc(Tp **p, size_t count)
{
    cudaError_t e = cudaMalloc((void **)p, std::max<size_t>(count, 1) * sizeof(Tp));
    if (e == cudaSuccess) e = cudaMemset(*p, 0, std::max<size_t>(count, 1) * sizeof(Tp));
    // The memset runs on the null stream and is asynchronous to the host; the handle's kernels run on their own
    // (possibly non-blocking) stream.  Without this wait a late memset can zero what a kernel already wrote
    // (observed: an all-zero u0 log once in ~100 runs).
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    return e;
}

// Set-up copies (descriptors, tables, constraint data, tracks).  cudaMemcpy runs on the null stream, with which the
// handle's non-blocking stream does not synchronise, and may return before the DMA of a pageable buffer has landed:
// wait for the device on both sides so that no launch on the handle's stream can overtake or be overtaken.
inline cudaError_t copy_sync(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind)
{
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(dst, src, bytes, kind);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    return e;
}

// ------------------------------------------------------------------ auxiliary kernels

// RD.shift_fill!(Z) / Altro.shift_fill!(conSet): z_k <- z_{k+1}, last knot kept. One CTA per instance.
__global__ void shift_fill_kernel(int n, int m, int N, int P, int ncon, const ConDesc *con, double *X, double *U,
                                  double *lam, int primal, int dual)
{
    extern __shared__ double buf[];
    const int inst = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    if (primal) {
        double *x = X + (size_t)inst * N * n, *u = U + (size_t)inst * (N - 1) * m;
        for (int i = tid; i < (N - 1) * n; i += T) buf[i] = x[i + n];
        __syncthreads();
        for (int i = tid; i < (N - 1) * n; i += T) x[i] = buf[i];
        __syncthreads();
        for (int i = tid; i < (N - 2) * m; i += T) buf[i] = u[i + m];
        __syncthreads();
        for (int i = tid; i < (N - 2) * m; i += T) u[i] = buf[i];
        __syncthreads();
    }
    if (dual) {
        double *l = lam + (size_t)inst * P;
        for (int i = tid; i < P; i += T) buf[i] = l[i];
        __syncthreads();
        for (int c = 0; c < ncon; ++c) {
            const int nk = con[c].k1 - con[c].k0, p = con[c].p, off = con[c].dual_off;
            for (int i = tid; i < (nk - 1) * p; i += T) l[off + i] = buf[off + i + p];
        }
    }
}

// Device-side MPC transition (random_linear_problem.jl:121-139, simple_rocket.jl:59-82):
// plant step with the first control (+ noise), reference window advanced along the track.
// Noise models: 0 additive w1*z; 1 random-linear z*|x|_inf*w1 (random_linear_problem.jl:129);
// 2 rocket: z*|x[0:n/2]|_2*w1 on positions, z*|x[n/2:n]|_2*w2 on velocities (simple_rocket.jl:63-70).
__global__ void mpc_transition_kernel(int n, int m, int N, const double *A, const double *Bm, const double *d,
                                      int dyn_per_knot, int dyn_per_instance, const int *sched, int sched_len,
                                      int dyn_slots, int step_abs, const double *X, const double *U,
                                      const double *noise, int noise_mode, double w1, double w2, double *x0,
                                      const double *trackX, const double *trackU, int Nt, int *kidx, double *xref,
                                      double *uref)
{
    __shared__ double scale[2];
    const int inst = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    size_t base = dyn_per_instance ? (size_t)inst * (sched ? (size_t)dyn_slots : (dyn_per_knot ? (size_t)(N - 1) : 1)) : 0;
    if (sched) base += sched[(size_t)inst * sched_len + min(step_abs, sched_len - N)];
    const double *A0 = A + base * n * n, *B0 = Bm + base * n * m, *d0 = d + base * n;
    const double *x = X + (size_t)inst * N * n, *u = U + (size_t)inst * (N - 1) * m;
    double *xo = x0 + (size_t)inst * n;
    for (int i = tid; i < n; i += T) {
        double acc = d0[i];
        for (int j = 0; j < n; ++j) acc = fma(A0[i * n + j], x[j], acc);
        for (int j = 0; j < m; ++j) acc = fma(B0[i * m + j], u[j], acc);
        xo[i] = acc;
    }
    __syncthreads();
    if (noise) {
        if (tid == 0) {
            double s0 = w1, s1 = w1;
            if (noise_mode == 1) {
                double mx = 0.0;
                for (int i = 0; i < n; ++i) mx = fmax(mx, fabs(xo[i]));
                s0 = s1 = mx * w1;
            } else if (noise_mode == 2) {
                double a = 0.0, b = 0.0;
                for (int i = 0; i < n / 2; ++i) a += xo[i] * xo[i];
                for (int i = n / 2; i < n; ++i) b += xo[i] * xo[i];
                s0 = sqrt(a) * w1;
                s1 = sqrt(b) * w2;
            }
            scale[0] = s0;
            scale[1] = s1;
        }
        __syncthreads();
        for (int i = tid; i < n; i += T) xo[i] += noise[(size_t)inst * n + i] * scale[(noise_mode == 2 && i >= n / 2) ? 1 : 0];
    }
    if (trackX) {
        const int k0 = kidx[inst] + 1;
        __syncthreads();
        if (tid == 0) kidx[inst] = k0;
        double *xr = xref + (size_t)inst * N * n, *ur = uref + (size_t)inst * (N - 1) * m;
        for (int i = tid; i < N * n; i += T) {
            int k = min(k0 + i / n, Nt - 1);
            xr[i] = trackX[(size_t)k * n + i % n];
        }
        for (int i = tid; i < (N - 1) * m; i += T) {
            int k = min(k0 + i / m, Nt - 2);
            ur[i] = trackU[(size_t)k * m + i % m];
        }
    }
}

__global__ void advance_kidx_kernel(int *kidx, int B, int steps)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) kidx[i] += steps;
}

// ------------------------------------------------------------------ FP64 peak microbenchmarks

__global__ void dfma_peak_kernel(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void dmma_peak_kernel(double *out, int iters)
{
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0, c7 = 0;
    for (int i = 0; i < iters; ++i) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c2), "+d"(c3) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c4), "+d"(c5) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c6), "+d"(c7) : "d"(a), "d"(b));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
}

// ------------------------------------------------------------------ kernel dispatch

}  // namespace

namespace altro {
// one translation unit per compiled dimension pair (inst_*.cu)
const void *kernel_6_3(int T);    // rocket
const void *kernel_12_12(int T);  // quadruped
const void *kernel_6_6(int T);    // grasp
const void *kernel_12_3(int T);   // flexible satellite
const void *kernel_12_6(int T);   // random linear (default)
const void *kernel_0_0(int T);    // run-time dimensions
const void *kernel_0_0_wide(int T);  // run-time dimensions, one CTA per SM (TMA-staged large-dimension layout)
bool lane_supported(int n, int m);  // lane-per-instance kernels (6, 3), (6, 6)
cudaError_t lane_launch(int n, int m, const LaneLaunch &a);
}  // namespace altro

namespace {

const void *find_kernel(int n, int m, int T)
{
    if (getenv("ALTRO_B200_GENERIC")) return kernel_0_0(T);
    if (n == 6 && m == 3) return kernel_6_3(T);
    if (n == 12 && m == 12) return kernel_12_12(T);
    if (n == 6 && m == 6) return kernel_6_6(T);
    if (n == 12 && m == 3) return kernel_12_3(T);
    if (n == 12 && m == 6) return kernel_12_6(T);
    return kernel_0_0(T);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is per-kernel, process-wide state shared by every handle that uses the
// same <NX, NU, T> instantiation: keep a running maximum and never lower it, so that a handle with a smaller horizon
// finalised later cannot invalidate the launches of an earlier one.
cudaError_t raise_smem_limit(const void *kernel, int bytes)
{
    static std::mutex mu;
    static std::map<const void *, int> cur;
    std::lock_guard<std::mutex> lock(mu);
    int &c = cur[kernel];
    if (bytes <= c) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) c = bytes;
    return e;
}

int finalize(altro_handle_t h)
{
    if (h->finalized) return ALTRO_OK;
    if (!h->have_dyn || !h->have_cost) return fail(h, ALTRO_ERR_STATE, "dynamics and cost must be set before solve");
    const int n = h->n, m = h->m, N = h->N, B = h->B;
    // dual / expansion offsets, device descriptors
    std::vector<ConDesc> cd(std::max<size_t>(h->cons.size(), 1));
    int P = 0, EX = 0, TG = 0;
    const int NT = n + n * n + m + m * m;
    std::vector<std::vector<int>> srcs(NT);
    for (size_t i = 0; i < h->cons.size(); ++i) {
        HostCon &c = h->cons[i];
        ConDesc &dsc = cd[i];
        memset(&dsc, 0, sizeof(dsc));
        dsc.sense = c.sense; dsc.side = c.side; dsc.k0 = c.k0; dsc.k1 = c.k1; dsc.p = c.p; dsc.w = c.w;
        dsc.per_knot = c.per_knot; dsc.per_instance = c.per_instance; dsc.rowsparse = c.rowsparse;
        dsc.dual_off = P;
        dsc.ex_off = EX;
        dsc.ex_stride = c.rowsparse ? 2 * c.w : c.w + c.w * (c.w + 1) / 2;
        dsc.track = c.track;
        dsc.tgt_off = TG;
        TG += dsc.ex_stride;
        // gather sources: (block, offset inside the block's per-knot expansion) for every target it touches
        const int ld = c.side == ALTRO_STATE ? n : m;
        const int ovec = c.side == ALTRO_STATE ? 0 : n + n * n, omat = ovec + ld;
        for (int e = 0; e < c.w; ++e) srcs[ovec + c.inds[e]].push_back(((int)i << 16) | e);
        if (c.rowsparse) {
            for (int e = 0; e < c.w; ++e) srcs[omat + c.inds[e] * ld + c.inds[e]].push_back(((int)i << 16) | (c.w + e));
        } else {
            for (int a = 0; a < c.w; ++a)
                for (int b = 0; b < c.w; ++b) {  // upper-triangle packed storage of the symmetric block
                    const int lo = std::min(a, b), hi = std::max(a, b);
                    srcs[omat + c.inds[a] * ld + c.inds[b]].push_back(((int)i << 16) | (c.w + lo * c.w - lo * (lo - 1) / 2 + (hi - lo)));
                }
        }
        dsc.G = c.G_dev; dsc.h = c.h_dev; dsc.rs_col = c.rs_col_dev; dsc.rs_coef = c.rs_coef_dev;
        for (int j = 0; j < c.w; ++j) dsc.inds[j] = c.inds[j];
        P += (c.k1 - c.k0) * c.p;
        EX += (c.k1 - c.k0) * dsc.ex_stride;
    }
    h->P = P;
    h->EX = EX;
    {  // Gather table: one 16-byte record {k0, k1, offset at knot 0, stride per knot} per source, in CSR order (per
       // target the sources stay in ascending block order), then gptr[NT+1], then the list of targets that are
       // refreshed at every knot: the two vectors and the matrix entries at least one block touches.
        std::vector<int> itab;
        std::vector<int> gptr(NT + 1, 0), tl;
        for (int t = 0; t < NT; ++t) {
            gptr[t + 1] = gptr[t] + (int)srcs[t].size();
            for (int src : srcs[t]) {
                const ConDesc &c = cd[src >> 16];
                itab.push_back(c.k0);
                itab.push_back(c.k1);
                itab.push_back(c.ex_off - c.k0 * c.ex_stride + (src & 0xffff));
                itab.push_back(c.ex_stride);
            }
            const bool vec = t < n || (t >= n + n * n && t < n + n * n + m);
            if (vec || !srcs[t].empty()) tl.push_back(t);
        }
        h->NSRC = gptr[NT];
        h->NTL = (int)tl.size();
        itab.insert(itab.end(), gptr.begin(), gptr.end());
        itab.insert(itab.end(), tl.begin(), tl.end());
        h->ITAB = (int)itab.size();
        CK(h, dalloc(&h->itab_dev, itab.size()));
        CK(h, copy_sync(h->itab_dev, itab.data(), itab.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    CK(h, dalloc(&h->con_dev, cd.size()));
    CK(h, copy_sync(h->con_dev, cd.data(), cd.size() * sizeof(ConDesc), cudaMemcpyHostToDevice));
    CK(h, dalloc(&h->lam, (size_t)B * std::max(P, 1)));  // copy_state moves B * max(P, 1) doubles
    CK(h, dalloc(&h->lam_snap, (size_t)B * std::max(P, 1)));
    // launch geometry
    cudaDeviceProp prop;
    CK(h, cudaGetDeviceProperties(&prop, h->device));
    h->num_sms = prop.multiProcessorCount;
    const int maxdim = std::max(n, m);
    int T = h->threads_req;
    if (const char *e = getenv("ALTRO_B200_THREADS")) T = atoi(e);
    if (T == 0)  // measured on B200 (scripts/dev_perf.py sweeps): 2 warps up to 16 dimensions (the second warp takes every
                 // other tensor tile and line-search trial), 4 when the horizon is long (flexible satellite), more for
                 // the big run-time sized ones
        T = maxdim <= 16 ? (N >= 60 ? 128 : 64) : maxdim <= 32 ? 128 : 256;
    if (T != 32 && T != 64 && T != 128 && T != 256) return fail(h, ALTRO_ERR_INVALID, "threads per instance must be 32, 64, 128 or 256");
    h->threads = T;
    h->dyn_in_smem = (!h->dyn_per_knot && !h->dyn_per_instance) ? 1 : 0;
    const int ncons = (int)h->cons.size();
    const size_t limit = (size_t)prop.sharedMemPerBlockOptin;
    // Fixed-dimension kernels keep everything (reference window, expansion blocks) in shared memory.  Problems that
    // do not fit run on the run-time sized kernel, which can leave the reference window in global memory (long
    // horizons; closed-loop runs then read it straight from the L2-resident track) and, longer still, the
    // expansion blocks too.
    h->kernel = find_kernel(n, m, T);
    h->ref_in_smem = 1;
    // Speculative line search (one trial per warp, see forward_pass_spec) needs one more candidate set per extra warp:
    // on unless that costs resident CTAs (measured: random_linear loses more from 6 -> 5 CTAs/SM than it gains).
    h->lay = make_layout(n, m, N, P, ncons, EX, 1, h->ITAB);
    h->spec = 0;
    if (T > 32) {
        const Layout ls = make_layout(n, m, N, P, ncons, EX, 1, h->ITAB, T / 32 - 1);
        int want = h->spec_req;
        if (const char *e = getenv("ALTRO_B200_SPEC")) want = atoi(e) ? 1 : 0;
        if (want < 0 && (size_t)ls.bytes <= limit) {
            int nb0 = 0, nb1 = 0;
            CK(h, raise_smem_limit(h->kernel, ls.bytes));
            CK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb0, h->kernel, T, (size_t)h->lay.bytes));
            CK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb1, h->kernel, T, (size_t)ls.bytes));
            want = nb1 >= nb0 ? 1 : 0;
        }
        if (want > 0 && (size_t)ls.bytes <= limit) { h->spec = 1; h->lay = ls; }
    }
    if ((size_t)h->lay.bytes > limit || !ALTRO_FIXED_ALL_SMEM) {
        if ((size_t)h->lay.bytes > limit) h->kernel = kernel_0_0(T);
        h->spec = 0;
        if (h->trackX || (size_t)h->lay.bytes > limit) {
            h->ref_in_smem = 0;
            h->lay = make_layout(n, m, N, P, ncons, EX, 0, h->ITAB);
        }
        if ((size_t)h->lay.bytes > limit && EX > 0) {
            CK(h, dalloc(&h->ex_glob, (size_t)B * EX));
            h->lay = make_layout(n, m, N, P, ncons, 0, 0, h->ITAB);
        }
    }
    // Run-time sized problems whose shared-memory-resident layout would leave one instance per SM run faster from the
    // workspace layout (several instances per SM hide each other's latency; measured n = 30, m = 10..25: 1.7-2x).
    const bool lonely = h->kernel == kernel_0_0(T) && 2 * (size_t)h->lay.bytes > limit;
    const char *big_env = getenv("ALTRO_B200_BIG");
    if ((size_t)h->lay.bytes > limit || (big_env ? atoi(big_env) != 0 : lonely)) {
        // Large state dimension: n-sized matrices and gains in a per-instance global workspace (make_layout_big).
        if (h->ex_glob == nullptr && EX > 0) CK(h, dalloc(&h->ex_glob, (size_t)B * EX));
        // TMA-staged operand panels (panel_gemm_fn, one CTA per SM with the whole register file): bit-identical to the
        // register-blocked tiles fed from L2 (two CTAs per SM); measured (profiles/r2_large_n.md, 296 instances):
        // n = 200 2915 vs 2126 solves/s, n = 128 5535 vs 5643, n = 100 8015 vs 8184, n = 64 15.2k vs 23.6k -- so the
        // staged path takes over from n = 160.  ALTRO_B200_TMA=0/1 forces either.
        int want_tma = (T == 256 && n >= 160) ? 1 : 0;
        if (const char *e = getenv("ALTRO_B200_TMA")) want_tma = (T == 256 && atoi(e)) ? 1 : 0;
        Layout lb = make_layout_big(n, m, N, P, ncons, 0, want_tma);
        if ((size_t)lb.bytes > limit && lb.tma) lb = make_layout_big(n, m, N, P, ncons, 0, 0);  // no room for the panel stages
        if ((size_t)lb.bytes <= limit) {
            h->lay = lb;
            h->kernel = (lb.tma && kernel_0_0_wide(T)) ? kernel_0_0_wide(T) : kernel_0_0(T);
            h->ref_in_smem = 0;
            h->dyn_in_smem = 0;
            h->spec = 0;
            CK(h, dalloc(&h->ws, (size_t)B * lb.ws_doubles));
        }
    }
    size_t smem = (size_t)h->lay.bytes;
    if (smem > limit) {
        char buf[256];
        snprintf(buf, sizeof buf, "problem needs %zu B of shared memory per instance, device offers %zu B "
                 "(n=%d m=%d N=%d P=%d): not supported by the shared-memory-resident kernel", smem, limit, n, m, N, P);
        return fail(h, ALTRO_ERR_UNSUPPORTED, buf);
    }
    h->smem = (int)smem;
    if (!h->kernel) return fail(h, ALTRO_ERR_UNSUPPORTED, "no kernel for this configuration");
    CK(h, raise_smem_limit(h->kernel, (int)smem));
    cudaFuncAttributes fa;
    CK(h, cudaFuncGetAttributes(&fa, h->kernel));
    h->regs = fa.numRegs;
    int nb = 0;
    CK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, h->kernel, T, smem));
    h->ctas_per_sm = nb;
    // lane-per-instance kernel for the small-dimension families (shared LTI model), unless the caller asked for a CTA
    // geometry
    h->lane_kern = nullptr;
    {
        int mode = h->kernel_mode;
        if (const char *e = getenv("ALTRO_B200_LANE")) mode = atoi(e) ? 2 : 1;
        const bool can = lane_supported(n, m) && h->dyn_in_smem && h->hA.size() == (size_t)n * n;
        if (mode == 2 && !can) return fail(h, ALTRO_ERR_UNSUPPORTED, "no lane-per-instance kernel for these dimensions / this dynamics layout");
        // opt-in only: measured 10x slower than the CTA kernel on the 4096-instance batches (see altro_lane.cuh)
        if (can && mode == 2) {
            h->lane_lpw = 8;
            if (const char *e = getenv("ALTRO_B200_LPW")) h->lane_lpw = atoi(e);
            if (h->lane_lpw != 8 && h->lane_lpw != 32)
                return fail(h, ALTRO_ERR_INVALID, "instances per warp must be 8 or 32");
            h->lane_lay = make_lane_layout(n, m, N, P);
            h->lane_stride = ((size_t)B + 31) / 32 * 32;
            CK(h, dalloc(&h->lane_ws, (size_t)h->lane_lay.total * h->lane_stride));
            h->lane_scratch = LANE_SCRATCH + N * (1 + ncons);
            h->lane_smem = (int)((size_t)std::max(ncons, 1) * sizeof(ConDesc) + (size_t)h->lane_scratch * h->lane_lpw * sizeof(double));
            if ((size_t)h->lane_smem > limit) return fail(h, ALTRO_ERR_UNSUPPORTED, "lane kernel scratch exceeds shared memory");
            LaneLaunch q{};
            q.lpw = h->lane_lpw;
            q.A = q.B = q.d = q.Q = q.R = q.Qf = h->hA.data();  // only resolving the kernel: contents unused
            const void *lk = nullptr;
            q.query = &lk;
            CK(h, lane_launch(n, m, q));
            CK(h, raise_smem_limit(lk, h->lane_smem));
            cudaFuncAttributes la;
            CK(h, cudaFuncGetAttributes(&la, lk));
            h->lane_regs = la.numRegs;
            h->lane_kern = lk;
        }
    }
    h->finalized = true;
    return ALTRO_OK;
}

int upload(altro_handle_t h, double *dst, const double *src, size_t count)
{
    CK(h, cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    return ALTRO_OK;
}

int download(altro_handle_t h, void *dst, const void *src, size_t bytes)
{
    CK(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    return ALTRO_OK;
}

}  // namespace

#define REQ(h)                                                       \
    do {                                                             \
        if (!(h)) return fail(nullptr, ALTRO_ERR_INVALID, "null handle"); \
        cudaError_t e_ = cudaSetDevice((h)->device);                 \
        if (e_ != cudaSuccess) return fail(h, ALTRO_ERR_CUDA, cudaGetErrorString(e_)); \
    } while (0)

extern "C" {

int altro_default_options(altro_opts_t *o)
{
    if (!o) return ALTRO_ERR_INVALID;
    o->constraint_tolerance = 1e-6;
    o->cost_tolerance = 1e-4;
    o->cost_tolerance_intermediate = 1e-4;
    o->gradient_tolerance = 10.0;
    o->gradient_tolerance_intermediate = 1.0;
    o->penalty_initial = 1.0;
    o->penalty_scaling = 10.0;
    o->penalty_max = 1e8;
    o->dual_max = 1e8;
    o->line_search_lower_bound = 1e-8;
    o->line_search_upper_bound = 10.0;
    o->max_cost_value = 1e8;
    o->max_state_value = 1e8;
    o->bp_reg_initial = 0.0;
    o->bp_reg_increase_factor = 1.6;
    o->bp_reg_max = 1e8;
    o->bp_reg_min = 1e-8;
    o->bp_reg_fp = 10.0;
    o->iterations = 1000;
    o->iterations_inner = 300;
    o->iterations_outer = 30;
    o->iterations_linesearch = 20;
    o->dJ_counter_limit = 10;
    o->reset_duals = 1;
    o->reset_penalties = 1;
    o->kickout_max_penalty = 0;
    o->dj_zero_converges = 1;
    o->first_step_unconditional = 1;
    o->soc_hess_exact = 1;
    o->soc_viol_proj = 1;
    return ALTRO_OK;
}

const char *altro_last_error(altro_handle_t h) { return h ? h->err.c_str() : g_err.c_str(); }

int altro_create(altro_handle_t *out, int device, int n, int m, int N, int batch, double dt)
{
    if (!out || n < 1 || m < 1 || N < 2 || batch < 1) return fail(nullptr, ALTRO_ERR_INVALID, "bad dimensions");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, ALTRO_ERR_CUDA, std::string("no CUDA device (this library has no CPU fallback): ") +
                                                 cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(nullptr, ALTRO_ERR_INVALID, "bad device index");
    altro_handle_t h = new altro_handle_s();
    h->device = device; h->n = n; h->m = m; h->N = N; h->B = batch; h->dt = dt;
    altro_default_options(&h->opts);
#define CKC(call)                                                                              \
    do {                                                                                       \
        cudaError_t e2_ = (call);                                                              \
        if (e2_ != cudaSuccess) {                                                              \
            fail(nullptr, ALTRO_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e2_)); \
            altro_destroy(h);                                                                  \
            return ALTRO_ERR_CUDA;                                                             \
        }                                                                                      \
    } while (0)
    CKC(cudaSetDevice(device));
    CKC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CKC(cudaEventCreate(&h->ev0));
    CKC(cudaEventCreate(&h->ev1));
    const size_t B = batch;
    CKC(dalloc(&h->Q, n)); CKC(dalloc(&h->R, m)); CKC(dalloc(&h->Qf, n));
    CKC(dalloc(&h->xref, B * N * n)); CKC(dalloc(&h->uref, B * (N - 1) * m)); CKC(dalloc(&h->x0, B * n));
    CKC(dalloc(&h->X, B * N * n)); CKC(dalloc(&h->U, B * (N - 1) * m));
    CKC(dalloc(&h->X_snap, B * N * n)); CKC(dalloc(&h->U_snap, B * (N - 1) * m));
    CKC(dalloc(&h->x0_snap, B * n)); CKC(dalloc(&h->xref_snap, B * N * n)); CKC(dalloc(&h->uref_snap, B * (N - 1) * m));
    CKC(dalloc(&h->kidx_snap, B));
    CKC(dalloc(&h->iters, B)); CKC(dalloc(&h->outer, B)); CKC(dalloc(&h->status, B)); CKC(dalloc(&h->trials, B));
    CKC(dalloc(&h->cost, B)); CKC(dalloc(&h->cost_al, B)); CKC(dalloc(&h->cmax, B)); CKC(dalloc(&h->penmax, B));
    CKC(dalloc(&h->t_ns, B)); CKC(dalloc(&h->noise, B * n)); CKC(dalloc(&h->kidx, B));
#undef CKC
    *out = h;
    return ALTRO_OK;
}

int altro_destroy(altro_handle_t h)
{
    if (!h) return ALTRO_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    void *ptrs[] = {h->A, h->Bm, h->d, h->Q, h->R, h->Qf, h->xref, h->uref, h->x0, h->X, h->U, h->lam, h->X_snap,
                    h->U_snap, h->lam_snap, h->x0_snap, h->xref_snap, h->uref_snap, h->kidx_snap, h->iters, h->outer, h->status, h->trials, h->cost, h->cost_al, h->cmax,
                    h->penmax, h->t_ns, h->x0_log, h->u0_log, h->phase, h->trace, h->con_dev, h->itab_dev, h->sched, h->ex_glob, h->ws, h->lane_ws, h->q_ctrl, h->admm_ws, h->admm_X, h->admm_U, h->admm_rp, h->admm_rd, h->admm_it, h->q_xref, h->q_uref, h->q_foot, h->q_contacts, h->q_t, h->q_curfoot, h->q_planner, h->trackX, h->trackU, h->noise, h->noise_bank, h->kidx};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    for (auto &c : h->cons) {
        if (c.G_dev) cudaFree(c.G_dev);
        if (c.h_dev) cudaFree(c.h_dev);
        if (c.rs_col_dev) cudaFree(c.rs_col_dev);
        if (c.rs_coef_dev) cudaFree(c.rs_coef_dev);
    }
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream && h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
    return ALTRO_OK;
}

int altro_set_stream(altro_handle_t h, void *s)
{
    REQ(h);
    CK(h, cudaStreamSynchronize(h->stream));
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)s;
    h->own_stream = false;
    return ALTRO_OK;
}

int altro_set_options(altro_handle_t h, const altro_opts_t *o)
{
    REQ(h);
    if (!o) return fail(h, ALTRO_ERR_INVALID, "null options");
    if (o->iterations_outer < 1 || o->iterations_inner < 1 || !(o->penalty_initial > 0.0))
        return fail(h, ALTRO_ERR_INVALID, "bad option values");
    if (!o->reset_penalties)
        return fail(h, ALTRO_ERR_UNSUPPORTED, "reset_penalties = false is not supported: penalties are not kept across "
                                              "solves (every benchmark of the reference leaves the default, true)");
    h->opts = *o;
    return ALTRO_OK;
}

int altro_set_dynamics(altro_handle_t h, int per_knot, int per_instance, const double *A, const double *Bm,
                       const double *d)
{
    REQ(h);
    if (!A || !Bm) return fail(h, ALTRO_ERR_INVALID, "null dynamics");
    const size_t cnt = (per_instance ? (size_t)h->B : 1) * (per_knot ? (size_t)(h->N - 1) : 1);
    if (h->have_dyn && (cnt != h->dyn_count || per_knot != h->dyn_per_knot || per_instance != h->dyn_per_instance)) {
        if (h->finalized) return fail(h, ALTRO_ERR_STATE, "dynamics layout cannot change after the first solve");
        cudaFree(h->A); cudaFree(h->Bm); cudaFree(h->d);
        h->A = h->Bm = h->d = nullptr;
        h->have_dyn = false;
    }
    const size_t n = h->n, m = h->m;
    if (!h->have_dyn) {
        CK(h, dalloc(&h->A, cnt * n * n));
        CK(h, dalloc(&h->Bm, cnt * n * m));
        CK(h, dalloc(&h->d, cnt * n));
        h->dyn_count = cnt; h->dyn_per_knot = per_knot; h->dyn_per_instance = per_instance;
        h->have_dyn = true;
    }
    if (cnt == 1) {  // shared LTI model: host copy for the lane kernel's constant parameters
        h->hA.assign(A, A + n * n);
        h->hB.assign(Bm, Bm + n * m);
        if (d) h->hd.assign(d, d + n);
        else h->hd.assign(n, 0.0);
    } else {
        h->hA.clear();
    }
    int rc = upload(h, h->A, A, cnt * n * n);
    if (rc) return rc;
    rc = upload(h, h->Bm, Bm, cnt * n * m);
    if (rc) return rc;
    if (d) return upload(h, h->d, d, cnt * n);
    CK(h, cudaMemsetAsync(h->d, 0, cnt * n * sizeof(double), h->stream));
    return ALTRO_OK;
}

int altro_set_dynamics_slots(altro_handle_t h, int nslots, const double *A, const double *Bm, const double *d,
                             const int *sched, int sched_len)
{
    REQ(h);
    if (!A || !Bm || !sched || nslots < 1 || sched_len < h->N) return fail(h, ALTRO_ERR_INVALID, "bad dynamics schedule");
    if (h->finalized && !h->sched) return fail(h, ALTRO_ERR_STATE, "dynamics layout cannot change after the first solve");
    const size_t cnt = (size_t)h->B * nslots, n = h->n, m = h->m;
    if (h->have_dyn && (cnt != h->dyn_count || !h->sched)) {
        cudaFree(h->A); cudaFree(h->Bm); cudaFree(h->d);
        h->A = h->Bm = h->d = nullptr;
        h->have_dyn = false;
    }
    if (!h->have_dyn) {
        CK(h, dalloc(&h->A, cnt * n * n));
        CK(h, dalloc(&h->Bm, cnt * n * m));
        CK(h, dalloc(&h->d, cnt * n));
        h->dyn_count = cnt; h->dyn_per_knot = 1; h->dyn_per_instance = 1;
        h->have_dyn = true;
    }
    if (h->sched && sched_len != h->sched_len) { cudaFree(h->sched); h->sched = nullptr; }
    if (!h->sched) CK(h, dalloc(&h->sched, (size_t)h->B * sched_len));
    h->dyn_slots = nslots; h->sched_len = sched_len;
    for (size_t i = 0; i < (size_t)h->B * sched_len; ++i)
        if (sched[i] < 0 || sched[i] >= nslots) return fail(h, ALTRO_ERR_INVALID, "schedule entry out of range");
    CK(h, cudaMemcpyAsync(h->sched, sched, (size_t)h->B * sched_len * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    int rc = upload(h, h->A, A, cnt * n * n);
    if (!rc) rc = upload(h, h->Bm, Bm, cnt * n * m);
    if (rc) return rc;
    if (d) return upload(h, h->d, d, cnt * n);
    CK(h, cudaMemsetAsync(h->d, 0, cnt * n * sizeof(double), h->stream));
    return ALTRO_OK;
}

int altro_set_cost_diag(altro_handle_t h, const double *Q, const double *R, const double *Qf)
{
    REQ(h);
    if (!Q || !R || !Qf) return fail(h, ALTRO_ERR_INVALID, "null weights");
    for (int i = 0; i < h->m; ++i)
        if (!(R[i] > 0.0)) return fail(h, ALTRO_ERR_INVALID, "R must be positive");
    h->hQ.assign(Q, Q + h->n);
    h->hR.assign(R, R + h->m);
    h->hQf.assign(Qf, Qf + h->n);
    int rc = upload(h, h->Q, Q, h->n);
    if (!rc) rc = upload(h, h->R, R, h->m);
    if (!rc) rc = upload(h, h->Qf, Qf, h->n);
    // the three vectors are tiny and caller-owned: make sure they are consumed before returning
    if (!rc) CK(h, cudaStreamSynchronize(h->stream));
    h->have_cost = true;
    return rc;
}

int altro_set_reference(altro_handle_t h, const double *Xref, const double *Uref)
{
    REQ(h);
    int rc = ALTRO_OK;
    if (Xref) rc = upload(h, h->xref, Xref, (size_t)h->B * h->N * h->n);
    if (!rc && Uref) rc = upload(h, h->uref, Uref, (size_t)h->B * (h->N - 1) * h->m);
    h->have_ref = true;
    return rc;
}

int altro_add_constraint(altro_handle_t h, int sense, int side, int k0, int k1, int p, int w, const int *inds,
                         int per_knot, int per_instance, const double *G, const double *hv, int *con_id)
{
    REQ(h);
    if (h->finalized) return fail(h, ALTRO_ERR_STATE, "constraints must be added before the first solve");
    if ((int)h->cons.size() >= MAX_CON) return fail(h, ALTRO_ERR_UNSUPPORTED, "too many constraint blocks");
    if (sense < 0 || sense > 2 || side < 0 || side > 1 || !inds || !G || !hv || p < 1 || w < 1)
        return fail(h, ALTRO_ERR_INVALID, "bad constraint arguments");
    const int kmax = side == ALTRO_CONTROL ? h->N - 1 : h->N;
    if (k0 < 0 || k1 > kmax || k1 <= k0) return fail(h, ALTRO_ERR_INVALID, "bad knot range (control blocks end at N-1)");
    if (w > MAX_W) return fail(h, ALTRO_ERR_UNSUPPORTED, "constraint index set wider than 32");
    const int lim = side == ALTRO_CONTROL ? h->m : h->n;
    for (int j = 0; j < w; ++j) {
        if (inds[j] < 0 || inds[j] >= lim) return fail(h, ALTRO_ERR_INVALID, "constraint index out of range");
        for (int q = 0; q < j; ++q)
            if (inds[q] == inds[j]) return fail(h, ALTRO_ERR_INVALID, "duplicate index in a constraint block");
    }
    if (sense == ALTRO_SECOND_ORDER_CONE && p < 2) return fail(h, ALTRO_ERR_INVALID, "a cone block needs >= 2 rows");
    HostCon c;
    c.sense = sense; c.side = side; c.k0 = k0; c.k1 = k1; c.p = p; c.w = w;
    c.per_knot = per_knot; c.per_instance = per_instance;
    c.inds.assign(inds, inds + w);
    const size_t cnt = (per_instance ? (size_t)h->B : 1) * (per_knot ? (size_t)(k1 - k0) : 1);
    c.g_count = cnt * p * w;
    c.h_count = cnt * p;
    // row-sparse detection (bounds): shared data, not a cone, every row has at most one nonzero
    c.rowsparse = 0;
    if (!per_knot && !per_instance && sense != ALTRO_SECOND_ORDER_CONE) {
        bool rs = true;
        std::vector<int> col(p, 0);
        std::vector<double> coef(p, 0.0);
        for (int r = 0; r < p && rs; ++r) {
            int nz = 0;
            for (int j = 0; j < w; ++j)
                if (G[r * w + j] != 0.0) { ++nz; col[r] = j; coef[r] = G[r * w + j]; }
            rs = nz <= 1;
        }
        if (rs) { c.rowsparse = 1; c.rs_col = col; c.rs_coef = coef; }
    }
    if (!c.rowsparse && (p > PMAX || w > DENSE_W))
        return fail(h, ALTRO_ERR_UNSUPPORTED, "dense constraint block larger than 8 rows x 8 indices");
    CK(h, dalloc(&c.G_dev, c.g_count));
    CK(h, dalloc(&c.h_dev, c.h_count));
    CK(h, copy_sync(c.G_dev, G, c.g_count * sizeof(double), cudaMemcpyHostToDevice));
    CK(h, copy_sync(c.h_dev, hv, c.h_count * sizeof(double), cudaMemcpyHostToDevice));
    if (c.rowsparse) {
        CK(h, dalloc(&c.rs_col_dev, p));
        CK(h, dalloc(&c.rs_coef_dev, p));
        CK(h, copy_sync(c.rs_col_dev, c.rs_col.data(), p * sizeof(int), cudaMemcpyHostToDevice));
        CK(h, copy_sync(c.rs_coef_dev, c.rs_coef.data(), p * sizeof(double), cudaMemcpyHostToDevice));
    }
    h->cons.push_back(c);
    if (con_id) *con_id = (int)h->cons.size() - 1;
    return ALTRO_OK;
}

int altro_add_track_constraint(altro_handle_t h, int sense, int side, int k0, int k1, int p, int w, const int *inds,
                               const double *G, const double *hv, int Nt, int *con_id)
{
    if (Nt < 1) return fail(h, ALTRO_ERR_INVALID, "empty constraint timeline");
    // registered as a shared block with Nt "knots" of data, then flagged as a timeline
    REQ(h);
    const int kmax = side == ALTRO_CONTROL ? h->N - 1 : h->N;
    if (k0 < 0 || k1 > kmax || k1 <= k0) return fail(h, ALTRO_ERR_INVALID, "bad knot range (control blocks end at N-1)");
    int id = -1;
    // temporarily widen the range so that add_constraint sizes the data as [Nt][p][w]
    const int saveN = h->N;
    h->N = Nt + 1 + k0;
    int rc = altro_add_constraint(h, sense, side, k0, k0 + Nt, p, w, inds, 1, 0, G, hv, &id);
    h->N = saveN;
    if (rc) return rc;
    HostCon &c = h->cons[id];
    c.k1 = k1;
    c.per_knot = 0;
    c.track = Nt;
    c.rowsparse = 0;
    if (p > PMAX || w > DENSE_W) return fail(h, ALTRO_ERR_UNSUPPORTED, "dense constraint block larger than 8 rows x 8 indices");
    if (con_id) *con_id = id;
    return ALTRO_OK;
}

int altro_set_track_index(altro_handle_t h, const int *kidx)
{
    REQ(h);
    if (!kidx) return fail(h, ALTRO_ERR_INVALID, "null index array");
    CK(h, cudaMemcpyAsync(h->kidx, kidx, (size_t)h->B * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return ALTRO_OK;
}

int altro_update_constraint_data(altro_handle_t h, int id, const double *G, const double *hv)
{
    REQ(h);
    if (id < 0 || id >= (int)h->cons.size()) return fail(h, ALTRO_ERR_INVALID, "bad constraint id");
    HostCon &c = h->cons[id];
    if (c.rowsparse && G) return fail(h, ALTRO_ERR_UNSUPPORTED, "G of a row-sparse (bound) block is fixed; update h only");
    int rc = ALTRO_OK;
    if (G) rc = upload(h, c.G_dev, G, c.g_count);
    if (!rc && hv) rc = upload(h, c.h_dev, hv, c.h_count);
    return rc;
}

int altro_set_x0(altro_handle_t h, const double *x0)
{
    REQ(h);
    if (!x0) return fail(h, ALTRO_ERR_INVALID, "null x0");
    h->have_x0 = true;
    return upload(h, h->x0, x0, (size_t)h->B * h->n);
}

int altro_set_trajectory(altro_handle_t h, const double *X, const double *U)
{
    REQ(h);
    int rc = ALTRO_OK;
    if (X) rc = upload(h, h->X, X, (size_t)h->B * h->N * h->n);
    if (!rc && U) rc = upload(h, h->U, U, (size_t)h->B * (h->N - 1) * h->m);
    return rc;
}

int altro_get_trajectory(altro_handle_t h, double *X, double *U)
{
    REQ(h);
    int rc = ALTRO_OK;
    if (X) rc = download(h, X, h->X, (size_t)h->B * h->N * h->n * sizeof(double));
    if (!rc && U) rc = download(h, U, h->U, (size_t)h->B * (h->N - 1) * h->m * sizeof(double));
    if (!rc) CK(h, cudaStreamSynchronize(h->stream));
    return rc;
}

int altro_dual_len(altro_handle_t h, int *P)
{
    REQ(h);
    int rc = finalize(h);
    if (rc) return rc;
    if (P) *P = h->P;
    return ALTRO_OK;
}

int altro_set_duals(altro_handle_t h, const double *lam)
{
    REQ(h);
    int rc = finalize(h);
    if (rc) return rc;
    return upload(h, h->lam, lam, (size_t)h->B * h->P);
}

int altro_get_duals(altro_handle_t h, double *lam)
{
    REQ(h);
    int rc = finalize(h);
    if (rc) return rc;
    rc = download(h, lam, h->lam, (size_t)h->B * h->P * sizeof(double));
    if (!rc) CK(h, cudaStreamSynchronize(h->stream));
    return rc;
}

int altro_shift_fill(altro_handle_t h, int primal, int dual)
{
    REQ(h);
    int rc = finalize(h);
    if (rc) return rc;
    const size_t words = std::max<size_t>({(size_t)h->N * h->n, (size_t)h->N * h->m, (size_t)h->P, 1});
    if (words * sizeof(double) > 48 * 1024)
        CK(h, cudaFuncSetAttribute(shift_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)(words * sizeof(double))));
    shift_fill_kernel<<<h->B, 128, words * sizeof(double), h->stream>>>(h->n, h->m, h->N, h->P, (int)h->cons.size(),
                                                                      h->con_dev, h->X, h->U, h->lam, primal, dual);
    CK(h, cudaGetLastError());
    return ALTRO_OK;
}

static int ensure_stat_capacity(altro_handle_t h, int slots)
{
    if (slots <= h->stat_cap) return ALTRO_OK;
    CK(h, cudaStreamSynchronize(h->stream));
    const size_t cnt = (size_t)slots * h->B;
    int **ip[] = {&h->iters, &h->outer, &h->status, &h->trials};
    double **dp[] = {&h->cost, &h->cost_al, &h->cmax, &h->penmax};
    for (auto p : ip) { cudaFree(*p); *p = nullptr; CK(h, dalloc(p, cnt)); }
    for (auto p : dp) { cudaFree(*p); *p = nullptr; CK(h, dalloc(p, cnt)); }
    cudaFree(h->t_ns); h->t_ns = nullptr;
    CK(h, dalloc(&h->t_ns, cnt));
    if (h->x0_log) cudaFree(h->x0_log);
    if (h->u0_log) cudaFree(h->u0_log);
    h->x0_log = h->u0_log = nullptr;
    CK(h, dalloc(&h->x0_log, cnt * h->n));
    CK(h, dalloc(&h->u0_log, cnt * h->m));
    h->stat_cap = slots;
    return ALTRO_OK;
}

static int launch_solve(altro_handle_t h, int steps, int shift)
{
    Params P;
    memset(&P, 0, sizeof(P));
    P.n = h->n; P.m = h->m; P.N = h->N; P.B = h->B; P.P = h->P; P.ncon = (int)h->cons.size(); P.EX = h->EX; P.ITAB = h->ITAB; P.NSRC = h->NSRC; P.NTL = h->NTL;
    P.inst_offset = 0;
    P.dt = h->dt;
    P.dyn_per_knot = h->dyn_per_knot; P.dyn_per_instance = h->dyn_per_instance; P.dyn_in_smem = h->dyn_in_smem; P.ref_in_smem = h->ref_in_smem;
    P.A = h->A; P.Bm = h->Bm; P.d = h->d;
    P.Q = h->Q; P.R = h->R; P.Qf = h->Qf;
    P.xref = h->xref; P.uref = h->uref; P.x0 = h->x0;
    P.X = h->X; P.U = h->U; P.lam = h->lam;
    P.iters = h->iters; P.outer = h->outer; P.status = h->status; P.trials = h->trials;
    P.cost = h->cost; P.cost_al = h->cost_al; P.cmax = h->cmax; P.penmax = h->penmax;
    P.t_ns = h->t_ns;
    P.trace = h->trace;
    P.trace_rows = h->trace_rows;
    P.con = h->con_dev;
    P.itab = h->itab_dev;
    P.lay = h->lay;
    P.spec = h->spec;
    P.o = h->opts;
    P.dyn_slots = h->dyn_slots; P.sched_len = h->sched_len; P.step0 = h->step_abs; P.dyn_sched = h->sched;
    P.steps = steps; P.shift = shift; P.noise_mode = h->noise_mode; P.Nt = h->Nt;
    P.noise_w1 = h->noise_w1; P.noise_w2 = h->noise_w2;
    if (steps > 0) {
        if (h->noise_bank) {
            if (steps > h->bank_steps) return fail(h, ALTRO_ERR_INVALID, "noise bank holds fewer steps than requested");
            if (h->bank_pos % h->bank_steps + steps > h->bank_steps) h->bank_pos = 0;  // keep one run contiguous
            P.noise = h->noise_bank + (size_t)(h->bank_pos % h->bank_steps) * h->B * h->n;
            h->bank_pos += steps;
        }
        P.trackX = h->trackX; P.trackU = h->trackU;
        P.x0_log = h->x0_log; P.u0_log = h->u0_log;
    }
    P.kidx = h->kidx;
    P.ex_glob = h->ex_glob;
    P.ws = h->ws;
    P.phase = h->phase;
    P.phase_detail = getenv("ALTRO_B200_PHASE_DETAIL") ? 1 : 0;
    void *args[] = {&P};
    // closed-loop runs of the shared-memory-resident CTA kernels go through the persistent grid + work queue
    int chunk = h->run_chunk;
    if (const char *e = getenv("ALTRO_B200_QUEUE")) chunk = atoi(e);
    const bool lane = h->lane_kern && !h->trace && !h->phase;
    const bool queued = steps > 0 && chunk > 0 && !lane && h->threads > 32 && !h->trace &&
                        (long long)h->B * ((steps + chunk - 1) / chunk) < (1ll << 31);
    int grid = h->B;
    if (queued) {
        if (!h->q_ctrl) CK(h, dalloc(&h->q_ctrl, (size_t)h->B + 2));
        CK(h, cudaMemsetAsync(h->q_ctrl, 0, ((size_t)h->B + 2) * sizeof(int), h->stream));
        P.q_chunk = chunk; P.q_head = h->q_ctrl; P.q_error = h->q_ctrl + 1; P.q_done = h->q_ctrl + 2;
        grid = std::min(h->B, std::max(1, h->ctas_per_sm) * h->num_sms);
    }
    if (h->trace)
        CK(h, cudaMemsetAsync(h->trace, 0, (size_t)h->B * h->trace_rows * TRACE_COLS * sizeof(double), h->stream));
    CK(h, cudaEventRecord(h->ev0, h->stream));
    if (lane) {  // (the per-iteration trace and the phase counters are CTA-kernel features)
        LaneLaunch a{};
        a.P = &P;
        a.A = h->hA.data(); a.B = h->hB.data(); a.d = h->hd.data(); a.Q = h->hQ.data(); a.R = h->hR.data(); a.Qf = h->hQf.data();
        a.L = h->lane_lay; a.ws = h->lane_ws; a.stride = h->lane_stride; a.lpw = h->lane_lpw;
        a.scratch_per_lane = h->lane_scratch; a.smem = (size_t)h->lane_smem; a.stream = h->stream;
        CK(h, lane_launch(h->n, h->m, a));
    } else {
        CK(h, cudaLaunchKernel(h->kernel, dim3(grid), dim3(h->threads), args, (size_t)h->smem, h->stream));
    }
    CK(h, cudaEventRecord(h->ev1, h->stream));
    h->step_abs += steps;
    if (steps > 0) {
        advance_kidx_kernel<<<(h->B + 255) / 256, 256, 0, h->stream>>>(h->kidx, h->B, steps);
        CK(h, cudaGetLastError());
    }
    return ALTRO_OK;
}

int altro_solve(altro_handle_t h)
{
    REQ(h);
    int rc = finalize(h);
    if (rc) return rc;
    if (!h->have_x0) return fail(h, ALTRO_ERR_STATE, "x0 not set");
    rc = launch_solve(h, 0, 0);
    if (!rc) { h->last_run_steps = 0; h->pending_transition = false; }
    return rc;
}

int altro_mpc_run(altro_handle_t h, int steps, int shift)
{
    REQ(h);
    int rc = finalize(h);
    if (rc) return rc;
    if (steps < 1) return fail(h, ALTRO_ERR_INVALID, "steps must be >= 1");
    if (h->pending_transition)  // the run's first transition takes x0 from X[1]: after altro_mpc_transition that would skip a state
        return fail(h, ALTRO_ERR_STATE, "altro_mpc_run after altro_mpc_transition: call altro_solve first");
    rc = ensure_stat_capacity(h, steps);
    if (rc) return rc;
    h->have_x0 = true;
    rc = launch_solve(h, steps, shift);
    if (!rc) h->last_run_steps = steps;
    return rc;
}

int altro_reserve_steps(altro_handle_t h, int steps)
{
    REQ(h);
    int rc = finalize(h);
    if (rc) return rc;
    if (steps < 1) return fail(h, ALTRO_ERR_INVALID, "steps must be >= 1");
    return ensure_stat_capacity(h, steps);
}

int altro_get_run_results(altro_handle_t h, int steps, int *iterations, int *iterations_outer, int *status,
                          int *ls_trials, double *cost, double *c_max, double *x0_log, double *u0_log,
                          long long *t_ns)
{
    REQ(h);
    if (steps < 1 || steps > h->last_run_steps) return fail(h, ALTRO_ERR_INVALID, "steps exceeds the last closed-loop run");
    if (h->q_ctrl) {
        int qerr = 0;
        CK(h, cudaMemcpyAsync(&qerr, h->q_ctrl + 1, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
        if (qerr) return fail(h, ALTRO_ERR_CUDA, "closed-loop run: a work item waited too long for its predecessor (run aborted)");
    }
    const size_t cnt = (size_t)steps * h->B;
    int rc = ALTRO_OK;
    if (iterations) rc |= download(h, iterations, h->iters, cnt * sizeof(int));
    if (iterations_outer) rc |= download(h, iterations_outer, h->outer, cnt * sizeof(int));
    if (status) rc |= download(h, status, h->status, cnt * sizeof(int));
    if (ls_trials) rc |= download(h, ls_trials, h->trials, cnt * sizeof(int));
    if (cost) rc |= download(h, cost, h->cost, cnt * sizeof(double));
    if (c_max) rc |= download(h, c_max, h->cmax, cnt * sizeof(double));
    if (x0_log && h->x0_log) rc |= download(h, x0_log, h->x0_log, cnt * h->n * sizeof(double));
    if (u0_log && h->u0_log) rc |= download(h, u0_log, h->u0_log, cnt * h->m * sizeof(double));
    if (t_ns) rc |= download(h, t_ns, h->t_ns, cnt * sizeof(long long));
    if (rc) return ALTRO_ERR_CUDA;
    CK(h, cudaStreamSynchronize(h->stream));
    return ALTRO_OK;
}

int altro_sync(altro_handle_t h)
{
    REQ(h);
    CK(h, cudaStreamSynchronize(h->stream));
    return ALTRO_OK;
}

int altro_get_stats(altro_handle_t h, int *iterations, int *iterations_outer, int *status, int *ls_trials,
                    double *cost, double *cost_al, double *c_max, double *penalty_max)
{
    REQ(h);
    const size_t B = h->B;
    const size_t at = h->last_run_steps > 0 ? (size_t)(h->last_run_steps - 1) * B : 0;  // after a run: its final solve
    int rc = ALTRO_OK;
    if (iterations) rc |= download(h, iterations, h->iters + at, B * sizeof(int));
    if (iterations_outer) rc |= download(h, iterations_outer, h->outer + at, B * sizeof(int));
    if (status) rc |= download(h, status, h->status + at, B * sizeof(int));
    if (ls_trials) rc |= download(h, ls_trials, h->trials + at, B * sizeof(int));
    if (cost) rc |= download(h, cost, h->cost + at, B * sizeof(double));
    if (cost_al) rc |= download(h, cost_al, h->cost_al + at, B * sizeof(double));
    if (c_max) rc |= download(h, c_max, h->cmax + at, B * sizeof(double));
    if (penalty_max) rc |= download(h, penalty_max, h->penmax + at, B * sizeof(double));
    if (rc) return ALTRO_ERR_CUDA;
    CK(h, cudaStreamSynchronize(h->stream));
    return ALTRO_OK;
}

int altro_get_timing(altro_handle_t h, double *device_ms, long long *per_instance_ns)
{
    REQ(h);
    CK(h, cudaStreamSynchronize(h->stream));
    if (device_ms) {
        float ms = 0.f;
        CK(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        *device_ms = ms;
    }
    if (per_instance_ns) {
        const size_t at = h->last_run_steps > 0 ? (size_t)(h->last_run_steps - 1) * h->B : 0;
        int rc = download(h, per_instance_ns, h->t_ns + at, (size_t)h->B * sizeof(long long));
        if (rc) return rc;
        CK(h, cudaStreamSynchronize(h->stream));
    }
    return ALTRO_OK;
}

int altro_set_trace(altro_handle_t h, int max_rows)
{
    REQ(h);
    CK(h, cudaStreamSynchronize(h->stream));
    if (h->trace) { cudaFree(h->trace); h->trace = nullptr; }
    h->trace_rows = 0;
    if (max_rows > 0) {
        CK(h, dalloc(&h->trace, (size_t)h->B * max_rows * TRACE_COLS));
        h->trace_rows = max_rows;
    }
    return ALTRO_OK;
}

int altro_get_trace(altro_handle_t h, double *out)
{
    REQ(h);
    if (!h->trace || !out) return fail(h, ALTRO_ERR_STATE, "tracing is not enabled");
    int rc = download(h, out, h->trace, (size_t)h->B * h->trace_rows * TRACE_COLS * sizeof(double));
    if (rc) return rc;
    CK(h, cudaStreamSynchronize(h->stream));
    return ALTRO_OK;
}

int altro_get_phase_cycles(altro_handle_t h, int enable, long long *out)
{
    REQ(h);
    CK(h, cudaStreamSynchronize(h->stream));
    if (out && h->phase) CK(h, copy_sync(out, h->phase, (size_t)h->B * 8 * sizeof(long long), cudaMemcpyDeviceToHost));
    if (enable && !h->phase) CK(h, dalloc(&h->phase, (size_t)h->B * 8));
    if (enable) CK(h, cudaMemsetAsync(h->phase, 0, (size_t)h->B * 8 * sizeof(long long), h->stream));
    if (!enable && h->phase) { cudaFree(h->phase); h->phase = nullptr; }
    return ALTRO_OK;
}

// (X, U, duals, x0, reference window, track index, noise-bank position): everything a closed-loop run mutates
static int copy_state(altro_handle_t h, bool save)
{
    const size_t B = h->B, n = h->n, m = h->m, N = h->N;
    struct { void *live, *snap; size_t bytes; } v[] = {
        {h->X, h->X_snap, B * N * n * sizeof(double)},       {h->U, h->U_snap, B * (N - 1) * m * sizeof(double)},
        {h->lam, h->lam_snap, B * std::max(h->P, 1) * sizeof(double)}, {h->x0, h->x0_snap, B * n * sizeof(double)},
        {h->xref, h->xref_snap, B * N * n * sizeof(double)}, {h->uref, h->uref_snap, B * (N - 1) * m * sizeof(double)},
        {h->kidx, h->kidx_snap, B * sizeof(int)}};
    for (auto &e : v)
        CK(h, cudaMemcpyAsync(save ? e.snap : e.live, save ? e.live : e.snap, e.bytes, cudaMemcpyDeviceToDevice, h->stream));
    if (save) { h->bank_pos_snap = h->bank_pos; h->step_abs_snap = h->step_abs; }
    else { h->bank_pos = h->bank_pos_snap; h->step_abs = h->step_abs_snap; }
    return ALTRO_OK;
}

int altro_snapshot(altro_handle_t h)
{
    REQ(h);
    int rc = finalize(h);
    if (rc) return rc;
    return copy_state(h, true);
}

int altro_restore(altro_handle_t h)
{
    REQ(h);
    int rc = finalize(h);
    if (rc) return rc;
    return copy_state(h, false);
}

int altro_set_track(altro_handle_t h, const double *tX, const double *tU, int Nt, const int *k_start)
{
    REQ(h);
    if (!tX || !tU || Nt < 2 || !k_start) return fail(h, ALTRO_ERR_INVALID, "bad track");
    if (h->trackX) { cudaFree(h->trackX); cudaFree(h->trackU); h->trackX = h->trackU = nullptr; }
    // device copies padded with N repeats of the last knot, so that every N-knot window is one contiguous slice
    std::vector<double> px((size_t)(Nt + h->N) * h->n), pu((size_t)(Nt - 1 + h->N) * h->m);
    for (int k = 0; k < Nt + h->N; ++k)
        memcpy(&px[(size_t)k * h->n], tX + (size_t)std::min(k, Nt - 1) * h->n, h->n * sizeof(double));
    for (int k = 0; k < Nt - 1 + h->N; ++k)
        memcpy(&pu[(size_t)k * h->m], tU + (size_t)std::min(k, Nt - 2) * h->m, h->m * sizeof(double));
    CK(h, dalloc(&h->trackX, px.size()));
    CK(h, dalloc(&h->trackU, pu.size()));
    CK(h, copy_sync(h->trackX, px.data(), px.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(h, copy_sync(h->trackU, pu.data(), pu.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(h, copy_sync(h->kidx, k_start, (size_t)h->B * sizeof(int), cudaMemcpyHostToDevice));
    h->Nt = Nt;
    return ALTRO_OK;
}

int altro_set_noise_model(altro_handle_t h, int mode, double w1, double w2)
{
    REQ(h);
    if (mode < 0 || mode > 2) return fail(h, ALTRO_ERR_INVALID, "noise mode must be 0, 1 or 2");
    h->noise_mode = mode; h->noise_w1 = w1; h->noise_w2 = w2;
    return ALTRO_OK;
}

int altro_get_x0(altro_handle_t h, double *x0)
{
    REQ(h);
    int rc = download(h, x0, h->x0, (size_t)h->B * h->n * sizeof(double));
    if (!rc) CK(h, cudaStreamSynchronize(h->stream));
    return rc;
}

int altro_set_noise_bank(altro_handle_t h, const double *noise, int steps)
{
    REQ(h);
    if (!noise || steps < 1) {
        CK(h, cudaStreamSynchronize(h->stream));
        if (h->noise_bank) { cudaFree(h->noise_bank); h->noise_bank = nullptr; }
        h->bank_steps = h->bank_pos = h->bank_cap = 0;
        return ALTRO_OK;
    }
    const size_t cnt = (size_t)steps * h->B * h->n;
    if (steps > h->bank_cap) {  // grow only: repeated uploads of the same size reuse the allocation
        CK(h, cudaStreamSynchronize(h->stream));
        if (h->noise_bank) { cudaFree(h->noise_bank); h->noise_bank = nullptr; }
        CK(h, dalloc(&h->noise_bank, cnt));
        h->bank_cap = steps;
    }
    h->bank_steps = steps;
    h->bank_pos = 0;
    return upload(h, h->noise_bank, noise, cnt);  // async on the handle's stream (DMA if the buffer is pinned)
}

int altro_mpc_transition(altro_handle_t h, const double *noise, int shift)
{
    REQ(h);
    int rc = finalize(h);
    if (rc) return rc;
    const double *nz = nullptr;
    if (noise) {
        rc = upload(h, h->noise, noise, (size_t)h->B * h->n);
        if (rc) return rc;
        nz = h->noise;
    } else if (h->noise_bank) {  // device-resident noise: no host traffic at all
        nz = h->noise_bank + (size_t)(h->bank_pos % h->bank_steps) * h->B * h->n;
        ++h->bank_pos;
    }
    mpc_transition_kernel<<<h->B, 64, 0, h->stream>>>(h->n, h->m, h->N, h->A, h->Bm, h->d, h->dyn_per_knot,
                                                     h->dyn_per_instance, h->sched, h->sched_len, h->dyn_slots,
                                                     h->step_abs, h->X, h->U, nz, h->noise_mode, h->noise_w1, h->noise_w2,
                                                     h->x0, h->trackX, h->trackU, h->Nt, h->kidx, h->xref, h->uref);
    CK(h, cudaGetLastError());
    h->have_x0 = true;
    h->pending_transition = true;
    h->step_abs += 1;
    if (!h->trackX) {  // the transition kernel advances kidx only together with the reference window
        advance_kidx_kernel<<<(h->B + 255) / 256, 256, 0, h->stream>>>(h->kidx, h->B, 1);
        CK(h, cudaGetLastError());
    }
    if (shift) return altro_shift_fill(h, 1, 1);
    return ALTRO_OK;
}

// ---- independent convex cross-check on the device (admm.cu)

int altro_admm_solve(altro_handle_t h, double rho, double eps, int max_iter, double *X, double *U, int *iterations,
                     double *r_prim, double *r_dual)
{
    REQ(h);
    int rc = finalize(h);
    if (rc) return rc;
    if (!(rho > 0.0) || !(eps > 0.0) || max_iter < 1) return fail(h, ALTRO_ERR_INVALID, "bad ADMM parameters");
    if (!h->have_x0) return fail(h, ALTRO_ERR_STATE, "x0 not set");
    for (auto &c : h->cons)
        if (c.sense == ALTRO_SECOND_ORDER_CONE && c.p > PMAX_ADMM) return fail(h, ALTRO_ERR_UNSUPPORTED, "cone with more than 8 rows");
    const size_t B = h->B, n = h->n, m = h->m, N = h->N;
    const size_t wsd = admm_workspace_doubles(h->n, h->m, h->N, h->P);
    if (!h->admm_ws) {
        CK(h, dalloc(&h->admm_ws, B * wsd));
        CK(h, dalloc(&h->admm_X, B * N * n)); CK(h, dalloc(&h->admm_U, B * (N - 1) * m));
        CK(h, dalloc(&h->admm_rp, B)); CK(h, dalloc(&h->admm_rd, B)); CK(h, dalloc(&h->admm_it, B));
    }
    AdmmParams P;
    memset(&P, 0, sizeof(P));
    P.n = h->n; P.m = h->m; P.N = h->N; P.B = h->B; P.Pd = h->P; P.ncon = (int)h->cons.size();
    P.dt = h->dt; P.rho = rho; P.eps = eps; P.max_iter = max_iter; P.adapt = 25;
    P.dyn_per_knot = h->dyn_per_knot; P.dyn_per_instance = h->dyn_per_instance; P.dyn_slots = h->dyn_slots;
    P.sched_len = h->sched_len; P.step0 = h->step_abs; P.dyn_sched = h->sched; P.kidx = h->kidx;
    P.A = h->A; P.Bm = h->Bm; P.d = h->d; P.Q = h->Q; P.R = h->R; P.Qf = h->Qf;
    P.xref = h->xref; P.uref = h->uref; P.x0 = h->x0; P.X = h->X; P.U = h->U;
    P.con = h->con_dev; P.ws = h->admm_ws; P.ws_doubles = wsd;
    P.Xout = h->admm_X; P.Uout = h->admm_U; P.rprim = h->admm_rp; P.rdual = h->admm_rd; P.iters = h->admm_it;
    CK(h, admm_launch(P, h->stream));
    if (X) rc |= download(h, X, h->admm_X, B * N * n * sizeof(double));
    if (U) rc |= download(h, U, h->admm_U, B * (N - 1) * m * sizeof(double));
    if (iterations) rc |= download(h, iterations, h->admm_it, B * sizeof(int));
    if (r_prim) rc |= download(h, r_prim, h->admm_rp, B * sizeof(double));
    if (r_dual) rc |= download(h, r_dual, h->admm_rd, B * sizeof(double));
    if (rc) return ALTRO_ERR_CUDA;
    CK(h, cudaStreamSynchronize(h->stream));
    return ALTRO_OK;
}

// ---- quadruped: the step before the solve path, on the device (quadruped.cu)

static int quad_prepare(altro_handle_t h)
{
    if (h->n != 12 || h->m != 12) return fail(h, ALTRO_ERR_INVALID, "quadruped kernels need n = m = 12");
    const size_t B = h->B, K = h->N - 1, cnt = B * K;
    if (h->have_dyn && (h->dyn_count != cnt || !h->dyn_per_knot || !h->dyn_per_instance || h->sched)) {
        if (h->finalized) return fail(h, ALTRO_ERR_STATE, "dynamics layout cannot change after the first solve");
        cudaFree(h->A); cudaFree(h->Bm); cudaFree(h->d);
        h->A = h->Bm = h->d = nullptr;
        h->have_dyn = false;
    }
    if (!h->have_dyn) {
        CK(h, dalloc(&h->A, cnt * 144));
        CK(h, dalloc(&h->Bm, cnt * 144));
        CK(h, dalloc(&h->d, cnt * 12));
        h->dyn_count = cnt; h->dyn_per_knot = 1; h->dyn_per_instance = 1;
        h->have_dyn = true;
        h->hA.clear();
    }
    if (!h->q_xref) {
        CK(h, dalloc(&h->q_xref, cnt * 12)); CK(h, dalloc(&h->q_uref, cnt * 12)); CK(h, dalloc(&h->q_foot, cnt * 12));
        CK(h, dalloc(&h->q_contacts, cnt * 4)); CK(h, dalloc(&h->q_t, B)); CK(h, dalloc(&h->q_curfoot, B * 12));
        CK(h, dalloc(&h->q_planner, B * 12));
    }
    return ALTRO_OK;
}

static int quad_body(altro_handle_t h, const double *J, double mass, QuadrupedBody *body)
{
    if (!J || !(mass > 0.0)) return fail(h, ALTRO_ERR_INVALID, "bad inertia / mass");
    for (int i = 0; i < 9; ++i) body->J[i] = J[i];
    const double *a = J;
    const double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
    if (!(fabs(det) > 0.0)) return fail(h, ALTRO_ERR_INVALID, "singular inertia");
    const double id = 1.0 / det;
    body->Jinv[0] = (a[4] * a[8] - a[5] * a[7]) * id; body->Jinv[1] = (a[2] * a[7] - a[1] * a[8]) * id; body->Jinv[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    body->Jinv[3] = (a[5] * a[6] - a[3] * a[8]) * id; body->Jinv[4] = (a[0] * a[8] - a[2] * a[6]) * id; body->Jinv[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    body->Jinv[6] = (a[3] * a[7] - a[4] * a[6]) * id; body->Jinv[7] = (a[1] * a[6] - a[0] * a[7]) * id; body->Jinv[8] = (a[0] * a[4] - a[1] * a[3]) * id;
    body->mass = mass;
    return ALTRO_OK;
}

int altro_quadruped_linearize(altro_handle_t h, const double *x_ref, int x_ref_per_knot, const double *u_ref,
                              int u_ref_per_knot, const double *foot, const double *contacts, const double *J,
                              double mass)
{
    REQ(h);
    if (!x_ref || !foot || !contacts) return fail(h, ALTRO_ERR_INVALID, "null quadruped inputs");
    int rc = quad_prepare(h);
    if (rc) return rc;
    QuadrupedBody body;
    rc = quad_body(h, J, mass, &body);
    if (rc) return rc;
    const size_t B = h->B, K = h->N - 1;
    rc = upload(h, h->q_xref, x_ref, B * (x_ref_per_knot ? K : 1) * 12);
    if (!rc && u_ref) rc = upload(h, h->q_uref, u_ref, B * (u_ref_per_knot ? K : 1) * 12);
    if (!rc) rc = upload(h, h->q_foot, foot, B * K * 12);
    if (!rc) rc = upload(h, h->q_contacts, contacts, B * K * 4);
    if (rc) return rc;
    CK(h, quadruped_linearize_launch((int)B, (int)K, h->q_xref, x_ref_per_knot, u_ref ? h->q_uref : nullptr, u_ref_per_knot,
                                     h->q_foot, h->q_contacts, body, h->dt, h->A, h->Bm, h->d, h->stream));
    return ALTRO_OK;
}

int altro_quadruped_tick(altro_handle_t h, const double *t, const double *x_ref, int x_ref_per_knot,
                         const double *cur_foot, int num_phases, const double *contact_phases,
                         const double *phase_times, double alpha, double foot_radius, const double *nom_foot,
                         const double *J, double mass)
{
    REQ(h);
    if (!t || !x_ref || !cur_foot || !contact_phases || !phase_times || !nom_foot || num_phases < 1 || num_phases > 8)
        return fail(h, ALTRO_ERR_INVALID, "bad gait description");
    int rc = quad_prepare(h);
    if (rc) return rc;
    QuadrupedBody body;
    rc = quad_body(h, J, mass, &body);
    if (rc) return rc;
    QuadrupedGait g;
    memset(&g, 0, sizeof(g));
    g.num_phases = num_phases;
    g.phase_length = 0.0;
    for (int i = 0; i < num_phases; ++i) {
        g.phase_times[i] = phase_times[i];
        g.phase_length += phase_times[i];
        for (int j = 0; j < 4; ++j) g.contact[i * 4 + j] = contact_phases[i * 4 + j];
    }
    g.alpha = alpha; g.foot_radius = foot_radius;
    for (int i = 0; i < 12; ++i) g.nom_foot[i] = nom_foot[i];
    const size_t B = h->B, K = h->N - 1;
    rc = upload(h, h->q_t, t, B);
    if (!rc) rc = upload(h, h->q_xref, x_ref, B * (x_ref_per_knot ? K : 1) * 12);
    if (!rc) rc = upload(h, h->q_curfoot, cur_foot, B * 12);
    if (rc) return rc;
    if (!h->q_planner_set) {  // planner_foot_loc starts at the current foot locations in the body frame (ControllerParams.jl:62-66)
        CK(h, cudaMemcpyAsync(h->q_planner, h->q_curfoot, B * 12 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        h->q_planner_set = true;
    }
    CK(h, quadruped_gait_launch((int)B, (int)K, h->q_t, h->q_xref, x_ref_per_knot, h->q_curfoot, g, h->dt, h->q_planner,
                                h->q_contacts, h->q_foot, h->stream));
    CK(h, quadruped_linearize_launch((int)B, (int)K, h->q_xref, x_ref_per_knot, nullptr, 0, h->q_foot, h->q_contacts, body,
                                     h->dt, h->A, h->Bm, h->d, h->stream));
    return ALTRO_OK;
}

int altro_quadruped_get_schedule(altro_handle_t h, double *contacts, double *foot)
{
    REQ(h);
    if (!h->q_foot) return fail(h, ALTRO_ERR_STATE, "no quadruped schedule on the device yet");
    const size_t cnt = (size_t)h->B * (h->N - 1);
    int rc = ALTRO_OK;
    if (contacts) rc = download(h, contacts, h->q_contacts, cnt * 4 * sizeof(double));
    if (!rc && foot) rc = download(h, foot, h->q_foot, cnt * 12 * sizeof(double));
    if (!rc) CK(h, cudaStreamSynchronize(h->stream));
    return rc;
}

int altro_get_dynamics(altro_handle_t h, double *A, double *Bm, double *d)
{
    REQ(h);
    if (!h->have_dyn) return fail(h, ALTRO_ERR_STATE, "dynamics not set");
    const size_t n = h->n, m = h->m;
    int rc = ALTRO_OK;
    if (A) rc = download(h, A, h->A, h->dyn_count * n * n * sizeof(double));
    if (!rc && Bm) rc = download(h, Bm, h->Bm, h->dyn_count * n * m * sizeof(double));
    if (!rc && d) rc = download(h, d, h->d, h->dyn_count * n * sizeof(double));
    if (!rc) CK(h, cudaStreamSynchronize(h->stream));
    return rc;
}

int altro_host_register(void *ptr, size_t bytes)
{
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) return fail(nullptr, ALTRO_ERR_CUDA, cudaGetErrorString(e));
    return ALTRO_OK;
}

int altro_host_unregister(void *ptr)
{
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) return fail(nullptr, ALTRO_ERR_CUDA, cudaGetErrorString(e));
    return ALTRO_OK;
}

int altro_set_launch_config(altro_handle_t h, int threads)
{
    REQ(h);
    if (h->finalized) return fail(h, ALTRO_ERR_STATE, "launch configuration is fixed after the first solve");
    if (threads != 0 && threads != 32 && threads != 64 && threads != 128 && threads != 256)
        return fail(h, ALTRO_ERR_INVALID, "threads per instance must be 0 (auto), 32, 64, 128 or 256");
    h->threads_req = threads;
    return ALTRO_OK;
}

int altro_set_run_queue(altro_handle_t h, int steps_per_item)
{
    REQ(h);
    if (steps_per_item < 0) return fail(h, ALTRO_ERR_INVALID, "steps per work item must be >= 0 (0 = one CTA per instance)");
    h->run_chunk = steps_per_item;
    return ALTRO_OK;
}

int altro_set_kernel_mode(altro_handle_t h, int mode)
{
    REQ(h);
    if (h->finalized) return fail(h, ALTRO_ERR_STATE, "launch configuration is fixed after the first solve");
    if (mode < 0 || mode > 2) return fail(h, ALTRO_ERR_INVALID, "kernel mode must be 0 (auto), 1 (CTA per instance) or 2 (lane per instance)");
    h->kernel_mode = mode;
    return ALTRO_OK;
}

int altro_get_kernel_mode(altro_handle_t h, int *mode, int *lane_regs, int *lane_smem, int *lpw)
{
    REQ(h);
    int rc = finalize(h);
    if (rc) return rc;
    if (mode) *mode = h->lane_kern ? 2 : 1;
    if (lpw) *lpw = h->lane_lpw;
    if (lane_regs) *lane_regs = h->lane_regs;
    if (lane_smem) *lane_smem = h->lane_smem;
    return ALTRO_OK;
}

int altro_set_line_search_mode(altro_handle_t h, int speculative)
{
    REQ(h);
    if (h->finalized) return fail(h, ALTRO_ERR_STATE, "launch configuration is fixed after the first solve");
    h->spec_req = speculative < 0 ? -1 : (speculative ? 1 : 0);
    return ALTRO_OK;
}

int altro_get_line_search_mode(altro_handle_t h, int *speculative)
{
    REQ(h);
    int rc = finalize(h);
    if (rc) return rc;
    if (speculative) *speculative = h->spec;
    return ALTRO_OK;
}

int altro_get_launch_info(altro_handle_t h, int *threads, int *smem, int *regs, int *ctas, int *sms)
{
    REQ(h);
    int rc = finalize(h);
    if (rc) return rc;
    if (threads) *threads = h->threads;
    if (smem) *smem = h->smem;
    if (regs) *regs = h->regs;
    if (ctas) *ctas = h->ctas_per_sm;
    if (sms) *sms = h->num_sms;
    return ALTRO_OK;
}

int altro_measure_peaks(int device, double *dfma, double *dmma, double *copy_gbs)
{
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, ALTRO_ERR_CUDA, cudaGetErrorString(e));
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
    double *out = nullptr;
    if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)) != cudaSuccess)
        return fail(nullptr, ALTRO_ERR_CUDA, "cudaMalloc");
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float ms = 0.f;
    if (dfma) {
        double best = 0.0;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(a);
            dfma_peak_kernel<<<blocks, threads>>>(out, iters);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms, a, b);
            double fl = 2.0 * 8.0 * iters * (double)blocks * threads;
            if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
        }
        *dfma = best;
    }
    if (dmma) {
        double best = 0.0;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(a);
            dmma_peak_kernel<<<blocks, threads>>>(out, iters / 4);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms, a, b);
            double fl = 2.0 * 8 * 8 * 4 * 4.0 * (iters / 4) * (double)blocks * (threads / 32);
            if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
        }
        *dmma = best;
    }
    if (copy_gbs) {
        const size_t bytes = (size_t)1 << 30;
        char *s = nullptr, *d = nullptr;
        if (cudaMalloc(&s, bytes) == cudaSuccess && cudaMalloc(&d, bytes) == cudaSuccess) {
            double best = 0.0;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(a);
                cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice);
                cudaEventRecord(b);
                cudaEventSynchronize(b);
                cudaEventElapsedTime(&ms, a, b);
                if (rep > 0) best = std::max(best, 2.0 * bytes / (ms * 1e-3) / 1e9);
            }
            *copy_gbs = best;
        } else *copy_gbs = 0.0;
        if (s) cudaFree(s);
        if (d) cudaFree(d);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(out);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(nullptr, ALTRO_ERR_CUDA, cudaGetErrorString(e));
    return ALTRO_OK;
}

}  // extern "C"
