// altro_kernels.cuh -- batched AL-iLQR (ALTRO) solve for sm_100a: one CTA per MPC instance.
//
// The whole solve! of one instance -- AL outer loop, iLQR inner loop, Riccati backward pass,
// line-searched forward rollouts, AL cost / expansion, dual and penalty updates (SURVEY.md 8a rows
// a1-a10; reference call sites random_linear_problem.jl:161, simple_rocket.jl:174, altro_solver.jl:72,
// grasp_mpc.jl:55, flexible_sat_mpc.jl:272) -- runs inside ONE kernel launch with the instance's
// trajectories, gains, duals and Riccati blocks resident in shared memory.  No host round trip per
// iteration; instances never synchronise with each other, so divergent iteration counts cost nothing
// but their own CTA's time and the hardware block scheduler balances the tail.
//
// Template <NX, NU, T>: state / control dimension (0 = run-time value from Params) and threads per
// instance (32 = one warp, __syncwarp only; 64..256 = CTA with __syncthreads).
#pragma once
#include <type_traits>

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/altro_b200.h"

namespace altro {

constexpr int MAX_CON = 16;  // constraint blocks per problem
constexpr int MAX_W = 32;    // index-set width of one block
constexpr int PMAX = 8;      // rows of a dense block (row-sparse blocks are unlimited)
constexpr int DENSE_W = 8;   // index-set width of a dense block
constexpr int TRACE_COLS = 10;  // outer, iter, J, dJ, grad, rho, dV1, dV2, ls trials so far, c_max (NaN inside an outer)

// One affine conic block c = G z[inds] + h (device view).
struct alignas(16) ConDesc {
    int sense, side, k0, k1, p, w;
    int per_knot, per_instance;
    int rowsparse;  // every row of G has <= 1 nonzero (bounds): value = h[r] + rs_coef[r] * z[inds[rs_col[r]]]
    int dual_off;   // offset of the block in the per-instance dual vector
    int ex_off;     // offset of the block in the per-instance expansion scratch
    int ex_stride;  // per-knot stride there: w + w(w+1)/2 (dense, upper triangle) or 2*w (row-sparse)
    int tgt_off;    // (reserved)
    int track;      // > 0: G, h are a shared timeline [track][p][w]; knot k reads row min(kidx + k, track - 1)
    const double *G, *h;
    const int *rs_col;
    const double *rs_coef;
    int inds[MAX_W];
};

// Shared-memory layout of one instance: offsets in doubles from the start of the dynamic shared memory, computed
// once on the host and passed in Params, so that the kernel forms every array pointer as base + constant-bank
// operand instead of re-deriving a chain of run-time products (11 % of all executed instructions before).
struct Layout {
    int Qd, Qfd, Rd, sA, sB, sd, X, U, Xb, Ub, xr, ur, K, dv, lam, mu, ex, S, SA, Qxx, SB, Qux, T1, Quu, L, s, Qx, Qu,
        t1, linv, red, bc, itm, cand, Qi, cd;  // cd: offset (in doubles) of the ConDesc array, followed by the gather tables
    int bytes;                       // total dynamic shared memory
    int big, ws_doubles;             // large state dimension: the n-sized matrices live in a global workspace (offsets into it)
    int tma, stage, stage_doubles, mbar;  // large state dimension: TMA-staged operand panels (see Ctx::panel_gemm)
};

#ifndef ALTRO_T128_CTAS
#define ALTRO_T128_CTAS 4  // resident CTAs per SM the 128-thread small-dimension kernels are register-capped for
#endif
#ifndef ALTRO_GEN256_CTAS
#define ALTRO_GEN256_CTAS 2  // resident CTAs per SM the 256-thread run-time sized kernel is register-capped for
#endif
#ifndef ALTRO_FIXED_ALL_SMEM
#define ALTRO_FIXED_ALL_SMEM 1
#endif

// Arrays whose size depends only on (n, m) come first: in the kernels instantiated for a fixed <NX, NU> their
// offsets are compile-time constants and every access folds into base + immediate.
__host__ __device__ constexpr Layout fixed_layout(int n, int m)
{
    Layout l{};
    int q = 0;
    l.Qd = q; q += n; l.Qfd = q; q += n; l.Rd = q; q += m;
    l.sA = q; q += n * n; l.sB = q; q += n * m; l.sd = q; q += n;  // shared LTI model, or the LTV knot in flight
    l.S = q; q += n * n; l.SA = q; q += n * n; l.Qxx = q; q += n * n; l.SB = q; q += n * m; l.Qux = q; q += m * n;
    l.T1 = q; q += m * n; l.Quu = q; q += m * m; l.L = q; q += m * m;
    l.s = q; q += n; l.Qx = q; q += n; l.Qu = q; q += m; l.t1 = q; q += m; l.linv = q; q += m;
    l.Qi = q; q += n + n * n + m + m * m;  // [Qx | Qxx | Qu | Quu] of the next knot: cost + AL expansion, gathered ahead
    l.mu = q; q += MAX_CON; l.bc = q; q += 24; l.red = q; q += 9;  // 8 broadcast slots + 16 speculative line-search results; 8 warps + 1
    l.X = q;
    return l;
}

// Large state dimensions (n >= ~60: S alone no longer fits next to the trajectories): the matrices with an n-sized
// side -- S, SA, Qxx, SB, Qux, T1, the gains K and the gathered expansion Qi -- move to a per-instance workspace in
// global memory (L2-resident), the dynamics are read in place, the gather tables stay in global memory; vectors,
// trajectories, duals, the m x m blocks and the descriptors stay in shared memory.  Same code, other pointers.
// shared-memory row stride of a staged panel: rows whose length is a multiple of 8 doubles are padded by 4 so that the
// four k-rows a DMMA fragment load touches fall into different banks (stride = 4 mod 8 doubles)
__host__ __device__ constexpr int panel_ld(int ld) { return (ld % 8 == 0) ? ld + 4 : ld; }
constexpr int PANEL_KC = 8;   // k-rows per staged operand panel (one bulk copy per row and operand, one lane each)
#ifndef ALTRO_PANEL_NST
#define ALTRO_PANEL_NST 5
#endif
constexpr int PANEL_NST = ALTRO_PANEL_NST;  // stages in flight (measured: 8 x 5 beats 16 x 3 at the same bytes in flight)
constexpr int PANEL_WT = 7;   // DMMA tile columns a consumer warp owns (32 x 56 accumulators = 112 registers; 200 = 4 x 7 x 8 - 24)

__host__ __device__ inline Layout make_layout_big(int n, int m, int N, int P, int ncon, int EX, int tma = 1)
{
    Layout l{};
    int q = 0, w = 0;
    auto take = [&](int count) { int at = q; q += count; return at; };
    auto wtake = [&](int count) { int at = w; w += count; return at; };
    l.Qd = take(n); l.Qfd = take(n); l.Rd = take(m); l.sd = take(n);
    l.Quu = take(m * m); l.L = take(m * m);
    l.s = take(n); l.Qx = take(n); l.Qu = take(m); l.t1 = take(m); l.linv = take(m);
    l.mu = take(MAX_CON); l.bc = take(24); l.red = take(9);
    l.X = take(N * n); l.U = take((N - 1) * m); l.Xb = take(N * n); l.Ub = take((N - 1) * m);
    l.xr = l.ur = -1;
    l.dv = take((N - 1) * m); l.lam = take(P); l.ex = take(EX); l.itm = take(N * (1 + ncon));
    l.cand = q;
    q += q & 1;
    // TMA-staged panels: rows must be 16-byte multiples and K a multiple of 4 (n % 4 == 0); each stage holds
    // PANEL_KC k-rows of the left operand and of one column window (<= 8 PANEL_WT columns) of the right operand
    // (+ 32 doubles of slack each: partial edge tiles read past the row end)
    l.tma = (tma && n % 4 == 0 && n >= 32 && (m <= 8 * PANEL_WT || m % 2 == 0)) ? 1 : 0;
    if (l.tma) {
        const int mx = n > m ? n : m, mx8 = (mx + 7) & ~7;
        l.stage_doubles = PANEL_KC * ((mx + 4) + ((mx8 < 8 * PANEL_WT ? mx8 : 8 * PANEL_WT) + 4)) + 64;
        l.stage = take(PANEL_NST * l.stage_doubles);
        l.mbar = take(2 * PANEL_NST);
    }
    l.cd = q;
    l.sA = l.sB = 0;  // unused: A_k, B_k are read where they are
    l.S = wtake(n * n); l.SA = wtake(n * n); l.Qxx = wtake(n * n); l.SB = wtake(n * m); l.Qux = wtake(m * n);
    l.T1 = wtake(m * n); l.Qi = wtake(n + n * n + m + m * m); l.K = wtake((N - 1) * m * n);
    w += w & 1;  // instances start on 16-byte boundaries (bulk copies)
    size_t b = (size_t)q * sizeof(double) + (size_t)(ncon > 0 ? ncon : 1) * sizeof(ConDesc);
    l.bytes = (int)((b + 15) & ~(size_t)15);
    l.big = 1;
    l.ws_doubles = w;
    return l;
}

__host__ __device__ inline Layout make_layout(int n, int m, int N, int P, int ncon, int EX, int ref_in_smem, int ITAB,
                                              int spec_sets = 0)
{
    Layout l = fixed_layout(n, m);
    int q = l.X;
    auto take = [&](int count) { int at = q; q += count; return at; };
    l.X = take(N * n); l.U = take((N - 1) * m); l.Xb = take(N * n); l.Ub = take((N - 1) * m);
    l.xr = ref_in_smem ? take(N * n) : -1;
    l.ur = ref_in_smem ? take((N - 1) * m) : -1;
    l.K = take((N - 1) * m * n); l.dv = take((N - 1) * m);
    l.lam = take(P);
    l.ex = take(EX);  // EX = 0 when the expansion blocks live in global memory
    l.itm = take(N * (1 + ncon));
    // speculative line search: one extra (Xb, Ub, itm) set per additional warp
    l.cand = take(spec_sets * (N * n + (N - 1) * m + N * (1 + ncon)));
    q += q & 1;  // the descriptors and the gather records behind them are read as 16-byte words
    l.cd = q;
    size_t b = (size_t)q * sizeof(double) + (size_t)(ncon > 0 ? ncon : 1) * sizeof(ConDesc) + (size_t)ITAB * sizeof(int);
    l.bytes = (int)((b + 15) & ~(size_t)15);
    l.big = 0;
    l.ws_doubles = 0;
    return l;
}

struct Params {
    int n, m, N, B, P, ncon, EX, ITAB;
    int NSRC, NTL;  // gather table: NSRC source records, NTL targets refreshed at every knot (see Ctx::gather)
    int inst_offset;
    double dt;
    int dyn_per_knot, dyn_per_instance, dyn_in_smem, ref_in_smem;
    // optional dynamics schedule (gait-scheduled LTV models, e.g. the quadruped): per instance `dyn_slots` models
    // A[B][slots].., and knot k of MPC step s uses slot dyn_sched[inst][min(step0 + s + k, sched_len - 1)]
    int dyn_slots, sched_len, step0;
    const int *dyn_sched;
    const double *A, *Bm, *d;
    const double *Q, *R, *Qf;
    double *xref, *uref, *x0, *X, *U, *lam;
    int *iters, *outer, *status, *trials;
    double *cost, *cost_al, *cmax, *penmax;
    long long *t_ns;
    double *trace;  // optional [B][trace_rows][TRACE_COLS] per-iteration log (verbose mode), or nullptr
    int trace_rows;
    // closed-loop MPC run: `steps` x {transition; solve!} per instance inside one launch (0 = one plain solve!)
    int steps, shift, noise_mode, Nt;
    double noise_w1, noise_w2;
    const double *noise;            // [steps][B][n] standard-normal samples, or nullptr
    const double *trackX, *trackU;  // reference track [Nt][n], [Nt-1][m], or nullptr
    const int *kidx;                // per-instance track index of the current window start
    double *x0_log, *u0_log;        // [steps][B][n], [steps][B][m]: closed-loop state and applied control
    double *ex_glob;                // expansion scratch [B][EX] in global memory when it does not fit in shared, or nullptr
    int phase_detail;               // 1: phase[] holds the backward-pass sub-phase split instead
    long long *phase;               // optional [B][8] cycle counters per phase (profiling aid), or nullptr
    const ConDesc *con;
    const int *itab;  // gather tables built by the host: source records, gptr[NT+1], per-knot refresh list
    double *ws;       // large state dimension: per-instance workspace [B][lay.ws_doubles]
    Layout lay;
    int spec;  // line-search trials evaluated concurrently, one per warp (0 = sequential)
    // closed-loop runs on a persistent grid: work items (instance, chunk of `q_chunk` steps) are handed out chunk-major
    // from q_head; q_done[inst] = number of steps of the instance that are complete and stored (see Ctx::solve)
    int q_chunk;
    int *q_head, *q_done, *q_error;
    altro_opts_t o;
};

// fragment prefetch depth of Ctx::tile_gemm (k-steps of operands in flight per warp).  Measured at n = 200 (profiles/r2_large_n.md):
// depth 1 (no explicit prefetch) 2076 solves/s, depth 2 / 3 / 4 1249 / 1329 / 1360 -- the extra live fragments spill.
#ifndef ALTRO_GEMM_DEPTH
#define ALTRO_GEMM_DEPTH 1
#endif

#ifdef __CUDACC__

// Constraint data and LTV dynamics always live in global memory: telling the compiler turns generic loads into LDG.
template <class Tp>
__device__ __forceinline__ const Tp *as_global(const Tp *p)
{
    __builtin_assume(__isGlobal(p));
    return p;
}

template <int T>
__device__ __forceinline__ void gsync()
{
    if (T == 32) __syncwarp();
    else __syncthreads();
}

// Canonical sum of arr[0..count): lane l of warp 0 adds arr[l], arr[l+32], ... in order, then an
// xor-butterfly over the 32 partials.  The summation order does not depend on T, so the CPU oracle
// can reproduce every cost / gradient value bit for bit.  arr must be visible to warp 0 on entry.
template <int T>
__device__ __forceinline__ double csum(const double *arr, int count, double *bc, int tid)
{
    double v = 0.0;
    if (T == 32 || tid < 32) {
#pragma unroll 1
        for (int i = tid; i < count; i += 32) v += arr[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    if (T == 32) return v;
    if (tid == 0) bc[2] = v;
    __syncthreads();
    v = bc[2];
    __syncthreads();
    return v;
}

// Max over the instance's T threads (order independent, exact); every thread gets the same value.
template <int T>
__device__ __forceinline__ double gmax(double v, double *red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (T == 32) return v;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = red[0];
#pragma unroll
    for (int i = 1; i < T / 32; ++i) s = fmax(s, red[i]);
    return s;
}

// LDL' of the m x m block L (4 < m <= 16) and [K | d] = -L^-1 [Qux | Qu], warp 0, kept out of line so that its
// register arrays get a register allocation of their own.  Element (i, j) of the un-normalised factor accumulates
// L_ij - sum_l X_il (X_jl r_l) over ascending l with reciprocal pivots r_l = 1/X_ll (the oracle's chains), in a
// right-looking order: once pivot l is known, all remaining entries take one independent fma each, so only
// pivot -> reciprocal -> product -> one fma is on the dependent path.
//  * m > 8 (quadruped, m = 12): EVERY lane factorises the whole block in its own registers -- no shuffle, no
//    shared-memory round trip (the row-per-lane version with shuffled pivot rows was half of a quadruped knot);
//  * m <= 8 (grasp, random linear, m = 6): lane i keeps row i and the pivot-row products travel by shuffle (measured:
//    the register-resident form costs the m = 6 kernels 20-30 % through the caller's spills around the call).
// Then lane c substitutes right-hand side c: forward y = L^-1 b, z = y r, backward (q descending).
template <int MM>
__device__ __noinline__ bool ldl_solve_medium(double *L, const double *Qux, const double *Qu, double *Kk, double *dk_, int n)
{
    const int lane = threadIdx.x & 31;
    if constexpr (MM > 8) {
        double X[MM * (MM + 1) / 2], rr[MM];
#define ALTRO_TRI(i, j) X[(i) * ((i) + 1) / 2 + (j)]
#pragma unroll
        for (int i = 0; i < MM; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) ALTRO_TRI(i, j) = L[i * MM + j];
        // software-pipelined: column l + 1 is updated first and the reciprocal of its pivot is started before the
        // updates of the columns behind it, so the 72-cycle reciprocal overlaps with them
        if (!(ALTRO_TRI(0, 0) > 0.0)) return true;  // (the same value in every lane)
        rr[0] = __drcp_rn(ALTRO_TRI(0, 0));
#pragma unroll
        for (int l = 0; l < MM - 1; ++l) {
            {
                const int j = l + 1;
                const double t = ALTRO_TRI(j, l) * rr[l];
#pragma unroll
                for (int i = j; i < MM; ++i) ALTRO_TRI(i, j) = fma(-ALTRO_TRI(i, l), t, ALTRO_TRI(i, j));
            }
            if (!(ALTRO_TRI(l + 1, l + 1) > 0.0)) return true;
            rr[l + 1] = __drcp_rn(ALTRO_TRI(l + 1, l + 1));
#pragma unroll
            for (int j = l + 2; j < MM; ++j) {
                const double t = ALTRO_TRI(j, l) * rr[l];
#pragma unroll
                for (int i = j; i < MM; ++i) ALTRO_TRI(i, j) = fma(-ALTRO_TRI(i, l), t, ALTRO_TRI(i, j));
            }
        }
#pragma unroll 1
        for (int c = lane; c <= n; c += 32) {
            double *bp = (c < n) ? Kk + c : dk_;
            const double *src = (c < n) ? Qux + c : Qu;
            const int st = (c < n) ? n : 1;
            // column-oriented in time, row chains in value: entry i still accumulates over ascending (forward) /
            // descending (backward) l, but once bb[l] is final all the entries behind it take one independent fma each
            double bb[MM], a2[MM];
#pragma unroll
            for (int i = 0; i < MM; ++i) bb[i] = -src[i * st];
#pragma unroll
            for (int l = 0; l < MM - 1; ++l) {
                const double t = bb[l] * rr[l];
#pragma unroll
                for (int i = l + 1; i < MM; ++i) bb[i] = fma(-ALTRO_TRI(i, l), t, bb[i]);
            }
#pragma unroll
            for (int i = 0; i < MM; ++i) { bb[i] = bb[i] * rr[i]; a2[i] = 0.0; }
#pragma unroll
            for (int l = MM - 1; l >= 0; --l) {
                bb[l] = fma(-rr[l], a2[l], bb[l]);
#pragma unroll
                for (int i = l - 1; i >= 0; --i) a2[i] = fma(ALTRO_TRI(l, i), bb[l], a2[i]);
            }
#pragma unroll
            for (int i = 0; i < MM; ++i) bp[i * st] = bb[i];
        }
#undef ALTRO_TRI
        return false;
    } else {
        bool bad = false;
        double row[MM], rr[MM];
        const int li = lane < MM ? lane : MM - 1;
#pragma unroll
        for (int j = 0; j < MM; ++j) row[j] = L[li * MM + j];
#pragma unroll
        for (int l = 0; l < MM; ++l) {
            const double piv = __shfl_sync(0xffffffffu, row[l], l);
            if (!(piv > 0.0)) { bad = true; break; }
            rr[l] = __drcp_rn(piv);
            const double t = row[l] * rr[l];
#pragma unroll
            for (int j = l + 1; j < MM; ++j) row[j] = fma(-row[l], __shfl_sync(0xffffffffu, t, j), row[j]);
        }
        if (bad) return true;
        if (lane < MM) {
#pragma unroll
            for (int j = 0; j < MM; ++j) L[lane * MM + j] = row[j];
        }
        __syncwarp();
#pragma unroll 1
        for (int c = lane; c <= n; c += 32) {
            double *bp = (c < n) ? Kk + c : dk_;
            const double *src = (c < n) ? Qux + c : Qu;
            const int st = (c < n) ? n : 1;
            double bb[MM];
#pragma unroll
            for (int i = 0; i < MM; ++i) {
                double acc = -src[i * st];
#pragma unroll
                for (int l = 0; l < i; ++l) acc = fma(-L[i * MM + l], bb[l] * rr[l], acc);
                bb[i] = acc;
            }
#pragma unroll
            for (int i = 0; i < MM; ++i) bb[i] = bb[i] * rr[i];
#pragma unroll
            for (int i = MM - 1; i >= 0; --i) {
                double acc2 = 0.0;
#pragma unroll
                for (int l = MM - 1; l > i; --l) acc2 = fma(L[l * MM + i], bb[l], acc2);
                bb[i] = fma(-rr[i], acc2, bb[i]);
            }
#pragma unroll
            for (int i = 0; i < MM; ++i) bp[i * st] = bb[i];
        }
        return false;
    }
}

// ---- TMA bulk copies and mbarriers (sm_90+ PTX): operand panels of the large-dimension GEMMs are fetched by the copy
// engine into shared memory while the warps run tensor tiles on the previous panel.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// the same on 32-bit shared-memory addresses (the panel loop keeps no generic pointers alive)
__device__ __forceinline__ void mbar_arrive_u32(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait_u32(uint32_t b, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(b), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// Per-instance context: shared-memory pointers and problem view.
// WIDE: the variant compiled into altro_solve_kernel_wide (whole register file, TMA-staged panels); the default kernels
// do not carry that code.
// ---- TMA-staged panel GEMM (large state dimension):
//     C(i,j) = Ci(i,j) + sum_l PA[l][i] PB[l][j]  [+ sum_l PA2[l][i] PB2[l][j]],   every sum over ascending l.
// The operands are "k-major" (row l of PA holds column l of the left factor; lda == M, ldb == Nc), so PANEL_KC
// k-rows of an operand are one contiguous block of global memory.  The last warp of the CTA is the PRODUCER: it
// streams the panels into a ring of PANEL_NST shared-memory stages with cp.async.bulk (one copy per k-row, issued
// by one lane each, or one copy per panel when the rows need no padding); full[] barriers count the transaction
// bytes, empty[] barriers one arrival per consumer warp.  The other NC = T/32 - 1 warps are CONSUMERS: C is cut
// into column windows; inside a window a consumer owns 32 rows x up to PANEL_WT = 7 DMMA tiles = up to 28 independent
// accumulator chains (4 A + 7 B fragment loads from shared memory feed 28 DMMAs per k-step).  With fewer than NC row
// blocks the warps share a row block column-wise and the window widens accordingly (B'SA, 25 x 200, is one window).
// A 200 x 200 x 200 product streams the left operand four times and the right operand once (1.6 MB from L2 instead
// of 7 MB for the register-blocked tiles that read their operands where they are).  Each element's chain is the same
// sequence of mma.sync.m8n8k4 steps over ascending l as in tile_gemm, so the bits do not change; K need not be a
// multiple of 4: the fragments of the last, partial k-step are zero beyond K, exactly like tile_gemm's operand
// functors.  S A uses PA = S: S is symmetric to the last bit (it is formed as (D + D')/2, commutative operations).
// A free, non-inlined function: the product loop gets the whole register file (128 accumulator registers) and the
// caller's live state is saved once per call instead of being spilled inside the loop.  Ci may be read transposed
// (ci_t; square C only); the result goes to Co -- with Avg as (Avg + C)/2, the symmetrisation of the cost-to-go
// Hessian fused into the last product -- and, if Co2 is given, Co2 = stored value + rho I (the regularised copy the
// factorisation works on); all row-major with ld = Nc.  Returns the number of panels streamed.
struct PanelArgs {
    int M, Nc, K, K2, ci_t;
    const double *PA, *PB, *PA2, *PB2, *Ci, *Avg;
    double *Co, *Co2;
    double rho;
};
template <int T>
__device__ __noinline__ unsigned panel_gemm_fn(double *stage0, uint64_t *full, int sd, unsigned pg_count, const PanelArgs &ga)
{
    constexpr int NC = T / 32 - 1, WT = PANEL_WT;
    const PanelArgs g = ga;  // a private copy: the stores below must not force the fields to be re-read
    const double *__restrict__ const Ci = g.Ci, *__restrict__ const Avg = g.Avg;
    double *__restrict__ const Co = g.Co, *__restrict__ const Co2 = g.Co2;
    const int M = g.M, Nc = g.Nc, lda = M, ldb = Nc;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, r = lane >> 2, q = lane & 3;
    uint64_t *empty = full + PANEL_NST;
    const int lsa = panel_ld(lda), boff = PANEL_KC * lsa + 32;
    const int ntc = (Nc + 7) >> 3, nrb = (M + 31) >> 5;
    const int csplit = max(1, NC / nrb);                       // warps that share a row block
    const int wcap = ((sd - 64) / PANEL_KC - lsa - 4) >> 3;     // tile columns of B a stage has room for
    const int ns = (ntc + min(WT * csplit, wcap) - 1) / min(WT * csplit, wcap), nbw = (ntc + ns - 1) / ns;
    const int nsub = min(csplit, nbw), tps = (nbw + nsub - 1) / nsub;
    const int bpw = nrb * nsub, rpw = (bpw + NC - 1) / NC, rounds = ns * rpw;
    const int lsb = ns == 1 ? panel_ld(ldb) : panel_ld(nbw * 8);  // row strides in shared memory
    const int nch1 = (g.K + PANEL_KC - 1) / PANEL_KC, nchunk = nch1 + (g.K2 + PANEL_KC - 1) / PANEL_KC;
    const int total = rounds * nchunk;
    if (warp == NC) {
        unsigned st = pg_count % PANEL_NST, use = pg_count / PANEL_NST;
        int R = 0, cc = 0;
#pragma unroll 1
        for (int gi = 0; gi < total; ++gi) {
            if (use > 0) mbar_wait(empty + st, (use - 1) & 1);
            const int second = cc >= nch1, c = second ? cc - nch1 : cc;
            const int rows = min(PANEL_KC, (second ? g.K2 : g.K) - c * PANEL_KC);
            const double *PA = second ? g.PA2 : g.PA, *PB = second ? g.PB2 : g.PB;
            const int cb0 = (R / rpw) * nbw * 8, wc = ns == 1 ? ldb : min(nbw * 8, ldb - cb0);
            double *As = stage0 + st * sd, *Bs = As + boff;
            const uint32_t ba = (uint32_t)(rows * lda * sizeof(double)), bb = (uint32_t)(rows * wc * sizeof(double));
            if (lane == 0) mbar_expect_tx(full + st, ba + bb);
            __syncwarp();
            if (lsa == lda) {
                if (lane == 0) bulk_g2s(As, PA + (size_t)c * PANEL_KC * lda, ba, full + st);
            } else if (lane < rows) {
                bulk_g2s(As + lane * lsa, PA + ((size_t)c * PANEL_KC + lane) * lda, (uint32_t)(lda * sizeof(double)), full + st);
            }
            if (ns == 1 && lsb == ldb) {
                if (lane == 16) bulk_g2s(Bs, PB + (size_t)c * PANEL_KC * ldb, bb, full + st);
            } else if (lane >= 16 && lane - 16 < rows) {
                bulk_g2s(Bs + (lane - 16) * lsb, PB + ((size_t)c * PANEL_KC + lane - 16) * ldb + cb0,
                         (uint32_t)(wc * sizeof(double)), full + st);
            }
            if (++cc == nchunk) { cc = 0; ++R; }
            if (++st == PANEL_NST) { st = 0; ++use; }
        }
    } else {
        unsigned st = pg_count % PANEL_NST, par = (pg_count / PANEL_NST) & 1;
        const uint32_t sbase = smem_u32(stage0), fbar = smem_u32(full), ebar = fbar + 8u * PANEL_NST;
        const uint32_t sdb = (uint32_t)sd * 8u, sa4 = (uint32_t)lsa * 32u, sb4 = (uint32_t)lsb * 32u;  // bytes
        const bool vec = (Nc & 1) == 0;  // rows start on 16-byte boundaries: two columns per load / store
#pragma unroll 1
        for (int R = 0; R < rounds; ++R) {
            const int w = R / rpw, b = (R - w * rpw) * NC + warp;
            const int rb = b / nsub, t0 = w * nbw + (b - rb * nsub) * tps;
            const int ntile = b < bpw ? min(tps, min((w + 1) * nbw, ntc) - t0) : 0;  // <= 0: nothing to do this round
            const int i0 = rb << 5, j0 = t0 << 3;
            double acc[4][WT][2];
            const bool civ = Ci && vec && !g.ci_t;
#pragma unroll
            for (int y = 0; y < WT; ++y)
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const int i = i0 + 8 * x + r, j = j0 + 8 * y + 2 * q;
                    const bool in = y < ntile && i < M;
                    if (civ) {  // two columns per load (j + 1 < Nc: Nc is even here)
                        double2 v = make_double2(0.0, 0.0);
                        if (in && j < Nc) v = *reinterpret_cast<const double2 *>(Ci + (size_t)i * Nc + j);
                        acc[x][y][0] = v.x;
                        acc[x][y][1] = v.y;
                    } else {
                        acc[x][y][0] = (Ci && in && j < Nc) ? Ci[g.ci_t ? (size_t)j * Nc + i : (size_t)i * Nc + j] : 0.0;
                        acc[x][y][1] = (Ci && in && j + 1 < Nc) ? Ci[g.ci_t ? (size_t)(j + 1) * Nc + i : (size_t)i * Nc + j + 1] : 0.0;
                    }
                }
            // the chunk loop keeps little state alive besides the accumulators: 32-bit shared-memory addresses, the ring
            // position and a row countdown
            const uint32_t a_off = (uint32_t)(q * lsa + i0 + r) * 8u, b_off = (uint32_t)(boff + q * lsb + (j0 - w * nbw * 8) + r) * 8u;
            int krem = g.K, second = 0;
#pragma unroll 1
            for (int cc = 0; cc < nchunk; ++cc) {
                mbar_wait_u32(fbar + 8u * st, par);
                const int rows = min(PANEL_KC, krem);
                if (ntile > 0) {
                    uint32_t pa = sbase + st * sdb + a_off, pb = sbase + st * sdb + b_off;
                    // every fragment load is unconditional (tile columns beyond ntile read stale shared memory that only
                    // feeds skipped DMMAs): conditionally written fragment arrays would live in local memory
                    auto kstep = [&](auto tail, bool live) {
                        double a[4], bf[WT];
#pragma unroll
                        for (int x = 0; x < 4; ++x) a[x] = lds_f64(pa + 64u * x);
#pragma unroll
                        for (int y = 0; y < WT; ++y) bf[y] = lds_f64(pb + 64u * y);
                        if constexpr (decltype(tail)::value) {
#pragma unroll
                            for (int x = 0; x < 4; ++x) a[x] = live ? a[x] : 0.0;
#pragma unroll
                            for (int y = 0; y < WT; ++y) bf[y] = live ? bf[y] : 0.0;
                        }
#pragma unroll
                        for (int y = 0; y < WT; ++y)
                            if (y < ntile) {
#pragma unroll
                                for (int x = 0; x < 4; ++x)
                                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                                 : "+d"(acc[x][y][0]), "+d"(acc[x][y][1]) : "d"(a[x]), "d"(bf[y]));
                            }
                        pa += sa4;
                        pb += sb4;
                    };
#pragma unroll 1
                    for (int kk = 0; kk < (rows >> 2); ++kk) kstep(std::false_type{}, true);
                    if (rows & 3) kstep(std::true_type{}, q < (rows & 3));  // k-rows beyond K contribute fma(0, 0, acc)
                }
                krem -= rows;
                if (krem == 0 && !second) { krem = g.K2; second = 1; }
                __syncwarp();
                if (lane == 0) mbar_arrive_u32(ebar + 8u * st);
                if (++st == PANEL_NST) { st = 0; par ^= 1; }
            }
            if (vec) {
                // two tile columns at a time: all loads of the batch first (Avg), then the arithmetic and the stores
#pragma unroll
                for (int y0 = 0; y0 < WT; y0 += 2) {
                    if (y0 >= ntile) break;
                    double2 o[2][4];
                    if (Avg) {
#pragma unroll
                        for (int yy = 0; yy < 2; ++yy)
#pragma unroll
                            for (int x = 0; x < 4; ++x) {
                                const int i = i0 + 8 * x + r, j = j0 + 8 * (y0 + yy) + 2 * q;
                                o[yy][x] = (y0 + yy < ntile && i < M && j < Nc)
                                               ? *reinterpret_cast<const double2 *>(Avg + (size_t)i * Nc + j) : make_double2(0.0, 0.0);
                            }
                    }
#pragma unroll
                    for (int yy = 0; yy < 2; ++yy)
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            const int i = i0 + 8 * x + r, j = j0 + 8 * (y0 + yy) + 2 * q;
                            if (y0 + yy >= ntile || i >= M || j >= Nc) continue;
                            double2 v = make_double2(acc[x][y0 + yy][0], acc[x][y0 + yy][1]);
                            if (Avg) {
                                v.x = 0.5 * (o[yy][x].x + v.x);
                                v.y = 0.5 * (o[yy][x].y + v.y);
                            }
                            *reinterpret_cast<double2 *>(Co + (size_t)i * Nc + j) = v;
                            if (Co2)
                                *reinterpret_cast<double2 *>(Co2 + (size_t)i * Nc + j) =
                                    make_double2(v.x + ((i == j) ? g.rho : 0.0), v.y + ((i == j + 1) ? g.rho : 0.0));
                        }
                }
            } else {
#pragma unroll
                for (int y = 0; y < WT; ++y)
                    if (y < ntile) {
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            const int i = i0 + 8 * x + r, j = j0 + 8 * y + 2 * q;
                            if (i >= M) continue;
                            const size_t e = (size_t)i * Nc + j;
                            if (j < Nc) {
                                const double v = Avg ? 0.5 * (Avg[e] + acc[x][y][0]) : acc[x][y][0];
                                Co[e] = v;
                                if (Co2) Co2[e] = v + ((i == j) ? g.rho : 0.0);
                            }
                            if (j + 1 < Nc) {
                                const double v = Avg ? 0.5 * (Avg[e + 1] + acc[x][y][1]) : acc[x][y][1];
                                Co[e + 1] = v;
                                if (Co2) Co2[e + 1] = v + ((i == j + 1) ? g.rho : 0.0);
                            }
                        }
                    }
            }
        }
    }
    // the results go to global memory through the generic proxy and are the next product's operands (async proxy)
    fence_proxy_async();
    __syncthreads();
    return (unsigned)total;
}

template <int NX, int NU, int T, bool WIDE = false>
struct Ctx {
    static constexpr bool ALL_SMEM = NX > 0 && NU > 0 && ALTRO_FIXED_ALL_SMEM;
    const Params &P;
    unsigned char *smem_base;
    double *specr;  // [2 W] (ok, J) of the speculative line-search trials
    int n, m, N, inst, tid, ncon;
    // shared memory
    double *Qd, *Qfd, *Rd, *sA, *sB, *sd;
    double *X, *U, *Xb, *Ub, *xr, *ur, *K, *dv, *lam, *mu, *ex;
    double *S, *SA, *Qxx, *SB, *Qux, *T1, *Quu, *L, *s, *Qx, *Qu, *t1, *ldiag, *linv, *red, *bc, *itm, *Qi;
    int4 *grec;
    int *gptr, *gtl;  // gather lists: for every entry of [Qx | Qxx | Qu | Quu] the (block << 16 | offset) sources
    int NT;
    long long ph_exp = 0, ph_roll = 0, ph_cost = 0;  // profiling aid (P.phase)
    long long bpc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    ConDesc *cd;
    size_t dyn_base;
    int dyn_k;
    const int *sched;  // this instance's dynamics schedule at the current MPC step, or nullptr
    int kcur;          // this instance's position on the shared timelines (reference track, track constraints)
    unsigned pg_count = 0;  // operand panels streamed so far (stage and mbarrier phase follow from it; same in every thread)

    __device__ Ctx(const Params &P_, unsigned char *raw) : P(P_), smem_base(raw)
    {
        n = NX ? NX : P.n;
        m = NU ? NU : P.m;
        N = P.N;
        ncon = P.ncon;
        tid = threadIdx.x;
        inst = blockIdx.x + P.inst_offset;
        double *sm = reinterpret_cast<double *>(raw);
        const Layout &l = P.lay;
        if constexpr (NX > 0 && NU > 0) {
            constexpr Layout f = fixed_layout(NX, NU);
            Qd = sm + f.Qd; Qfd = sm + f.Qfd; Rd = sm + f.Rd; sA = sm + f.sA; sB = sm + f.sB; sd = sm + f.sd;
            S = sm + f.S; SA = sm + f.SA; Qxx = sm + f.Qxx; SB = sm + f.SB; Qux = sm + f.Qux; T1 = sm + f.T1;
            Quu = sm + f.Quu; L = sm + f.L; s = sm + f.s; Qx = sm + f.Qx; Qu = sm + f.Qu; t1 = sm + f.t1;
            linv = sm + f.linv; mu = sm + f.mu; bc = sm + f.bc; red = sm + f.red; Qi = sm + f.Qi; X = sm + f.X;
        } else {
            // run-time sized kernel: large problems keep their n-sized matrices in a global workspace
            // (scratch of one solve, nothing in it survives a step: it belongs to the CTA, not to the instance, so that
            // the persistent grid of a queued closed-loop run needs one workspace per resident CTA)
            double *big = l.big ? P.ws + (size_t)blockIdx.x * l.ws_doubles : sm;
            Qd = sm + l.Qd; Qfd = sm + l.Qfd; Rd = sm + l.Rd; sA = big + l.sA; sB = big + l.sB; sd = sm + l.sd;
            S = big + l.S; SA = big + l.SA; Qxx = big + l.Qxx; SB = big + l.SB; Qux = big + l.Qux; T1 = big + l.T1;
            Quu = sm + l.Quu; L = sm + l.L; s = sm + l.s; Qx = sm + l.Qx; Qu = sm + l.Qu; t1 = sm + l.t1;
            linv = sm + l.linv; mu = sm + l.mu; bc = sm + l.bc; red = sm + l.red; Qi = big + l.Qi; X = sm + l.X;
        }
        U = sm + l.U; Xb = sm + l.Xb; Ub = sm + l.Ub;
        K = ((NX == 0 && l.big) ? P.ws + (size_t)blockIdx.x * l.ws_doubles : sm) + l.K; dv = sm + l.dv; lam = sm + l.lam;
        if constexpr (ALL_SMEM) {
            // fixed-dimension kernels keep the reference window and the expansion blocks in shared memory, always
            // (the host routes problems that do not fit to the run-time sized kernel): every access is an LDS/STS
            xr = sm + l.xr; ur = sm + l.ur; ex = sm + l.ex;
        } else {
            if (P.ref_in_smem) { xr = sm + l.xr; ur = sm + l.ur; }
            else {  // long horizons: read the reference through L1/L2 instead
                xr = P.xref + (size_t)inst * N * n;
                ur = P.uref + (size_t)inst * (N - 1) * m;
            }
            ex = P.ex_glob ? P.ex_glob + (size_t)inst * P.EX : sm + l.ex;  // longer still: expansion blocks in global memory
        }
        ldiag = nullptr; itm = sm + l.itm; specr = bc + 8;
        cd = reinterpret_cast<ConDesc *>(sm + l.cd);
        NT = n + n * n + m + m * m;
        grec = (NX == 0 && l.big) ? reinterpret_cast<int4 *>(const_cast<int *>(P.itab))  // too long for shared memory
                                  : reinterpret_cast<int4 *>(cd + (ncon > 0 ? ncon : 1));
        gptr = reinterpret_cast<int *>(grec + P.NSRC);
        gtl = gptr + NT + 1;
        dyn_base = P.dyn_per_instance ? (size_t)inst * (P.dyn_sched ? (size_t)P.dyn_slots : (P.dyn_per_knot ? (size_t)(N - 1) : 1)) : 0;
        dyn_k = P.dyn_per_knot ? 1 : 0;
        sched = nullptr;
        kcur = P.kidx ? P.kidx[inst] : 0;
        set_step(0);
    }

    // (re)binds the context to instance i: persistent CTAs walk over many instances
    __device__ __forceinline__ void bind(int i)
    {
        inst = i + P.inst_offset;
        dyn_base = P.dyn_per_instance ? (size_t)inst * (P.dyn_sched ? (size_t)P.dyn_slots : (P.dyn_per_knot ? (size_t)(N - 1) : 1)) : 0;
        kcur = P.kidx ? P.kidx[inst] : 0;
        if constexpr (!ALL_SMEM) {
            if (!P.ref_in_smem) {
                xr = P.xref + (size_t)inst * N * n;
                ur = P.uref + (size_t)inst * (N - 1) * m;
            }
            if (P.ex_glob) ex = P.ex_glob + (size_t)inst * P.EX;
        }
        set_step(0);
    }

    __device__ __forceinline__ void set_step(int st)
    {
        if (P.dyn_sched) sched = P.dyn_sched + (size_t)inst * P.sched_len + min(P.step0 + st, P.sched_len - N);
    }
    __device__ __forceinline__ size_t dyn_index(int k) const
    {
        return dyn_base + (sched ? (size_t)sched[k] : (size_t)dyn_k * k);
    }
    __device__ __forceinline__ const double *Ak(int k) const
    {
        return P.dyn_in_smem ? sA : P.A + dyn_index(k) * n * n;
    }
    __device__ __forceinline__ const double *Bk(int k) const
    {
        return P.dyn_in_smem ? sB : P.Bm + dyn_index(k) * n * m;
    }
    __device__ __forceinline__ const double *dk(int k) const
    {
        return P.dyn_in_smem ? sd : P.d + dyn_index(k) * n;
    }
    __device__ __forceinline__ size_t con_idx(const ConDesc &c, int k) const
    {
        if (c.track) return (size_t)min(kcur + k, c.track - 1);
        size_t idx = c.per_instance ? (size_t)inst * (c.per_knot ? (size_t)(c.k1 - c.k0) : 1) : 0;
        return idx + (c.per_knot ? (size_t)(k - c.k0) : 0);
    }

    // ---------------------------------------------------------------- load / store
    // Everything that does not depend on the instance: weights, shared LTI model, descriptors, gather tables, the cost
    // Hessian entries no block touches.
    __device__ void load_static()
    {
        if constexpr (WIDE) {
            if (P.lay.big && P.lay.tma && tid == 0) {  // panel pipeline barriers (panel_gemm)
                uint64_t *mb = reinterpret_cast<uint64_t *>(reinterpret_cast<double *>(smem_base) + P.lay.mbar);
                for (int i = 0; i < PANEL_NST; ++i) { mbar_init(mb + i, 1); mbar_init(mb + PANEL_NST + i, T / 32 - 1); }
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
        }
#pragma unroll 1
        for (int i = tid; i < n; i += T) { Qd[i] = P.Q[i]; Qfd[i] = P.Qf[i]; }
#pragma unroll 1
        for (int i = tid; i < m; i += T) Rd[i] = P.R[i];
        if (P.dyn_in_smem) {  // shared LTI model
#pragma unroll 1
            for (int i = tid; i < n * n; i += T) sA[i] = P.A[i];
#pragma unroll 1
            for (int i = tid; i < n * m; i += T) sB[i] = P.Bm[i];
#pragma unroll 1
            for (int i = tid; i < n; i += T) sd[i] = P.d[i];
        }
        const int words = ncon * (int)(sizeof(ConDesc) / sizeof(int));
        const int *src = reinterpret_cast<const int *>(P.con);
        int *dst = reinterpret_cast<int *>(cd);
#pragma unroll 1
        for (int i = tid; i < words; i += T) dst[i] = src[i];
        if (!(NX == 0 && P.lay.big))
#pragma unroll 1
            for (int i = tid; i < P.ITAB; i += T) reinterpret_cast<int *>(grec)[i] = P.itab[i];
#pragma unroll 1
        for (int t = tid; t < NT; t += T) {  // matrix entries no block touches keep the cost Hessian for the whole launch
            double base = 0.0;
            if (t >= n && t < n + n * n) { const int e = t - n, i = e / n, j = e - i * n; base = (i == j) ? P.dt * P.Q[i] : 0.0; }
            else if (t >= n + n * n + m) { const int e = t - (n + n * n + m), i = e / m, j = e - i * m; base = (i == j) ? P.dt * P.R[i] : 0.0; }
            Qi[t] = base;
        }
    }

    // The instance's state: warm start, duals, reference window.  CG = true reads through L2 (ld.global.cg): in a queued
    // run another SM wrote these rows a moment ago and this SM's L1 may still hold an older copy.
    template <bool CG>
    __device__ void load_state()
    {
        auto rd_ = [](const double *p) { return CG ? __ldcg(p) : *p; };
        if (P.steps > 0) {  // closed-loop run: start from the previous solution (its x_1 is the next x_0)
            const double *gX = P.X + (size_t)inst * N * n;
#pragma unroll 1
            for (int i = tid; i < N * n; i += T) X[i] = rd_(gX + i);
        } else {
            const double *gx0 = P.x0 + (size_t)inst * n;
#pragma unroll 1
            for (int i = tid; i < n; i += T) X[i] = gx0[i];
        }
        const double *gU = P.U + (size_t)inst * (N - 1) * m;
#pragma unroll 1
        for (int i = tid; i < (N - 1) * m; i += T) U[i] = rd_(gU + i);
        const double *gxr = P.xref + (size_t)inst * N * n, *gur = P.uref + (size_t)inst * (N - 1) * m;
        if (ALL_SMEM || P.ref_in_smem) {
#pragma unroll 1
            for (int i = tid; i < N * n; i += T) xr[i] = rd_(gxr + i);
#pragma unroll 1
            for (int i = tid; i < (N - 1) * m; i += T) ur[i] = rd_(gur + i);
        }
        const double *gl = P.lam + (size_t)inst * P.P;
        const bool rd = P.o.reset_duals != 0;
#pragma unroll 1
        for (int i = tid; i < P.P; i += T) lam[i] = rd ? 0.0 : rd_(gl + i);
#pragma unroll 1
        for (int i = tid; i < MAX_CON; i += T) mu[i] = P.o.penalty_initial;
    }

    __device__ void load()
    {
        load_static();
        load_state<false>();
        gsync<T>();
    }

    __device__ void store()
    {
        double *gX = P.X + (size_t)inst * N * n, *gU = P.U + (size_t)inst * (N - 1) * m;
#pragma unroll 1
        for (int i = tid; i < N * n; i += T) gX[i] = X[i];
#pragma unroll 1
        for (int i = tid; i < (N - 1) * m; i += T) gU[i] = U[i];
        double *gl = P.lam + (size_t)inst * P.P;
#pragma unroll 1
        for (int i = tid; i < P.P; i += T) gl[i] = lam[i];
        if (P.steps > 0) {  // closed-loop run: the handle's x0 and reference window follow the plant
#pragma unroll 1
            for (int i = tid; i < n; i += T) P.x0[(size_t)inst * n + i] = X[i];
            if (P.trackX) {
                double *gxr = P.xref + (size_t)inst * N * n, *gur = P.uref + (size_t)inst * (N - 1) * m;
#pragma unroll 1
                for (int i = tid; i < N * n; i += T) gxr[i] = xr[i];
#pragma unroll 1
                for (int i = tid; i < (N - 1) * m; i += T) gur[i] = ur[i];
            }
        }
    }

    // ---------------------------------------------------------------- constraint values
    // Row r of block c at knot k evaluated on z (x_k or u_k): TO.evaluate.
    // The block's slice of z, fetched once per (block, knot) item for narrow dense blocks (cones, pyramids: w <= ZW) so
    // that the rows are pure G loads + fma instead of a dependent index -> value load chain per term.
    static constexpr int ZW = 4;
    __device__ __forceinline__ void load_zl(const ConDesc &c, const double *z, double (&zl)[ZW]) const
    {
#pragma unroll
        for (int j = 0; j < ZW; ++j) zl[j] = (!c.rowsparse && j < c.w) ? z[c.inds[j]] : 0.0;
    }
    __device__ __forceinline__ double row_value(const ConDesc &c, const double *G, const double *h, const double *z,
                                                const double (&zl)[ZW], int r) const
    {
        if (c.rowsparse) return fma(as_global(c.rs_coef)[r], z[c.inds[as_global(c.rs_col)[r]]], h[r]);
        double acc = h[r];
        const double *g = G + r * c.w;
        if (c.w <= ZW) {  // same ascending-index fma chain, operands already in registers
#pragma unroll
            for (int j = 0; j < ZW; ++j)
                if (j < c.w) acc = fma(g[j], zl[j], acc);
            return acc;
        }
#pragma unroll 1
        for (int j = 0; j < c.w; ++j) acc = fma(g[j], z[c.inds[j]], acc);
        return acc;
    }

    // AL penalty term of block ci at knot k (Altro cost!(J, conval)); SURVEY.md A.3.
    __device__ double con_cost(int ci, int k, const double *Xc, const double *Uc) const
    {
        const ConDesc &c = cd[ci];
        if (k < c.k0 || k >= c.k1) return 0.0;
        const size_t di = con_idx(c, k);
        const double *G = as_global(c.G) + di * c.p * c.w, *h = as_global(c.h) + di * c.p;
        const double *z = c.side == ALTRO_STATE ? Xc + k * n : Uc + k * m;
        double zl[ZW];
        load_zl(c, z, zl);
        const double *l = lam + c.dual_off + (k - c.k0) * c.p;
        const double mu_c = mu[ci];
        double J = 0.0;
        if (c.sense == ALTRO_EQUALITY) {
#pragma unroll 1
            for (int r = 0; r < c.p; ++r) {
                double v = row_value(c, G, h, z, zl, r);
                J += l[r] * v + 0.5 * mu_c * v * v;
            }
        } else if (c.sense == ALTRO_INEQUALITY) {
#pragma unroll 1
            for (int r = 0; r < c.p; ++r) {
                double v = row_value(c, G, h, z, zl, r);
                bool act = (v >= 0.0) || (l[r] > 0.0);
                J += l[r] * v + (act ? 0.5 * mu_c * v * v : 0.0);
            }
        } else {
            double a2 = 0.0, t = 0.0, nl = 0.0;
#pragma unroll 1
            for (int r = 0; r < c.p; ++r) {
                double lb = l[r] - mu_c * row_value(c, G, h, z, zl, r);
                nl += l[r] * l[r];
                if (r < c.p - 1) a2 += lb * lb;
                else t = lb;
            }
            double a = sqrt(a2), np;
            if (a <= -t) np = 0.0;
            else if (a <= t) np = a2 + t * t;
            else np = 0.5 * (a + t) * (a + t);
            J = (np - nl) / (2.0 * mu_c);
        }
        return J;
    }

    __device__ double stage_cost(int k, const double *Xc, const double *Uc) const
    {
        const double *x = Xc + k * n, *r = xr + k * n;
        double J = 0.0;
        if (k == N - 1) {
            for (int i = 0; i < n; ++i) { double e = x[i] - r[i]; J += 0.5 * Qfd[i] * e * e; }
            return J;
        }
        for (int i = 0; i < n; ++i) { double e = x[i] - r[i]; J += 0.5 * Qd[i] * e * e; }
        const double *u = Uc + k * m, *q = ur + k * m;
        for (int i = 0; i < m; ++i) { double e = u[i] - q[i]; J += 0.5 * Rd[i] * e * e; }
        return J * P.dt;
    }

    // AL cost of a trajectory: sum over (knot, piece) work items, one per thread.
    __device__ double al_cost(const double *Xc, const double *Uc) const
    {
        const int items = N * (1 + ncon);
#pragma unroll 1
        for (int it = tid; it < items; it += T) {  // piece-major: neighbouring lanes share the piece (no divergence)
            const int j = it / N, k = it - j * N;
            itm[it] = (j == 0) ? stage_cost(k, Xc, Uc) : con_cost(j - 1, k, Xc, Uc);
        }
        gsync<T>();
        return csum<T>(itm, items, bc, tid);
    }

    __device__ double objective_cost() const
    {
#pragma unroll 1
        for (int k = tid; k < N; k += T) itm[k] = stage_cost(k, X, U);
        gsync<T>();
        return csum<T>(itm, N, bc, tid);
    }

    // max_violation
    __device__ double max_violation() const
    {
        double v = 0.0;
#pragma unroll 1
        for (int it = tid; it < ncon * N; it += T) {  // (block, knot) work items, block-major
            const int ci = it / N, k = it - ci * N;
            const ConDesc &c = cd[ci];
            if (k >= c.k0 && k < c.k1) {
                const size_t di = con_idx(c, k);
                const double *G = as_global(c.G) + di * c.p * c.w, *h = as_global(c.h) + di * c.p;
                const double *z = c.side == ALTRO_STATE ? X + k * n : U + k * m;
                double zl[ZW];
                load_zl(c, z, zl);
                if (c.sense == ALTRO_EQUALITY) {
#pragma unroll 1
                    for (int r = 0; r < c.p; ++r) v = fmax(v, fabs(row_value(c, G, h, z, zl, r)));
                } else if (c.sense == ALTRO_INEQUALITY) {
#pragma unroll 1
                    for (int r = 0; r < c.p; ++r) v = fmax(v, row_value(c, G, h, z, zl, r));
                } else {
                    double a2 = 0.0, t = 0.0;
#pragma unroll 1
                    for (int r = 0; r < c.p; ++r) {
                        double cv = row_value(c, G, h, z, zl, r);
                        if (r < c.p - 1) a2 += cv * cv;
                        else t = cv;
                    }
                    double a = sqrt(a2);
                    if (!P.o.soc_viol_proj) {
                        v = fmax(v, a - t);
                    } else if (a <= -t) {  // projection is 0: distance = |c|_inf
#pragma unroll 1
                        for (int r = 0; r < c.p; ++r) v = fmax(v, fabs(row_value(c, G, h, z, zl, r)));
                    } else if (a > t) {  // c - Pi(c) = ((1-cf) v, t - cf a)
                        double cf = 0.5 * (1.0 + t / a);
#pragma unroll 1
                        for (int r = 0; r < c.p - 1; ++r) v = fmax(v, fabs((1.0 - cf) * row_value(c, G, h, z, zl, r)));
                        v = fmax(v, fabs(t - cf * a));
                    }
                }
            }
        }
        return gmax<T>(v, red);
    }

    // dual_update!: one (block, knot) per thread.
    __device__ void dual_update()
    {
#pragma unroll 1
        for (int it = tid; it < ncon * N; it += T) {  // (block, knot) work items, block-major
            const int ci = it / N, k = it - ci * N;
            const ConDesc &c = cd[ci];
            const double mu_c = mu[ci];
            if (k >= c.k0 && k < c.k1) {
                const size_t di = con_idx(c, k);
                const double *G = as_global(c.G) + di * c.p * c.w, *h = as_global(c.h) + di * c.p;
                const double *z = c.side == ALTRO_STATE ? X + k * n : U + k * m;
                double zl[ZW];
                load_zl(c, z, zl);
                double *l = lam + c.dual_off + (k - c.k0) * c.p;
                if (c.sense == ALTRO_EQUALITY) {
#pragma unroll 1
                    for (int r = 0; r < c.p; ++r)
                        l[r] = fmin(fmax(l[r] + mu_c * row_value(c, G, h, z, zl, r), -P.o.dual_max), P.o.dual_max);
                } else if (c.sense == ALTRO_INEQUALITY) {
#pragma unroll 1
                    for (int r = 0; r < c.p; ++r)
                        l[r] = fmin(fmax(l[r] + mu_c * row_value(c, G, h, z, zl, r), 0.0), P.o.dual_max);
                } else {
                    double a2 = 0.0, t = 0.0;
#pragma unroll 1
                    for (int r = 0; r < c.p; ++r) {
                        double lb = l[r] - mu_c * row_value(c, G, h, z, zl, r);
                        l[r] = lb;
                        if (r < c.p - 1) a2 += lb * lb;
                        else t = lb;
                    }
                    double a = sqrt(a2);
                    if (a <= -t) {
#pragma unroll 1
                        for (int r = 0; r < c.p; ++r) l[r] = 0.0;
                    } else if (a > t) {
                        double cf = 0.5 * (1.0 + t / a);
#pragma unroll 1
                        for (int r = 0; r < c.p - 1; ++r) l[r] *= cf;
                        l[c.p - 1] = cf * a;
                    }
                }
            }
        }
        gsync<T>();
    }

    // ---------------------------------------------------------------- AL expansion
    // Gradient g[w] and Hessian (w x w dense, or w diagonal for row-sparse blocks) of the AL term of
    // every (block, knot) into the expansion scratch; one work item per thread.
    __device__ void expand_constraints()
    {
#pragma unroll 1
        for (int it = tid; it < ncon * N; it += T) {  // (block, knot) work items, block-major
            const int ci = it / N, k = it - ci * N;
            const ConDesc &c = cd[ci];
            const double mu_c = mu[ci];
            const int w = c.w, p = c.p;
            if (k >= c.k0 && k < c.k1) {
                const size_t di = con_idx(c, k);
                const double *G = as_global(c.G) + di * p * w, *h = as_global(c.h) + di * p;
                const double *z = c.side == ALTRO_STATE ? X + k * n : U + k * m;
                double zl[ZW];
                load_zl(c, z, zl);
                const double *l = lam + c.dual_off + (k - c.k0) * p;
                double *g = ex + c.ex_off + (k - c.k0) * c.ex_stride;
                double *H = g + w;
                if (c.rowsparse) {
#pragma unroll 1
                    for (int j = 0; j < 2 * w; ++j) g[j] = 0.0;
#pragma unroll 1
                    for (int r = 0; r < p; ++r) {
                        double v = row_value(c, G, h, z, zl, r), cf = as_global(c.rs_coef)[r];
                        bool act = c.sense == ALTRO_EQUALITY || (v >= 0.0) || (l[r] > 0.0);
                        int col = as_global(c.rs_col)[r];
                        g[col] += cf * (l[r] + (act ? mu_c * v : 0.0));
                        H[col] += act ? cf * cf * mu_c : 0.0;
                    }
                } else if (c.sense != ALTRO_SECOND_ORDER_CONE) {
                    double y[PMAX], D[PMAX];
#pragma unroll 1
                    for (int r = 0; r < p; ++r) {
                        double v = row_value(c, G, h, z, zl, r);
                        bool act = c.sense == ALTRO_EQUALITY || (v >= 0.0) || (l[r] > 0.0);
                        y[r] = l[r] + (act ? mu_c * v : 0.0);
                        D[r] = act ? mu_c : 0.0;
                    }
#pragma unroll 1
                    for (int j = 0; j < w; ++j) {
                        double acc = 0.0;
#pragma unroll 1
                        for (int r = 0; r < p; ++r) acc = fma(G[r * w + j], y[r], acc);
                        g[j] = acc;
                    }
#pragma unroll 1
                    for (int i = 0, at = 0; i < w; ++i)
#pragma unroll 1
                        for (int j = i; j < w; ++j, ++at) {
                            double acc = 0.0;
#pragma unroll 1
                            for (int r = 0; r < p; ++r) acc = fma(G[r * w + i] * D[r], G[r * w + j], acc);
                            H[at] = acc;
                        }
                } else {
                    // lb = lam - mu c ; Pi(lb) ; g = -G' Pi(lb) ; H = mu G' dPi(lb) G  (structured)
                    double lb[PMAX], q[DENSE_W];
                    double a2 = 0.0;
#pragma unroll 1
                    for (int r = 0; r < p; ++r) {
                        lb[r] = l[r] - mu_c * row_value(c, G, h, z, zl, r);
                        if (r < p - 1) a2 += lb[r] * lb[r];
                    }
                    const double t = lb[p - 1], a = sqrt(a2);
                    const double *gt = G + (p - 1) * w;
                    if (a <= -t) {
#pragma unroll 1
                        for (int j = 0; j < c.ex_stride; ++j) g[j] = 0.0;
                    } else if (a <= t) {
#pragma unroll 1
                        for (int j = 0; j < w; ++j) {
                            double acc = 0.0;
#pragma unroll 1
                            for (int r = 0; r < p; ++r) acc = fma(G[r * w + j], lb[r], acc);
                            g[j] = -acc;
                        }
#pragma unroll 1
                        for (int i = 0, at = 0; i < w; ++i)
#pragma unroll 1
                            for (int j = i; j < w; ++j, ++at) {
                                double acc = 0.0;
#pragma unroll 1
                                for (int r = 0; r < p; ++r) acc = fma(G[r * w + i], G[r * w + j], acc);
                                H[at] = mu_c * acc;
                            }
                    } else {
                        const double ia = 1.0 / a, cf = 0.5 * (1.0 + t * ia);
                        const double cx = P.o.soc_hess_exact ? cf : cf * cf;
                        // q = G' [xhat; 1],  xhat = v / a
#pragma unroll 1
                        for (int j = 0; j < w; ++j) {
                            double acc = 0.0;
#pragma unroll 1
                            for (int r = 0; r < p - 1; ++r) acc = fma(G[r * w + j], lb[r], acc);
                            q[j] = acc * ia;  // omega_hat
                        }
#pragma unroll 1
                        for (int i = 0, at = 0; i < w; ++i)
#pragma unroll 1
                            for (int j = i; j < w; ++j, ++at) {
                                double gg = 0.0;
#pragma unroll 1
                                for (int r = 0; r < p - 1; ++r) gg = fma(G[r * w + i], G[r * w + j], gg);
                                double qi = q[i] + gt[i], qj = q[j] + gt[j];
                                H[at] = mu_c * (cx * (gg - q[i] * q[j]) + 0.5 * qi * qj);
                            }
#pragma unroll 1
                        for (int j = 0; j < w; ++j) g[j] = -cf * a * (q[j] + gt[j]);
                    }
                }
            }
        }
        gsync<T>();
    }

    // Sum of the AL expansion entries that land on target t of [Qx | Qxx | Qu | Quu] at knot k, added to `base`
    // in ascending block order (the order the oracle's scatter uses).
    __device__ __forceinline__ double gather(int t, int k, double base) const
    {
#pragma unroll 1
        for (int q = gptr[t]; q < gptr[t + 1]; ++q) {
            const int4 r = grec[q];  // {k0, k1, offset of the entry at knot 0, stride per knot}
            if (k >= r.x && k < r.y) base += ex[r.z + k * r.w];
        }
        return base;
    }

    // Everything knot k needs that does not depend on the cost-to-go: its LTV dynamics staged into shared memory and
    // [Qx | Qxx | Qu | Quu] = cost expansion + AL expansion gathered into Qi.  Done by threads t0, t0+stride, ...:
    // the warps that would idle while warp 0 factorises Quu of knot k+1 prepare knot k (backward_pass, P3).
    __device__ __forceinline__ void prep_knot(int k, int t0, int stride)
    {
        if (!P.dyn_in_smem && !(NX == 0 && P.lay.big)) {  // d_k is not needed by the backward pass
            const double *gA = as_global(P.A) + dyn_index(k) * n * n;
            const double *gB = as_global(P.Bm) + dyn_index(k) * n * m;
#pragma unroll 2
            for (int i = t0; i < n * n; i += stride) sA[i] = gA[i];
#pragma unroll 2
            for (int i = t0; i < n * m; i += stride) sB[i] = gB[i];
        }
        const int oQxx = n, oQu = n + n * n, oQuu = n + n * n + m;
#pragma unroll 1
        for (int q = t0; q < P.NTL; q += stride) {  // the vectors and the matrix entries some block touches
            const int t = gtl[q];
            double base;
            if (t < oQxx) base = P.dt * Qd[t] * (X[k * n + t] - xr[k * n + t]);
            else if (t < oQu) {
                const int e = t - oQxx, i = e / n, j = e - i * n;
                base = (i == j) ? P.dt * Qd[i] : 0.0;
            } else if (t < oQuu) {
                const int i = t - oQu;
                base = P.dt * Rd[i] * (U[k * m + i] - ur[k * m + i]);
            } else {
                const int e = t - oQuu, i = e / m, j = e - i * m;
                base = (i == j) ? P.dt * Rd[i] : 0.0;
            }
            Qi[t] = gather(t, k, base);
        }
    }

    // ---------------------------------------------------------------- backward pass (A.7)
    __device__ void reg_increase(double &rho, double &drho) const
    {
        drho = fmax(drho * P.o.bp_reg_increase_factor, P.o.bp_reg_increase_factor);
        rho = fmax(rho * drho, P.o.bp_reg_min);
    }
    __device__ void reg_decrease(double &rho, double &drho) const
    {
        drho = fmin(drho / P.o.bp_reg_increase_factor, 1.0 / P.o.bp_reg_increase_factor);
        double r = rho * drho;
        rho = (r > P.o.bp_reg_min) ? r : 0.0;
    }

    // ---- FP64 tensor-core tiles.  One warp owns one 8x8 output tile:  C += sum_k A(i,k) B(k,j), k ascending.
    // mma.sync.m8n8k4.f64 accumulates exactly like the scalar chain  acc = fma(a_k, b_k, acc), k = 0..3, starting
    // from C (probed on B200: scripts/probes/dmma_order.cu, 1.28 M elements bit-identical), so the CPU oracle's
    // plain fma loops reproduce every tile bit for bit.  Operands outside the matrix are fetched as 0.0.
    // Fragment layout: lane l holds A[l/4][l%4], B[l%4][l/4], C[l/4][2(l%4)] and C[l/4][2(l%4)+1].
    template <class FA, class FB>
    __device__ __forceinline__ void mma_chain(double &c0, double &c1, int K, FA a_at, FB b_at) const
    {
        const int lane = tid & 31, r = lane >> 2, q = lane & 3;
        for (int k0 = 0; k0 < K; k0 += 4) {
            const double a = a_at(r, k0 + q), b = b_at(k0 + q, r);
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0), "+d"(c1)
                         : "d"(a), "d"(b));
        }
    }

    // Two tiles interleaved: the chains of tile 0 and tile 1 are independent, so issuing their k-steps alternately hides
    // the shared-memory and DMMA latency of one behind the other (a single chain is LDS -> DMMA -> LDS -> DMMA ...,
    // about 60 cycles per k-step with nothing else to issue).  Each chain is unchanged.  `two` is warp-uniform; the
    // operand functors take the slot (0 / 1) first and must be safe to evaluate for slot 1 even when it is unused.
    template <class FA, class FB>
    __device__ __forceinline__ void mma_chain2(double (&c)[2][2], bool two, int K, FA a_at, FB b_at) const
    {
        const int lane = tid & 31, r = lane >> 2, q = lane & 3;
        for (int k0 = 0; k0 < K; k0 += 4) {
            const double a0 = a_at(0, r, k0 + q), b0 = b_at(0, k0 + q, r);
            const double a1 = a_at(1, r, k0 + q), b1 = b_at(1, k0 + q, r);
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[0][0]), "+d"(c[0][1]) : "d"(a0), "d"(b0));
            if (two)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[1][0]), "+d"(c[1][1]) : "d"(a1), "d"(b1));
        }
    }
    // The tiles of one product, dealt to the warps round-robin continuing a running tile index `goff` (so that products
    // with a single tile do not all land on warp 0), two of a warp's tiles at a time.  body(t0, t1, two).
    template <class F>
    __device__ __forceinline__ void tile_pairs(int nt, int goff, F body) const
    {
        constexpr int NW = T / 32;
        const int warp = tid >> 5;
        for (int t = (warp + NW - goff % NW) % NW; t < nt; t += 2 * NW) body(t, t + NW < nt ? t + NW : t, t + NW < nt);
    }

    // Register-blocked tile GEMM for the large-dimension path:  C(i,j) = init(i,j) + sum_l a1(i,l) b1(l,j)
    // [+ sum_l a2(i,l) b2(l,j)], every chain over ascending l exactly like mma_chain.  A warp owns a 16 x 32 block of C
    // (2 x 4 DMMA tiles: two A and four B fragments feed eight independent accumulator chains per k-step), the
    // blocks are dealt round-robin to the warps.  Operand functors return 0.0 outside the matrix.
    template <class FI, class FA1, class FB1, class FA2, class FB2, class FS>
    __device__ __forceinline__ void tile_gemm(int M, int Nc, FI init, int K1, FA1 a1, FB1 b1, int K2, FA2 a2, FB2 b2,
                                              FS store) const
    {
        // a warp owns 16 x (8 NB) of C: 2 A and NB B fragments feed 2 NB accumulator chains.  Measured: 16 x 32 blocks
        // pay from n = 128 up, 16 x 16 below (more blocks to deal out to the warps).
        if (n >= 128) tile_gemm_nb(std::integral_constant<int, 4>{}, M, Nc, init, K1, a1, b1, K2, a2, b2, store);
        else tile_gemm_nb(std::integral_constant<int, 2>{}, M, Nc, init, K1, a1, b1, K2, a2, b2, store);
    }
    template <int NB, class FI, class FA1, class FB1, class FA2, class FB2, class FS>
    __device__ __forceinline__ void tile_gemm_nb(std::integral_constant<int, NB>, int M, int Nc, FI init, int K1, FA1 a1,
                                                 FB1 b1, int K2, FA2 a2, FB2 b2, FS store) const
    {
        const int lane = tid & 31, r = lane >> 2, q = lane & 3, fc = 2 * q;
        const int nt = (Nc + 8 * NB - 1) / (8 * NB), nblk = ((M + 15) >> 4) * nt;
#pragma unroll 1
        for (int t = tid >> 5; t < nblk; t += T / 32) {
            const int i0 = (t / nt) << 4, j0 = (t - (t / nt) * nt) * (8 * NB);
            double c[2][NB][2];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    c[a][b][0] = init(i0 + 8 * a + r, j0 + 8 * b + fc);
                    c[a][b][1] = init(i0 + 8 * a + r, j0 + 8 * b + fc + 1);
                }
            auto step = [&](double x0, double x1, const double (&y)[NB]) {
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(c[0][b][0]), "+d"(c[0][b][1]) : "d"(x0), "d"(y[b]));
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(c[1][b][0]), "+d"(c[1][b][1]) : "d"(x1), "d"(y[b]));
                }
            };
            // Software pipeline: the fragments of k-step s + D - 1 are requested before the DMMAs of k-step s are issued,
            // so D - 1 k-steps of tensor work cover the L2 latency of the operand loads (the operands of this path live in
            // the per-instance global workspace).  The chain order of every element is unchanged.
            auto chain = [&](int K, auto &&fa, auto &&fb) {
                constexpr int D = ALTRO_GEMM_DEPTH;
                double xa[D][2], ya[D][NB];
                auto fetch = [&](int k, double (&x)[2], double (&y)[NB]) {
                    x[0] = fa(i0 + r, k + q);
                    x[1] = fa(i0 + 8 + r, k + q);
#pragma unroll
                    for (int b = 0; b < NB; ++b) y[b] = fb(k + q, j0 + 8 * b + r);
                };
#pragma unroll
                for (int s = 0; s < D - 1; ++s)
                    if (4 * s < K) fetch(4 * s, xa[s], ya[s]);
#pragma unroll 1
                for (int k0 = 0; k0 < K; k0 += 4 * D) {
#pragma unroll
                    for (int s = 0; s < D; ++s) {
                        const int k = k0 + 4 * s;
                        if (k < K) {
                            if (k + 4 * (D - 1) < K) fetch(k + 4 * (D - 1), xa[(s + D - 1) % D], ya[(s + D - 1) % D]);
                            step(xa[s][0], xa[s][1], ya[s]);
                        }
                    }
                }
            };
            chain(K1, a1, b1);
            chain(K2, a2, b2);
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    store(i0 + 8 * a + r, j0 + 8 * b + fc, c[a][b][0]);
                    store(i0 + 8 * a + r, j0 + 8 * b + fc + 1, c[a][b][1]);
                }
        }
    }

    // see panel_gemm_fn
    __device__ __forceinline__ void panel_gemm(const PanelArgs &g)
    {
        double *stage0 = reinterpret_cast<double *>(smem_base) + P.lay.stage;
        uint64_t *full = reinterpret_cast<uint64_t *>(reinterpret_cast<double *>(smem_base) + P.lay.mbar);
        pg_count += panel_gemm_fn<T>(stage0, full, P.lay.stage_doubles, pg_count, g);
    }
    __device__ __forceinline__ void panel_gemm(int M, int Nc, int K, const double *PA, const double *PB, const double *Ci,
                                               double *Co, double *Co2 = nullptr, double rho2 = 0.0)
    {
        PanelArgs g{M, Nc, K, 0, 0, PA, PB, nullptr, nullptr, Ci, nullptr, Co, Co2, rho2};
        panel_gemm(g);
    }

    // Returns false if Quu could not be made positive definite.
    __device__ bool backward_pass(double &rho, double &drho, double &dV1, double &dV2)
    {
        {
            const long long ce = clock64();
            expand_constraints();
            ph_exp += clock64() - ce;
        }
        constexpr int NW = T / 32;
        const int warp = tid >> 5, lane = tid & 31;
        const int fr = lane >> 2, fc = 2 * (lane & 3);  // this lane's row / first column inside a tile
        const int tn = (n + 7) >> 3, tm = (m + 7) >> 3, tn1 = (n + 8) >> 3, tm1 = (m + 8) >> 3;
        const int oQxx = n, oQu = n + n * n, oQuu = n + n * n + m;
        const double *A = sA, *Bm = sB;  // shared memory (LTV knots are staged by prep_knot), except for large problems
        for (;;) {
            bool restart = false;
            double a1 = 0.0, a2 = 0.0;  // dV accumulators, kept by thread T-1
            // terminal cost-to-go: S = Qf + state-side AL Hessian, s = Qf (x - xref) + AL gradient
            for (int t = tid; t < oQu; t += T) {
                if (t < n) s[t] = gather(t, N - 1, Qfd[t] * (X[(N - 1) * n + t] - xr[(N - 1) * n + t]));
                else {
                    const int e = t - n, i = e / n, j = e - i * n;
                    S[e] = gather(t, N - 1, (i == j) ? Qfd[i] : 0.0);
                }
            }
            prep_knot(N - 2, tid, T);
            gsync<T>();
            for (int k = N - 2; k >= 0; --k) {
#ifdef ALTRO_PHASE_TIMERS
                long long tq = clock64(), tq2;
#else
                long long tq = 0, tq2 = 0;
#endif
#ifdef ALTRO_PHASE_TIMERS  // development builds only (make EXTRA=-DALTRO_PHASE_TIMERS): 7 clock reads per knot
#define ALTRO_TICK(slot) do { tq2 = clock64(); bpc[slot] += tq2 - tq; tq = tq2; } while (0)
#else
#define ALTRO_TICK(slot) do { (void)tq; (void)tq2; } while (0)
#endif
                if (NX == 0 && P.lay.big) {  // large state dimension: the dynamics are read where they are
                    A = P.A + dyn_index(k) * n * n;
                    Bm = P.Bm + dyn_index(k) * n * m;
                }
                constexpr bool MAYBE_BIG = NX == 0;
                const bool big = MAYBE_BIG && P.lay.big;
                auto zero_a = [](int, int) { return 0.0; };
                // P1: SA = S A, SB = S B (tensor tiles); A_k, B_k and Qi were prepared during the previous knot
                const bool tma = WIDE && big && P.lay.tma;
                if (tma) {
                    if constexpr (WIDE) {
                        fence_proxy_async();  // S (and the first knot's terminal S) was written with ordinary stores
                        __syncthreads();
                        panel_gemm(n, n, n, S, A, nullptr, SA);
                        panel_gemm(n, m, n, S, Bm, nullptr, SB);
                    }
                } else if (big) {
                    tile_gemm(n, n, [](int, int) { return 0.0; }, n,
                              [&](int i, int l) { return (i < n && l < n) ? S[i * n + l] : 0.0; },
                              [&](int l, int j) { return (l < n && j < n) ? A[l * n + j] : 0.0; }, 0, zero_a, zero_a,
                              [&](int i, int j, double v) { if (i < n && j < n) SA[i * n + j] = v; });
                    tile_gemm(n, m, [](int, int) { return 0.0; }, n,
                              [&](int i, int l) { return (i < n && l < n) ? S[i * n + l] : 0.0; },
                              [&](int l, int j) { return (l < n && j < m) ? Bm[l * m + j] : 0.0; }, 0, zero_a, zero_a,
                              [&](int i, int j, double v) { if (i < n && j < m) SB[i * m + j] = v; });
                }
                if constexpr (NX > 0) {  // compiled dimensions: two tiles of a product interleaved per warp
                    // SA = S A (tn x tn tiles), then SB = S B (tn x tm tiles)
                    tile_pairs(tn * tn, 0, [&](int t0, int t1, bool two) {
                        const int r0[2] = {(t0 / tn) << 3, (t1 / tn) << 3}, c0[2] = {(t0 % tn) << 3, (t1 % tn) << 3};
                        double c[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
                        mma_chain2(c, two, n,
                                   [&](int u, int ii, int l) { ii += r0[u]; return (ii < n && l < n) ? S[ii * n + l] : 0.0; },
                                   [&](int u, int l, int jj) { jj += c0[u]; return (l < n && jj < n) ? A[l * n + jj] : 0.0; });
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int i = r0[u] + fr, j = c0[u] + fc;
                            if ((u == 0 || two) && i < n) {
                                if (j < n) SA[i * n + j] = c[u][0];
                                if (j + 1 < n) SA[i * n + j + 1] = c[u][1];
                            }
                        }
                    });
                    tile_pairs(tn * tm, tn * tn, [&](int t0, int t1, bool two) {
                        const int r0[2] = {(t0 / tm) << 3, (t1 / tm) << 3}, c0[2] = {(t0 % tm) << 3, (t1 % tm) << 3};
                        double c[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
                        mma_chain2(c, two, n,
                                   [&](int u, int ii, int l) { ii += r0[u]; return (ii < n && l < n) ? S[ii * n + l] : 0.0; },
                                   [&](int u, int l, int jj) { jj += c0[u]; return (l < n && jj < m) ? Bm[l * m + jj] : 0.0; });
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int i = r0[u] + fr, j = c0[u] + fc;
                            if ((u == 0 || two) && i < n) {
                                if (j < m) SB[i * m + j] = c[u][0];
                                if (j + 1 < m) SB[i * m + j + 1] = c[u][1];
                            }
                        }
                    });
                } else {
                // run-time dimensions: one tile at a time (measured: the interleaved form costs this kernel 10 % at small n)
                for (int t = warp; !big && t < tn * (tn + tm); t += NW) {
                    const int r0 = (t / (tn + tm)) << 3, ct = t % (tn + tm);
                    const int i = r0 + fr;
                    double d0 = 0.0, d1 = 0.0;
                    if (ct < tn) {
                        const int c0 = ct << 3, j = c0 + fc;
                        mma_chain(d0, d1, n,
                                  [&](int ii, int l) { ii += r0; return (ii < n && l < n) ? S[ii * n + l] : 0.0; },
                                  [&](int l, int jj) { jj += c0; return (l < n && jj < n) ? A[l * n + jj] : 0.0; });
                        if (i < n) {
                            if (j < n) SA[i * n + j] = d0;
                            if (j + 1 < n) SA[i * n + j + 1] = d1;
                        }
                    } else {
                        const int c0 = (ct - tn) << 3, j = c0 + fc;
                        mma_chain(d0, d1, n,
                                  [&](int ii, int l) { ii += r0; return (ii < n && l < n) ? S[ii * n + l] : 0.0; },
                                  [&](int l, int jj) { jj += c0; return (l < n && jj < m) ? Bm[l * m + jj] : 0.0; });
                        if (i < n) {
                            if (j < m) SB[i * m + j] = d0;
                            if (j + 1 < m) SB[i * m + j + 1] = d1;
                        }
                    }
                }
                }
                gsync<T>();
                ALTRO_TICK(0);
                // P2: [Qxx | Qx] += A'[SA | s],  Qux = B'SA,  [Quu | Qu] += B'[SB | s]   (chains start from the
                //     cost + AL expansion already in Qxx / Qx / Quu / Qu)
                const int nt_xx = tn * tn1, nt_ux = tm * tn, nt_uu = tm * tm1;
                if (tma) {
                    if constexpr (WIDE) {
                        panel_gemm(n, n, n, A, SA, Qi + oQxx, Qxx);
                        panel_gemm(m, n, n, Bm, SA, nullptr, Qux);
                        panel_gemm(m, m, n, Bm, SB, Qi + oQuu, Quu, L, rho);
                        // the vector columns of the two tiles: Qx = Qi + A's, Qu = Qi + B's (same chains, one thread each)
                        for (int i = tid; i < n + m; i += T) {
                            if (i < n) {
                                double acc = Qi[i];
#pragma unroll 8
                                for (int l = 0; l < n; ++l) acc = fma(A[l * n + i], s[l], acc);
                                Qx[i] = acc;
                            } else {
                                const int iu = i - n;
                                double acc = Qi[oQu + iu];
#pragma unroll 8
                                for (int l = 0; l < n; ++l) acc = fma(Bm[l * m + iu], s[l], acc);
                                Qu[iu] = acc;
                            }
                        }
                    }
                } else if (big) {
                    tile_gemm(n, n + 1,
                              [&](int i, int j) { return i < n ? (j < n ? Qi[oQxx + i * n + j] : (j == n ? Qi[i] : 0.0)) : 0.0; }, n,
                              [&](int i, int l) { return (i < n && l < n) ? A[l * n + i] : 0.0; },
                              [&](int l, int j) { return (l < n && j <= n) ? (j < n ? SA[l * n + j] : s[l]) : 0.0; },
                              0, zero_a, zero_a,
                              [&](int i, int j, double v) { if (i < n) { if (j < n) Qxx[i * n + j] = v; else if (j == n) Qx[i] = v; } });
                    tile_gemm(m, n, [](int, int) { return 0.0; }, n,
                              [&](int i, int l) { return (i < m && l < n) ? Bm[l * m + i] : 0.0; },
                              [&](int l, int j) { return (l < n && j < n) ? SA[l * n + j] : 0.0; }, 0, zero_a, zero_a,
                              [&](int i, int j, double v) { if (i < m && j < n) Qux[i * n + j] = v; });
                    tile_gemm(m, m + 1,
                              [&](int i, int j) { return i < m ? (j < m ? Qi[oQuu + i * m + j] : (j == m ? Qi[oQu + i] : 0.0)) : 0.0; }, n,
                              [&](int i, int l) { return (i < m && l < n) ? Bm[l * m + i] : 0.0; },
                              [&](int l, int j) { return (l < n && j <= m) ? (j < m ? SB[l * m + j] : s[l]) : 0.0; },
                              0, zero_a, zero_a,
                              [&](int i, int j, double v) {
                                  if (i < m) {
                                      if (j < m) { Quu[i * m + j] = v; L[i * m + j] = v + ((i == j) ? rho : 0.0); }
                                      else if (j == m) Qu[i] = v;
                                  }
                              });
                }
                if constexpr (NX > 0) {
                    tile_pairs(nt_xx, 0, [&](int t0, int t1, bool two) {
                        const int r0[2] = {(t0 / tn1) << 3, (t1 / tn1) << 3}, c0[2] = {(t0 % tn1) << 3, (t1 % tn1) << 3};
                        double c[2][2];
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int i = r0[u] + fr, j = c0[u] + fc;
                            c[u][0] = (i < n) ? (j < n ? Qi[oQxx + i * n + j] : (j == n ? Qi[i] : 0.0)) : 0.0;
                            c[u][1] = (i < n) ? (j + 1 < n ? Qi[oQxx + i * n + j + 1] : (j + 1 == n ? Qi[i] : 0.0)) : 0.0;
                        }
                        mma_chain2(c, two, n,
                                   [&](int u, int ii, int l) { ii += r0[u]; return (ii < n && l < n) ? A[l * n + ii] : 0.0; },
                                   [&](int u, int l, int jj) {
                                       jj += c0[u];
                                       const double *p = jj < n ? SA + l * n + jj : s + l;
                                       return (l < n && jj <= n) ? *p : 0.0;
                                   });
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int i = r0[u] + fr, j = c0[u] + fc;
                            if ((u == 0 || two) && i < n) {
                                if (j < n) Qxx[i * n + j] = c[u][0]; else if (j == n) Qx[i] = c[u][0];
                                if (j + 1 < n) Qxx[i * n + j + 1] = c[u][1]; else if (j + 1 == n) Qx[i] = c[u][1];
                            }
                        }
                    });
                    tile_pairs(nt_ux, nt_xx, [&](int t0, int t1, bool two) {
                        const int r0[2] = {(t0 / tn) << 3, (t1 / tn) << 3}, c0[2] = {(t0 % tn) << 3, (t1 % tn) << 3};
                        double c[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
                        mma_chain2(c, two, n,
                                   [&](int u, int ii, int l) { ii += r0[u]; return (ii < m && l < n) ? Bm[l * m + ii] : 0.0; },
                                   [&](int u, int l, int jj) { jj += c0[u]; return (l < n && jj < n) ? SA[l * n + jj] : 0.0; });
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int i = r0[u] + fr, j = c0[u] + fc;
                            if ((u == 0 || two) && i < m) {
                                if (j < n) Qux[i * n + j] = c[u][0];
                                if (j + 1 < n) Qux[i * n + j + 1] = c[u][1];
                            }
                        }
                    });
                    tile_pairs(nt_uu, nt_xx + nt_ux, [&](int t0, int t1, bool two) {
                        const int r0[2] = {(t0 / tm1) << 3, (t1 / tm1) << 3}, c0[2] = {(t0 % tm1) << 3, (t1 % tm1) << 3};
                        double c[2][2];
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int i = r0[u] + fr, j = c0[u] + fc;
                            c[u][0] = (i < m) ? (j < m ? Qi[oQuu + i * m + j] : (j == m ? Qi[oQu + i] : 0.0)) : 0.0;
                            c[u][1] = (i < m) ? (j + 1 < m ? Qi[oQuu + i * m + j + 1] : (j + 1 == m ? Qi[oQu + i] : 0.0)) : 0.0;
                        }
                        mma_chain2(c, two, n,
                                   [&](int u, int ii, int l) { ii += r0[u]; return (ii < m && l < n) ? Bm[l * m + ii] : 0.0; },
                                   [&](int u, int l, int jj) {
                                       jj += c0[u];
                                       const double *p = jj < m ? SB + l * m + jj : s + l;
                                       return (l < n && jj <= m) ? *p : 0.0;
                                   });
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int i = r0[u] + fr, j = c0[u] + fc;
                            if ((u == 0 || two) && i < m) {  // L = Quu + rho I: regularised copy for the factorisation
                                if (j < m) { Quu[i * m + j] = c[u][0]; L[i * m + j] = c[u][0] + ((i == j) ? rho : 0.0); }
                                else if (j == m) Qu[i] = c[u][0];
                                if (j + 1 < m) { Quu[i * m + j + 1] = c[u][1]; L[i * m + j + 1] = c[u][1] + ((i == j + 1) ? rho : 0.0); }
                                else if (j + 1 == m) Qu[i] = c[u][1];
                            }
                        }
                    });
                } else {
                for (int t = warp; !big && t < nt_xx + nt_ux + nt_uu; t += NW) {
                    if (t < nt_xx) {
                        const int r0 = (t / tn1) << 3, c0 = (t % tn1) << 3;
                        const int i = r0 + fr, j = c0 + fc;
                        double d0 = (i < n) ? (j < n ? Qi[oQxx + i * n + j] : (j == n ? Qi[i] : 0.0)) : 0.0;
                        double d1 = (i < n) ? (j + 1 < n ? Qi[oQxx + i * n + j + 1] : (j + 1 == n ? Qi[i] : 0.0)) : 0.0;
                        mma_chain(d0, d1, n,
                                  [&](int ii, int l) { ii += r0; return (ii < n && l < n) ? A[l * n + ii] : 0.0; },
                                  [&](int l, int jj) {
                                      jj += c0;
                                      const double *p = jj < n ? SA + l * n + jj : s + l;
                                      return (l < n && jj <= n) ? *p : 0.0;
                                  });
                        if (i < n) {
                            if (j < n) Qxx[i * n + j] = d0; else if (j == n) Qx[i] = d0;
                            if (j + 1 < n) Qxx[i * n + j + 1] = d1; else if (j + 1 == n) Qx[i] = d1;
                        }
                    } else if (t < nt_xx + nt_ux) {
                        const int u = t - nt_xx, r0 = (u / tn) << 3, c0 = (u % tn) << 3;
                        const int i = r0 + fr, j = c0 + fc;
                        double d0 = 0.0, d1 = 0.0;
                        mma_chain(d0, d1, n,
                                  [&](int ii, int l) { ii += r0; return (ii < m && l < n) ? Bm[l * m + ii] : 0.0; },
                                  [&](int l, int jj) { jj += c0; return (l < n && jj < n) ? SA[l * n + jj] : 0.0; });
                        if (i < m) {
                            if (j < n) Qux[i * n + j] = d0;
                            if (j + 1 < n) Qux[i * n + j + 1] = d1;
                        }
                    } else {
                        const int u = t - nt_xx - nt_ux, r0 = (u / tm1) << 3, c0 = (u % tm1) << 3;
                        const int i = r0 + fr, j = c0 + fc;
                        double d0 = (i < m) ? (j < m ? Qi[oQuu + i * m + j] : (j == m ? Qi[oQu + i] : 0.0)) : 0.0;
                        double d1 = (i < m) ? (j + 1 < m ? Qi[oQuu + i * m + j + 1] : (j + 1 == m ? Qi[oQu + i] : 0.0)) : 0.0;
                        mma_chain(d0, d1, n,
                                  [&](int ii, int l) { ii += r0; return (ii < m && l < n) ? Bm[l * m + ii] : 0.0; },
                                  [&](int l, int jj) {
                                      jj += c0;
                                      const double *p = jj < m ? SB + l * m + jj : s + l;
                                      return (l < n && jj <= m) ? *p : 0.0;
                                  });
                        if (i < m) {  // L = Quu + rho I: regularised copy for the factorisation
                            if (j < m) { Quu[i * m + j] = d0; L[i * m + j] = d0 + ((i == j) ? rho : 0.0); }
                            else if (j == m) Qu[i] = d0;
                            if (j + 1 < m) { Quu[i * m + j + 1] = d1; L[i * m + j + 1] = d1 + ((i == j + 1) ? rho : 0.0); }
                            else if (j + 1 == m) Qu[i] = d1;
                        }
                    }
                }
                }
                gsync<T>();
                ALTRO_TICK(2);
                // P3 + P4: LDL' of Quu + rho I, kept as the un-normalised lower factor X (L = X diag(r)) and the
                //     reciprocal pivots r = 1/D (one division per column, none per entry), then
                //     [K | d] = -(Quu + rho I)^-1 [Qux | Qu]: forward y = L^-1 b, z = y r, backward
                //     x_i = z_i - r_i sum_{q>i} X[q][i] x_q (q descending).  One right-hand side per thread.
                bool bad = false;
                double *Kk = K + k * m * n, *dk_ = dv + k * m;
                if (NU > 0 && NU <= 16) {
                    // Small / medium control dimension: warp 0 factorises and substitutes in registers while the other
                    // warps (if any) prepare the next knot -- its dynamics and gathered expansion do not depend on
                    // the cost-to-go, so the gather and the global-memory latency hide behind the factorisation.
                    if (warp == 0) {
                        if (NU <= 4) {
                            constexpr int MM = (NU > 0 && NU <= 4) ? NU : 1;
                            double Xr[MM * MM], rr[MM];
#pragma unroll
                            for (int j = 0; j < MM; ++j) {
#pragma unroll
                                for (int i = j; i < MM; ++i) {
                                    double acc = L[i * MM + j];
#pragma unroll
                                    for (int l = 0; l < j; ++l) acc = fma(-Xr[i * MM + l], Xr[j * MM + l] * rr[l], acc);
                                    Xr[i * MM + j] = acc;
                                }
                                if (!(Xr[j * MM + j] > 0.0)) { bad = true; break; }
                                rr[j] = __drcp_rn(Xr[j * MM + j]);
                            }
                            if (!bad) {
#pragma unroll 1
                                for (int c = lane; c <= n; c += 32) {
                                    double *bp = (c < n) ? Kk + c : dk_;
                                    const double *src = (c < n) ? Qux + c : Qu;
                                    const int st = (c < n) ? n : 1;
                                    double bb[MM];
#pragma unroll
                                    for (int i = 0; i < MM; ++i) {
                                        double acc = -src[i * st];
#pragma unroll
                                        for (int l = 0; l < i; ++l) acc = fma(-Xr[i * MM + l], bb[l] * rr[l], acc);
                                        bb[i] = acc;
                                    }
#pragma unroll
                                    for (int i = 0; i < MM; ++i) bb[i] = bb[i] * rr[i];
#pragma unroll
                                    for (int i = MM - 1; i >= 0; --i) {
                                        double acc2 = 0.0;
#pragma unroll
                                        for (int l = MM - 1; l > i; --l) acc2 = fma(Xr[l * MM + i], bb[l], acc2);
                                        bb[i] = fma(-rr[i], acc2, bb[i]);
                                    }
#pragma unroll
                                    for (int i = 0; i < MM; ++i) bp[i * st] = bb[i];
                                }
                            }
                        } else {
                            bad = ldl_solve_medium<(NU > 4 && NU <= 16) ? NU : 5>(L, Qux, Qu, Kk, dk_, n);
                        }
                        if (T > 32 && lane == 0) bc[5] = bad ? 1.0 : 0.0;
                        if (T == 32 && !bad && k > 0) prep_knot(k - 1, tid, T);
                    } else if (k > 0) {
                        prep_knot(k - 1, tid - 32, T - 32);
                    }
                    ALTRO_TICK(1);  // warp 0: its own factorisation + substitution
                    if (T > 32) {
                        gsync<T>();
                        bad = bc[5] != 0.0;
                    }
                    ALTRO_TICK(3);  // wait for the warps that prepared the next knot
                } else {
                    if (m <= 32) {  // warp 0 alone, one lane per row, one __syncwarp per column
                        if (warp == 0) {
                            double rprev = 0.0;  // reciprocal pivot of the previous column (linv[j-1] is still in flight)
                            for (int j = 0; j < m; ++j) {
                                if (lane >= j && lane < m) {
                                    double acc = L[lane * m + j];
                                    for (int l = 0; l < j; ++l)
                                        acc = fma(-L[lane * m + l], L[j * m + l] * (l == j - 1 ? rprev : linv[l]), acc);
                                    L[lane * m + j] = acc;
                                }
                                __syncwarp();
                                const double piv = L[j * m + j];
                                if (!(piv > 0.0)) { bad = true; break; }
                                rprev = __drcp_rn(piv);
                                if (lane == 0) linv[j] = rprev;
                            }
                            if (lane == 0) bc[5] = bad ? 1.0 : 0.0;
                        }
                        gsync<T>();
                        bad = bc[5] != 0.0;
                    } else {
                        for (int j = 0; j < m; ++j) {
                            for (int i = j + tid; i < m; i += T) {
                                double acc = L[i * m + j];
                                for (int l = 0; l < j; ++l) acc = fma(-L[i * m + l], L[j * m + l] * linv[l], acc);
                                L[i * m + j] = acc;
                            }
                            gsync<T>();
                            const double piv = L[j * m + j];
                            if (!(piv > 0.0)) { bad = true; break; }  // same value in every thread
                            if (tid == 0) linv[j] = __drcp_rn(piv);
                            gsync<T>();
                        }
                    }
                    ALTRO_TICK(3);
                    if (!bad) {
                        for (int c = tid; c <= n; c += T) {  // serial substitution per right-hand side, in place
                            double *bp = (c < n) ? Kk + c : dk_;
                            const double *src = (c < n) ? Qux + c : Qu;
                            const int st = (c < n) ? n : 1;
                            for (int i = 0; i < m; ++i) {
                                double acc = -src[i * st];
                                for (int l = 0; l < i; ++l) acc = fma(-L[i * m + l], bp[l * st] * linv[l], acc);
                                bp[i * st] = acc;
                            }
                            for (int i = 0; i < m; ++i) bp[i * st] = bp[i * st] * linv[i];
                            for (int i = m - 1; i >= 0; --i) {
                                double acc2 = 0.0;
                                for (int l = m - 1; l > i; --l) acc2 = fma(L[l * m + i], bp[l * st], acc2);
                                bp[i * st] = fma(-linv[i], acc2, bp[i * st]);
                            }
                        }
                    }
                    if (!bad && k > 0) prep_knot(k - 1, tid, T);
                }
                if (bad) { restart = true; break; }
                gsync<T>();
                ALTRO_TICK(4);
                // P5: [T1 | t1] = Quu [K | d] + [Qux | Qu]
                if (big)
                    tile_gemm(m, n + 1,
                              [&](int i, int j) { return i < m ? (j < n ? Qux[i * n + j] : (j == n ? Qu[i] : 0.0)) : 0.0; }, m,
                              [&](int i, int l) { return (i < m && l < m) ? Quu[i * m + l] : 0.0; },
                              [&](int l, int j) { return (l < m && j <= n) ? (j < n ? Kk[l * n + j] : dk_[l]) : 0.0; },
                              0, zero_a, zero_a,
                              [&](int i, int j, double v) { if (i < m) { if (j < n) T1[i * n + j] = v; else if (j == n) t1[i] = v; } });
                for (int t = warp; !big && t < tm * tn1; t += NW) {
                    const int r0 = (t / tn1) << 3, c0 = (t % tn1) << 3;
                    const int i = r0 + fr, j = c0 + fc;
                    double d0 = (i < m) ? (j < n ? Qux[i * n + j] : (j == n ? Qu[i] : 0.0)) : 0.0;
                    double d1 = (i < m) ? (j + 1 < n ? Qux[i * n + j + 1] : (j + 1 == n ? Qu[i] : 0.0)) : 0.0;
                    mma_chain(d0, d1, m,
                              [&](int ii, int l) { ii += r0; return (ii < m && l < m) ? Quu[ii * m + l] : 0.0; },
                              [&](int l, int jj) {
                                  jj += c0;
                                  const double *p = jj < n ? Kk + l * n + jj : dk_ + l;
                                  return (l < m && jj <= n) ? *p : 0.0;
                              });
                    __syncwarp();
                    if (i < m) {
                        if (j < n) T1[i * n + j] = d0; else if (j == n) t1[i] = d0;
                        if (j + 1 < n) T1[i * n + j + 1] = d1; else if (j + 1 == n) t1[i] = d1;
                    }
                }
                gsync<T>();
                ALTRO_TICK(5);
                // P6: [S' | s] = [Qxx | Qx] + K'[T1 | t1] + Qux'[K | d]; the transposed tile is chained in the same
                //     lanes so that S = (S' + S'^T)/2 needs no second pass.  dV += [d'Qu, 1/2 d'Quu d].
                if (tma) {
                    // the same chains through the staged panels: D = Qxx + K'T1 + Qux'K into SA, E = Qxx' + T1'K + K'Qux
                    // with the symmetrisation S = (D + E)/2 fused into the last store; the vector column by plain chains
                    if constexpr (WIDE) {
                        fence_proxy_async();  // K, T1 were written with ordinary stores
                        __syncthreads();
                        panel_gemm(PanelArgs{n, n, m, m, 0, Kk, T1, Qux, Kk, Qxx, nullptr, SA, nullptr, 0.0});
                        panel_gemm(PanelArgs{n, n, m, m, 1, T1, Kk, Kk, Qux, Qxx, SA, S, nullptr, 0.0});
#pragma unroll 1
                        for (int i = tid; i < n; i += T) {
                            double acc = Qx[i];
#pragma unroll 8
                            for (int l = 0; l < m; ++l) acc = fma(Kk[l * n + i], t1[l], acc);
#pragma unroll 8
                            for (int l = 0; l < m; ++l) acc = fma(Qux[l * n + i], dk_[l], acc);
                            s[i] = acc;
                        }
                    }
                } else if (big) {
                    // D = [Qxx | Qx] + K'[T1 | t1] + Qux'[K | d] into SA (free since P2) and s; E = Qxx' + T1'K + K'Qux
                    // into S (not read since P1); then S = (D + E)/2 element by element.
                    tile_gemm(n, n + 1,
                              [&](int i, int j) { return i < n ? (j < n ? Qxx[i * n + j] : (j == n ? Qx[i] : 0.0)) : 0.0; }, m,
                              [&](int i, int l) { return (i < n && l < m) ? Kk[l * n + i] : 0.0; },
                              [&](int l, int j) { return (l < m && j <= n) ? (j < n ? T1[l * n + j] : t1[l]) : 0.0; }, m,
                              [&](int i, int l) { return (i < n && l < m) ? Qux[l * n + i] : 0.0; },
                              [&](int l, int j) { return (l < m && j <= n) ? (j < n ? Kk[l * n + j] : dk_[l]) : 0.0; },
                              [&](int i, int j, double v) { if (i < n) { if (j < n) SA[i * n + j] = v; else if (j == n) s[i] = v; } });
                    tile_gemm(n, n, [&](int i, int j) { return (i < n && j < n) ? Qxx[j * n + i] : 0.0; }, m,
                              [&](int i, int l) { return (i < n && l < m) ? T1[l * n + i] : 0.0; },
                              [&](int l, int j) { return (l < m && j < n) ? Kk[l * n + j] : 0.0; }, m,
                              [&](int i, int l) { return (i < n && l < m) ? Kk[l * n + i] : 0.0; },
                              [&](int l, int j) { return (l < m && j < n) ? Qux[l * n + j] : 0.0; },
                              [&](int i, int j, double v) { if (i < n && j < n) S[i * n + j] = v; });
                    gsync<T>();
#pragma unroll 2
                    for (int e = tid; e < n * n; e += T) S[e] = 0.5 * (SA[e] + S[e]);
                }
                for (int t = warp; !big && t < tn * tn1; t += NW) {
                    const int r0 = (t / tn1) << 3, c0 = (t % tn1) << 3;
                    const int i = r0 + fr, j = c0 + fc;
                    double d0 = (i < n) ? (j < n ? Qxx[i * n + j] : (j == n ? Qx[i] : 0.0)) : 0.0;
                    double d1 = (i < n) ? (j + 1 < n ? Qxx[i * n + j + 1] : (j + 1 == n ? Qx[i] : 0.0)) : 0.0;
                    double e0 = (i < n && j < n) ? Qxx[j * n + i] : 0.0;
                    double e1 = (i < n && j + 1 < n) ? Qxx[(j + 1) * n + i] : 0.0;
                    auto Kt = [&](int ii, int l) { ii += r0; return (ii < n && l < m) ? Kk[l * n + ii] : 0.0; };
                    auto Qt = [&](int ii, int l) { ii += r0; return (ii < n && l < m) ? Qux[l * n + ii] : 0.0; };
                    auto Tt = [&](int ii, int l) { ii += r0; return (ii < n && l < m) ? T1[l * n + ii] : 0.0; };
                    auto T1c = [&](int l, int jj) {
                        jj += c0;
                        const double *p = jj < n ? T1 + l * n + jj : t1 + l;
                        return (l < m && jj <= n) ? *p : 0.0;
                    };
                    auto Kc = [&](int l, int jj) {
                        jj += c0;
                        const double *p = jj < n ? Kk + l * n + jj : dk_ + l;
                        return (l < m && jj <= n) ? *p : 0.0;
                    };
                    auto Kn = [&](int l, int jj) { jj += c0; return (l < m && jj < n) ? Kk[l * n + jj] : 0.0; };
                    auto Qn = [&](int l, int jj) { jj += c0; return (l < m && jj < n) ? Qux[l * n + jj] : 0.0; };
                    mma_chain(d0, d1, m, Kt, T1c);  // + K' T1
                    mma_chain(d0, d1, m, Qt, Kc);   // + Qux' K
                    mma_chain(e0, e1, m, Tt, Kn);   // transposed: + T1' K  (= (K'T1)' element for element)
                    mma_chain(e0, e1, m, Kt, Qn);   //             + K' Qux
                    if (i < n) {
                        if (j < n) S[i * n + j] = 0.5 * (d0 + e0); else if (j == n) s[i] = d0;
                        if (j + 1 < n) S[i * n + j + 1] = 0.5 * (d1 + e1); else if (j + 1 == n) s[i] = d1;
                    }
                }
                if (tid == T - 1) {
                    for (int i = 0; i < m; ++i) {
                        a1 = fma(dk_[i], Qu[i], a1);
                        a2 = fma(0.5 * dk_[i], t1[i] - Qu[i], a2);
                    }
                }
                gsync<T>();
                ALTRO_TICK(6);
            }
            if (restart) {
                reg_increase(rho, drho);
                if (rho > P.o.bp_reg_max) return false;
                continue;
            }
            if (tid == T - 1) { bc[0] = a1; bc[1] = a2; }
            gsync<T>();
            dV1 = bc[0];
            dV2 = bc[1];
            gsync<T>();
            reg_decrease(rho, drho);
            return true;
        }
    }

    // ---------------------------------------------------------------- rollouts (A.6, A.8)
    template <bool DS>
    __device__ __forceinline__ void rollout_open_loop_impl()
    {
        for (int k = 0; k < N - 1; ++k) {
            const double *A = DS ? sA : as_global(P.A) + dyn_index(k) * n * n;
            const double *Bm = DS ? sB : as_global(P.Bm) + dyn_index(k) * n * m;
            const double *d = DS ? sd : as_global(P.d) + dyn_index(k) * n;
            for (int i = tid; i < n; i += T) {
                double acc = d[i];
                for (int j = 0; j < n; ++j) acc = fma(A[i * n + j], X[k * n + j], acc);
                for (int j = 0; j < m; ++j) acc = fma(Bm[i * m + j], U[k * m + j], acc);
                X[(k + 1) * n + i] = acc;
            }
            gsync<T>();
        }
    }
    __device__ void rollout_open_loop()
    {
        if (P.dyn_in_smem) rollout_open_loop_impl<true>();
        else rollout_open_loop_impl<false>();
    }

    // Closed-loop rollout with step alpha into Xb, Ub.  Returns 0 if a state leaves the box, 2 if the trial reproduces
    // (X, U) bit for bit -- every smaller step then does too (|alpha d| is below half an ulp of u at every knot and dx
    // stays exactly 0), so the line search can stop without rolling them out -- and 1 otherwise.
    template <bool DS>
    __device__ __forceinline__ int rollout_alpha_impl(double alpha)
    {
        for (int i = tid; i < n; i += T) Xb[i] = X[i];
        gsync<T>();
        double flag = 0.0;  // 2: out of the box, 1: differs from the current trajectory
        for (int k = 0; k < N - 1; ++k) {
            const double *A = DS ? sA : as_global(P.A) + dyn_index(k) * n * n;
            const double *Bm = DS ? sB : as_global(P.Bm) + dyn_index(k) * n * m;
            const double *d = DS ? sd : as_global(P.d) + dyn_index(k) * n;
            const double *Kk = K + k * m * n;
            for (int i = tid; i < m; i += T) {
                double acc = fma(alpha, dv[k * m + i], U[k * m + i]);
                for (int j = 0; j < n; ++j) acc = fma(Kk[i * n + j], Xb[k * n + j] - X[k * n + j], acc);
                Ub[k * m + i] = acc;
                if (!(acc == U[k * m + i])) flag = fmax(flag, 1.0);
            }
            gsync<T>();
            for (int i = tid; i < n; i += T) {
                double acc = d[i];
                for (int j = 0; j < n; ++j) acc = fma(A[i * n + j], Xb[k * n + j], acc);
                for (int j = 0; j < m; ++j) acc = fma(Bm[i * m + j], Ub[k * m + j], acc);
                Xb[(k + 1) * n + i] = acc;
                if (!(fabs(acc) <= P.o.max_state_value)) flag = 2.0;
                else if (!(acc == X[(k + 1) * n + i])) flag = fmax(flag, 1.0);
            }
            gsync<T>();
        }
        const double f = gmax<T>(flag, red);
        return f == 2.0 ? 0 : (f == 1.0 ? 1 : 2);
    }
    __device__ int rollout_alpha(double alpha)
    {
        return P.dyn_in_smem ? rollout_alpha_impl<true>(alpha) : rollout_alpha_impl<false>(alpha);
    }

    __device__ void copy_traj(double *Xd, double *Ud, const double *Xs, const double *Us)
    {
#pragma unroll 1
        for (int i = tid; i < N * n; i += T) Xd[i] = Xs[i];
#pragma unroll 1
        for (int i = tid; i < (N - 1) * m; i += T) Ud[i] = Us[i];
        gsync<T>();
    }

    __device__ double forward_pass(double dV1, double dV2, double J_prev, double &rho, double &drho, int &trials)
    {
        if constexpr (T > 32) {
            if (P.spec) return forward_pass_spec(dV1, dV2, J_prev, rho, drho, trials);
        }
        double J = INFINITY, alpha = 1.0, z = -1.0;
        int iter = 0;
        while ((z <= P.o.line_search_lower_bound || z > P.o.line_search_upper_bound) && J >= J_prev) {
            if (iter > P.o.iterations_linesearch) {
                copy_traj(Xb, Ub, X, U);
                J = al_cost(Xb, Ub);
                reg_increase(rho, drho);
                rho += P.o.bp_reg_fp;
                break;
            }
            const long long cr = clock64();
            const int ok = rollout_alpha(alpha);
            ph_roll += clock64() - cr;
            ++trials;
            if (!ok) { ++iter; alpha *= 0.5; continue; }
            const long long cc = clock64();
            J = al_cost(Xb, Ub);
            ph_cost += clock64() - cc;
            double expected = -alpha * (dV1 + alpha * dV2);
            z = expected > 0.0 ? (J_prev - J) / expected : -1.0;
            ++iter;
            alpha *= 0.5;
            // bit-identical trial: every remaining (smaller) step would reproduce it and fail the same test
            if (ok == 2) iter = P.o.iterations_linesearch + 1;
        }
        return J;
    }

    // Speculative line search: warp w rolls out and costs the step alpha 2^-w into its own candidate buffers while
    // the other warps do theirs, then every thread replays the sequential acceptance test over the results in order.
    // Same accepted step, same trajectory, same trial count as forward_pass (a trial's value does not depend on
    // the number of threads that compute it); the trials past the accepted one are wasted work traded for latency.
    __device__ double forward_pass_spec(double dV1, double dV2, double J_prev, double &rho, double &drho, int &trials)
    {
        constexpr int W = T / 32;
        const int warp = tid >> 5, lane = tid & 31;
        const int set = N * n + (N - 1) * m + N * (1 + ncon);
        double *Xb0 = reinterpret_cast<double *>(smem_base) + P.lay.Xb, *Ub0 = reinterpret_cast<double *>(smem_base) + P.lay.Ub;
        double *itm0 = itm;
        double *cs = reinterpret_cast<double *>(smem_base) + P.lay.cand + (warp > 0 ? (warp - 1) * set : 0);
        Ctx<NX, NU, 32> v(P, smem_base);
        v.tid = lane;
        // the view works on THIS context's instance (persistent CTAs are re-bound, see bind()), not on blockIdx's
        v.inst = inst; v.dyn_base = dyn_base; v.kcur = kcur; v.sched = sched; v.xr = xr; v.ur = ur; v.ex = ex;
        if (warp > 0) { v.Xb = cs; v.Ub = cs + N * n; v.itm = cs + N * n + (N - 1) * m; }
        else { v.Xb = Xb0; v.Ub = Ub0; v.itm = itm0; }
        double J = INFINITY, alpha = 1.0, z = -1.0;
        int iter = 0, chosen = -1;
        const double lo = P.o.line_search_lower_bound, hi = P.o.line_search_upper_bound;
        while ((z <= lo || z > hi) && J >= J_prev) {
            if (iter > P.o.iterations_linesearch) {
                Xb = Xb0; Ub = Ub0;
                copy_traj(Xb, Ub, X, U);
                J = al_cost(Xb, Ub);
                reg_increase(rho, drho);
                rho += P.o.bp_reg_fp;
                return J;
            }
            const long long cr = clock64();
            {
                double aw = alpha;
                for (int q = 0; q < warp; ++q) aw *= 0.5;
                double ok = 0.0, Jw = 0.0;
                if (iter + warp <= P.o.iterations_linesearch) {
                    ok = (double)v.rollout_alpha(aw);  // 0 out of the box, 1 ok, 2 ok and bit-identical to (X, U)
                    if (ok != 0.0) Jw = v.al_cost(v.Xb, v.Ub);
                }
                if (lane == 0) { specr[2 * warp] = ok; specr[2 * warp + 1] = Jw; }
            }
            __syncthreads();
            ph_roll += clock64() - cr;
            for (int w = 0; w < W; ++w) {
                if (!((z <= lo || z > hi) && J >= J_prev)) break;
                if (iter > P.o.iterations_linesearch) break;
                ++trials;
                if (specr[2 * w] != 0.0) {
                    J = specr[2 * w + 1];
                    const double expected = -alpha * (dV1 + alpha * dV2);
                    z = expected > 0.0 ? (J_prev - J) / expected : -1.0;
                    chosen = w;
                }
                ++iter;
                alpha *= 0.5;
                if (specr[2 * w] == 2.0) iter = P.o.iterations_linesearch + 1;  // see forward_pass
            }
            __syncthreads();
        }
        // the accepted candidate becomes (Xb, Ub); the caller copies it into (X, U)
        if (chosen > 0) {
            double *cc = reinterpret_cast<double *>(smem_base) + P.lay.cand + (chosen - 1) * set;
            Xb = cc; Ub = cc + N * n;
        } else { Xb = Xb0; Ub = Ub0; }
        return J;
    }

    __device__ double gradient_todorov() const
    {
#pragma unroll 1
        for (int k = tid; k < N - 1; k += T) {
            double mx = 0.0;
#pragma unroll 1
            for (int i = 0; i < m; ++i) mx = fmax(mx, fabs(dv[k * m + i]) / (fabs(U[k * m + i]) + 1.0));
            itm[k] = mx;
        }
        gsync<T>();
        return csum<T>(itm, N - 1, bc, tid) / (double)(N - 1);
    }

    // ---------------------------------------------------------------- solve! (A.4 - A.6)
    // Warm-started MPC transition in shared memory (random_linear_problem.jl:121-139, simple_rocket.jl:59-82):
    // x0 <- x_1 of the last solution (= plant step with its first control) + noise, reference window advanced
    // along the track, RD.shift_fill!(Z) on the controls and Altro.shift_fill!(conSet) on the duals.
    __device__ void transition(int st)
    {
        const double *z = P.noise ? P.noise + ((size_t)st * P.B + inst) * n : nullptr;
        if (z && tid == 0) {
            const double *xo = X + n;
            double s0 = P.noise_w1, s1 = P.noise_w1;
            if (P.noise_mode == 1) {
                double mx = 0.0;
#pragma unroll 1
                for (int i = 0; i < n; ++i) mx = fmax(mx, fabs(xo[i]));
                s0 = s1 = mx * P.noise_w1;
            } else if (P.noise_mode == 2) {
                double a = 0.0, b = 0.0;
#pragma unroll 1
                for (int i = 0; i < n / 2; ++i) a += xo[i] * xo[i];
#pragma unroll 1
                for (int i = n / 2; i < n; ++i) b += xo[i] * xo[i];
                s0 = sqrt(a) * P.noise_w1;
                s1 = sqrt(b) * P.noise_w2;
            }
            bc[3] = s0;
            bc[4] = s1;
        }
        gsync<T>();
#pragma unroll 1
        for (int i = tid; i < n; i += T) {
            double v = X[n + i];
            if (z) v += z[i] * bc[(P.noise_mode == 2 && i >= n / 2) ? 4 : 3];
            X[i] = v;
        }
        if (P.trackX) {
            const int k0 = P.kidx[inst] + st + 1;
            if (ALL_SMEM || P.ref_in_smem) {  // the device track is padded with N copies of its last knot: one contiguous slice
                const double *wx = P.trackX + (size_t)min(k0, P.Nt) * n, *wu = P.trackU + (size_t)min(k0, P.Nt - 1) * m;
#pragma unroll 1
                for (int i = tid; i < N * n; i += T) xr[i] = wx[i];
#pragma unroll 1
                for (int i = tid; i < (N - 1) * m; i += T) ur[i] = wu[i];
            } else {
                xr = const_cast<double *>(P.trackX) + (size_t)min(k0, P.Nt) * n;
                ur = const_cast<double *>(P.trackU) + (size_t)min(k0, P.Nt - 1) * m;
            }
        }
        if (P.shift) {
            // in-place left shifts in chunks of T: all reads of a chunk precede its writes, later chunks are untouched
#pragma unroll 1
            for (int base = 0; base < (N - 2) * m; base += T) {
                const int i = base + tid;
                double v = (i < (N - 2) * m) ? U[i + m] : 0.0;
                gsync<T>();
                if (i < (N - 2) * m) U[i] = v;
                gsync<T>();
            }
#pragma unroll 1
            for (int ci = 0; ci < ncon; ++ci) {
                const int cnt = (cd[ci].k1 - cd[ci].k0 - 1) * cd[ci].p, p = cd[ci].p;
                double *l = lam + cd[ci].dual_off;
#pragma unroll 1
                for (int base = 0; base < cnt; base += T) {
                    const int i = base + tid;
                    double v = (i < cnt) ? l[i + p] : 0.0;
                    gsync<T>();
                    if (i < cnt) l[i] = v;
                    gsync<T>();
                }
            }
        }
        gsync<T>();
    }

    // ---------------------------------------------------------------- solve! (A.4 - A.6)
    __device__ __forceinline__ void solve_core(int slot)
    {
        const altro_opts_t &o = P.o;
        long long ph[4] = {0, 0, 0, 0};
        const long long cs = clock64();
        int iters = 0, outer_done = 0, status = ALTRO_UNSOLVED, trials = 0;
        double cmax = INFINITY, J = 0.0, pen_max = 0.0;
#pragma unroll 1
        for (int outer = 1; outer <= o.iterations_outer; ++outer) {
            outer_done = outer;
            const bool last = (outer == o.iterations_outer) || ncon == 0;
            const double ctol = last ? o.cost_tolerance : o.cost_tolerance_intermediate;
            const double gtol = last ? o.gradient_tolerance : o.gradient_tolerance_intermediate;
            double rho = o.bp_reg_initial, drho = 0.0;
            int dJ_zero = 0;
            long long c0 = clock64();
            rollout_open_loop();
            double J_prev = al_cost(X, U);
            J = J_prev;
            // the initial rollout's cost is not a line-search reference (see altro_opts_t)
            if (o.first_step_unconditional) J_prev = INFINITY;
            ph[0] += clock64() - c0;
#pragma unroll 1
            for (int it = 0; it < o.iterations_inner; ++it) {
                double dV1, dV2;
                c0 = clock64();
                if (!backward_pass(rho, drho, dV1, dV2)) { status = ALTRO_NOT_PD; break; }
                long long c1 = clock64();
                ph[1] += c1 - c0;
                J = forward_pass(dV1, dV2, J_prev, rho, drho, trials);
                ph[2] += clock64() - c1;
                if (J > o.max_cost_value || !(J == J)) { status = ALTRO_MAXIMUM_COST; break; }
                copy_traj(X, U, Xb, Ub);
                double dJ = fabs(J - J_prev);
                J_prev = J;
                double grad = gradient_todorov();
                ++iters;
                dJ_zero = (dJ == 0.0) ? dJ_zero + 1 : 0;
                if (P.trace && tid == 0 && iters <= P.trace_rows) {
                    double *tr = P.trace + ((size_t)inst * P.trace_rows + (iters - 1)) * TRACE_COLS;
                    tr[0] = outer; tr[1] = iters; tr[2] = J; tr[3] = dJ; tr[4] = grad; tr[5] = rho;
                    tr[6] = dV1; tr[7] = dV2; tr[8] = trials; tr[9] = nan("");
                }
                bool small = o.dj_zero_converges ? (dJ >= 0.0 && dJ < ctol) : (dJ > 0.0 && dJ < ctol);
                if (small && grad < gtol) { status = ALTRO_SOLVE_SUCCEEDED; break; }
                if (iters >= o.iterations) { status = ALTRO_MAX_ITERATIONS; break; }
                if (dJ_zero > o.dJ_counter_limit) { status = ALTRO_NO_PROGRESS; break; }
            }
            if (status > ALTRO_SOLVE_SUCCEEDED) break;
            cmax = max_violation();
            if (P.trace && tid == 0 && iters >= 1 && iters <= P.trace_rows)
                P.trace[((size_t)inst * P.trace_rows + (iters - 1)) * TRACE_COLS + 9] = cmax;
            pen_max = 0.0;
#pragma unroll 1
            for (int c = 0; c < ncon; ++c) pen_max = fmax(pen_max, mu[c]);
            if (cmax < o.constraint_tolerance) break;
            if (o.kickout_max_penalty && pen_max >= o.penalty_max) break;
            dual_update();
            if (tid < ncon) mu[tid] = fmin(mu[tid] * o.penalty_scaling, o.penalty_max);
            gsync<T>();
            if (outer == o.iterations_outer) status = ALTRO_MAX_ITERATIONS_OUTER;
        }
        cmax = max_violation();
        if (status <= ALTRO_SOLVE_SUCCEEDED)
            status = (cmax < o.constraint_tolerance) ? ALTRO_SOLVE_SUCCEEDED : ALTRO_UNSOLVED;
        double Jobj = objective_cost();
        if (tid == 0) {
            const size_t at = (size_t)slot * P.B + inst;
            P.iters[at] = iters;
            P.outer[at] = outer_done;
            P.status[at] = status;
            P.trials[at] = trials;
            P.cost[at] = Jobj;
            P.cost_al[at] = J;
            P.cmax[at] = cmax;
            P.penmax[at] = pen_max;
            if (P.phase) {  // cycles: initial rollout+cost, backward pass (incl. expansion), forward pass, whole solve
                long long *pp = P.phase + (size_t)inst * 8;
                if (!P.phase_detail) { pp[0] += ph[0]; pp[1] += ph[1]; pp[2] += ph[2]; pp[3] += clock64() - cs; }
                if (P.phase_detail) { for (int q = 0; q < 7; ++q) { pp[q] += bpc[q]; bpc[q] = 0; } pp[7] += iters; }
                else { pp[4] += ph_exp; pp[5] += ph_roll; pp[6] += ph_cost; }
                ph_exp = ph_roll = ph_cost = 0;
            }
        }
    }

    // steps [s0, s1) of the closed-loop run of the bound instance (or the one plain solve! when P.steps == 0)
    __device__ __forceinline__ void run_steps(int s0, int s1, long long t0)
    {
#pragma unroll 1
        for (int st = s0; st < s1; ++st) {
            if (P.steps > 0) {
                set_step(st + 1);  // the window moves one knot forward with every transition
                kcur = (P.kidx ? P.kidx[inst] : 0) + st + 1;
                transition(st);
                if (st > s0) {  // every solve! starts from reset penalties (and duals, if asked); load_state did it for s0
                    if (P.o.reset_duals)
#pragma unroll 1
                        for (int i = tid; i < P.P; i += T) lam[i] = 0.0;
#pragma unroll 1
                    for (int i = tid; i < MAX_CON; i += T) mu[i] = P.o.penalty_initial;
                    gsync<T>();
                }
                if (P.x0_log)
#pragma unroll 1
                    for (int i = tid; i < n; i += T) P.x0_log[((size_t)st * P.B + inst) * n + i] = X[i];
            }
            solve_core(st);
            if (P.steps > 0 && P.u0_log)
#pragma unroll 1
                for (int i = tid; i < m; i += T) P.u0_log[((size_t)st * P.B + inst) * m + i] = U[i];
            if (tid == 0 && P.t_ns) {
                long long t1v;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1v));
                P.t_ns[(size_t)st * P.B + inst] = t1v - t0;
                t0 = t1v;
            }
        }
    }

    // One plain solve! / one closed-loop run of the CTA's own instance (P.q_head == nullptr), or the closed-loop run on a
    // persistent grid (one CTA per resident slot): the CTAs pull (instance, chunk of q_chunk steps) items from an atomic
    // counter, chunk-major, so that every slot stays busy until the whole batch is done -- no wave quantisation (4096
    // instances over 1184 slots = 3.46 waves) and no waiting for the slot that happened to get the longest chains.  An
    // item needs the instance's previous chunk: it was handed out a whole round (B items) earlier, so it is normally
    // long finished; otherwise thread 0 waits on q_done[inst].  State travels through global memory (a few KB per item,
    // L2-resident), written with a release fence and read through L2.
    // Both modes share ONE call site of run_steps(): with two, the compiler keeps solve_core() out of line and the whole
    // Ctx moves to local memory (measured: 3.3x slower at T = 128).
    __device__ __forceinline__ void solve()
    {
        const bool queued = T > 32 && P.q_head != nullptr;
        const int chunk = queued ? P.q_chunk : 0;
        const long long total = queued ? (long long)((P.steps + chunk - 1) / chunk) * P.B : 0;
        if (queued) {
            load_static();
            gsync<T>();
        }
#pragma unroll 1
        for (;;) {
            int s0 = 0, s1 = P.steps > 0 ? P.steps : 1, item = inst;
            if (queued) {
                if (tid == 0) {
                    const long long t = (long long)atomicAdd(P.q_head, 1);
                    int go = t < total ? 1 : 0;
                    if (go) {
                        const int i = (int)(t % P.B), c = (int)(t / P.B);
                        if (c > 0) {  // wait for the instance's previous chunk (bounded: a lost item must not hang the GPU)
                            const long long w0 = clock64();
                            while (atomicAdd(P.q_done + i, 0) < c * chunk) {
                                __nanosleep(256);
                                if (clock64() - w0 > (1ll << 33)) { atomicExch(P.q_error, 1); go = 0; break; }
                            }
                        }
                        reinterpret_cast<int *>(bc)[0] = i;
                        reinterpret_cast<int *>(bc)[1] = c;
                    }
                    reinterpret_cast<int *>(bc)[2] = go;
                }
                __syncthreads();
                const int go = reinterpret_cast<int *>(bc)[2], c = reinterpret_cast<int *>(bc)[1];
                item = reinterpret_cast<int *>(bc)[0];
                __syncthreads();
                if (!go) break;
                __threadfence();  // acquire: the previous chunk's rows are visible before they are read
                s0 = c * chunk;
                s1 = min(P.steps, s0 + chunk);
            }
            long long t0 = 0;
            if (tid == 0 && P.t_ns) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            if (queued) {
                bind(item);
                load_state<true>();
                gsync<T>();
            } else {
                load();
            }
            run_steps(s0, s1, t0);
            store();
            if (!queued) break;
            __threadfence();  // release: rows first, then the step count
            __syncthreads();
            if (tid == 0) atomicExch(P.q_done + item, s1);
        }
    }
};

template <int NX, int NU, int T>
// Registers: tiny problems (rocket, grasp) are capped at 128 so that 16 warps fit an SM; 12-dimensional and run-time
// sized problems are shared-memory-limited to a few CTAs per SM anyway and get the full register file (no spills).
// (measured, scripts/dev_perf.py): quadruped (12,12) is fastest with the whole register file at 4 CTAs/SM; the other
// 12-dimensional and run-time sized problems with a 168-register cap (6 CTAs/SM at T = 64).
__global__ void __launch_bounds__(T, ((NX == 12 && NU == 12) ? (256 / T > 0 ? 256 / T : 1)
                                      : (NX == 0 && T == 256) ? ALTRO_GEN256_CTAS : (NX >= 12 || NX == 0) ? (384 / T > 0 ? 384 / T : 1) : (T == 128 ? ALTRO_T128_CTAS : 512 / T))) altro_solve_kernel(const __grid_constant__ Params P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Ctx<NX, NU, T> ctx(P, smem_raw);
    ctx.solve();
}

// The same run-time sized solve with the whole register file (one CTA per SM): the TMA-staged large-dimension layout is
// shared-memory-limited to one CTA per SM anyway, and its 32 x 32 register blocks want the registers.
template <int T>
__global__ void __launch_bounds__(T, 1) altro_solve_kernel_wide(const __grid_constant__ Params P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Ctx<0, 0, T, true> ctx(P, smem_raw);
    ctx.solve();
}

#endif  // __CUDACC__

}  // namespace altro
