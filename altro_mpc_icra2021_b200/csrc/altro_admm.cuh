// altro_admm.cuh -- parameters of the batched ADMM cross-check (admm.cu).
#pragma once
#include "altro_kernels.cuh"

namespace altro {

constexpr int PMAX_ADMM = 8;

struct AdmmParams {
    int n, m, N, B, Pd, ncon;
    double dt, rho, eps;
    int max_iter, adapt;  // adapt > 0: rebalance rho every `adapt` iterations
    int dyn_per_knot, dyn_per_instance, dyn_slots, sched_len, step0;
    const int *dyn_sched, *kidx;
    const double *A, *Bm, *d, *Q, *R, *Qf, *xref, *uref, *x0, *X, *U;
    const ConDesc *con;
    double *ws;
    size_t ws_doubles;
    double *Xout, *Uout, *rprim, *rdual;
    int *iters;
};

size_t admm_workspace_doubles(int n, int m, int N, int Pd);
cudaError_t admm_launch(const AdmmParams &P, cudaStream_t stream);

}  // namespace altro
