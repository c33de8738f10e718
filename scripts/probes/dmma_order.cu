// Probe: is mma.sync.m8n8k4.f64 bit-identical to a sequential fma chain over k = 0..3 starting from C?
// (decides whether FP64 tensor-core tiles can be used under the repo's bit-parity contract)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

__global__ void k(const double *A, const double *B, const double *C, double *D, int ntiles)
{
    int l = threadIdx.x;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const double *a = A + t * 32, *b = B + t * 32, *c = C + t * 64;
        double av = a[(l >> 2) * 4 + (l & 3)];        // A[row=l/4][col=l%4], 8x4 row-major
        double bv = b[(l & 3) * 8 + (l >> 2)];        // B[row=l%4][col=l/4], 4x8 row-major storage
        double c0 = c[(l >> 2) * 8 + 2 * (l & 3)], c1 = c[(l >> 2) * 8 + 2 * (l & 3) + 1];
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0), "+d"(c1) : "d"(av), "d"(bv));
        D[t * 64 + (l >> 2) * 8 + 2 * (l & 3)] = c0;
        D[t * 64 + (l >> 2) * 8 + 2 * (l & 3) + 1] = c1;
    }
}

int main()
{
    const int nt = 20000;
    double *A, *B, *C, *D;
    cudaMallocManaged(&A, nt * 32 * 8); cudaMallocManaged(&B, nt * 32 * 8);
    cudaMallocManaged(&C, nt * 64 * 8); cudaMallocManaged(&D, nt * 64 * 8);
    srand(1);
    auto rnd = [](int t) { double v = (rand() / (double)RAND_MAX - 0.5) * 2; int e = (t % 7) * 6 - 18; return ldexp(v, rand() % 2 ? e : -e / 2); };
    for (int i = 0; i < nt * 32; ++i) { A[i] = rnd(i / 32); B[i] = rnd(i / 32 + 3); }
    for (int i = 0; i < nt * 64; ++i) C[i] = (i % 5 == 0) ? 0.0 : rnd(i / 64 + 1);
    k<<<256, 32>>>(A, B, C, D, nt);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("cuda error\n"); return 1; }
    long long n_seq = 0, n_rev = 0, n_unfused = 0, n_pair = 0, total = 0;
    for (int t = 0; t < nt; ++t)
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 8; ++j) {
                const double *a = A + t * 32 + i * 4, *b = B + t * 32;
                double c = C[t * 64 + i * 8 + j], d = D[t * 64 + i * 8 + j];
                double s = c; for (int kk = 0; kk < 4; ++kk) s = fma(a[kk], b[kk * 8 + j], s);
                double r = c; for (int kk = 3; kk >= 0; --kk) r = fma(a[kk], b[kk * 8 + j], r);
                double u = c; for (int kk = 0; kk < 4; ++kk) u = u + a[kk] * b[kk * 8 + j];
                double p = fma(a[0], b[j], fma(a[1], b[8 + j], 0.0)) + fma(a[2], b[16 + j], fma(a[3], b[24 + j], 0.0)) + c;
                ++total; n_seq += (s == d); n_rev += (r == d); n_unfused += (u == d); n_pair += (p == d);
            }
    printf("elements %lld: == sequential fma chain k=0..3 from C: %lld | reversed: %lld | unfused: %lld | pairwise: %lld\n",
           total, n_seq, n_rev, n_unfused, n_pair);
    return 0;
}
