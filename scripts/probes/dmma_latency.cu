// Probe: dependent-issue latency of mma.sync.m8n8k4.f64 (DMMA) vs fma.rn.f64 (DFMA) vs shared-memory load on B200,
// one warp alone on an SM.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_latency dmma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void lat(double *out, long long *cyc, int iters)
{
    __shared__ double sm[1024];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double a = 1.0 + 1e-9 * lane, b = 1.0 - 1e-9 * lane, c0 = 0.0, c1 = 0.0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    long long t1 = clock64();
    double f = 0.5;
    for (int i = 0; i < iters; ++i) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f) : "d"(a), "d"(b));
    long long t2 = clock64();
    // two independent DMMA chains interleaved (does a second chain hide the latency?)
    double d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0;
    for (int i = 0; i < iters; ++i) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(e0), "+d"(e1) : "d"(b), "d"(a));
    }
    long long t3 = clock64();
    // pointer-chasing shared-memory load
    int idx = lane;
    for (int i = 0; i < iters; ++i) idx = (int)sm[idx & 1023] + lane;
    long long t4 = clock64();
    double r = 3.0 + lane;
    for (int i = 0; i < iters; ++i) r = __drcp_rn(r) + 2.0;
    long long t5 = clock64();
    double sh = a;
    for (int i = 0; i < iters; ++i) sh = __shfl_sync(0xffffffffu, sh, (lane + 1) & 31) + 1.0;
    long long t6 = clock64();
    if (threadIdx.x == 0) {
        cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5;
    }
    out[threadIdx.x] = c0 + c1 + f + d0 + d1 + e0 + e1 + idx + r + sh;
}

int main()
{
    double *out; long long *cyc, h[6];
    cudaMalloc(&out, 1024 * sizeof(double)); cudaMalloc(&cyc, 6 * sizeof(long long));
    const int iters = 4096;
    for (int warps = 1; warps <= 4; warps *= 2) {
        lat<<<1, 32 * warps>>>(out, cyc, iters);
        lat<<<1, 32 * warps>>>(out, cyc, iters);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        printf("warps %d: DMMA dependent %.1f cyc, DFMA dependent %.1f cyc, 2 interleaved DMMA chains %.1f cyc/pair, LDS chase(+cvt+add) %.1f, drcp+add %.1f, shfl+add %.1f\n",
               warps, (double)h[0] / iters, (double)h[1] / iters, (double)h[2] / iters, (double)h[3] / iters, (double)h[4] / iters, (double)h[5] / iters);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
