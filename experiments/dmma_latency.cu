// dmma_latency.cu -- what one warp pays for fp64 work on sm_100a: dependent DFMA chain, dependent mma.m8n8k4.f64 chain,
// an LDS -> 2x DMMA -> STS -> __syncwarp round (the shape of one Riccati sub-step) and rsqrt.  One warp, one block.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_latency dmma_latency.cu && ./dmma_latency
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void k(double* out, long long* cyc, int iters)
{
    __shared__ double sA[64], sB[64], sC[64];
    const int lane = threadIdx.x;
    sA[lane] = 1.0 + 1e-3 * lane; sA[lane + 32] = 0.5; sB[lane] = 1e-3 * lane; sB[lane + 32] = 0.25; sC[lane] = 0; sC[lane + 32] = 0;
    __syncwarp();
    double a = sA[lane], b = sB[lane], c0 = 0, c1 = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) c0 = fma(a, c0, b);
    long long t1 = clock64();
    for (int i = 0; i < iters; i++) dmma(c0, c1, a, b);
    long long t2 = clock64();
    for (int i = 0; i < iters; i++)
    {
        const double x = sA[(lane + i) & 63], y = sB[(lane * 3 + i) & 63];
        double d0 = sC[lane], d1 = sC[lane + 32];
        dmma(d0, d1, x, y);
        dmma(d0, d1, y, x);
        sC[lane] = d0; sC[lane + 32] = d1;
        __syncwarp();
    }
    long long t3 = clock64();
    double r = 2.0 + lane;
    for (int i = 0; i < iters; i++) r = rsqrt(r + 1.5);
    long long t4 = clock64();
    double q = 2.0 + lane;
    for (int i = 0; i < iters; i++) q = 1.0 / (q + 1.5);
    long long t5 = clock64();
    // 6-term dot product from shared memory, result stored, syncwarp (the current sub-step)
    for (int i = 0; i < iters; i++)
    {
        double s = sC[lane];
#pragma unroll
        for (int j = 0; j < 6; j++) s = fma(sA[(lane + j) & 63], sB[(i + j * 5) & 63], s);
        sC[lane] = s;
        __syncwarp();
    }
    long long t6 = clock64();
    out[lane] = c0 + c1 + r + q + sC[lane];
    if (lane == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5; }
}

int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMallocManaged(&cyc, 6 * 8);
    const int iters = 4096;
    k<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize();
    k<<<1, 32>>>(out, cyc, iters);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("failed\n"); return 1; }
    const char* name[6] = {"dependent DFMA", "dependent DMMA m8n8k4", "LDS + 2 DMMA + STS + syncwarp", "dependent rsqrt(fp64)", "dependent 1/x (fp64)",
                           "LDS 6-term dot + STS + syncwarp"};
    for (int i = 0; i < 6; i++) printf("%-34s %8.1f cycles\n", name[i], (double) cyc[i] / iters);
    return 0;
}
