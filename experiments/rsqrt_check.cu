// rsqrt_check.cu -- accuracy of usvmpc::drsqrt_pos (hardware seed + two Newton steps) against 1 / sqrt in higher precision
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../mpc_collisionavoidance_b200/csrc -o rsqrt_check rsqrt_check.cu
#include <cmath>
#include <cstdio>
#include "cta_compat.h"

__global__ void k(const double* x, double* y, double* z, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { y[i] = usvmpc::drsqrt_pos(x[i]); z[i] = rsqrt(x[i]); }
}

int main()
{
    const int n = 1 << 20;
    double *x, *y, *z;
    cudaMallocManaged(&x, n * 8); cudaMallocManaged(&y, n * 8); cudaMallocManaged(&z, n * 8);
    unsigned long long s = 88172645463325252ull;
    for (int i = 0; i < n; i++)
    {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const double m = 1.0 + (double) (s >> 11) / 9007199254740992.0;      // [1, 2)
        const int e = (int) ((s >> 3) % 601) - 300;                          // 2^-300 .. 2^300
        x[i] = ldexp(m, e);
    }
    k<<<n / 256, 256>>>(x, y, z, n);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("failed\n"); return 1; }
    double worst = 0, worst_lib = 0;
    for (int i = 0; i < n; i++)
    {
        const long double ref = 1.0L / sqrtl((long double) x[i]);
        const double ulp = ldexp(1.0, ilogb((double) ref) - 52);
        const double e1 = fabs((double) ((long double) y[i] - ref)) / ulp, e2 = fabs((double) ((long double) z[i] - ref)) / ulp;
        if (e1 > worst) worst = e1;
        if (e2 > worst_lib) worst_lib = e2;
    }
    printf("drsqrt_pos: max error %.3f ulp; rsqrt(double): %.3f ulp  (%d values, 2^-300 .. 2^300)\n", worst, worst_lib, n);
    return worst < 2.5 ? 0 : 1;
}
