// FP64 dependent-issue latency and per-warp throughput on sm_100a (B200): cycles per DFMA for
// ILP = 1, 2, 4, 8 independent chains in ONE warp on one SM.  Build: nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double *out, long long *cyc, int iters, double m, double b)
{
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) a[i] = fma(a[i], m, b);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += a[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
void run(int warps)
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, sizeof(double) * 1024); cudaMalloc(&cyc, 8);
    const int iters = 4000;
    chain<ILP><<<1, 32 * warps>>>(out, cyc, iters, 0.999999, 1e-9);
    chain<ILP><<<1, 32 * warps>>>(out, cyc, iters, 0.999999, 1e-9);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("warps/SM %2d  ILP %d : %.2f cycles per DFMA per warp (%.2f cycles per dependent step)\n", warps, ILP,
           (double)h / (iters * 16.0 * ILP), (double)h / (iters * 16.0));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int w : {1, 4, 8, 16}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
    return 0;
}
