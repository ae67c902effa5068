// Does a warp with only 16 active lanes issue FP64 instructions faster than a full warp on sm_100a (B200)?
// The FP64 pipe of a sub-partition is 16 lanes wide: a full warp's DFMA takes two passes.  If a half-empty warp took
// one, two half-warps per sub-partition would double the instruction rate of the one-warp-per-sub-partition kernels
// (cfg 2, cfg 4).  Build: nvcc -arch=sm_100a -O3 -o fp64_halfwarp fp64_halfwarp.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double *out, long long *cyc, int iters, double m, double b, int active_lanes)
{
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = threadIdx.x + i;
    const bool on = (threadIdx.x & 31) < active_lanes;
    long long t0 = clock64();
    if (on) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 16; u++) {
#pragma unroll
                for (int i = 0; i < ILP; i++) a[i] = fma(a[i], m, b);
            }
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
void run(int warps, int lanes)
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, sizeof(double) * 1024); cudaMalloc(&cyc, 8);
    const int iters = 4000;
    chain<ILP><<<1, 32 * warps>>>(out, cyc, iters, 0.999999, 1e-9, lanes);
    chain<ILP><<<1, 32 * warps>>>(out, cyc, iters, 0.999999, 1e-9, lanes);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // warps per SM are spread over the 4 sub-partitions: warps/4 per sub-partition
    printf("warps/SM %2d (%.1f per sub-partition)  active lanes %2d  ILP %d : %.2f cycles per DFMA per warp -> %.3f warp-DFMA per cycle per sub-partition\n",
           warps, warps / 4.0, lanes, ILP, (double)h / (iters * 16.0 * ILP), (warps / 4.0) * (iters * 16.0 * ILP) / (double)h);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int lanes : {32, 16, 8}) {
        for (int w : {4, 8, 16}) { run<1>(w, lanes); run<4>(w, lanes); run<8>(w, lanes); }
    }
    return 0;
}
