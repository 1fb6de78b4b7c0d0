#include <cstdio>
#include <cuda_runtime.h>
#define W 8
__device__ inline unsigned gshift() { return (threadIdx.x & 31u) & ~(unsigned)(W - 1); }
__device__ inline unsigned gmask() { return ((1u << W) - 1u) << gshift(); }
__device__ inline unsigned gballot(bool p) { return (__ballot_sync(gmask(), p) >> gshift()) & ((1u << W) - 1u); }
__global__ void k(int *out, int mode)
{
    const int lane = threadIdx.x & (W - 1), grp = threadIdx.x / W;
    int acc = 0;
    const int iters = 10 + grp * 7;  // groups of one warp run different trip counts
    for (int it = 0; it < iters; ++it) {
        // per-lane divergent work
        int x = 0;
        for (int k = 0; k < (lane * 3 + it) % 11; ++k) x += k * lane;
        if (mode & 1) acc += __popc(gballot((x & 1) != 0));
        if (mode & 2) acc += __shfl_sync(gmask(), x, it % W, W);
        if (mode & 4) __syncwarp(gmask());
        if (mode & 8) acc += __shfl_xor_sync(gmask(), x, 4, W);
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
}
int main()
{
    int *d;
    cudaMalloc(&d, 4 * 384 * 4);
    for (int mode = 1; mode <= 15; mode = mode * 2 + (mode == 8 ? -1 : 0)) {
        k<<<4, 384>>>(d, mode);
        cudaError_t e = cudaDeviceSynchronize();
        printf("mode %d: %s\n", mode, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        if (mode == 15) break;
    }
    return 0;
}
