// micro-benchmark (scratch): latencies that decide the link-update kernel's memory protocol on B200
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double ld_ca(const double *p){ double v; asm volatile("ld.global.ca.f64 %0,[%1];":"=d"(v):"l"(p):"memory"); return v; }
__device__ __forceinline__ double ld_cg(const double *p){ double v; asm volatile("ld.global.cg.f64 %0,[%1];":"=d"(v):"l"(p):"memory"); return v; }
__device__ __forceinline__ double ld_rlx_cta(const double *p){ double v; asm volatile("ld.relaxed.cta.global.f64 %0,[%1];":"=d"(v):"l"(p):"memory"); return v; }
__device__ __forceinline__ double ld_rlx_gpu(const double *p){ double v; asm volatile("ld.relaxed.gpu.global.f64 %0,[%1];":"=d"(v):"l"(p):"memory"); return v; }
__device__ __forceinline__ void st_wt(double *p, double v){ asm volatile("st.global.f64 [%0],%1;"::"l"(p),"d"(v):"memory"); }
// out[k]: cycles per iteration of test k
__global__ void k_lat(double *buf, long long *out, int iters)
{
    __shared__ double sm[64];
    if (threadIdx.x != 0) return;
    double acc = 0; long long t0, t1;
    // 0: dependent ld.ca chain on one location (L1 hit)
    buf[0] = 0.0; acc += ld_ca(buf);
    t0 = clock64(); for (int i = 0; i < iters; ++i) { acc += ld_ca(buf + (int)acc); } t1 = clock64(); out[0] = (t1 - t0) / iters;
    // 1: dependent ld.cg chain (L2 hit)
    t0 = clock64(); for (int i = 0; i < iters; ++i) { acc += ld_cg(buf + (int)acc); } t1 = clock64(); out[1] = (t1 - t0) / iters;
    // 2: store then ld.ca of the same location (does a store keep the L1 line valid?)
    t0 = clock64(); for (int i = 0; i < iters; ++i) { st_wt(buf + 8 + (int)acc, acc); acc += ld_ca(buf + 8 + (int)acc); } t1 = clock64(); out[2] = (t1 - t0) / iters;
    // 3: store to location A, then ld.ca of another, previously cached location in the same 128-byte line (other sector)
    acc += ld_ca(buf + 16 + 4);
    t0 = clock64(); for (int i = 0; i < iters; ++i) { st_wt(buf + 16 + (int)acc, acc); acc += ld_ca(buf + 16 + 4 + (int)acc); } t1 = clock64(); out[3] = (t1 - t0) / iters;
    // 4: __threadfence() alone (MEMBAR.SC.GPU + CCTL.IVALL) after one store
    t0 = clock64(); for (int i = 0; i < iters; ++i) { st_wt(buf + 32, acc); __threadfence(); } t1 = clock64(); out[4] = (t1 - t0) / iters;
    // 5: fence.release.gpu (MEMBAR.ALL.GPU) after one store
    t0 = clock64(); for (int i = 0; i < iters; ++i) { st_wt(buf + 32, acc); asm volatile("fence.release.gpu;":::"memory"); } t1 = clock64(); out[5] = (t1 - t0) / iters;
    // 6: __threadfence_block() after one store
    t0 = clock64(); for (int i = 0; i < iters; ++i) { st_wt(buf + 32, acc); __threadfence_block(); } t1 = clock64(); out[6] = (t1 - t0) / iters;
    // 7: ld.ca of a cached location right after __threadfence() (CCTL.IVALL => miss)
    t0 = clock64(); for (int i = 0; i < iters; ++i) { __threadfence(); acc += ld_ca(buf + (int)acc); } t1 = clock64(); out[7] = (t1 - t0) / iters;
    // 8: ld.ca right after fence.release.gpu (no invalidate => hit)
    t0 = clock64(); for (int i = 0; i < iters; ++i) { asm volatile("fence.release.gpu;":::"memory"); acc += ld_ca(buf + (int)acc); } t1 = clock64(); out[8] = (t1 - t0) / iters;
    // 9: ld.relaxed.cta chain (does it hit L1?)
    t0 = clock64(); for (int i = 0; i < iters; ++i) { acc += ld_rlx_cta(buf + (int)acc); } t1 = clock64(); out[9] = (t1 - t0) / iters;
    // 10: ld.relaxed.gpu chain
    t0 = clock64(); for (int i = 0; i < iters; ++i) { acc += ld_rlx_gpu(buf + (int)acc); } t1 = clock64(); out[10] = (t1 - t0) / iters;
    // 11: shared memory through a generic pointer
    sm[0] = 0.0; volatile double *g = sm;
    t0 = clock64(); for (int i = 0; i < iters; ++i) { acc += g[(int)acc]; } t1 = clock64(); out[11] = (t1 - t0) / iters;
    // 12: fence.release.sys after one store
    t0 = clock64(); for (int i = 0; i < iters; ++i) { st_wt(buf + 32, acc); asm volatile("fence.release.sys;":::"memory"); } t1 = clock64(); out[12] = (t1 - t0) / iters;
    // 13: __threadfence_system after one store
    t0 = clock64(); for (int i = 0; i < iters; ++i) { st_wt(buf + 32, acc); __threadfence_system(); } t1 = clock64(); out[13] = (t1 - t0) / iters;
    // 14: fp64 division chain, 15: log chain, 16: sin chain
    double x = 1.0000001 + acc;
    t0 = clock64(); for (int i = 0; i < iters; ++i) { x = 3.0 / (x + 1.0); } t1 = clock64(); out[14] = (t1 - t0) / iters;
    t0 = clock64(); for (int i = 0; i < iters; ++i) { x = log(x + 2.0); } t1 = clock64(); out[15] = (t1 - t0) / iters;
    t0 = clock64(); for (int i = 0; i < iters; ++i) { x = sin(x + 2.0); } t1 = clock64(); out[16] = (t1 - t0) / iters;
    buf[40] = acc + x;
}
// visibility test: warp 0 of block 0 writes data with plain stores + fence.cta + flag; warp 1 of the same block polls the
// flag with ld.relaxed.cta (strong: ptxas must keep the loop; hits L1) and reads the data with weak ld.ca after a
// fence.cta. Counts stale reads and measures the hand-off latency. mode 1: the same across two BLOCKS with gpu-scope
// release + ld.relaxed.gpu / ld.cg (the path for region-boundary state).
__global__ void k_vis(double *buf, long long *out, int iters, int mode)
{
    const int w = mode ? (int)blockIdx.x : (int)(threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (lane || (mode && threadIdx.x >= 32)) return;
    double *flag0 = buf + 64, *flag1 = buf + 96, *data = buf + 128;
    long long stale = 0;
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        if (w == 0) {
            st_wt(data, (double)i); st_wt(data + 20, (double)i);
            if (mode) asm volatile("fence.release.gpu;":::"memory"); else __threadfence_block();
            st_wt(flag0, (double)i);
            if (mode) { while (ld_rlx_gpu(flag1) < (double)i) {} } else { while (ld_rlx_cta(flag1) < (double)i) {} }
        } else {
            if (mode) { while (ld_rlx_gpu(flag0) < (double)i) {} } else { while (ld_rlx_cta(flag0) < (double)i) {} }
            __threadfence_block();
            const double a = mode ? ld_cg(data) : ld_ca(data), b = mode ? ld_cg(data + 20) : ld_ca(data + 20);
            if (a != (double)i || b != (double)i) ++stale;
            if (mode) asm volatile("fence.release.gpu;":::"memory"); else __threadfence_block();
            st_wt(flag1, (double)i);
        }
    }
    long long t1 = clock64();
    if (w == 0) out[20 + 2 * mode] = (t1 - t0) / iters;
    else out[21 + 2 * mode] = stale;
}
int main()
{
    double *buf; long long *out, h[32] = {0};
    cudaMalloc(&buf, 1 << 20); cudaMemset(buf, 0, 1 << 20); cudaMalloc(&out, sizeof(h)); cudaMemset(out, 0, sizeof(h));
    k_lat<<<1, 32>>>(buf, out, 2000);
    k_vis<<<1, 64>>>(buf, out, 20000, 0);
    cudaDeviceSynchronize();
    cudaMemset(buf, 0, 4096);
    k_vis<<<2, 32>>>(buf, out, 20000, 1);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    const char *names[] = {"ld.ca chain", "ld.cg chain", "st+ld.ca same loc", "st+ld.ca same line", "st+__threadfence", "st+fence.release.gpu",
        "st+__threadfence_block", "__threadfence+ld.ca", "fence.release.gpu+ld.ca", "ld.relaxed.cta chain", "ld.relaxed.gpu chain", "generic smem chain",
        "st+fence.release.sys", "st+__threadfence_system", "fp64 div", "fp64 log", "fp64 sin"};
    printf("err=%s\n", cudaGetErrorString(e));
    for (int i = 0; i < 17; ++i) printf("%-28s %lld cycles\n", names[i], h[i]);
    printf("intra-block ping-pong via L1: %lld cycles per round trip, stale reads %lld\n", h[20], h[21]);
    printf("inter-block ping-pong via L2: %lld cycles per round trip, stale reads %lld\n", h[22], h[23]);
    return 0;
}
