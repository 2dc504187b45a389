// Micro-benchmark: tcgen05.ld (32x32b.x16 / .x32) throughput per SM as a function of the number of warps issuing.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
#define LD16(taddr, v)                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),  \
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr))
#define ST16(taddr, v)                                                                              \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), \
                   "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory")
__global__ void bench(int nwarps, int iters, int store, long long *out, unsigned *sink)
{
    __shared__ unsigned tslot;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tb = tslot;
    const int w = threadIdx.x >> 5;
    unsigned acc = 0;
    long long t0 = 0, t1 = 0;
    if (w < nwarps) {
        const unsigned taddr = tb + ((unsigned)((w & 3) * 32) << 16) + (unsigned)((w >> 2) * 64);
        unsigned v[16];
        for (int i = 0; i < 16; i++) v[i] = threadIdx.x + i;
        ST16(taddr, v); ST16(taddr + 16, v); ST16(taddr + 32, v); ST16(taddr + 48, v);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        t0 = clock64();
        for (int it = 0; it < iters; it++) {
            if (store) {
                ST16(taddr, v); ST16(taddr + 16, v); ST16(taddr + 32, v); ST16(taddr + 48, v);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            } else {
                unsigned a[16], b[16], c[16], d[16];
                LD16(taddr, a); LD16(taddr + 16, b); LD16(taddr + 32, c); LD16(taddr + 48, d);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += a[0] + b[5] + c[7] + d[15];
            }
        }
        t1 = clock64();
    }
    if (acc == 0x12345) sink[0] = acc;
    if (threadIdx.x == 0) { out[0] = t1 - t0; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb));
}
int main()
{
    long long *d, h;
    unsigned *sink;
    cudaMalloc(&d, 64); cudaMalloc(&sink, 64);
    for (int store = 0; store < 2; store++)
        for (int nw : {1, 4, 8, 16}) {
            const int iters = 200;
            bench<<<1, 512>>>(nw, iters, store, d, sink);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            const double bytes = (double)nw * 32 * 64 * 4 * iters;
            printf("%s warps=%2d: %lld cyc, %.1f B/clk/SM (%.1f cyc per 64-col x 32-lane block) %s\n", store ? "STTM" : "LDTM", nw, h,
                   bytes / h, (double)h / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
