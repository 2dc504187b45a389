// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16, A from TMEM or smem) as a function of N and of the
// number of independent accumulators the issue order rotates over.   nvcc -arch=sm_100a -o umma_bench umma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long make_desc(unsigned saddr, unsigned lbo, unsigned sbo)
{
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr & 0x3FFFFu) >> 4);
    d |= (unsigned long long)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (unsigned long long)((sbo >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    return d;
}
__device__ __forceinline__ unsigned make_idesc(int M, int N)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}
template <int NACC, int A_SMEM>
__global__ void bench(int N, int nmma, int dstride, long long *out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned tslot;
    __shared__ unsigned long long bar;
    const unsigned sb = smem_u32(smem);
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) reinterpret_cast<unsigned *>(smem)[i] = 0x3F803F80u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tb = tslot;
    bool leader = threadIdx.x == 0;
#ifdef USE_ELECT
    leader = false;
    if (threadIdx.x < 32) {
        unsigned pred;
        asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
        leader = pred != 0;
    }
#endif
    if (leader) {
        const unsigned idesc = make_idesc(128, N);
        const unsigned long long bd = make_desc(sb + 32768, (unsigned)N * 16, 128);
        const unsigned long long ad = make_desc(sb, 2048, 128);
        unsigned phase = 0;
        for (int rep = 0; rep < 3; rep++) {
            const long long t0 = clock64();
            for (int i = 0; i < nmma; i += 8) {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const unsigned d = tb + (unsigned)((j % NACC) * dstride);
                    const unsigned acc = (i + j) >= NACC ? 1u : 0u;
                    if (A_SMEM)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                     ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
                    else
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                     ::"r"(d), "r"(tb + 480u), "l"(bd), "r"(idesc), "r"(acc) : "memory");
                }
            }
            const long long t1 = clock64();
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)), "r"(phase) : "memory");
            phase ^= 1u;
            const long long t2 = clock64();
            out[rep * 2] = t1 - t0;  // issue
            out[rep * 2 + 1] = t2 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb));
}
int main()
{
    long long *d, h[6];
    cudaMalloc(&d, 64);
    cudaFuncSetAttribute(bench<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(bench<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(bench<4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(bench<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(bench<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(bench<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const int Ns[] = {16, 64, 128, 256};
    for (int a_smem = 0; a_smem < 2; a_smem++)
        for (int N : Ns)
            for (int nacc = 1; nacc <= 4; nacc *= 2) {
                if (N * nacc > 448) continue;
                for (int nmma : {8, 64}) {
#define RUN(NA, AS) bench<NA, AS><<<1, 128, 65536>>>(N, nmma, N, d)
                    if (a_smem) { if (nacc == 1) RUN(1, 1); else if (nacc == 2) RUN(2, 1); else RUN(4, 1); }
                    else { if (nacc == 1) RUN(1, 0); else if (nacc == 2) RUN(2, 0); else RUN(4, 0); }
                    cudaError_t e = cudaDeviceSynchronize();
                    cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
                    printf("A=%s N=%3d nacc=%d nmma=%2d: issue %5lld total %5lld cyc  (%.1f cyc/mma)  %s\n", a_smem ? "smem" : "tmem", N, nacc,
                           nmma, h[4], h[5], (double)h[5] / nmma, e == cudaSuccess ? "" : cudaGetErrorString(e));
                }
            }
    return 0;
}
