// Pairwise-MLP neural force field over the neighbor tensor, fused forward + input-gradient on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), sm_100a.
//
// Model (BASELINE config 3; the per-pair analogue of /root/reference htf/test-py/build_examples.py:231-241
// `RBF` + examples/08's Dense stack):
//   r      = nlist_rinv's safe norm of d_ij                       (htf/simmodel.py:630-631)
//   phi_c  = exp(-(r - mu_c)^2 / gap), mu = linspace(0, r_cut, 32) (RBFExpansion, htf/layers.py:7-49)
//   h1 = tanh(W1 phi + b1), h2 = tanh(W2 h1 + b2), h3 = tanh(W3 h2 + b3), u = w4 . h3 + b4   (64 wide)
//   e_i = 1/2 sum_j u_ij over the non-padded slots;  F_i = sum_j 2 d(sum e)/d d_ij = sum_j du/dr (d+1e-7)/r
//   (compute_nlist_forces convention, htf/simmodel.py:542-550).
//
// One CTA (128 threads) walks tiles of 128 pairs.  Thread t owns pair t of the tile = TMEM lane t.
// Six GEMMs per tile, M = 128: three forward (K = 32, 64, 64 -> N = 64) and three for the input
// gradient (K = 64 -> N = 64, 64, 32).  Activations live on chip for the whole chain: the A operand of
// every GEMM is written by the epilogue of the previous one into shared memory (bf16, UMMA canonical
// K-major layout, no swizzle), the weights (both W and W^T, 40 KB bf16) stay resident in shared memory,
// accumulators are read back from TMEM with tcgen05.ld.  tanh runs as tanh.approx.bf16x2 (two per MUFU).
// HBM traffic is the 16*K bytes/row read + 16 bytes/row written, the tensor pipe and the SIMT epilogue
// are the bound.
#include "common.cuh"

#include <cuda_bf16.h>

namespace {

constexpr int MLP_F = 32;       // radial basis features
constexpr int MLP_H = 64;       // hidden width
constexpr int MLP_TM = 128;     // pairs per tile = UMMA M
constexpr int MLP_THREADS = 128;

// ---- packed parameter blob (device), built by mlp_pack_kernel ----
// bf16 canonical layouts: element (row, k) of an [R x Kt] K-major operand sits at
//   (k/8) * (R/8)*128 + (row/8) * 128 + (row%8) * 16 + (k%8) * 2      [bytes]
// i.e. 8x8 core matrices, row groups contiguous (SBO = 128 B), K groups LBO = R*16 B apart.
constexpr int OFF_B1 = 0;                         // W1   : N=64 x K=32   (forward  layer 1)
constexpr int OFF_B2 = OFF_B1 + 64 * 32 * 2;      // W2   : 64 x 64
constexpr int OFF_B3 = OFF_B2 + 64 * 64 * 2;      // W3   : 64 x 64
constexpr int OFF_B4 = OFF_B3 + 64 * 64 * 2;      // W3^T : 64 x 64       (gradient through layer 3)
constexpr int OFF_B5 = OFF_B4 + 64 * 64 * 2;      // W2^T : 64 x 64
constexpr int OFF_B6 = OFF_B5 + 64 * 64 * 2;      // W1^T : N=32 x K=64
constexpr int OFF_FP = OFF_B6 + 32 * 64 * 2;      // fp32: b1[64] b2[64] b3[64] w4[64] b4 (+3 pad)
constexpr int MLP_PACKED_BYTES = OFF_FP + (4 * 64 + 4) * 4;

// raw fp32 parameter blob (torch.nn.Linear layout, [out][in]):
//   W1[64][32] b1[64] W2[64][64] b2[64] W3[64][64] b3[64] w4[64] b4[1]
constexpr int RAW_W1 = 0, RAW_B1 = RAW_W1 + 64 * 32, RAW_W2 = RAW_B1 + 64, RAW_B2 = RAW_W2 + 64 * 64,
              RAW_W3 = RAW_B2 + 64, RAW_B3 = RAW_W3 + 64 * 64, RAW_W4 = RAW_B3 + 64, RAW_B4 = RAW_W4 + 64,
              RAW_COUNT = RAW_B4 + 1;

__host__ __device__ constexpr int canon_off(int row, int k, int rows)
{
    return (k / 8) * (rows / 8) * 128 + (row / 8) * 128 + (row % 8) * 16 + (k % 8) * 2;
}

__global__ void mlp_pack_kernel(const float *__restrict__ raw, unsigned char *__restrict__ packed)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    auto put = [&](int off, int row, int k, int rows, float v) {
        *reinterpret_cast<__nv_bfloat16 *>(packed + off + canon_off(row, k, rows)) = __float2bfloat16(v);
    };
    if (t < 64 * 32) {                      // W1[o][i]: forward B1 (N=o, K=i); gradient B6 (N=i, K=o)
        const int o = t / 32, i = t % 32;
        const float v = raw[RAW_W1 + t];
        put(OFF_B1, o, i, 64, v);
        put(OFF_B6, i, o, 32, v);
    }
    if (t < 64 * 64) {
        const int o = t / 64, i = t % 64;
        const float v2 = raw[RAW_W2 + t], v3 = raw[RAW_W3 + t];
        put(OFF_B2, o, i, 64, v2);
        put(OFF_B5, i, o, 64, v2);
        put(OFF_B3, o, i, 64, v3);
        put(OFF_B4, i, o, 64, v3);
    }
    float *fp = reinterpret_cast<float *>(packed + OFF_FP);
    if (t < 64) {
        fp[t] = raw[RAW_B1 + t];
        fp[64 + t] = raw[RAW_B2 + t];
        fp[128 + t] = raw[RAW_B3 + t];
        fp[192 + t] = raw[RAW_W4 + t];
    }
    if (t == 0) fp[256] = raw[RAW_B4];
}

// ---- shared memory map of the main kernel (dynamic, 1024-aligned base) ----
constexpr int SM_W = 0;                                   // packed parameters (MLP_PACKED_BYTES)
constexpr int SM_A0 = (MLP_PACKED_BYTES + 127) / 128 * 128;   // Phi      [128 x 32] bf16   8 KB
constexpr int SM_A1 = SM_A0 + 128 * 32 * 2;               // H1       [128 x 64] bf16  16 KB
constexpr int SM_A2 = SM_A1 + 128 * 64 * 2;               // H2       [128 x 64]
constexpr int SM_A3 = SM_A2 + 128 * 64 * 2;               // Delta    [128 x 64] (reused for delta3, delta2, delta1)
constexpr int SM_BAR = SM_A3 + 128 * 64 * 2;              // mbarrier (8 B) + tmem base (4 B)
constexpr int MLP_SMEM = SM_BAR + 16;

// ---- PTX wrappers ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned long long make_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes)
{
    // UMMA shared-memory descriptor, SWIZZLE_NONE: [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version 1
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr & 0x3FFFFu) >> 4);
    d |= (unsigned long long)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (unsigned long long)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    return d;
}

__device__ __forceinline__ unsigned make_idesc(int M, int N)
{
    // kind::f16: D = F32 (1 << 4), A = BF16 (1 << 7), B = BF16 (1 << 10), both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
    return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc,
                                          unsigned idesc, unsigned accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_commit(unsigned bar_s)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_s) : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned bar_s, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar_s), "r"(parity) : "memory");
}

#define TMEM_LD16(taddr, v, o)                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
                 : "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]),    \
                   "=r"(v[o + 6]), "=r"(v[o + 7]), "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]),  \
                   "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15])                                  \
                 : "r"((taddr) + (o)))

__device__ __forceinline__ unsigned pack_bf16x2(float lo, float hi)
{
    unsigned r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));     // first source -> upper half
    return r;
}
__device__ __forceinline__ unsigned tanh_bf16x2(unsigned x)
{
    unsigned r;
    asm("tanh.approx.bf16x2 %0, %1;" : "=r"(r) : "r"(x));
    return r;
}
__device__ __forceinline__ float bf16_lo(unsigned v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(unsigned v) { return __uint_as_float(v & 0xffff0000u); }

// 16-byte store of 8 bf16 of this thread's row into a [128 x Kt] canonical A tile: K group kg
__device__ __forceinline__ void st_row8(unsigned tile_s, int t, int kg, unsigned a, unsigned b, unsigned c, unsigned d)
{
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tile_s + (unsigned)kg * 2048u + (unsigned)t * 16u),
                 "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void ld_row8(unsigned tile_s, int t, int kg, unsigned &a, unsigned &b, unsigned &c, unsigned &d)
{
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
                 : "r"(tile_s + (unsigned)kg * 2048u + (unsigned)t * 16u) : "memory");
}

struct MlpParams {
    const float4 *nlist;
    long long npairs;      // rows * K
    int K;
    const unsigned char *packed;
    float gap, inv_gap;    // RBF centre spacing
    float4 *fe;            // [rows]: += (Fx, Fy, Fz, e)   (zeroed by the launcher)
};

// one GEMM of the chain: D[128 x N] (TMEM) = A[128 x Kt] (smem) * B[N x Kt]^T (smem); issued by one thread
__device__ __forceinline__ void issue_gemm(unsigned tmem_d, unsigned a_s, unsigned b_s, int N, int Kt, unsigned bar_s)
{
    const unsigned idesc = make_idesc(MLP_TM, N);
    const unsigned a_lbo = MLP_TM * 16, b_lbo = (unsigned)N * 16;
    for (int k = 0; k < Kt / 16; k++) {
        const unsigned long long ad = make_desc(a_s + (unsigned)k * 2u * a_lbo, a_lbo, 128);
        const unsigned long long bd = make_desc(b_s + (unsigned)k * 2u * b_lbo, b_lbo, 128);
        umma_bf16(tmem_d, ad, bd, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(bar_s);
}

__global__ void __launch_bounds__(MLP_THREADS, 2) mlp_force_kernel(const MlpParams p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const unsigned sbase = smem_u32(smem);
    const unsigned bar_s = sbase + SM_BAR;
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(smem + SM_BAR + 8);

    // parameters -> shared memory (resident for the whole kernel)
    for (int i = t; i < MLP_PACKED_BYTES / 16; i += MLP_THREADS)
        reinterpret_cast<uint4 *>(smem + SM_W)[i] = __ldg(reinterpret_cast<const uint4 *>(p.packed) + i);
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {                    // TMEM: 64 fp32 columns x 128 lanes for the accumulator
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // parameter stores visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_d = *tmem_slot;
    const unsigned taddr = tmem_d + ((unsigned)(warp * 32) << 16);   // this warp's 32 lanes
    const float *fp = reinterpret_cast<const float *>(smem + SM_W + OFF_FP);
    unsigned phase = 0;

    const long long ntiles = (p.npairs + MLP_TM - 1) / MLP_TM;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // ---- pair -> features ----
        const long long pair = tile * MLP_TM + t;
        const bool inb = pair < p.npairs;
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
        if (inb) d = __ldg(p.nlist + pair);
        const float ax = d.x + 1e-7f, ay = d.y + 1e-7f, az = d.z + 1e-7f;
        const float r2 = ax * ax + ay * ay + az * az;
        const float r = sqrtf(r2);
        const bool valid = inb && r > 3e-6f;
        {
            unsigned w[4];
#pragma unroll
            for (int kg = 0; kg < MLP_F / 8; kg++) {
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int c = kg * 8 + 2 * q;
                    const float u0 = r - (float)c * p.gap, u1 = r - (float)(c + 1) * p.gap;
                    w[q] = pack_bf16x2(__expf(-u0 * u0 * p.inv_gap), __expf(-u1 * u1 * p.inv_gap));
                }
                st_row8(sbase + SM_A0, t, kg, w[0], w[1], w[2], w[3]);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();

        unsigned v[64];
        float uval = 0.f;
        // ================= forward =================
#pragma unroll 1
        for (int layer = 0; layer < 3; layer++) {
            if (t == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned a_s = sbase + (layer == 0 ? SM_A0 : (layer == 1 ? SM_A1 : SM_A2));
                const unsigned b_s = sbase + SM_W + (layer == 0 ? OFF_B1 : (layer == 1 ? OFF_B2 : OFF_B3));
                issue_gemm(tmem_d, a_s, b_s, MLP_H, layer == 0 ? MLP_F : MLP_H, bar_s);
            }
            mbar_wait(bar_s, phase);
            phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            TMEM_LD16(taddr, v, 0); TMEM_LD16(taddr, v, 16); TMEM_LD16(taddr, v, 32); TMEM_LD16(taddr, v, 48);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const float *bias = fp + layer * 64;
            const unsigned out_s = sbase + (layer == 0 ? SM_A1 : (layer == 1 ? SM_A2 : SM_A3));
#pragma unroll
            for (int kg = 0; kg < 8; kg++) {
                unsigned h[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int c = kg * 8 + 2 * q;
                    h[q] = tanh_bf16x2(pack_bf16x2(__uint_as_float(v[c]) + bias[c], __uint_as_float(v[c + 1]) + bias[c + 1]));
                }
                if (layer < 2) {
                    st_row8(out_s, t, kg, h[0], h[1], h[2], h[3]);
                } else {
                    // last hidden layer: u = w4 . h3 + b4 and delta3 = (1 - h3^2) w4 (A operand of the first gradient GEMM)
                    unsigned dl[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int c = kg * 8 + 2 * q;
                        const float h0 = bf16_lo(h[q]), h1 = bf16_hi(h[q]);
                        const float w0 = fp[192 + c], w1 = fp[192 + c + 1];
                        uval += h0 * w0 + h1 * w1;
                        dl[q] = pack_bf16x2((1.f - h0 * h0) * w0, (1.f - h1 * h1) * w1);
                    }
                    st_row8(out_s, t, kg, dl[0], dl[1], dl[2], dl[3]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
        }
        uval += fp[256];

        // ================= gradient w.r.t. the features =================
#pragma unroll 1
        for (int layer = 0; layer < 3; layer++) {
            const int N = layer == 2 ? MLP_F : MLP_H;
            if (t == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned b_s = sbase + SM_W + (layer == 0 ? OFF_B4 : (layer == 1 ? OFF_B5 : OFF_B6));
                issue_gemm(tmem_d, sbase + SM_A3, b_s, N, MLP_H, bar_s);
            }
            mbar_wait(bar_s, phase);
            phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            TMEM_LD16(taddr, v, 0); TMEM_LD16(taddr, v, 16);
            if (layer < 2) { TMEM_LD16(taddr, v, 32); TMEM_LD16(taddr, v, 48); }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (layer < 2) {
                // delta_l = (1 - h_l^2) * g_l, h_l read back from the forward A tile of the next layer
                const unsigned h_s = sbase + (layer == 0 ? SM_A2 : SM_A1);
#pragma unroll
                for (int kg = 0; kg < 8; kg++) {
                    unsigned h[4], dl[4];
                    ld_row8(h_s, t, kg, h[0], h[1], h[2], h[3]);
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int c = kg * 8 + 2 * q;
                        const float h0 = bf16_lo(h[q]), h1 = bf16_hi(h[q]);
                        dl[q] = pack_bf16x2((1.f - h0 * h0) * __uint_as_float(v[c]), (1.f - h1 * h1) * __uint_as_float(v[c + 1]));
                    }
                    st_row8(sbase + SM_A3, t, kg, dl[0], dl[1], dl[2], dl[3]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
        }

        // ---- du/dr = sum_c g_c dphi_c/dr, pair force, row sums ----
        float dudr = 0.f;
#pragma unroll
        for (int c = 0; c < MLP_F; c++) {
            const float uu = r - (float)c * p.gap;
            const float phi = __expf(-uu * uu * p.inv_gap);
            dudr += __uint_as_float(v[c]) * (-2.f * uu * p.inv_gap) * phi;
        }
        const float coef = valid ? dudr / r : 0.f;
        float fx = coef * ax, fy = coef * ay, fz = coef * az, en = valid ? 0.5f * uval : 0.f;
        if ((p.K & 31) == 0) {
            // the 32 pairs of a warp belong to one row: warp reduction, one atomic per component
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                fx += __shfl_xor_sync(HTF_FULL, fx, o);
                fy += __shfl_xor_sync(HTF_FULL, fy, o);
                fz += __shfl_xor_sync(HTF_FULL, fz, o);
                en += __shfl_xor_sync(HTF_FULL, en, o);
            }
            const long long first = tile * MLP_TM + warp * 32;
            if (lane == 0 && first < p.npairs) {
                float *dst = reinterpret_cast<float *>(p.fe + first / p.K);
                atomicAdd(dst + 0, fx); atomicAdd(dst + 1, fy); atomicAdd(dst + 2, fz); atomicAdd(dst + 3, en);
            }
        } else if (inb) {
            float *dst = reinterpret_cast<float *>(p.fe + pair / p.K);
            atomicAdd(dst + 0, fx); atomicAdd(dst + 1, fy); atomicAdd(dst + 2, fz); atomicAdd(dst + 3, en);
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem_d));
}

}  // namespace

int htf_mlp_packed_bytes_host() { return MLP_PACKED_BYTES; }
int htf_mlp_raw_count_host() { return RAW_COUNT; }

cudaError_t htf_launch_mlp_pack(htf_ctx *ctx, const float *raw, unsigned char *packed, cudaStream_t st)
{
    mlp_pack_kernel<<<(64 * 64 + 255) / 256, 256, 0, st>>>(raw, packed);
    ctx->launches += 1;
    return cudaGetLastError();
}

cudaError_t htf_launch_mlp(htf_ctx *ctx, const float4 *nlist, int64_t rows, int K, const unsigned char *packed,
                           float rbf_high, float4 *fe, cudaStream_t st)
{
    if (rows <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(fe, 0, sizeof(float4) * (size_t)rows, st);
    if (e != cudaSuccess) return e;
    static bool configured = false;
    if (!configured) {
        e = cudaFuncSetAttribute(mlp_force_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MLP_SMEM);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    MlpParams p;
    p.nlist = nlist; p.npairs = (long long)rows * K; p.K = K; p.packed = packed;
    p.gap = rbf_high / (float)(MLP_F - 1); p.inv_gap = 1.0f / p.gap; p.fe = fe;
    const long long ntiles = (p.npairs + MLP_TM - 1) / MLP_TM;
    long long grid = 2LL * ctx->sm_count;
    if (grid > ntiles) grid = ntiles;
    mlp_force_kernel<<<(unsigned)grid, MLP_THREADS, MLP_SMEM, st>>>(p);
    ctx->launches += 1;
    return cudaGetLastError();
}
