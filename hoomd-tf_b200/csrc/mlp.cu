// Pairwise-MLP neural force field over the neighbor tensor: energy and its radial derivative in ONE forward
// sweep on the 5th-generation tensor cores (tcgen05.mma, operands and accumulators in TMEM), sm_100a.
//
// Model (BASELINE config 3; the per-pair analogue of /root/reference htf/test-py/build_examples.py:231-241
// `RBF` + examples/08's Dense stack):
//   r      = nlist_rinv's safe norm of d_ij                       (htf/simmodel.py:630-631)
//   phi_c  = exp(-(r - mu_c)^2 / gap), mu = linspace(0, r_cut, 32) (RBFExpansion, htf/layers.py:7-49)
//   h1 = tanh(W1 phi + b1), h2 = tanh(W2 h1 + b2), h3 = tanh(W3 h2 + b3), u = w4 . h3 + b4   (64 wide)
//   e_i = 1/2 sum_j u_ij over the non-padded slots;  F_i = sum_j 2 d(sum e)/d d_ij = sum_j du/dr (d+1e-7)/r
//   (compute_nlist_forces convention, htf/simmodel.py:542-550).
//
// The network input is the scalar r, so du/dr is propagated FORWARD next to the activations (tangent mode):
//   h' = (1 - h^2) (W h'_prev),  phi'_c = -2 (r - mu_c)/gap phi_c,  du/dr = w4 . h3'
// Every layer is two M=128 UMMA chains against the same weight tile: Z = [H | 1] [W | b]^T (the bias rides on a
// constant "ones" K-block, split hi+lo bf16) and Z' = H' W^T.  The last Dense(1) is one more N=16 UMMA pair, so no
// dot product is left on the SIMT side: u and du/dr are one TMEM column each.
//
// One CTA per SM, 512 threads, two tile slots of 128 pairs.  Everything between the UMMAs lives in TMEM: a slot
// owns 192 columns (Z 64 | Z' 64 fp32 accumulators, H 32 | H' 32 = the next layer's bf16 A operands) and the
// tensor core reads A straight from TMEM (tcgen05.mma [d], [a], b-desc): activations never touch shared memory
// or HBM, the only shared-memory operand traffic is the resident weight tiles (29 KB bf16).  A slot is served by
// two warpgroups that split the columns (TMEM lane = pair, warps w and w+4 reach the same lanes): 32 of the 64
// outputs each per layer, and the lower / upper 16 radial basis centres.  The pairs arrive by TMA bulk copies
// (cp.async.bulk + mbarrier, double buffered, one tile ahead).  The epilogues of the two slots alternate through
// a named-barrier token, so one slot's UMMA round trip (~400 cycles for 9 UMMAs) hides under the other's epilogue;
// the MUFU pipe (192 MUFU.TANH per pair at 16 lanes/clk/SM) is the bound.  The radial basis is a multiplicative
// recurrence outwards from the middle centres (8 MUFU.EX2 per pair instead of 32), the tangent update is bf16x2
// HFMA2/HMUL2.  HBM traffic is the 16 B/pair read + 16 B/row accumulated.
//
// Measured on B200 (tools/micro/umma_bench.cu): a UMMA M=128 K=16 with A in TMEM costs 36 cycles at N=64 and 13 at
// N=16, but only when the issuing lane is chosen with elect.sync -- behind a `threadIdx.x == 0` branch ptxas wraps
// every UTCHMMA in an ELECT/BRA loop and the issue rate drops to one per ~47 cycles.
#include "common.cuh"

#include <cuda_bf16.h>
#include <cstdlib>

namespace {

constexpr int MLP_F = 32;       // radial basis features
constexpr int MLP_H = 64;       // hidden width
constexpr int MLP_TM = 128;     // pairs per tile = UMMA M
constexpr int MLP_SLOTS = 2;    // tile slots per CTA
#ifndef HTF_MLP_PARTS
#define HTF_MLP_PARTS 2
#endif
constexpr int MLP_PARTS = HTF_MLP_PARTS;        // warpgroups per slot: each takes 64 / MLP_PARTS output columns per layer
                                                // (measured at 1M x 64: 2 -> 6.7 ms, 4 -> 7.9 ms: more barrier participants,
                                                //  duplicated per-thread set-up, no gain in latency hiding)
constexpr int MLP_THREADS = MLP_SLOTS * MLP_PARTS * MLP_TM;
constexpr int PART_COLS = MLP_H / MLP_PARTS;    // fp32 accumulator columns per thread and layer
constexpr int PART_PAIRS = MLP_F / 2 / MLP_PARTS;   // radial basis centre pairs per thread
static_assert(MLP_PARTS == 2 || MLP_PARTS == 4, "two or four warpgroups per slot");
// TMEM columns: per slot Z | Z' (fp32 accumulators) and H | H' (bf16 pairs, A operands); one shared ones block
constexpr int TM_Z = 0, TM_ZP = 64, TM_H = 128, TM_HP = 160, TM_SLOT = 192, TM_ONES = MLP_SLOTS * TM_SLOT;

// ---- packed parameter blob (device), built by mlp_pack_kernel ----
// bf16 canonical layouts: element (row, k) of an [R x Kt] K-major operand sits at
//   (k/8) * (R/8)*128 + (row/8) * 128 + (row%8) * 16 + (k%8) * 2      [bytes]
// i.e. 8x8 core matrices, row groups contiguous (SBO = 128 B), K groups LBO = R*16 B apart.
// Every weight tile carries 16 extra K columns: column Kt = bf16(bias), Kt+1 = bf16(bias - bf16(bias)), rest 0.
constexpr int OFF_B1 = 0;                               // [W1 | b1] : N=64 x K=32+16   (K order: rbf_centre_of_k)
constexpr int OFF_B2 = OFF_B1 + 64 * (32 + 16) * 2;     // [W2 | b2] : 64 x 64+16
constexpr int OFF_B3 = OFF_B2 + 64 * (64 + 16) * 2;     // [W3 | b3] : 64 x 64+16
constexpr int OFF_B4 = OFF_B3 + 64 * (64 + 16) * 2;     // [w4 | b4] : N=16 x 64+16, rows 1..15 zero
constexpr int MLP_PACKED_BYTES = OFF_B4 + 16 * (64 + 16) * 2;
// K order of layer 1: the kernel emits the centres below 16 from the middle outwards, (14,15), (12,13), ...
__host__ __device__ constexpr int rbf_centre_of_k(int k) { return k < 16 ? 14 - 2 * (k / 2) + (k % 2) : k; }

// raw fp32 parameter blob (torch.nn.Linear layout, [out][in]):
//   W1[64][32] b1[64] W2[64][64] b2[64] W3[64][64] b3[64] w4[64] b4[1]
constexpr int RAW_W1 = 0, RAW_B1 = RAW_W1 + 64 * 32, RAW_W2 = RAW_B1 + 64, RAW_B2 = RAW_W2 + 64 * 64,
              RAW_W3 = RAW_B2 + 64, RAW_B3 = RAW_W3 + 64 * 64, RAW_W4 = RAW_B3 + 64, RAW_B4 = RAW_W4 + 64,
              RAW_COUNT = RAW_B4 + 1;

// value of element (row, k) of a [W | b] tile with Kt weight columns
__device__ __forceinline__ float wb_value(const float *w, const float *b, int row, int k, int Kt)
{
    if (k < Kt) return w[row * Kt + (Kt == MLP_F ? rbf_centre_of_k(k) : k)];
    const float bv = b[row];
    const float hi = __bfloat162float(__float2bfloat16(bv));
    return k == Kt ? hi : (k == Kt + 1 ? bv - hi : 0.f);
}

__global__ void mlp_pack_kernel(const float *__restrict__ raw, unsigned char *__restrict__ packed)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;      // one bf16 element of the blob: invert the layout
    if (e >= MLP_PACKED_BYTES / 2) return;
    const int off = e * 2;
    const int base = off < OFF_B2 ? OFF_B1 : (off < OFF_B3 ? OFF_B2 : (off < OFF_B4 ? OFF_B3 : OFF_B4));
    const int rows = base == OFF_B4 ? 16 : 64;
    const int rel = off - base;
    const int kg = rel / (rows * 16), rem = rel % (rows * 16);
    const int row = (rem / 128) * 8 + (rem % 128) / 16, k = kg * 8 + (rem % 16) / 2;
    float v;
    if (base == OFF_B1) v = wb_value(raw + RAW_W1, raw + RAW_B1, row, k, 32);
    else if (base == OFF_B2) v = wb_value(raw + RAW_W2, raw + RAW_B2, row, k, 64);
    else if (base == OFF_B3) v = wb_value(raw + RAW_W3, raw + RAW_B3, row, k, 64);
    else v = row == 0 ? wb_value(raw + RAW_W4, raw + RAW_B4, 0, k, 64) : 0.f;
    reinterpret_cast<__nv_bfloat16 *>(packed)[e] = __float2bfloat16(v);
}

// ---- shared memory map of the main kernel (dynamic) ----
constexpr int SM_W = 0;                                       // packed parameters (MLP_PACKED_BYTES)
constexpr int SM_PAIR = (MLP_PACKED_BYTES + 127) / 128 * 128; // float4[MLP_SLOTS][2][128]: TMA-staged pairs, double buffered
constexpr int SM_BAR = SM_PAIR + MLP_SLOTS * 2 * MLP_TM * 16; // mbarriers: UMMA[MLP_SLOTS], pairs[MLP_SLOTS][2]; tmem base
constexpr int MLP_SMEM = SM_BAR + 64;

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

// D[128 x N] (TMEM) (+)= A[128 x 16] (TMEM, lane = row, 8 columns of bf16 pairs) * B[N x 16]^T (shared memory)
__device__ __forceinline__ void umma_bf16(unsigned tmem_d, unsigned tmem_a, unsigned long long bdesc, unsigned idesc,
                                          unsigned accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
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

// one lane of a converged warp.  A leader chosen this way lets ptxas issue tcgen05.mma back to back; a plain
// `threadIdx.x == 0` branch wraps every UTCHMMA in an ELECT/BRA loop (~47 cycles per instruction, measured).
__device__ __forceinline__ bool elect_one()
{
    unsigned pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_load(unsigned dst_s, const void *src, unsigned bytes, unsigned bar_s)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_s), "l"(src), "r"(bytes), "r"(bar_s) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar_s)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_s) : "memory");
}

// slot-local barrier: the warpgroups of one slot, named barrier 1 + slot
__device__ __forceinline__ void slot_sync(int slot)
{
    asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "n"(MLP_PARTS * MLP_TM) : "memory");
}
// epilogue token: named barriers 3 + slot, all threads take part = the waiting slot (bar.sync) + the releasing slot (bar.arrive)
__device__ __forceinline__ void token_wait(int slot) { asm volatile("bar.sync %0, %1;" ::"r"(slot + 3), "n"(MLP_THREADS) : "memory"); }
__device__ __forceinline__ void token_arrive(int slot) { asm volatile("bar.arrive %0, %1;" ::"r"(slot + 3), "n"(MLP_THREADS) : "memory"); }

#define TMEM_LD16(taddr, v, o)                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
                 : "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]),    \
                   "=r"(v[o + 6]), "=r"(v[o + 7]), "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]),  \
                   "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15])                                  \
                 : "r"(taddr))
#define TMEM_ST16(taddr, v, o)                                                                              \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
                 ::"r"(taddr), "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]),  \
                   "r"(v[o + 6]), "r"(v[o + 7]), "r"(v[o + 8]), "r"(v[o + 9]), "r"(v[o + 10]), "r"(v[o + 11]),           \
                   "r"(v[o + 12]), "r"(v[o + 13]), "r"(v[o + 14]), "r"(v[o + 15]) : "memory")
#define TMEM_ST4(taddr, v)                                                                                  \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"                               \
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory")
#define TMEM_ST8(taddr, v)                                                                                  \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"                   \
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory")

__device__ __forceinline__ unsigned tmem_ld1(unsigned taddr)
{
    unsigned r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr));
    return r;
}

__device__ __forceinline__ unsigned pack_bf16x2(float lo, float hi)
{
    unsigned r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));     // first source -> upper half
    return r;
}
// volatile: the MUFU burst stays between token_wait and token_arrive (the compiler may move everything else)
__device__ __forceinline__ float tanh_approx(float x)
{
    float r;
    asm volatile("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// (h*h - 1) * g on bf16 pairs = MINUS the tangent (1 - h^2) g: one HFMA2 (exact h^2 - 1, rounded once) and one
// HMUL2.  The sign flips once per layer instead of costing a negation per element; after the three hidden
// layers the Dense(1) output is -du/dr.
__device__ __forceinline__ unsigned neg_tangent_bf16x2(unsigned h, unsigned g)
{
    unsigned d, r;
    const unsigned minus_one = 0xBF80BF80u;
    asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(h), "r"(h), "r"(minus_one));
    asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(d), "r"(g));
    return r;
}

// packed fp32 pairs (FMUL2 on sm_100a)
__device__ __forceinline__ unsigned long long pk2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned bf16x2_of(unsigned long long v)
{
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return pack_bf16x2(lo, hi);
}

struct MlpParams {
    const float4 *nlist;   // [rows][K] slots, or (compact) the compacted valid pairs (dx, dy, dz, row index as int bits)
    long long npairs;      // rows * K
    const long long *npairs_dev;   // compact: the number of valid pairs (device value written by the compaction pre-pass)
    int compact;
    int K;
    const unsigned char *packed;
    float gap, inv_gap;    // RBF centre spacing
    float kk;              // exp(-8 gap): ratio of successive stride-2 feature ratios
    float4 *fe;            // [rows]: += (Fx, Fy, Fz, e)   (zeroed by the launcher)
};

// One layer of the chain for one slot, issued by one elected lane (tz = the slot's first TMEM column):
//   Z  [128 x N] = [H | ones] [W | b]^T     (ksteps K-steps of 16 + the bias step)
//   Z' [128 x N] =  H' W^T
__device__ __forceinline__ void issue_layer(unsigned tz, unsigned ones_t, unsigned b_s, int N, int ksteps, unsigned bar_s)
{
    const unsigned idesc = make_idesc(MLP_TM, N);
    const unsigned b_lbo = (unsigned)N * 16;
#pragma unroll 1
    for (int k = 0; k < ksteps; k++) {
        const unsigned long long bd = make_desc(b_s + (unsigned)k * 2u * b_lbo, b_lbo, 128);
        umma_bf16(tz + TM_Z, tz + TM_H + 8u * k, bd, idesc, k > 0 ? 1u : 0u);
        umma_bf16(tz + TM_ZP, tz + TM_HP + 8u * k, bd, idesc, k > 0 ? 1u : 0u);
    }
    umma_bf16(tz + TM_Z, ones_t, make_desc(b_s + (unsigned)ksteps * 2u * b_lbo, b_lbo, 128), idesc, 1u);
    umma_commit(bar_s);
}

// row sums of a finished tile: pair force from (u, du/dr), warp reduction when a warp's 32 pairs share a row
__device__ __forceinline__ void finish_tile(const MlpParams &p, long long npairs, long long tile, int lt, float ax, float ay, float az,
                                            float w, float r, float uval, float dudr)
{
    const int lane = lt & 31;
    const long long pair = tile * MLP_TM + lt;
    const bool inb = pair < npairs, valid = inb && r > 3e-6f;
    const float coef = valid ? dudr / r : 0.f;
    float fx = coef * ax, fy = coef * ay, fz = coef * az, en = valid ? 0.5f * uval : 0.f;
    if (p.compact) {
        // compacted pairs carry their row; rows are non-decreasing along the list, so a warp holds a few row segments:
        // segmented inclusive scan, the last lane of every segment adds the segment's sums to its row
        const int rowid = valid ? __float_as_int(w) : -1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float tx = __shfl_up_sync(HTF_FULL, fx, o), ty = __shfl_up_sync(HTF_FULL, fy, o);
            const float tz_ = __shfl_up_sync(HTF_FULL, fz, o), te = __shfl_up_sync(HTF_FULL, en, o);
            const int tr = __shfl_up_sync(HTF_FULL, rowid, o);
            if (lane >= o && tr == rowid) { fx += tx; fy += ty; fz += tz_; en += te; }
        }
        const int nxt = __shfl_down_sync(HTF_FULL, rowid, 1);
        if (rowid >= 0 && (lane == 31 || nxt != rowid)) {
            float *dst = reinterpret_cast<float *>(p.fe + rowid);
            atomicAdd(dst + 0, fx); atomicAdd(dst + 1, fy); atomicAdd(dst + 2, fz); atomicAdd(dst + 3, en);
        }
        return;
    }
    if ((p.K & 31) == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            fx += __shfl_xor_sync(HTF_FULL, fx, o);
            fy += __shfl_xor_sync(HTF_FULL, fy, o);
            fz += __shfl_xor_sync(HTF_FULL, fz, o);
            en += __shfl_xor_sync(HTF_FULL, en, o);
        }
        const long long first = pair - lane;
        if (lane == 0 && first < npairs) {
            // K is a multiple of 32 here: row = (first / 32) / (K / 32), in 32 bits whenever it fits
            const long long f32 = first >> 5;
            const long long row = f32 < (1ll << 32) ? (long long)((unsigned)f32 / (unsigned)(p.K >> 5)) : f32 / (p.K >> 5);
            float *dst = reinterpret_cast<float *>(p.fe + row);
            atomicAdd(dst + 0, fx); atomicAdd(dst + 1, fy); atomicAdd(dst + 2, fz); atomicAdd(dst + 3, en);
        }
    } else if (inb) {
        float *dst = reinterpret_cast<float *>(p.fe + pair / p.K);
        atomicAdd(dst + 0, fx); atomicAdd(dst + 1, fy); atomicAdd(dst + 2, fz); atomicAdd(dst + 3, en);
    }
}

__global__ void __launch_bounds__(MLP_THREADS, 1) mlp_force_kernel(const MlpParams p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const int t = threadIdx.x;
    const int slot = t / (MLP_PARTS * MLP_TM), part = (t >> 7) % MLP_PARTS, lt = t & 127, lw = lt >> 5;
    const bool slot_warp0 = (t % (MLP_PARTS * MLP_TM)) < 32;   // issues this slot's UMMAs and TMA copies
    const unsigned sbase = smem_u32(smem);
    const unsigned bar_s = sbase + SM_BAR + 8u * slot;                     // UMMA completion
    const unsigned ldbar_s = sbase + SM_BAR + 8u * MLP_SLOTS + 16u * slot; // pair buffers 0, 1
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(smem + SM_BAR + 8 * MLP_SLOTS * 3);

    // parameters -> shared memory (resident for the whole kernel)
    for (int i = t; i < MLP_PACKED_BYTES / 16; i += MLP_THREADS)
        reinterpret_cast<uint4 *>(smem + SM_W)[i] = __ldg(reinterpret_cast<const uint4 *>(p.packed) + i);
    if (t == 0) {
        for (int s = 0; s < MLP_SLOTS * 3; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sbase + SM_BAR + 8u * s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (t < 32) {                       // the whole TMEM (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // parameter stores visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = *tmem_slot;
    const unsigned lane_off = (unsigned)(lw * 32) << 16;                   // this warp's 32 TMEM lanes
    const unsigned tz = tmem_base + (unsigned)slot * TM_SLOT;              // this slot's columns
    const unsigned w_s = sbase + SM_W;
    if (t < MLP_TM) {                   // ones block: K columns 0,1 = 1 (bias hi + lo), the other 14 zero
        unsigned one[8] = {0x3F803F80u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        TMEM_ST8(tmem_base + TM_ONES + lane_off, one);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const long long npairs = p.npairs_dev ? *p.npairs_dev : p.npairs;
    const long long ntiles = (npairs + MLP_TM - 1) / MLP_TM;
    const long long stride = (long long)gridDim.x * MLP_SLOTS;
    long long tile = (long long)blockIdx.x * MLP_SLOTS + slot;
    const unsigned pair_s = sbase + SM_PAIR + (unsigned)slot * (2 * MLP_TM * 16);
    // one elected lane per slot stages the pairs of a tile (whole tile, or the ragged last one, or nothing)
    auto stage_pairs = [&](long long tl, int buf) {
        long long n = npairs - tl * MLP_TM;
        n = n < 0 ? 0 : (n > MLP_TM ? MLP_TM : n);
        if (n > 0) tma_load(pair_s + (unsigned)buf * (MLP_TM * 16), p.nlist + tl * MLP_TM, (unsigned)n * 16u, ldbar_s + 8u * buf);
        else mbar_arrive(ldbar_s + 8u * buf);
    };
    if (slot_warp0 && elect_one()) stage_pairs(tile, 0);
    if (slot == 1) token_arrive(0);     // slot 0 owns the epilogue token first
    unsigned phase = 0;
    // the lower half finishes tile i while the first UMMAs of tile i+1 run: its pair and (u, du/dr) wait here
    long long pend_tile = -1;
    float pend_ax = 0.f, pend_ay = 0.f, pend_az = 0.f, pend_w = 0.f, pend_r = 1.f, pend_u = 0.f, pend_du = 0.f;
    // both slots run the same number of rounds (a slot past the end works on an all-padding tile): the epilogue
    // token alternates strictly between them
    int round = 0;
    for (long long tile0 = (long long)blockIdx.x * MLP_SLOTS; tile0 < ntiles; tile0 += stride, tile += stride, round++) {
        const int buf = round & 1;
        if (slot_warp0) {
            if (elect_one()) stage_pairs(tile + stride, buf ^ 1);         // next round's pairs
            __syncwarp();
        }
        mbar_wait(ldbar_s + 8u * buf, (unsigned)(round >> 1) & 1u);
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tile * MLP_TM + lt < npairs) d = *reinterpret_cast<const float4 *>(smem + SM_PAIR + ((size_t)(slot * 2 + buf) * MLP_TM + lt) * 16);
        const float ax = d.x + 1e-7f, ay = d.y + 1e-7f, az = d.z + 1e-7f;
        const float r = sqrtf(ax * ax + ay * ay + az * az);

        // ---- radial basis and its derivative: this warpgroup's 32 / MLP_PARTS centres by recurrence away from the
        //      middle: upper parts (16,17), (18,19), ...; lower parts (14,15), (12,13), ...  (K order: rbf_centre_of_k)
        {
            const float g = p.gap, s = p.inv_gap;
            const bool up = part >= MLP_PARTS / 2;
            const int blk = part % (MLP_PARTS / 2);                       // which run of PART_PAIRS pairs inside the half
            const float sgn = up ? 1.f : -1.f;
            const float c0 = up ? (float)(16 + 2 * PART_PAIRS * blk) : (float)(14 - 2 * PART_PAIRS * blk);
            const float ua = r - c0 * g, ub = ua - g;
            unsigned long long ph = pk2(__expf(-ua * ua * s), __expf(-ub * ub * s));
            unsigned long long ratio = pk2(__expf(sgn * 4.f * ua - 4.f * g), __expf(sgn * 4.f * ub - 4.f * g));   // phi_{c+-2}/phi_c
            const unsigned long long kk2 = pk2(p.kk, p.kk);
            const float wbase = 2.f * c0 - 2.f * s * r, sgn4 = 4.f * sgn;  // phi'_c = (2c - 2 r/gap) phi_c
            unsigned f[PART_PAIRS], fd[PART_PAIRS];
#pragma unroll
            for (int i = 0; i < PART_PAIRS; i++) {
                const float wa = fmaf(sgn4, (float)i, wbase);
                f[i] = bf16x2_of(ph);
                fd[i] = bf16x2_of(mul2(ph, pk2(wa, wa + 2.f)));
                if (i < PART_PAIRS - 1) { ph = mul2(ph, ratio); ratio = mul2(ratio, kk2); }
            }
            if (PART_PAIRS == 8) {
                TMEM_ST8(tz + TM_H + 8u * part + lane_off, f);
                TMEM_ST8(tz + TM_HP + 8u * part + lane_off, fd);
            } else {
                TMEM_ST4(tz + TM_H + 4u * part + lane_off, f);
                TMEM_ST4(tz + TM_HP + 4u * part + lane_off, fd);
            }
        }

        // ---- three hidden layers and Dense(1): UMMA pair -> epilogue back into the A columns ----
#pragma unroll 1
        for (int layer = 0; layer < 4; layer++) {
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            slot_sync(slot);
            if (slot_warp0) {
                if (elect_one()) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const unsigned b_s = w_s + (layer == 0 ? OFF_B1 : (layer == 1 ? OFF_B2 : (layer == 2 ? OFF_B3 : OFF_B4)));
                    issue_layer(tz, tmem_base + TM_ONES, b_s, layer == 3 ? 16 : MLP_H, layer == 0 ? MLP_F / 16 : MLP_H / 16, bar_s);
                }
                __syncwarp();
            }
            if (layer == 0 && part == 0 && pend_tile >= 0)                // previous tile's row sums, under this tile's UMMAs
                finish_tile(p, npairs, pend_tile, lt, pend_ax, pend_ay, pend_az, pend_w, pend_r, pend_u, pend_du);
            mbar_wait(bar_s, phase);
            phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (layer == 3) break;
            // One slot's epilogue at a time; the other one is inside its UMMA round trip.  Measured at 1M x 64:
            // 6.7 ms with the token, 8.2 ms with free-running slots (they fall into phase and the MUFU pipe idles
            // through both UMMA round trips), 8.5 ms when only the MUFU burst is serialised.
            token_wait(slot);
            unsigned z[PART_COLS], zp[PART_COLS], h[PART_COLS / 2], hp[PART_COLS / 2];
            const unsigned zc = tz + lane_off + (unsigned)(PART_COLS * part);    // this warpgroup's output columns
            TMEM_LD16(zc + TM_Z, z, 0);
            if (PART_COLS == 32) TMEM_LD16(zc + TM_Z + 16, z, PART_COLS - 16);
            TMEM_LD16(zc + TM_ZP, zp, 0);
            if (PART_COLS == 32) TMEM_LD16(zc + TM_ZP + 16, zp, PART_COLS - 16);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int q = 0; q < PART_COLS; q++) z[q] = __float_as_uint(tanh_approx(__uint_as_float(z[q])));
#pragma unroll
            for (int q = 0; q < PART_COLS / 2; q++) {
                h[q] = pack_bf16x2(__uint_as_float(z[2 * q]), __uint_as_float(z[2 * q + 1]));
                hp[q] = neg_tangent_bf16x2(h[q], pack_bf16x2(__uint_as_float(zp[2 * q]), __uint_as_float(zp[2 * q + 1])));
            }
            if (PART_COLS == 32) {
                TMEM_ST16(tz + TM_H + 16u * part + lane_off, h, 0);
                TMEM_ST16(tz + TM_HP + 16u * part + lane_off, hp, 0);
            } else {
                TMEM_ST8(tz + TM_H + 8u * part + lane_off, h);
                TMEM_ST8(tz + TM_HP + 8u * part + lane_off, hp);
            }
            token_arrive(1 - slot);
        }
        if (part == 0) {
            pend_u = __uint_as_float(tmem_ld1(tz + TM_Z + lane_off));     // u     = w4 . h3 + b4
            pend_du = -__uint_as_float(tmem_ld1(tz + TM_ZP + lane_off));  // du/dr = w4 . h3' (three sign flips, see neg_tangent)
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            pend_tile = tile; pend_ax = ax; pend_ay = ay; pend_az = az; pend_w = d.w; pend_r = r;
        }
    }
    if (part == 0 && pend_tile >= 0) finish_tile(p, npairs, pend_tile, lt, pend_ax, pend_ay, pend_az, pend_w, pend_r, pend_u, pend_du);
    // drain the pair copy staged for the round that never ran, then release the TMEM
    mbar_wait(ldbar_s + 8u * (round & 1), (unsigned)(round >> 1) & 1u);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

// ---- compaction pre-pass: the valid (non-padded) slots of the neighbor tensor, in order, as (dx, dy, dz, row) ----
// About a third of a dense fluid's slots are zero padding, and a padded slot costs the MLP kernel exactly as much as
// a real pair.  Three small passes (count per 1024-slot block, one-block scan, scatter) cost ~0.5 ms at 1M x 64 and
// take a third off the MLP kernel.
constexpr int CP_BLOCK = 1024;          // slots per block of the count / scatter kernels (256 threads x 4)

__device__ __forceinline__ bool slot_valid(const float4 d)
{
    const float ax = d.x + 1e-7f, ay = d.y + 1e-7f, az = d.z + 1e-7f;
    return sqrtf(ax * ax + ay * ay + az * az) > 3e-6f;
}

__global__ void __launch_bounds__(256) mlp_count_kernel(const float4 *__restrict__ nlist, long long slots, int *__restrict__ blk_cnt)
{
    __shared__ int wsum[8];
    const long long base = (long long)blockIdx.x * CP_BLOCK;
    int c = 0;
#pragma unroll
    for (int it = 0; it < 4; it++) {
        const long long s = base + it * 256 + threadIdx.x;
        const bool v = s < slots && slot_valid(__ldg(nlist + s));
        c += __popc(__ballot_sync(HTF_FULL, v));
    }
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) t += wsum[w];
        blk_cnt[blockIdx.x] = t;
    }
}

// one block: exclusive scan of blk_cnt[nb] in place, grand total -> *total
__global__ void __launch_bounds__(1024) mlp_scan_kernel(int *__restrict__ blk_cnt, int nb, long long *__restrict__ total)
{
    __shared__ long long wtot[32];
    __shared__ long long carry_s;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nb; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const int v = i < nb ? blk_cnt[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(HTF_FULL, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wtot[w] = incl;
        __syncthreads();
        long long basew = carry_s;
        for (int q = 0; q < w; q++) basew += wtot[q];
        if (i < nb) blk_cnt[i] = (int)(basew + incl - v);          // offsets fit 31 bits: slots < 2^31 is checked on the host
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = basew + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

__global__ void __launch_bounds__(256) mlp_scatter_kernel(const float4 *__restrict__ nlist, long long slots, int K,
                                                          const int *__restrict__ blk_off, float4 *__restrict__ out)
{
    __shared__ int wsum[4][8];
    const long long base = (long long)blockIdx.x * CP_BLOCK;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float4 d[4];
    unsigned m[4];
#pragma unroll
    for (int it = 0; it < 4; it++) {
        const long long s = base + it * 256 + threadIdx.x;
        d[it] = s < slots ? __ldg(nlist + s) : make_float4(0.f, 0.f, 0.f, 0.f);
        m[it] = __ballot_sync(HTF_FULL, s < slots && slot_valid(d[it]));
        if (lane == 0) wsum[it][w] = __popc(m[it]);
    }
    __syncthreads();
    int off = blk_off[blockIdx.x];
#pragma unroll
    for (int it = 0; it < 4; it++) {
        int o = off;
        for (int q = 0; q < w; q++) o += wsum[it][q];
        if (m[it] & (1u << lane)) {
            const long long s = base + it * 256 + threadIdx.x;
            float4 v = d[it];
            v.w = __int_as_float((int)(s / K));
            out[o + __popc(m[it] & ((1u << lane) - 1u))] = v;
        }
#pragma unroll
        for (int q = 0; q < 8; q++) off += wsum[it][q];
    }
}

// ---- the same compaction from the builder's per-row neighbor counts: the valid slots of a row are its first
//      min(count, K), so neither pass reads the padding (the pre-pass drops from two reads + one write of the
//      tensor to 4 bytes per row + one read and one write of the valid slots) ----
constexpr int CR_ROWS = 64;             // rows per block of the row-wise kernels (8 warps x 8 rows)

__global__ void __launch_bounds__(256) mlp_rowsum_kernel(const int *__restrict__ row_count, long long rows, int K,
                                                         int *__restrict__ blk_cnt)
{
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // one thread per block of CR_ROWS rows
    const long long r0 = b * CR_ROWS;
    if (r0 >= rows) return;
    int t = 0;
    for (int i = 0; i < CR_ROWS && r0 + i < rows; i++) t += min(__ldg(row_count + r0 + i), K);
    blk_cnt[b] = t;
}

__global__ void __launch_bounds__(256) mlp_rowscatter_kernel(const float4 *__restrict__ nlist, const int *__restrict__ row_count,
                                                             long long rows, int K, const int *__restrict__ blk_off,
                                                             float4 *__restrict__ out)
{
    __shared__ int s_off[CR_ROWS + 1];
    const long long r0 = (long long)blockIdx.x * CR_ROWS;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (w < 2) {                                            // offsets of the block's rows: two 32-lane scans
        const long long r = r0 + threadIdx.x;
        const int c = r < rows ? min(__ldg(row_count + r), K) : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(HTF_FULL, incl, o);
            if (lane >= o) incl += t;
        }
        s_off[1 + threadIdx.x] = incl;                      // inclusive within the half; the second half is shifted below
    }
    if (threadIdx.x == 0) s_off[0] = 0;
    __syncthreads();
    const int half = s_off[32];
    __syncthreads();
    if (w == 1) s_off[1 + threadIdx.x] += half;
    __syncthreads();
    const int base = blk_off[blockIdx.x];
    for (int i = w; i < CR_ROWS; i += 8) {                  // a warp copies a row's valid slots, 32 at a time
        const long long r = r0 + i;
        if (r >= rows) break;
        const int o = s_off[i], c = s_off[i + 1] - o;
        const float4 *src = nlist + r * K;
        for (int j = lane; j < c; j += 32) {
            float4 v = __ldg(src + j);
            v.w = __int_as_float((int)r);
            out[base + o + j] = v;
        }
    }
}

}  // namespace

int htf_mlp_packed_bytes_host() { return MLP_PACKED_BYTES; }
int htf_mlp_raw_count_host() { return RAW_COUNT; }

cudaError_t htf_launch_mlp_pack(htf_ctx *ctx, const float *raw, unsigned char *packed, cudaStream_t st)
{
    mlp_pack_kernel<<<(MLP_PACKED_BYTES / 2 + 255) / 256, 256, 0, st>>>(raw, packed);
    ctx->launches += 1;
    return cudaGetLastError();
}

cudaError_t htf_launch_mlp(htf_ctx *ctx, const float4 *nlist, int64_t rows, int K, const unsigned char *packed,
                           float rbf_high, float4 *fe, cudaStream_t st, const int32_t *row_count)
{
    if (rows <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(fe, 0, sizeof(float4) * (size_t)rows, st);
    if (e != cudaSuccess) return e;
    // large tensors: compact the valid slots first (scratch = one more copy of the tensor, owned by the context)
    const long long slots = (long long)rows * K;
    bool compact = slots >= (1ll << 20) && slots < (1ll << 31);
    if (const char *env = getenv("HTF_MLP_COMPACT")) compact = atoi(env) != 0 && slots < (1ll << 31);
    const bool by_rows = compact && row_count != nullptr;
    const int nb = by_rows ? (int)((rows + CR_ROWS - 1) / CR_ROWS) : (int)((slots + CP_BLOCK - 1) / CP_BLOCK);
    if (compact) {
        if (slots + MLP_TM > ctx->mlp_pairs_cap) {
            if (ctx->d_mlp_pairs) cudaFree(ctx->d_mlp_pairs);
            ctx->d_mlp_pairs = nullptr; ctx->mlp_pairs_cap = 0;
            if ((e = cudaMalloc(reinterpret_cast<void **>(&ctx->d_mlp_pairs), sizeof(float4) * (size_t)(slots + MLP_TM))) != cudaSuccess) return e;
            ctx->mlp_pairs_cap = slots + MLP_TM;
        }
        if (nb + 2 > ctx->mlp_blk_cap) {
            if (ctx->d_mlp_blk) cudaFree(ctx->d_mlp_blk);
            ctx->d_mlp_blk = nullptr; ctx->mlp_blk_cap = 0;
            if ((e = cudaMalloc(reinterpret_cast<void **>(&ctx->d_mlp_blk), sizeof(int) * (size_t)(nb + 2) + 16)) != cudaSuccess) return e;
            ctx->mlp_blk_cap = nb + 2;
        }
        long long *total = reinterpret_cast<long long *>(ctx->d_mlp_blk + ((nb + 2 + 1) / 2) * 2);   // 8-byte aligned tail
        if (by_rows) mlp_rowsum_kernel<<<(nb + 255) / 256, 256, 0, st>>>(row_count, rows, K, ctx->d_mlp_blk);
        else mlp_count_kernel<<<nb, 256, 0, st>>>(nlist, slots, ctx->d_mlp_blk);
        mlp_scan_kernel<<<1, 1024, 0, st>>>(ctx->d_mlp_blk, nb, total);
        if (by_rows) mlp_rowscatter_kernel<<<nb, 256, 0, st>>>(nlist, row_count, rows, K, ctx->d_mlp_blk, ctx->d_mlp_pairs);
        else mlp_scatter_kernel<<<nb, 256, 0, st>>>(nlist, slots, K, ctx->d_mlp_blk, ctx->d_mlp_pairs);
        ctx->launches += 3;
    }
    static bool configured_dev[HTF_MAX_DEVICES] = {false};  // the attribute is per device
    bool &configured = configured_dev[htf_current_device_slot()];
    if (!configured) {
        e = cudaFuncSetAttribute(mlp_force_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MLP_SMEM);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    MlpParams p;
    p.nlist = compact ? ctx->d_mlp_pairs : nlist; p.npairs = slots; p.K = K; p.packed = packed;
    p.compact = compact ? 1 : 0;
    p.npairs_dev = compact ? reinterpret_cast<const long long *>(ctx->d_mlp_blk + ((nb + 2 + 1) / 2) * 2) : nullptr;
    p.gap = rbf_high / (float)(MLP_F - 1); p.inv_gap = 1.0f / p.gap; p.kk = expf(-8.0f * p.gap); p.fe = fe;
    const long long ntiles = (p.npairs + MLP_TM - 1) / MLP_TM;
    long long grid = ctx->sm_count;
    if (grid > (ntiles + MLP_SLOTS - 1) / MLP_SLOTS) grid = (ntiles + MLP_SLOTS - 1) / MLP_SLOTS;
    mlp_force_kernel<<<(unsigned)grid, MLP_THREADS, MLP_SMEM, st>>>(p);
    ctx->launches += 1;
    return cudaGetLastError();
}
