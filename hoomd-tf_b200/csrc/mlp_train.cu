// Training step of the pairwise-MLP force field (BASELINE config 4, online force matching): gradient of the
// force-matching loss with respect to the ~10.5k parameters, a fused Adam update, and the loss itself.
//
// Replaces what Keras does behind `model.train_on_batch(x=inputs, y=labels)` in the reference's label / training
// mode (/root/reference htf/tensorflowcompute.py:346-370; labels = HOOMD's forces, htf/TensorflowCompute.cc:177-187,
// :251-269): TensorFlow differentiates the MSE between compute_nlist_forces' output [N,4] and the labels, i.e. it
// back-propagates THROUGH the force gradient (a double backward).  Here that is one hand-written sweep.
//
// Math (oracle: oracle.pairwise_mlp_train_grads, pinned against torch's double backward to 1e-15).  Per pair p of row
// i: r, u(r), u'(r) from the value / tangent chain of mlp.cu; F_i = sum_j u' a/r, e_i = 1/2 sum_j u.  With
// L = mean_{i,c} (pred - label)^2 the parameter gradient is sum_p [ g_p du'_p/dtheta + h_p du_p/dtheta ],
// g_p = (dF_i . a_p / r_p) / (2N), h_p = de_i / (4N).  The reverse sweep through the tangent chain, per layer with
// outputs h = tanh(z), h' = (1 - h^2) z' and adjoints (hb, hb'):
//     zb' = (1 - h^2) hb',   zb = (1 - h^2) hb - 2 h h' hb',
//     dW += zb^T h_in + zb'^T h'_in,   db += sum zb,   (hb, hb')_in = (zb W, zb' W).
// The predictions (dF, de) come from the inference kernel (mlp_force_kernel) run first; this kernel re-runs the
// forward per tile -- activations never go to HBM.
//
// Implementation: warp-level tensor-core MMAs (mma.sync m16n8k16, bf16 operands, fp32 accumulation).  A block walks
// chunks of 1024 neighbor slots, compacts each chunk's VALID slots (padding is about a third of a dense fluid's tensor
// and would cost as many MMAs as real pairs) and tiles them 64 at a time.  A block of four warps owns a tile; warp w
// runs the forward and backward CHAINS of slots 16w..16w+15 entirely in registers (the accumulator fragment of one layer is the A fragment of the next), and owns rows 16w..16w+15 of every
// weight-gradient accumulator, which stay in registers across all tiles of the block.  The weight gradients contract
// over the SLOT index, so their operands are transposed views of the activations: the chains leave h, h' and zb, zb'
// in shared memory as [slot][feature] bf16 and the MMAs read them with ldmatrix.trans.  The bias gradient rides on a
// constant ones column appended to h.  Per-block partial gradients are written out and summed by a second kernel in
// block order, so the result is bit-reproducible run to run.
//
// This is the first correct CUDA path of config 4; it uses the legacy warp-MMA pipe, not tcgen05 (the weight gradient
// needs the slot-major operands that TMEM-resident activations do not offer without a transpose through shared
// memory).  DESIGN.md states its measured fraction of the bf16 peak.
#include "common.cuh"

#include <cuda_bf16.h>

namespace {

constexpr int TR_TILE = 64;          // neighbor slots per block tile
constexpr int TR_WARPS = 4;
constexpr int TR_THREADS = TR_WARPS * 32;
constexpr int LD64 = 72;             // row stride (bf16) of 64-wide arrays: 64 features | ones/zero column | 7 zeros
constexpr int LD32 = 40;             // row stride of the 32-wide radial basis arrays

// raw fp32 parameter blob (include/htf_b200.h): W1[64][32] b1[64] W2[64][64] b2[64] W3[64][64] b3[64] w4[64] b4[1]
constexpr int T_W1 = 0, T_B1 = T_W1 + 64 * 32, T_W2 = T_B1 + 64, T_B2 = T_W2 + 64 * 64, T_W3 = T_B2 + 64,
              T_B3 = T_W3 + 64 * 64, T_W4 = T_B3 + 64, T_B4 = T_W4 + 64, T_COUNT = T_B4 + 1;

// shared memory map (bytes)
constexpr int S_W1 = 0;                              // bf16 [64][LD32]   W1[out][in]
constexpr int S_W2 = S_W1 + 64 * LD32 * 2;           // bf16 [64][LD64]   W2[out][in]
constexpr int S_W3 = S_W2 + 64 * LD64 * 2;
constexpr int S_W2T = S_W3 + 64 * LD64 * 2;          // bf16 [64][LD64]   W2[out][in] stored as [in][out]
constexpr int S_W3T = S_W2T + 64 * LD64 * 2;
constexpr int S_BIAS = S_W3T + 64 * LD64 * 2;        // fp32 b1[64] b2[64] b3[64] w4[64]
constexpr int S_A0 = S_BIAS + 4 * 64 * 4;            // bf16 [TILE][LD32] phi ; then phi'
constexpr int S_A0P = S_A0 + TR_TILE * LD32 * 2;
constexpr int S_A1 = S_A0P + TR_TILE * LD32 * 2;     // h1, h1', h2, h2' : bf16 [TILE][LD64]
constexpr int S_A1P = S_A1 + TR_TILE * LD64 * 2;
constexpr int S_A2 = S_A1P + TR_TILE * LD64 * 2;
constexpr int S_A2P = S_A2 + TR_TILE * LD64 * 2;
constexpr int S_ZB = S_A2P + TR_TILE * LD64 * 2;     // zb, zb' of the layer being differentiated
constexpr int S_ZBP = S_ZB + TR_TILE * LD64 * 2;
constexpr int S_RED = S_ZBP + TR_TILE * LD64 * 2;    // fp32 [TR_WARPS][65]: w4 / b4 gradient partials of the warps
constexpr int TR_CHUNK = 1024;                       // slots per chunk: the valid ones are compacted before they are tiled
constexpr int S_SEG = S_RED + TR_WARPS * 65 * 4 + 16;   // u16 [TR_CHUNK]: per-warp compacted slot offsets; then i32 [TR_WARPS] counts
constexpr int S_LIST = S_SEG + TR_CHUNK * 2;         // u16 [TR_CHUNK]: the chunk's valid slots, in order
constexpr int S_CNT = S_LIST + TR_CHUNK * 2;         // i32 [TR_WARPS]
constexpr int TR_SMEM = S_CNT + TR_WARPS * 4;

struct TrainParams {
    const float4 *nlist;       // [rows][K]
    long long rows;
    int K;
    const float *raw;          // fp32 parameters
    const float4 *pred;        // [rows] forces + energy of the current parameters (mlp_force_kernel)
    const float4 *labels;      // [rows]
    float rbf_high;
    float inv_2n, inv_4n;      // 1 / (2 N_total), 1 / (4 N_total)
    float *partial;            // [gridDim.x][T_COUNT]
};

__device__ __forceinline__ unsigned pack_bf16(float lo, float hi)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);          // .x = lo -> low 16 bits = the lower column
    return *reinterpret_cast<unsigned *>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(unsigned v)
{
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162 *>(&v));
}
__device__ __forceinline__ float tanh_fast(float x)
{
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// D += A (16x16, row) * B (16x8, col), bf16 in, fp32 accumulate
__device__ __forceinline__ void mma16816(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(unsigned (&r)[4], unsigned saddr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ldmatrix_x2_trans(unsigned &r0, unsigned &r1, unsigned saddr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(saddr));
}

// One Dense layer of the value and tangent chains for the warp's 16 slots:
//   z[nt] (+bias) = sum_ks A[ks] * W^T,  zp[nt] = sum_ks Ap[ks] * W^T;   W is bf16 [64][ld] (row = output) in shared memory.
template <int KS>
__device__ __forceinline__ void dense_fwd(float (&z)[8][4], float (&zp)[8][4], const unsigned (&a)[4][4], const unsigned (&ap)[4][4],
                                          const __nv_bfloat16 *W, int ld, const float *bias, int g, int t)
{
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
        const float b0 = bias[8 * nt + 2 * t], b1 = bias[8 * nt + 2 * t + 1];
        z[nt][0] = b0; z[nt][1] = b1; z[nt][2] = b0; z[nt][3] = b1;
        zp[nt][0] = zp[nt][1] = zp[nt][2] = zp[nt][3] = 0.f;
        const __nv_bfloat16 *wrow = W + (8 * nt + g) * ld + 2 * t;
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
            const unsigned w0 = *reinterpret_cast<const unsigned *>(wrow + 16 * ks);
            const unsigned w1 = *reinterpret_cast<const unsigned *>(wrow + 16 * ks + 8);
            mma16816(z[nt], a[ks], w0, w1);
            mma16816(zp[nt], ap[ks], w0, w1);
        }
    }
}

// h = tanh(z), h' = (1 - h^2) z' in place; the bf16 copies become the next layer's A fragments and (optionally) go to
// shared memory as [slot][feature]
__device__ __forceinline__ void activate(float (&z)[8][4], float (&zp)[8][4], unsigned (&a)[4][4], unsigned (&ap)[4][4],
                                         __nv_bfloat16 *dst, __nv_bfloat16 *dstp, int row_lo, int t)
{
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float h = tanh_fast(z[nt][i]);
            z[nt][i] = h;
            zp[nt][i] = (1.0f - h * h) * zp[nt][i];
        }
        const unsigned lo = pack_bf16(z[nt][0], z[nt][1]), hi = pack_bf16(z[nt][2], z[nt][3]);
        const unsigned lop = pack_bf16(zp[nt][0], zp[nt][1]), hip = pack_bf16(zp[nt][2], zp[nt][3]);
        // accumulator tile nt -> A fragment of k-step nt/2: registers (0,1) from the even tile, (2,3) from the odd one
        a[nt >> 1][(nt & 1) * 2 + 0] = lo; a[nt >> 1][(nt & 1) * 2 + 1] = hi;
        ap[nt >> 1][(nt & 1) * 2 + 0] = lop; ap[nt >> 1][(nt & 1) * 2 + 1] = hip;
        if (dst) {
            *reinterpret_cast<unsigned *>(dst + row_lo * LD64 + 8 * nt + 2 * t) = lo;
            *reinterpret_cast<unsigned *>(dst + (row_lo + 8) * LD64 + 8 * nt + 2 * t) = hi;
            *reinterpret_cast<unsigned *>(dstp + row_lo * LD64 + 8 * nt + 2 * t) = lop;
            *reinterpret_cast<unsigned *>(dstp + (row_lo + 8) * LD64 + 8 * nt + 2 * t) = hip;
        }
    }
}

// adjoints of the pre-activations from the adjoints (hb, hbp) of the layer's outputs (h, hp), all in accumulator layout:
//   zb' = (1 - h^2) hb',  zb = (1 - h^2) hb - 2 h h' hb'.   Results overwrite hb / hbp, go to shared memory as bf16
//   [slot][feature] (operands of the weight gradient) and into A fragments (operands of the next input gradient).
__device__ __forceinline__ void adjoint_preact(float (&hb)[8][4], float (&hbp)[8][4], const float (&h)[8][4], const float (&hp)[8][4],
                                               unsigned (&a)[4][4], unsigned (&ap)[4][4], __nv_bfloat16 *zb_s, __nv_bfloat16 *zbp_s,
                                               int row_lo, int t)
{
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float s = 1.0f - h[nt][i] * h[nt][i];
            const float zbp = s * hbp[nt][i];
            const float zb = s * hb[nt][i] - 2.0f * h[nt][i] * hp[nt][i] * hbp[nt][i];
            hb[nt][i] = zb; hbp[nt][i] = zbp;
        }
        const unsigned lo = pack_bf16(hb[nt][0], hb[nt][1]), hi = pack_bf16(hb[nt][2], hb[nt][3]);
        const unsigned lop = pack_bf16(hbp[nt][0], hbp[nt][1]), hip = pack_bf16(hbp[nt][2], hbp[nt][3]);
        a[nt >> 1][(nt & 1) * 2 + 0] = lo; a[nt >> 1][(nt & 1) * 2 + 1] = hi;
        ap[nt >> 1][(nt & 1) * 2 + 0] = lop; ap[nt >> 1][(nt & 1) * 2 + 1] = hip;
        *reinterpret_cast<unsigned *>(zb_s + row_lo * LD64 + 8 * nt + 2 * t) = lo;
        *reinterpret_cast<unsigned *>(zb_s + (row_lo + 8) * LD64 + 8 * nt + 2 * t) = hi;
        *reinterpret_cast<unsigned *>(zbp_s + row_lo * LD64 + 8 * nt + 2 * t) = lop;
        *reinterpret_cast<unsigned *>(zbp_s + (row_lo + 8) * LD64 + 8 * nt + 2 * t) = hip;
    }
}

// (hb, hbp)_in = (zb W, zb' W): contraction over the OUTPUT index, so B comes from the transposed weights Wt[in][out]
__device__ __forceinline__ void input_grad(float (&hb)[8][4], float (&hbp)[8][4], const unsigned (&a)[4][4], const unsigned (&ap)[4][4],
                                           const __nv_bfloat16 *Wt, int g, int t)
{
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
        hb[nt][0] = hb[nt][1] = hb[nt][2] = hb[nt][3] = 0.f;
        hbp[nt][0] = hbp[nt][1] = hbp[nt][2] = hbp[nt][3] = 0.f;
        const __nv_bfloat16 *wrow = Wt + (8 * nt + g) * LD64 + 2 * t;
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            const unsigned w0 = *reinterpret_cast<const unsigned *>(wrow + 16 * ks);
            const unsigned w1 = *reinterpret_cast<const unsigned *>(wrow + 16 * ks + 8);
            mma16816(hb[nt], a[ks], w0, w1);
            mma16816(hbp[nt], ap[ks], w0, w1);
        }
    }
}

// reload a layer's outputs (bf16 [slot][feature] in shared memory) in accumulator layout
__device__ __forceinline__ void reload_act(float (&h)[8][4], float (&hp)[8][4], const __nv_bfloat16 *src, const __nv_bfloat16 *srcp,
                                           int row_lo, int t)
{
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
        float2 v;
        v = unpack_bf16(*reinterpret_cast<const unsigned *>(src + row_lo * LD64 + 8 * nt + 2 * t)); h[nt][0] = v.x; h[nt][1] = v.y;
        v = unpack_bf16(*reinterpret_cast<const unsigned *>(src + (row_lo + 8) * LD64 + 8 * nt + 2 * t)); h[nt][2] = v.x; h[nt][3] = v.y;
        v = unpack_bf16(*reinterpret_cast<const unsigned *>(srcp + row_lo * LD64 + 8 * nt + 2 * t)); hp[nt][0] = v.x; hp[nt][1] = v.y;
        v = unpack_bf16(*reinterpret_cast<const unsigned *>(srcp + (row_lo + 8) * LD64 + 8 * nt + 2 * t)); hp[nt][2] = v.x; hp[nt][3] = v.y;
    }
}

// dW[16w..16w+15][0..NIN) += zb^T h_in + zb'^T h'_in over the tile's 64 slots, bias gradient on the ones column:
// acc[nt] = columns 8nt..8nt+7, acc[NIN/8] = (sum zb, 0, ...).  Operands are [slot][feature] arrays read transposed.
template <int NIN, int LDIN>
__device__ __forceinline__ void weight_grad(float (&acc)[NIN / 8 + 1][4], unsigned zb_s, unsigned zbp_s, unsigned in_s, unsigned inp_s,
                                            int warp, int lane)
{
    const int r8 = lane & 7, sel = lane >> 3;
#pragma unroll
    for (int ks = 0; ks < TR_TILE / 16; ks++) {
        // A = zb^T: m = output feature (16 of this warp), k = slot.  Matrices: (k 0-7, m 0-7), (k 0-7, m 8-15), (k 8-15, m 0-7), (k 8-15, m 8-15)
        unsigned a[4], ap[4];
        const unsigned aoff = (unsigned)(((16 * ks + (sel >> 1) * 8 + r8) * LD64 + 16 * warp + (sel & 1) * 8) * 2);
        ldmatrix_x4_trans(a, zb_s + aoff);
        ldmatrix_x4_trans(ap, zbp_s + aoff);
#pragma unroll
        for (int np = 0; np < NIN / 16; np++) {
            // B = h_in: k = slot, n = input feature.  Matrices: (k 0-7, n 0-7), (k 8-15, n 0-7), (k 0-7, n 8-15), (k 8-15, n 8-15)
            unsigned b[4], bp[4];
            const unsigned boff = (unsigned)(((16 * ks + (sel & 1) * 8 + r8) * LDIN + 16 * np + (sel >> 1) * 8) * 2);
            ldmatrix_x4_trans(b, in_s + boff);
            ldmatrix_x4_trans(bp, inp_s + boff);
            mma16816(acc[2 * np], a, b[0], b[1]);
            mma16816(acc[2 * np], ap, bp[0], bp[1]);
            mma16816(acc[2 * np + 1], a, b[2], b[3]);
            mma16816(acc[2 * np + 1], ap, bp[2], bp[3]);
        }
        // ones column (feature NIN of the unprimed input): lanes 0-15 supply the two matrices' row addresses
        unsigned o0, o1;
        ldmatrix_x2_trans(o0, o1, in_s + (unsigned)(((16 * ks + ((lane >> 3) & 1) * 8 + r8) * LDIN + NIN) * 2));
        mma16816(acc[NIN / 8], a, o0, o1);
    }
}

__global__ void __launch_bounds__(TR_THREADS, 2) mlp_train_kernel(const TrainParams P)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    __nv_bfloat16 *W1s = reinterpret_cast<__nv_bfloat16 *>(smem + S_W1), *W2s = reinterpret_cast<__nv_bfloat16 *>(smem + S_W2),
                  *W3s = reinterpret_cast<__nv_bfloat16 *>(smem + S_W3), *W2t = reinterpret_cast<__nv_bfloat16 *>(smem + S_W2T),
                  *W3t = reinterpret_cast<__nv_bfloat16 *>(smem + S_W3T);
    float *bias_s = reinterpret_cast<float *>(smem + S_BIAS);        // b1 | b2 | b3 | w4
    __nv_bfloat16 *A0 = reinterpret_cast<__nv_bfloat16 *>(smem + S_A0), *A0p = reinterpret_cast<__nv_bfloat16 *>(smem + S_A0P),
                  *A1 = reinterpret_cast<__nv_bfloat16 *>(smem + S_A1), *A1p = reinterpret_cast<__nv_bfloat16 *>(smem + S_A1P),
                  *A2 = reinterpret_cast<__nv_bfloat16 *>(smem + S_A2), *A2p = reinterpret_cast<__nv_bfloat16 *>(smem + S_A2P),
                  *ZB = reinterpret_cast<__nv_bfloat16 *>(smem + S_ZB), *ZBp = reinterpret_cast<__nv_bfloat16 *>(smem + S_ZBP);
    float *red_s = reinterpret_cast<float *>(smem + S_RED);
    const unsigned sb = (unsigned)__cvta_generic_to_shared(smem);

    // ---- parameters -> bf16 tiles (and transposes) in shared memory; pad columns of the activation arrays ----
    for (int e = tid; e < 64 * 32; e += TR_THREADS) W1s[(e >> 5) * LD32 + (e & 31)] = __float2bfloat16(P.raw[T_W1 + e]);
    for (int e = tid; e < 64 * 64; e += TR_THREADS) {
        const int o = e >> 6, i = e & 63;
        const __nv_bfloat16 w2 = __float2bfloat16(P.raw[T_W2 + e]), w3 = __float2bfloat16(P.raw[T_W3 + e]);
        W2s[o * LD64 + i] = w2; W3s[o * LD64 + i] = w3;
        W2t[i * LD64 + o] = w2; W3t[i * LD64 + o] = w3;
    }
    for (int e = tid; e < 64; e += TR_THREADS) {
        bias_s[e] = P.raw[T_B1 + e]; bias_s[64 + e] = P.raw[T_B2 + e]; bias_s[128 + e] = P.raw[T_B3 + e]; bias_s[192 + e] = P.raw[T_W4 + e];
    }
    for (int e = tid; e < TR_TILE * 8; e += TR_THREADS) {
        const int r = e >> 3, c = e & 7;
        const __nv_bfloat16 one = __float2bfloat16(c == 0 ? 1.0f : 0.0f), zero = __float2bfloat16(0.0f);
        A0[r * LD32 + 32 + c] = one; A0p[r * LD32 + 32 + c] = zero;
        A1[r * LD64 + 64 + c] = one; A1p[r * LD64 + 64 + c] = zero;
        A2[r * LD64 + 64 + c] = one; A2p[r * LD64 + 64 + c] = zero;
        ZB[r * LD64 + 64 + c] = zero; ZBp[r * LD64 + 64 + c] = zero;
    }
    __syncthreads();

    // weight-gradient accumulators of this warp's 16 output rows, kept across all tiles
    float acc1[5][4], acc2[9][4], acc3[9][4], gw4[8][2], gb4 = 0.f;
#pragma unroll
    for (int i = 0; i < 5; i++) acc1[i][0] = acc1[i][1] = acc1[i][2] = acc1[i][3] = 0.f;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        acc2[i][0] = acc2[i][1] = acc2[i][2] = acc2[i][3] = 0.f;
        acc3[i][0] = acc3[i][1] = acc3[i][2] = acc3[i][3] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) gw4[i][0] = gw4[i][1] = 0.f;

    const long long slots = P.rows * P.K;
    const long long nchunks = (slots + TR_CHUNK - 1) / TR_CHUNK;
    const float gapf = P.rbf_high / 31.0f, inv_gap = 31.0f / P.rbf_high;
    const int row_lo = 16 * warp + g;                       // this lane's two slots inside the tile: row_lo, row_lo + 8
    unsigned short *seg_s = reinterpret_cast<unsigned short *>(smem + S_SEG);
    unsigned short *list_s = reinterpret_cast<unsigned short *>(smem + S_LIST);
    int *cnt_s = reinterpret_cast<int *>(smem + S_CNT);

    for (long long chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
      // ---- compact the chunk's valid (non-padded) slots, order preserved: about a third of a dense fluid's slots are
      //      padding, and a padded slot would cost exactly as many MMAs as a real pair ----
      const long long base = chunk * TR_CHUNK;
      {
        int cnt = 0;
#pragma unroll 1
        for (int it = 0; it < TR_CHUNK / TR_THREADS; it++) {
            const int off = warp * (TR_CHUNK / TR_WARPS) + it * 32 + lane;
            const long long s = base + off;
            bool valid = false;
            if (s < slots) {
                const float4 d = __ldg(P.nlist + s);
                const float ax = d.x + 1e-7f, ay = d.y + 1e-7f, az = d.z + 1e-7f;
                valid = sqrtf(ax * ax + ay * ay + az * az) > 3e-6f;
            }
            const unsigned m = __ballot_sync(HTF_FULL, valid);
            if (valid) seg_s[warp * (TR_CHUNK / TR_WARPS) + cnt + __popc(m & ((1u << lane) - 1u))] = (unsigned short)off;
            cnt += __popc(m);
        }
        if (lane == 0) cnt_s[warp] = cnt;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; w++) woff += cnt_s[w];
        for (int i = lane; i < cnt; i += 32) list_s[woff + i] = seg_s[warp * (TR_CHUNK / TR_WARPS) + i];
        __syncthreads();
      }
      const int nvalid = cnt_s[0] + cnt_s[1] + cnt_s[2] + cnt_s[3];
      const int ntiles = (nvalid + TR_TILE - 1) / TR_TILE;
      __syncthreads();                                      // everyone has the counts before a later chunk rewrites them

      for (int tile = 0; tile < ntiles; tile++) {
        // ---- the lane's two slots: geometry and loss weights ----
        float rr[2], gp[2], hw[2];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int li = tile * TR_TILE + row_lo + 8 * q;
            const long long s = li < nvalid ? base + list_s[li] : slots;
            float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 df = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s < slots) {
                d = __ldg(P.nlist + s);
                const long long row = s / P.K;
                const float4 pr = __ldg(P.pred + row), lb = __ldg(P.labels + row);
                df = make_float4(pr.x - lb.x, pr.y - lb.y, pr.z - lb.z, pr.w - lb.w);
            }
            const float ax = d.x + 1e-7f, ay = d.y + 1e-7f, az = d.z + 1e-7f;
            const float r = sqrtf(ax * ax + ay * ay + az * az);
            const bool valid = r > 3e-6f;
            rr[q] = r;
            gp[q] = valid ? (df.x * ax + df.y * ay + df.z * az) / r * P.inv_2n : 0.f;
            hw[q] = valid ? df.w * P.inv_4n : 0.f;
        }
        // ---- radial basis (value and d/dr) in A-fragment layout; copy to shared memory ----
        unsigned a[4][4], ap[4][4];
#pragma unroll
        for (int ks = 0; ks < 2; ks++) {
#pragma unroll
            for (int half = 0; half < 2; half++) {           // columns 16ks + 8half + 2t, +1
                const int c0 = 16 * ks + 8 * half + 2 * t;
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const float u0 = rr[q] - (float)c0 * gapf, u1 = rr[q] - (float)(c0 + 1) * gapf;
                    const float p0 = __expf(-u0 * u0 * inv_gap), p1 = __expf(-u1 * u1 * inv_gap);
                    const float d0 = -2.0f * u0 * inv_gap * p0, d1 = -2.0f * u1 * inv_gap * p1;
                    const unsigned v = pack_bf16(p0, p1), vp = pack_bf16(d0, d1);
                    a[ks][half * 2 + q] = v; ap[ks][half * 2 + q] = vp;
                    *reinterpret_cast<unsigned *>(A0 + (row_lo + 8 * q) * LD32 + c0) = v;
                    *reinterpret_cast<unsigned *>(A0p + (row_lo + 8 * q) * LD32 + c0) = vp;
                }
            }
        }
        // ---- forward chains ----
        float z[8][4], zp[8][4];
        dense_fwd<2>(z, zp, a, ap, W1s, LD32, bias_s, g, t);
        activate(z, zp, a, ap, A1, A1p, row_lo, t);
        dense_fwd<4>(z, zp, a, ap, W2s, LD64, bias_s + 64, g, t);
        activate(z, zp, a, ap, A2, A2p, row_lo, t);
        dense_fwd<4>(z, zp, a, ap, W3s, LD64, bias_s + 128, g, t);
        activate(z, zp, a, ap, nullptr, nullptr, row_lo, t);              // h3, h3' stay in registers (z, zp)

        // ---- last layer: u = w4 . h3 + b4, u' = w4 . h3'.  Its gradient and the adjoints of (h3, h3') ----
        float hb[8][4], hbp[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            const float w0 = bias_s[192 + 8 * nt + 2 * t], w1 = bias_s[192 + 8 * nt + 2 * t + 1];
            gw4[nt][0] += hw[0] * z[nt][0] + gp[0] * zp[nt][0] + hw[1] * z[nt][2] + gp[1] * zp[nt][2];
            gw4[nt][1] += hw[0] * z[nt][1] + gp[0] * zp[nt][1] + hw[1] * z[nt][3] + gp[1] * zp[nt][3];
            hb[nt][0] = hw[0] * w0; hb[nt][1] = hw[0] * w1; hb[nt][2] = hw[1] * w0; hb[nt][3] = hw[1] * w1;
            hbp[nt][0] = gp[0] * w0; hbp[nt][1] = gp[0] * w1; hbp[nt][2] = gp[1] * w0; hbp[nt][3] = gp[1] * w1;
        }
        if (t == 0) gb4 += hw[0] + hw[1];

        // ---- layer 3 ----
        adjoint_preact(hb, hbp, z, zp, a, ap, ZB, ZBp, row_lo, t);
        input_grad(hb, hbp, a, ap, W3t, g, t);                            // adjoints of (h2, h2')
        __syncthreads();                                                  // zb / zb' of all four warps, and A2 / A2p
        weight_grad<64, LD64>(acc3, sb + S_ZB, sb + S_ZBP, sb + S_A2, sb + S_A2P, warp, lane);
        __syncthreads();
        // ---- layer 2 ----
        reload_act(z, zp, A2, A2p, row_lo, t);
        adjoint_preact(hb, hbp, z, zp, a, ap, ZB, ZBp, row_lo, t);
        input_grad(hb, hbp, a, ap, W2t, g, t);                            // adjoints of (h1, h1')
        __syncthreads();
        weight_grad<64, LD64>(acc2, sb + S_ZB, sb + S_ZBP, sb + S_A1, sb + S_A1P, warp, lane);
        __syncthreads();
        // ---- layer 1 ----
        reload_act(z, zp, A1, A1p, row_lo, t);
        adjoint_preact(hb, hbp, z, zp, a, ap, ZB, ZBp, row_lo, t);
        __syncthreads();
        weight_grad<32, LD32>(acc1, sb + S_ZB, sb + S_ZBP, sb + S_A0, sb + S_A0P, warp, lane);
        __syncthreads();                                                  // A0.. and ZB are rewritten by the next tile
      }
    }

    // ---- write this block's partial gradient (plain stores: every element has one owner) ----
    float *out = P.partial + (size_t)blockIdx.x * T_COUNT;
    const int o_lo = 16 * warp + g, o_hi = o_lo + 8;
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
        out[T_W1 + o_lo * 32 + 8 * nt + 2 * t] = acc1[nt][0]; out[T_W1 + o_lo * 32 + 8 * nt + 2 * t + 1] = acc1[nt][1];
        out[T_W1 + o_hi * 32 + 8 * nt + 2 * t] = acc1[nt][2]; out[T_W1 + o_hi * 32 + 8 * nt + 2 * t + 1] = acc1[nt][3];
    }
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
        out[T_W2 + o_lo * 64 + 8 * nt + 2 * t] = acc2[nt][0]; out[T_W2 + o_lo * 64 + 8 * nt + 2 * t + 1] = acc2[nt][1];
        out[T_W2 + o_hi * 64 + 8 * nt + 2 * t] = acc2[nt][2]; out[T_W2 + o_hi * 64 + 8 * nt + 2 * t + 1] = acc2[nt][3];
        out[T_W3 + o_lo * 64 + 8 * nt + 2 * t] = acc3[nt][0]; out[T_W3 + o_lo * 64 + 8 * nt + 2 * t + 1] = acc3[nt][1];
        out[T_W3 + o_hi * 64 + 8 * nt + 2 * t] = acc3[nt][2]; out[T_W3 + o_hi * 64 + 8 * nt + 2 * t + 1] = acc3[nt][3];
    }
    if (t == 0) {                                            // column 0 of the ones tile = the bias gradient
        out[T_B1 + o_lo] = acc1[4][0]; out[T_B1 + o_hi] = acc1[4][2];
        out[T_B2 + o_lo] = acc2[8][0]; out[T_B2 + o_hi] = acc2[8][2];
        out[T_B3 + o_lo] = acc3[8][0]; out[T_B3 + o_hi] = acc3[8][2];
    }
    // w4 / b4: sum over the eight lanes that share t (they hold different slots), then over the four warps
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            gw4[nt][0] += __shfl_xor_sync(HTF_FULL, gw4[nt][0], o);
            gw4[nt][1] += __shfl_xor_sync(HTF_FULL, gw4[nt][1], o);
        }
        if (g == 0) { red_s[warp * 65 + 8 * nt + 2 * t] = gw4[nt][0]; red_s[warp * 65 + 8 * nt + 2 * t + 1] = gw4[nt][1]; }
    }
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) gb4 += __shfl_xor_sync(HTF_FULL, gb4, o);
    if (lane == 0) red_s[warp * 65 + 64] = gb4;
    __syncthreads();
    if (tid < 65) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < TR_WARPS; w++) s += red_s[w * 65 + tid];
        out[T_W4 + tid] = s;                                 // T_B4 == T_W4 + 64
    }
}

// grads[e] = sum over blocks (in block order) of partial[b][e]
__global__ void __launch_bounds__(256) mlp_grad_reduce_kernel(const float *__restrict__ partial, int nblocks, float *__restrict__ grads)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= T_COUNT) return;
    float s = 0.f;
    for (int b = 0; b < nblocks; b++) s += partial[(size_t)b * T_COUNT + e];
    grads[e] = s;
}

// loss = mean over rows x 4 of (pred - labels)^2: per-block partial sums in double, summed in block order
__global__ void __launch_bounds__(256) mse_partial_kernel(const float4 *__restrict__ pred, const float4 *__restrict__ labels,
                                                          long long rows, double *__restrict__ partial)
{
    __shared__ double s_w[8];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < rows; i += (long long)gridDim.x * 256) {
        const float4 p = pred[i], l = labels[i];
        const float dx = p.x - l.x, dy = p.y - l.y, dz = p.z - l.z, dw = p.w - l.w;
        acc += (double)(dx * dx + dy * dy + dz * dz + dw * dw);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(HTF_FULL, acc, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; w++) s += s_w[w];
        partial[blockIdx.x] = s;
    }
}
__global__ void mse_final_kernel(const double *__restrict__ partial, int n, double scale, float *__restrict__ loss)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s = 0.0;
    for (int i = 0; i < n; i++) s += partial[i];
    *loss = (float)(s * scale);
}

// Keras Adam (tf.keras.optimizers.Adam behind train_on_batch): t = step count after the increment (device scalar)
__global__ void __launch_bounds__(256) adam_kernel(float *__restrict__ params, const float *__restrict__ grads, float *__restrict__ m,
                                                   float *__restrict__ v, float *__restrict__ t_p, int n, float lr, float beta1, float beta2,
                                                   float eps)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float t = *t_p + 1.0f;
    if (i < n) {
        const float g = grads[i];
        const float mi = beta1 * m[i] + (1.0f - beta1) * g;
        const float vi = beta2 * v[i] + (1.0f - beta2) * g * g;
        const float lr_t = lr * sqrtf(1.0f - powf(beta2, t)) / (1.0f - powf(beta1, t));
        params[i] -= lr_t * mi / (sqrtf(vi) + eps);
        m[i] = mi; v[i] = vi;
    }
}
__global__ void adam_tick_kernel(float *t_p) { if (threadIdx.x == 0 && blockIdx.x == 0) *t_p += 1.0f; }

}  // namespace

int htf_mlp_train_partial_floats(int sm_count) { return 2 * sm_count * T_COUNT; }

cudaError_t htf_launch_mlp_train(htf_ctx *ctx, const float4 *nlist, int64_t rows, int K, const float *raw, float rbf_high,
                                 const float4 *pred, const float4 *labels, int64_t n_total, float *partial, double *loss_partial,
                                 float *grads, float *loss, cudaStream_t st)
{
    static bool configured[HTF_MAX_DEVICES] = {false};
    bool &conf = configured[htf_current_device_slot()];
    if (!conf) {
        cudaError_t e = cudaFuncSetAttribute(mlp_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TR_SMEM);
        if (e != cudaSuccess) return e;
        conf = true;
    }
    TrainParams P;
    P.nlist = nlist; P.rows = rows; P.K = K; P.raw = raw; P.pred = pred; P.labels = labels; P.rbf_high = rbf_high;
    P.inv_2n = (float)(1.0 / (2.0 * (double)n_total)); P.inv_4n = (float)(1.0 / (4.0 * (double)n_total));
    P.partial = partial;
    const long long nchunks = (rows * K + TR_CHUNK - 1) / TR_CHUNK;
    int grid = 2 * ctx->sm_count;
    if ((long long)grid > nchunks) grid = (int)(nchunks > 0 ? nchunks : 1);
    mlp_train_kernel<<<grid, TR_THREADS, TR_SMEM, st>>>(P);
    mlp_grad_reduce_kernel<<<(T_COUNT + 255) / 256, 256, 0, st>>>(partial, grid, grads);
    ctx->launches += 2;
    if (loss) {
        const int lb = 2 * ctx->sm_count;
        mse_partial_kernel<<<lb, 256, 0, st>>>(pred, labels, rows, loss_partial);
        mse_final_kernel<<<1, 32, 0, st>>>(loss_partial, lb, 1.0 / (4.0 * (double)n_total), loss);
        ctx->launches += 2;
    }
    return cudaGetLastError();
}

cudaError_t htf_launch_adam(htf_ctx *ctx, float *params, const float *grads, float *m, float *v, float *t, int n, float lr,
                            float beta1, float beta2, float eps, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    adam_kernel<<<(n + 255) / 256, 256, 0, st>>>(params, grads, m, v, t, n, lr, beta1, beta2, eps);
    adam_tick_kernel<<<1, 32, 0, st>>>(t);
    ctx->launches += 2;
    return cudaGetLastError();
}

int htf_mlp_train_smem_bytes() { return TR_SMEM; }
