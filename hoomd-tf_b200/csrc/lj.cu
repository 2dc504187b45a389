// One streaming pass over the neighbor tensor: LJModel forces + per-particle energy +
// virial, and (optionally, fused into the same read) the compute_rdf histogram.
//
// Replaces, for the built-in LJ model, the TensorFlow graph
//   nlist_rinv (/root/reference htf/simmodel.py:618-635) -> pair energy
//   (htf/test-py/build_examples.py:67-77) -> tf.gradients + *2 + reduce_sum
//   (htf/simmodel.py:542-550) -> _add_energy (:558-578) -> _compute_virial (:509-523),
// the TfToHoomd copies (htf/tf2hoomd_op/tf2hoomd.cc:48-59) and the 3x3 -> 6 virial scatter
// (htf/TensorflowCompute.cu:41-71); and for the RDF masked_nlist + tf.norm +
// tf.histogram_fixed_width (htf/simmodel.py:638-693).
//
// HBM-bound: 16*K bytes read per row, 16 (+24) written.  LPR lanes share a row, each lane
// streams float4 slots LPR apart (a warp request covers 32/LPR rows x 128 B = whole cache
// lines), partial sums are combined with xor shuffles inside the LPR-lane group.
#include "common.cuh"
#include <cstdlib>

namespace {

#ifndef HTF_LJ_THREADS
#define HTF_LJ_THREADS 128      // A/B at 1 M x 64 (4 lanes per row, counts): 64 -> 0.145, 128 -> 0.146, 256 -> 0.150, 512 -> 0.174 ms; 64 loses on the histogram
#endif
constexpr int LJ_THREADS = HTF_LJ_THREADS;
#ifndef HTF_LJ_UNROLL
#define HTF_LJ_UNROLL 4
#endif
constexpr int LJ_UNROLL = HTF_LJ_UNROLL;          // independent 16-byte loads in flight per lane
constexpr int RDF_MAX_BINS = 1024;    // nbins + 2 <= RDF_MAX_BINS (shared-memory histogram)
#ifndef HTF_HSPLIT
#define HTF_HSPLIT 2
#endif
constexpr int HSPLIT_MAX = HTF_HSPLIT;    // private histograms per warp (power of two), fewer when nbins is large

__device__ __forceinline__ float4 ld_stream(const float4 *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

struct PairParams {
    const float4 *nlist;
    long long rows;
    int K;
    const int *row_count;      // nullable [rows]: neighbors of the row (may exceed K); slots >= min(count, K) are the
                               // builder's zero padding and are not read
    float4 *fe;
    float *virial;
    int vcomp;                 // 0, 6 or 9
    // rdf
    const float *thr;          // [nb + 1]: thr[b], b = 1..nb-1, ascending thresholds in rsq space
    int nb;                    // nbins + 2
    float r_lo, inv_step;      // first guess only; the thresholds decide
    const float *row_type;     // type of row r at row_type[r * row_type_stride]
    long long row_type_stride;
    int type_i, type_j;
    unsigned long long *bins;
    // smooth coordination-number collective variable: s(r) = 1 / (1 + (r/r0)^6)
    float cv_inv_r0;
    float4 *cv_row;            // [rows]: (sum_j ds/dd_ij (x,y,z), sum_j s(r_ij))
    double *cv_sum;            // += sum over rows of the coordination number
    // slab mode (pipelined step): the pass walks cell-sorted slots [*slot_lo, *slot_hi) and handles row
    // row_map[slot] - map_row_lo when that particle belongs to [map_row_lo, map_row_hi)
    const int *row_map;
    const int *slot_lo, *slot_hi;
    long long map_row_lo, map_row_hi;
    int blocks_per_sm;         // host only: grid cap of a slab pass
    int hsplit;                // private histograms per warp (set by the launcher)
};

template <int LPR, bool FORCES, bool VIRIAL, bool RDF, bool CV>
__global__ void __launch_bounds__(LJ_THREADS) pair_pass_kernel(const PairParams p)
{
    extern __shared__ int s_hist[];             // RDF: [warps][nb] private histograms + thr copy
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int sub = lane % LPR;
    constexpr int RPW = 32 / LPR;               // rows per warp
    constexpr int WARPS = LJ_THREADS / 32;
    const int K = p.K;

    int *my_hist = nullptr;
    float *s_thr = nullptr;
    unsigned bin0 = 0;                          // lane-private count of bin 0 (padded slots land here)
    double cv_acc = 0.0;                        // CV: coordination numbers of the rows this lane reported
    if (RDF) {
        // HSPLIT private histograms per warp (lane mod HSPLIT): fewer lanes of one atomic instruction on the same bin
        my_hist = s_hist + (warp * p.hsplit + (lane & (p.hsplit - 1))) * p.nb;
        s_thr = reinterpret_cast<float *>(s_hist + WARPS * p.hsplit * p.nb);
        for (int q = threadIdx.x; q < WARPS * p.hsplit * p.nb; q += LJ_THREADS) s_hist[q] = 0;
        for (int q = threadIdx.x; q <= p.nb; q += LJ_THREADS) s_thr[q] = p.thr[q];
        __syncthreads();
    }

    long long nrows = p.rows, slot0 = 0;
    if (p.row_map) {
        slot0 = __ldg(p.slot_lo);
        nrows = (long long)__ldg(p.slot_hi) - slot0;
    }
    const long long groups = (nrows + RPW - 1) / RPW;       // one warp-iteration = RPW rows
    for (long long gi = (long long)blockIdx.x * WARPS + warp; gi < groups; gi += (long long)gridDim.x * WARPS) {
        // plain mode walks the rows from the last to the first: with spatially coherent particle order the rows the
        // builder wrote last are still in the 126 MB L2 (measured: 188.7 -> 186.3 us at 1 M x 64)
        long long row = (p.row_map ? gi : groups - 1 - gi) * RPW + lane / LPR;
        bool active = row < nrows;
        if (p.row_map) {
            const long long o = active ? (long long)__ldg(p.row_map + slot0 + row) : -1;
            active = o >= p.map_row_lo && o < p.map_row_hi;
            row = o - p.map_row_lo;
        }
        const float4 *rp = p.nlist + (active ? row : 0) * K;
        int kv = K;                                 // valid slots of this lane's row
        if (p.row_count) kv = active ? min(__ldg(p.row_count + row), K) : 0;
        float fx = 0.f, fy = 0.f, fz = 0.f, en = 0.f;
        float vxx = 0.f, vxy = 0.f, vxz = 0.f, vyy = 0.f, vyz = 0.f, vzz = 0.f;
        float cn = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
        bool row_in_rdf = false;
        if (RDF) {
            row_in_rdf = active;
            if (active && p.type_i >= 0) row_in_rdf = (__ldg(p.row_type + row * p.row_type_stride) == (float)p.type_i);
        }
        // slots at or past kmax are padding in every row of this warp: whole iterations are skipped (K = 96 at
        // liquid density: one of three)
        const int kmax = p.row_count ? __reduce_max_sync(HTF_FULL, kv) : K;
        for (int s0 = sub; s0 - sub < kmax; s0 += LPR * LJ_UNROLL) {
            float4 d[LJ_UNROLL];
#pragma unroll
            for (int u = 0; u < LJ_UNROLL; u++) {
                const int s = s0 + u * LPR;
                d[u] = (active && s < kv) ? ld_stream(rp + s) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            // scheduling fence: in one basic block ptxas sinks each load next to its use in the (branch-free) body and
            // the lane has one 16-byte request in flight instead of four (measured: 0.28 ms instead of 0.2 ms)
            if (p.vcomp < 0) continue;                      // never true: a basic-block boundary between loads and math
#pragma unroll
            for (int u = 0; u < LJ_UNROLL; u++) {
                const float dx = d[u].x, dy = d[u].y, dz = d[u].z;
                float r_guess = 0.f;                     // RDF: any r within a fraction of a bin of the exact one
                if (FORCES) {
                    // Branch-free: a padded slot (d = 0, rt = 1.7e-7) runs the same instructions with s = 0.
                    // sqrt and the reciprocal are MUFU seeds + one Newton step each (<= 1 ulp, no slow-path calls):
                    // the error of 1/(rt + 3e-6) (nlist_rinv) is amplified 13x by s^13, so the seed alone (1 ulp)
                    // would already spend 1.5e-6 of the 1e-5 contract; the 1/rt of the gradient enters linearly and
                    // keeps the 2-ulp MUFU.RSQ.
                    const float ax = dx + 1e-7f, ay = dy + 1e-7f, az = dz + 1e-7f;
                    const float rt2 = fmaxf(ax * ax + ay * ay + az * az, 1e-30f);
                    float irt;
                    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(irt) : "f"(rt2));
                    float rt = rt2 * irt;
                    rt = fmaf(fmaf(-rt, rt, rt2), 0.5f * irt, rt);
                    r_guess = rt;
                    const bool pair = rt > 3e-6f;
                    const float a = rt + 3e-6f;
                    float si;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(si) : "f"(a));
                    si = fmaf(si, fmaf(-a, si, 1.0f), si);
                    si = pair ? si : 0.0f;
                    const float s2 = si * si, s6 = s2 * s2 * s2;
                    en += 2.0f * (s6 * s6 - s6);
                    const float coef = (24.0f * s6 * si - 48.0f * s6 * s6 * si) * irt;
                    const float px = coef * ax, py = coef * ay, pz = coef * az;
                    fx += px; fy += py; fz += pz;
                    if (VIRIAL) {
                        // |F_pair| / (2 |d|) = |coef| |a| / (2 |d|); |a| and |d| differ by the 1e-7 offset of
                        // safe_norm only (< 2.2e-7 relative for r >= 0.8), far inside the 1e-5 contract
                        const float w = 0.5f * fabsf(coef);
                        const float wx = w * dx, wy = w * dy, wz = w * dz;
                        vxx += wx * dx; vxy += wx * dy; vxz += wx * dz;
                        vyy += wy * dy; vyz += wy * dz; vzz += wz * dz;
                    }
                    if (CV) {
                        // s = 1/(1 + x^6), x = rt/r0;  ds/dd = -6 x^6 s^2 / rt^2 * a
                        const float x = rt * p.cv_inv_r0, x2 = x * x, x6 = x2 * x2 * x2;
                        float sw = __fdividef(1.0f, 1.0f + x6);      // 1-ulp reciprocal: 1e-7 relative on s
                        sw = pair ? sw : 0.0f;
                        cn += sw;
                        const float cg = -6.0f * x6 * sw * sw * irt * irt;
                        gx += cg * ax; gy += cg * ay; gz += cg * az;
                    }
                }
                if (RDF) {
                    const int s = s0 + u * LPR;
                    if (row_in_rdf && s < kv) {
                        // type filter (masked_nlist): a slot of another type counts with d = 0.  The untyped histogram
                        // (warp-uniform test) skips the three mask multiplications
                        const bool typed = p.type_j >= 0;
                        float m = 1.0f, x = dx, y = dy, z = dz;
                        if (typed) {
                            m = (d[u].w == (float)p.type_j) ? 1.0f : 0.0f;
                            x = __fmul_rn(dx, m); y = __fmul_rn(dy, m); z = __fmul_rn(dz, m);
                        }
                        const float q = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
                        // first guess from an approximate r (off by far less than a bin), then the exact threshold
                        // table decides: thr[g] <= q < thr[g+1], thr[0] = 0, thr[nb] = +inf, so one step each way is
                        // enough and the clamps make both look-ups safe
                        if (!FORCES) r_guess = q * rsqrtf(fmaxf(q, 1e-30f));
                        int g = (int)((r_guess - p.r_lo) * p.inv_step);
                        g = max(0, min(p.nb - 1, g));
                        if (typed && m == 0.0f) g = 0;                      // masked entry: q = 0 whatever r was
                        g += (q >= s_thr[g + 1]) ? 1 : 0;
                        g -= (q < s_thr[g]) ? 1 : 0;
                        if (g == 0) bin0++;
                        else atomicAdd(my_hist + g, 1);
                    }
                }
            }
        }
        if (RDF && row_in_rdf && sub == 0) bin0 += (unsigned)(K - kv);      // the skipped padding: q = 0, bin 0
        if (FORCES) {
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) {
                fx += __shfl_xor_sync(HTF_FULL, fx, o);
                fy += __shfl_xor_sync(HTF_FULL, fy, o);
                fz += __shfl_xor_sync(HTF_FULL, fz, o);
                en += __shfl_xor_sync(HTF_FULL, en, o);
                if (VIRIAL) {
                    vxx += __shfl_xor_sync(HTF_FULL, vxx, o);
                    vxy += __shfl_xor_sync(HTF_FULL, vxy, o);
                    vxz += __shfl_xor_sync(HTF_FULL, vxz, o);
                    vyy += __shfl_xor_sync(HTF_FULL, vyy, o);
                    vyz += __shfl_xor_sync(HTF_FULL, vyz, o);
                    vzz += __shfl_xor_sync(HTF_FULL, vzz, o);
                }
            }
            if (active && sub == 0) p.fe[row] = make_float4(fx, fy, fz, en);
            if (CV) {
#pragma unroll
                for (int o = LPR / 2; o > 0; o >>= 1) {
                    cn += __shfl_xor_sync(HTF_FULL, cn, o);
                    gx += __shfl_xor_sync(HTF_FULL, gx, o);
                    gy += __shfl_xor_sync(HTF_FULL, gy, o);
                    gz += __shfl_xor_sync(HTF_FULL, gz, o);
                }
                if (active && sub == 0) {
                    p.cv_row[row] = make_float4(gx, gy, gz, cn);
                    cv_acc += (double)cn;
                }
            }
            if (VIRIAL && active) {
                if (p.vcomp == 6) {
                    // xx,xy,xz,yy,yz,zz  (htf/TensorflowCompute.cc:294-299)
                    // LPR < 6 lanes (K < 24) each take several components
                    for (int c = sub; c < 6; c += LPR) {
                        const float v = c == 0 ? vxx : c == 1 ? vxy : c == 2 ? vxz : c == 3 ? vyy : c == 4 ? vyz : vzz;
                        p.virial[row * 6 + c] = -v;
                    }
                } else {
                    for (int c = sub; c < 9; c += LPR) {
                        const int k = c / 3, l = c % 3;
                        const int a = min(k, l), b2 = max(k, l);
                        const float v = a == 0 ? (b2 == 0 ? vxx : b2 == 1 ? vxy : vxz) : a == 1 ? (b2 == 1 ? vyy : vyz) : vzz;
                        p.virial[row * 9 + c] = -v;
                    }
                }
            }
        }
    }
    if (CV) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cv_acc += __shfl_xor_sync(HTF_FULL, cv_acc, o);
        if (lane == 0 && cv_acc != 0.0) atomicAdd(p.cv_sum, cv_acc);
    }
    if (RDF) {
        // bin 0 holds every padded slot: reduce it in registers, one shared atomic per warp
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bin0 += __shfl_xor_sync(HTF_FULL, bin0, o);
        if (lane == 0 && bin0) atomicAdd(my_hist, (int)bin0);
        __syncthreads();
        for (int b = threadIdx.x; b < p.nb; b += LJ_THREADS) {
            unsigned long long t = 0;
#pragma unroll
            for (int w = 0; w < WARPS * p.hsplit; w++) t += (unsigned)s_hist[w * p.nb + b];
            if (t) atomicAdd(p.bins + b, t);
        }
    }
}

template <bool FORCES, bool VIRIAL, bool RDF, bool CV = false>
cudaError_t launch_pair(htf_ctx *ctx, const PairParams &p, cudaStream_t st)
{
    // lanes per row: 8 keeps every warp request on whole 128-byte lines; small K uses fewer
    const int K = p.K;
    int hs = HSPLIT_MAX;
    while (hs > 1 && sizeof(int) * ((size_t)(LJ_THREADS / 32) * hs * p.nb + p.nb + 1) > 40 * 1024) hs /= 2;
    PairParams pp = p;
    pp.hsplit = hs;
    const size_t smem = RDF ? sizeof(int) * ((size_t)(LJ_THREADS / 32) * hs * p.nb + p.nb + 1) : 0;
    // 4 lanes per row: a warp request covers 8 rows x 64 B (whole sectors), and the per-warp set-up and the shuffle
    // reductions are shared by 8 rows instead of 4 (measured at 1 M x 64: 0.180 vs 0.197 ms, with the builder's row
    // counts 0.149 vs 0.191 ms; 8 lanes per row can be asked for with HTF_LJ_LPR=8)
    int lpr = (K >= 12) ? 4 : (K >= 6 ? 2 : 1);
    static const int lpr_env = [] { const char *e = getenv("HTF_LJ_LPR"); return e ? atoi(e) : 0; }();
    if (lpr_env == 8 && K >= 24) lpr = 8;
    if (lpr_env == 2 && K >= 6) lpr = 2;
    const long long rpw = 32 / lpr;
    const long long groups = (p.rows + rpw - 1) / rpw;
    const long long blocks_needed = (groups + LJ_THREADS / 32 - 1) / (LJ_THREADS / 32);
    long long grid = blocks_needed;
    const long long persistent = (long long)ctx->sm_count * (2048 / LJ_THREADS);     // a full complement of threads per SM
    if ((RDF || CV) && grid > persistent) grid = persistent;       // fewer histogram / CV flushes
    // (a one-wave persistent grid for the plain LJ pass, which amortises the ~60 set-up instructions of a warp, was
    // measured: 0.198 ms either way at 1 M x 64, slower without the virial -- not adopted)
    if (p.row_map && p.blocks_per_sm > 0 && grid > (long long)ctx->sm_count * p.blocks_per_sm)
        grid = (long long)ctx->sm_count * p.blocks_per_sm;         // slab pass next to a running build: a slice of every SM
    if (grid < 1) grid = 1;
    switch (lpr) {
    case 8: pair_pass_kernel<8, FORCES, VIRIAL, RDF, CV><<<(unsigned)grid, LJ_THREADS, smem, st>>>(pp); break;
    case 4: pair_pass_kernel<4, FORCES, VIRIAL, RDF, CV><<<(unsigned)grid, LJ_THREADS, smem, st>>>(pp); break;
    case 2: pair_pass_kernel<2, FORCES, VIRIAL, RDF, CV><<<(unsigned)grid, LJ_THREADS, smem, st>>>(pp); break;
    default: pair_pass_kernel<1, FORCES, VIRIAL, RDF, CV><<<(unsigned)grid, LJ_THREADS, smem, st>>>(pp); break;
    }
    ctx->launches += 1;
    return cudaGetLastError();
}

}  // namespace

static void set_slab(PairParams &p, const HtfSlab *slab)
{
    p.row_map = nullptr; p.slot_lo = p.slot_hi = nullptr; p.map_row_lo = p.map_row_hi = 0; p.blocks_per_sm = 0;
    p.row_count = nullptr;
    if (!slab) return;
    p.blocks_per_sm = slab->blocks_per_sm;
    p.row_map = slab->sorted_idx; p.slot_lo = slab->slot_lo; p.slot_hi = slab->slot_hi;
    p.map_row_lo = slab->row_lo; p.map_row_hi = slab->row_hi;
}

cudaError_t htf_launch_lj(htf_ctx *ctx, const float4 *nlist, int64_t rows, int K, float4 *fe, float *virial,
                          int vcomp, const float *rdf_thr, int nb, const float *row_type, long long row_type_stride,
                          int type_i, int type_j, unsigned long long *bins, cudaStream_t st, const HtfSlab *slab,
                          const int32_t *row_count)
{
    if (rows <= 0) return cudaSuccess;
    PairParams p;
    set_slab(p, slab);
    p.row_count = row_count;
    p.nlist = nlist; p.rows = rows; p.K = K; p.fe = fe; p.virial = virial; p.vcomp = virial ? vcomp : 0;
    p.thr = rdf_thr; p.nb = nb; p.r_lo = ctx->rdf_lo;
    p.inv_step = (nb > 0 && ctx->rdf_hi > ctx->rdf_lo) ? (float)nb / (ctx->rdf_hi - ctx->rdf_lo) : 0.f;
    p.row_type = row_type; p.row_type_stride = row_type_stride; p.type_i = type_i; p.type_j = type_j; p.bins = bins;
    p.cv_inv_r0 = 0.f; p.cv_row = nullptr; p.cv_sum = nullptr;
    const bool rdf = bins != nullptr;
    if (rdf && nb > RDF_MAX_BINS) return cudaErrorInvalidValue;
    if (virial) return rdf ? launch_pair<true, true, true>(ctx, p, st) : launch_pair<true, true, false>(ctx, p, st);
    return rdf ? launch_pair<true, false, true>(ctx, p, st) : launch_pair<true, false, false>(ctx, p, st);
}

cudaError_t htf_launch_rdf(htf_ctx *ctx, const float4 *nlist, int64_t rows, int K, const float *row_type,
                           long long row_type_stride, const float *thr, int nb, int type_i, int type_j,
                           unsigned long long *bins, cudaStream_t st)
{
    if (rows <= 0) return cudaSuccess;
    if (nb > RDF_MAX_BINS) return cudaErrorInvalidValue;
    PairParams p;
    set_slab(p, nullptr);
    p.nlist = nlist; p.rows = rows; p.K = K; p.fe = nullptr; p.virial = nullptr; p.vcomp = 0;
    p.thr = thr; p.nb = nb; p.r_lo = ctx->rdf_lo;
    p.inv_step = (ctx->rdf_hi > ctx->rdf_lo) ? (float)nb / (ctx->rdf_hi - ctx->rdf_lo) : 0.f;
    p.row_type = row_type; p.row_type_stride = row_type_stride; p.type_i = type_i; p.type_j = type_j; p.bins = bins;
    p.cv_inv_r0 = 0.f; p.cv_row = nullptr; p.cv_sum = nullptr;
    return launch_pair<false, false, true>(ctx, p, st);
}


// LJ forces + virial + smooth coordination CV (+ RDF) in one pass: the EDS-biased model of BASELINE config 5
cudaError_t htf_launch_lj_cv(htf_ctx *ctx, const float4 *nlist, int64_t rows, int K, float4 *fe, float *virial,
                             int vcomp, float r0, float4 *cv_row, double *cv_sum, const float *rdf_thr, int nb,
                             unsigned long long *bins, cudaStream_t st, const HtfSlab *slab, const int32_t *row_count)
{
    if (rows <= 0) return cudaSuccess;
    PairParams p;
    set_slab(p, slab);
    p.row_count = row_count;
    p.nlist = nlist; p.rows = rows; p.K = K; p.fe = fe; p.virial = virial; p.vcomp = virial ? vcomp : 0;
    p.thr = rdf_thr; p.nb = nb; p.r_lo = ctx->rdf_lo;
    p.inv_step = (nb > 0 && ctx->rdf_hi > ctx->rdf_lo) ? (float)nb / (ctx->rdf_hi - ctx->rdf_lo) : 0.f;
    p.row_type = nullptr; p.row_type_stride = 0; p.type_i = -1; p.type_j = -1; p.bins = bins;
    p.cv_inv_r0 = 1.0f / r0; p.cv_row = cv_row; p.cv_sum = cv_sum;
    if (bins && nb > RDF_MAX_BINS) return cudaErrorInvalidValue;
    if (virial) return bins ? launch_pair<true, true, true, true>(ctx, p, st) : launch_pair<true, true, false, true>(ctx, p, st);
    return bins ? launch_pair<true, false, true, true>(ctx, p, st) : launch_pair<true, false, false, true>(ctx, p, st);
}

// ---- EDS bias update (htf/layers.py:142-195 EDSLayer.call), one thread: the whole per-step state machine -- Welford
// mean / ssd over the second half of the period, one tf.compat.v1 Adam step at n == period - 1 -- as a single launch
// on device-resident state instead of ~50 element-wise launches.  fp32, same operation order as the reference. ----
namespace {
__global__ void eds_step_kernel(const float *__restrict__ cv_p, const float *__restrict__ set_point, float *mean, float *ssd,
                                int *n_p, float *alpha, float *adam_m, float *adam_v, float *adam_t, int period, float lr,
                                float cv_scale)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float cv = *cv_p;
    const int n = *n_p, h = period / 2;
    float mu = *mean, s2 = *ssd;
    if (n == 0) { mu = 0.f; s2 = 0.f; }                                    // layers.py:161-165
    if (n > h) {                                                            // :169-178
        const float delta = cv - mu;
        mu = mu + delta / (float)(n - h);
        s2 = s2 + delta * (cv - mu);
    }
    if (n == period - 1) {                                                  // :181-190
        float g = -2.0f * (mu - *set_point) * s2 / (float)period / 2.0f;
        g = g / cv_scale;
        const float t = *adam_t + 1.0f;
        const float m = 0.9f * *adam_m + 0.1f * g;
        const float v = 0.999f * *adam_v + 0.001f * (g * g);
        const float lr_t = lr * sqrtf(1.0f - powf(0.999f, t)) / (1.0f - powf(0.9f, t));
        *alpha = *alpha - lr_t * m / (sqrtf(v) + 1e-8f);
        *adam_t = t; *adam_m = m; *adam_v = v;
    }
    *mean = mu; *ssd = s2;
    *n_p = (n + 1) % period;                                                // :193
}
}  // namespace

cudaError_t htf_launch_eds_step(htf_ctx *ctx, const float *cv, const float *set_point, float *mean, float *ssd, int *n,
                                float *alpha, float *adam_m, float *adam_v, float *adam_t, int period, float lr,
                                float cv_scale, cudaStream_t st)
{
    eds_step_kernel<<<1, 32, 0, st>>>(cv, set_point, mean, ssd, n, alpha, adam_m, adam_v, adam_t, period, lr, cv_scale);
    ctx->launches += 1;
    return cudaGetLastError();
}
