// Padded neighbor tensor [rows, K, 4] from the cell-sorted positions.
//
// Replaces prepareNeighbors (/root/reference htf/TensorflowCompute.cc:304-374, CPU rule
// "skip if rsq > rc^2"), its GPU twin htf_gpu_reshape_nlist_kernel
// (htf/TensorflowCompute.cu:80-151: one thread per row, 16-byte stores strided by 16*K)
// and the cudaMemset before it (.cu:180).
//
// Work decomposition: one warp per cell.  The warp stages the positions of the 3x3x3
// stencil (9 contiguous runs of the cell-sorted array when the x-neighbours do not wrap)
// into shared memory once, then walks the rows of its cell four at a time: every lane
// holds one candidate, tests it against the four rows, and hits are compacted with
// ballot/popc into a per-row shared-memory buffer.  A finished row leaves the SM as one
// contiguous K*16-byte coalesced store, zero padding included -- there is no memset pass
// and no partial-sector traffic.
//
// Arithmetic is the oracle's, bit for bit: d = p_j - p_i, compare-and-shift minimum image
// (HOOMD BoxDim::minImage CPU branch), rsq = (dx*dx + dy*dy) + dz*dz with explicit
// round-to-nearest mul/add (never contracted to FMA), keep iff rsq <= rc^2.
//
// The hot loop is kept branch-light: the staged candidate list is padded to a multiple of
// 32 with +inf sentinels (never a hit), rows missing from the last group of four sit at
// +inf too, each row buffer has 32 spare slots so a chunk can be appended without a bounds
// check, and the modulo-K overflow rule of the reference (htf/TensorflowCompute.cc:370)
// lives in a separate slow path entered only once a row has more than K neighbors.
#include "common.cuh"

#include <math_constants.h>

namespace {

constexpr int RPP = 4;            // rows tested against each staged candidate chunk
constexpr int SPARE = 32;         // spare slots per row buffer (one chunk of hits)

struct NlistParams {
    CellGrid g;
    const int *cell_start;
    const int *sorted_idx;
    const float4 *spos;
    int n_all;
    int row_lo, row_hi;
    int K;
    float rc2;
    int map_type_start;
    int cap;             // candidates staged per pass, multiple of 32
    float4 *out;
    int *idx_out;
    int *count_out;
    int *overflow;
};

__device__ __forceinline__ float wrap_axis(float d, float lo, float hi, float L)
{
    // if (d >= hi) d -= L; else if (d < lo) d += L;   (d - (-L) == d + L exactly)
    float adj = (d >= hi) ? L : ((d < lo) ? -L : 0.0f);
    return __fsub_rn(d, adj);
}

__device__ __forceinline__ void cp_async16(unsigned dst_s, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst_s), "l"(src) : "memory");
}

__device__ __forceinline__ void cp_async4(unsigned dst_s, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_s), "l"(src) : "memory");
}

__device__ __forceinline__ void sts128(unsigned addr, float x, float y, float z, float w)
{
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

struct RowState {
    float x[RPP], y[RPP], z[RPP], t[RPP];
    int self_rel[RPP];      // index of the row's own particle inside the staged window (or -1)
    unsigned wp[RPP];       // shared-space byte address of the row's next free slot
    unsigned lim[RPP];      // shared-space byte address of the row's slot K (fast path may not pass it)
};

// pair test shared by the fast and the slow path
template <bool WRAP, bool MAPPED>
__device__ __forceinline__ bool pair_hit(const NlistParams &p, const float4 &c, const RowState &rs, int r, int tl,
                                         float &dx, float &dy, float &dz)
{
    dx = __fsub_rn(c.x, rs.x[r]);
    dy = __fsub_rn(c.y, rs.y[r]);
    dz = __fsub_rn(c.z, rs.z[r]);
    if (WRAP) {
        dz = wrap_axis(dz, p.g.lo[2], p.g.hi[2], p.g.L[2]);
        dy = wrap_axis(dy, p.g.lo[1], p.g.hi[1], p.g.L[1]);
        dx = wrap_axis(dx, p.g.lo[0], p.g.hi[0], p.g.L[0]);
    }
    const float rsq = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    // rsq <= rc2 is !(rsq > rc2) for every non-NaN rsq; the +inf sentinels give inf/NaN -> no hit
    bool hit = (rsq <= p.rc2) & (tl != rs.self_rel[r]);
    if (MAPPED) hit = hit && (((int)c.w >= p.map_type_start) == ((int)rs.t[r] >= p.map_type_start));
    return hit;
}

// Fast path: appends without bounds checks (row buffers have SPARE extra slots).  Returns the
// chunk offset at which some row exceeded K (the caller switches to the slow path there), or
// mround when the window is exhausted.
template <bool WRAP, bool MAPPED, bool WITH_IDX>
__device__ __forceinline__ int test_fast(const NlistParams &p, const float4 *cand, const int *candidx, int mround,
                                         RowState &rs, unsigned rowbuf_s, int *rowidx, int stride, int lane)
{
    const unsigned lt = (1u << lane) - 1u;
    int t0 = 0;
    for (; t0 < mround; t0 += 32) {
        const int tl = t0 + lane;
        const float4 c = cand[tl];
        int cj = 0;
        if (WITH_IDX) cj = candidx[tl];
        bool ovf = false;
#pragma unroll
        for (int r = 0; r < RPP; r++) {
            float dx, dy, dz;
            const bool hit = pair_hit<WRAP, MAPPED>(p, c, rs, r, tl, dx, dy, dz);
            const unsigned mask = __ballot_sync(HTF_FULL, hit);
            const unsigned a = rs.wp[r] + (unsigned)__popc(mask & lt) * 16u;
            if (hit) {
                sts128(a, dx, dy, dz, c.w);
                if (WITH_IDX) rowidx[(a - rowbuf_s) >> 4] = cj;
            }
            rs.wp[r] += (unsigned)__popc(mask) * 16u;
            ovf |= rs.wp[r] > rs.lim[r];
        }
        if (ovf) { t0 += 32; break; }
    }
    return t0;
}

// Slow path (some row overflowed K): slot index wraps modulo K like htf/TensorflowCompute.cc:370;
// when one chunk holds more than K hits only the last writer of a slot may store.  Always applies
// the minimum image (a no-op for interior cells) so that a single copy of this cold code exists.
template <bool MAPPED, bool WITH_IDX>
__device__ __forceinline__ void test_slow(const NlistParams &p, const float4 *cand, const int *candidx, int t_begin,
                                          int mround, const RowState &rs, int (&cnt)[RPP], float4 *rowbuf, int *rowidx,
                                          int stride, int lane)
{
    const unsigned lt = (1u << lane) - 1u;
    const int K = p.K;
    for (int t0 = t_begin; t0 < mround; t0 += 32) {
        const int tl = t0 + lane;
        const float4 c = cand[tl];
        int cj = 0;
        if (WITH_IDX) cj = candidx[tl];
#pragma unroll
        for (int r = 0; r < RPP; r++) {
            float dx, dy, dz;
            const bool hit = pair_hit<true, MAPPED>(p, c, rs, r, tl, dx, dy, dz);
            const unsigned mask = __ballot_sync(HTF_FULL, hit);
            const int nh = __popc(mask);
            const int rank = __popc(mask & lt);
            if (hit && (rank + K >= nh)) {
                const int q = (cnt[r] + rank) % K;
                rowbuf[r * stride + q] = make_float4(dx, dy, dz, c.w);
                if (WITH_IDX) rowidx[r * stride + q] = cj;
            }
            cnt[r] += nh;
        }
    }
}

// after the chunk that pushed a row past K: move entries K.. down to (q mod K), in order
template <bool WITH_IDX>
__device__ __forceinline__ void fold_overflow(const NlistParams &p, const int (&cnt)[RPP], float4 *rowbuf, int *rowidx,
                                              int stride, int lane)
{
    __syncwarp();
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < RPP; r++)
            for (int q = p.K; q < cnt[r]; q++) {
                rowbuf[r * stride + q % p.K] = rowbuf[r * stride + q];
                if (WITH_IDX) rowidx[r * stride + q % p.K] = rowidx[r * stride + q];
            }
    }
    __syncwarp();
}

template <bool WITH_IDX, bool MAPPED>
__global__ void __launch_bounds__(128) nlist_build_kernel(const NlistParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int cell = blockIdx.x * wpb + warp;
    if (cell >= p.g.ncell) return;

    const int K = p.K;
    const int stride = K + SPARE;        // row buffer pitch in slots
    // per-warp carve-up: cand[cap] f4 | rowbuf[RPP*stride] f4 | (candidx[cap] i32 | rowidx[RPP*stride] i32)
    const size_t per_warp = (size_t)(p.cap + RPP * stride) * (WITH_IDX ? 20 : 16);
    unsigned char *base = smem_raw + per_warp * warp;
    float4 *cand = reinterpret_cast<float4 *>(base);
    float4 *rowbuf = cand + p.cap;
    int *candidx = reinterpret_cast<int *>(rowbuf + RPP * stride);
    int *rowidx = candidx + p.cap;

    const int b = __ldg(p.cell_start + cell), e = __ldg(p.cell_start + cell + 1);
    if (e == b) return;
    const bool full = (p.row_lo == 0 && p.row_hi == p.n_all);
    if (!full) {                         // skip cells without a row of this shard
        bool any = false;
        for (int s = b + lane; s < e; s += 32) {
            int o = __ldg(p.sorted_idx + s);
            any |= (o >= p.row_lo && o < p.row_hi);
        }
        if (!__any_sync(HTF_FULL, any)) return;
    }

    // ---- stencil runs: lane q < 27 describes stencil cell (q%3, (q/3)%3, q/9) ----
    const int nx = p.g.n[0], ny = p.g.n[1], nz = p.g.n[2];
    const int cx = cell % nx, cy = (cell / nx) % ny, cz = cell / (nx * ny);
    const bool xmerge = (nx > 3) && (cx >= 1) && (cx <= nx - 2);
    // minimum image can be skipped only when no stencil cell is reached through the
    // periodic boundary and |d| stays well below L/2 (n >= 5 cells per dimension)
    const bool wrap = !((nx >= 5 && cx >= 1 && cx <= nx - 2) && (ny >= 5 && cy >= 1 && cy <= ny - 2) &&
                        (nz >= 5 && cz >= 1 && cz <= nz - 2));
    int rb = 0, rl = 0;                  // this lane's run: begin slot, length
    if (lane < 27) {
        const int qx = lane % 3, qy = (lane / 3) % 3, qz = lane / 9;
        bool ok = true;
        int sx, sy, sz;
        if (nx <= 3) { sx = qx; ok &= (qx < nx); } else sx = (cx + qx - 1 + nx) % nx;
        if (ny <= 3) { sy = qy; ok &= (qy < ny); } else sy = (cy + qy - 1 + ny) % ny;
        if (nz <= 3) { sz = qz; ok &= (qz < nz); } else sz = (cz + qz - 1 + nz) % nz;
        if (xmerge) ok &= (qx == 0);
        if (ok) {
            const int c0 = (sz * ny + sy) * nx + sx;
            rb = __ldg(p.cell_start + c0);
            rl = __ldg(p.cell_start + c0 + (xmerge ? 3 : 1)) - rb;
        }
    }
    int incl = rl;                        // inclusive scan of run lengths -> virtual offsets
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(HTF_FULL, incl, o);
        if (lane >= o) incl += t;
    }
    const int roff = incl - rl;
    const int m = __shfl_sync(HTF_FULL, incl, 31);
    const unsigned runmask = __ballot_sync(HTF_FULL, rl > 0);
    // the run that holds this cell's own particles -> virtual index of "self" for each row
    const unsigned selfmask = __ballot_sync(HTF_FULL, rl > 0 && rb <= b && b < rb + rl);
    const int qself = __ffs(selfmask) - 1;
    const int self_base = __shfl_sync(HTF_FULL, roff - rb, qself);

    const int cap = p.cap;
    const int npass = (m + cap - 1) / cap;

    // stage window `pass` of the virtual candidate list; returns its length rounded up to 32
    // (the tail is filled with +inf sentinels, which can never be a hit)
    const unsigned cand_s = (unsigned)__cvta_generic_to_shared(cand);
    const unsigned candidx_s = (unsigned)__cvta_generic_to_shared(candidx);
    auto stage = [&](int pass) {
        const int w0 = pass * cap, w1 = min(m, w0 + cap);
        unsigned rm = runmask;
        while (rm) {
            const int q = __ffs(rm) - 1;
            rm &= rm - 1;
            const int qb = __shfl_sync(HTF_FULL, rb, q);
            const int ql = __shfl_sync(HTF_FULL, rl, q);
            const int qo = __shfl_sync(HTF_FULL, roff, q);
            const int lo = max(qo, w0), hi = min(qo + ql, w1);
            // asynchronous global->shared copies: every run is in flight before the single wait below
            for (int t = lo + lane; t < hi; t += 32) {
                const int s = qb + (t - qo);
                cp_async16(cand_s + (unsigned)(t - w0) * 16u, p.spos + s);
                if (WITH_IDX) cp_async4(candidx_s + (unsigned)(t - w0) * 4u, p.sorted_idx + s);
            }
        }
        const int len = w1 - w0, mround = (len + 31) & ~31;
        if (len + lane < mround) {
            cand[len + lane] = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, 0.f);
            if (WITH_IDX) candidx[len + lane] = -1;
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        return mround;
    };

    int mround = 0;
    if (npass == 1) mround = stage(0);

    const unsigned rowbuf_s = (unsigned)__cvta_generic_to_shared(rowbuf);
    for (int s0 = b; s0 < e; s0 += RPP) {
        RowState rs;
        int orig[RPP], cnt[RPP];
        bool anyrow = false;
#pragma unroll
        for (int r = 0; r < RPP; r++) {
            const int s = min(s0 + r, e - 1);                    // both loads are independent of each other
            const int o = __ldg(p.sorted_idx + s);
            float4 pi = __ldg(p.spos + s);
            const bool ok = (s0 + r < e) && o >= p.row_lo && o < p.row_hi;
            orig[r] = ok ? o : -1;
            if (!ok) pi = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, 0.f);           // never hits
            rs.x[r] = pi.x; rs.y[r] = pi.y; rs.z[r] = pi.z; rs.t[r] = pi.w;
            rs.wp[r] = rowbuf_s + (unsigned)(r * stride) * 16u;
            rs.lim[r] = rs.wp[r] + (unsigned)K * 16u;
            cnt[r] = 0;
            anyrow |= ok;
        }
        if (!anyrow) continue;

        bool overflowed = false;
        for (int pass = 0; pass < npass; pass++) {
            if (npass > 1) { __syncwarp(); mround = stage(pass); }
#pragma unroll
            for (int r = 0; r < RPP; r++) {
                const int rel = self_base + s0 + r - pass * cap;       // own particle inside this window?
                rs.self_rel[r] = (orig[r] >= 0 && rel >= 0 && rel < mround) ? rel : -1;
            }
            int t = 0;
            if (!overflowed) {
                t = wrap ? test_fast<true, MAPPED, WITH_IDX>(p, cand, candidx, mround, rs, rowbuf_s, rowidx, stride, lane)
                         : test_fast<false, MAPPED, WITH_IDX>(p, cand, candidx, mround, rs, rowbuf_s, rowidx, stride, lane);
                int cmax = 0;
#pragma unroll
                for (int r = 0; r < RPP; r++) {
                    cnt[r] = (int)((rs.wp[r] - rowbuf_s) >> 4) - r * stride;
                    cmax = max(cmax, cnt[r]);
                }
                if (cmax > K) {
                    overflowed = true;
                    fold_overflow<WITH_IDX>(p, cnt, rowbuf, rowidx, stride, lane);
                }
            }
            if (overflowed && t < mround)
                test_slow<MAPPED, WITH_IDX>(p, cand, candidx, t, mround, rs, cnt, rowbuf, rowidx, stride, lane);
        }
        __syncwarp();

        // ---- flush: one contiguous K*16 B store per row, zero padded ----
#pragma unroll
        for (int r = 0; r < RPP; r++) {
            if (orig[r] < 0) continue;
            const size_t row = (size_t)(orig[r] - p.row_lo);
            float4 *dst = p.out + row * K;
            const int c = cnt[r];
            for (int s = lane; s < K; s += 32) {
                float4 v = (s < c) ? rowbuf[r * stride + s] : make_float4(0.f, 0.f, 0.f, 0.f);
                dst[s] = v;
                if (WITH_IDX) p.idx_out[row * K + s] = (s < c) ? rowidx[r * stride + s] : -1;
            }
            if (lane == 0) {
                if (p.count_out) p.count_out[row] = c;
                if (p.overflow && c >= K) atomicMax(p.overflow, c);
            }
        }
        __syncwarp();
    }
}

template <bool WITH_IDX, bool MAPPED>
cudaError_t launch_variant(const NlistParams &p, int grid, int wpb, size_t smem, cudaStream_t st)
{
    static size_t configured = 0;        // per instantiation: largest dynamic smem opted in so far
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(nlist_build_kernel<WITH_IDX, MAPPED>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    nlist_build_kernel<WITH_IDX, MAPPED><<<grid, wpb * 32, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace

cudaError_t htf_launch_nlist(htf_ctx *ctx, int64_t row_lo, int64_t row_hi, float4 *out, int32_t *idx_out,
                             int32_t *count_out, int32_t *overflow, cudaStream_t st)
{
    if (row_hi <= row_lo) return cudaSuccess;
    NlistParams p;
    p.g = ctx->grid;
    p.cell_start = ctx->d_cell_start;
    p.sorted_idx = ctx->d_sorted_idx;
    p.spos = ctx->d_spos;
    p.n_all = (int)ctx->n_binned;
    p.row_lo = (int)row_lo;
    p.row_hi = (int)row_hi;
    p.K = ctx->K;
    p.rc2 = ctx->r_cut * ctx->r_cut;
    p.map_type_start = ctx->map_type_start;
    p.out = out;
    p.idx_out = idx_out;
    p.count_out = count_out;
    p.overflow = overflow;

    // stencil population: mean + 5 sigma (Poisson) + slack, rounded to a chunk
    const CellGrid &g = ctx->grid;
    const int stencil = min(g.n[0], 3) * min(g.n[1], 3) * min(g.n[2], 3);
    const double mean = (double)stencil * (double)ctx->n_binned / (double)g.ncell;
    int cap = (int)(mean + 5.0 * sqrt(mean > 1.0 ? mean : 1.0)) + 32;
    cap = (cap + 31) / 32 * 32;
    const bool with_idx = idx_out != nullptr;
    const bool mapped = ctx->map_type_start >= 0;
    const int bytes_per = with_idx ? 20 : 16;
    const int stride = p.K + SPARE;
    int wpb = 4;
    const size_t smem_max = 200 * 1024;
    // keep the per-block footprint within the opt-in limit; shrink the staging window first
    // (the kernel re-stages in passes), then the block
    while ((size_t)(cap + RPP * stride) * bytes_per * wpb > smem_max) {
        if (cap > 64) cap = max(64, cap / 2 / 32 * 32);
        else if (wpb > 1) wpb /= 2;
        else return cudaErrorInvalidValue;
    }
    p.cap = cap;
    const size_t smem = (size_t)(cap + RPP * stride) * bytes_per * wpb;
    const int grid = (g.ncell + wpb - 1) / wpb;
    ctx->launches += 1;
    if (with_idx) return mapped ? launch_variant<true, true>(p, grid, wpb, smem, st)
                                : launch_variant<true, false>(p, grid, wpb, smem, st);
    return mapped ? launch_variant<false, true>(p, grid, wpb, smem, st)
                  : launch_variant<false, false>(p, grid, wpb, smem, st);
}
