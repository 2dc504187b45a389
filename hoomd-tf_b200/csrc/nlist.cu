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
#include "common.cuh"

namespace {

constexpr int ROWS_PER_PASS = 4;   // rows tested against each staged candidate chunk

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

template <bool WRAP, bool WITH_IDX>
__device__ __forceinline__ void test_chunks(const NlistParams &p, const float4 *cand, const int *candidx,
                                            int mstage, int vbase, const float (&pix)[ROWS_PER_PASS],
                                            const float (&piy)[ROWS_PER_PASS], const float (&piz)[ROWS_PER_PASS],
                                            const float (&pit)[ROWS_PER_PASS], const int (&vself)[ROWS_PER_PASS],
                                            const bool (&rvalid)[ROWS_PER_PASS], int (&cnt)[ROWS_PER_PASS],
                                            float4 *rowbuf, int *rowidx, int lane)
{
    const int K = p.K;
    const unsigned lt = (1u << lane) - 1u;
    for (int t0 = 0; t0 < mstage; t0 += 32) {
        const int t = t0 + lane;
        const bool valid = t < mstage;
        float4 c = valid ? cand[t] : make_float4(0.f, 0.f, 0.f, 0.f);
        int cj = 0;
        if (WITH_IDX) cj = valid ? candidx[t] : -1;
#pragma unroll
        for (int r = 0; r < ROWS_PER_PASS; r++) {
            if (!rvalid[r]) continue;                       // warp-uniform
            float dx = __fsub_rn(c.x, pix[r]);
            float dy = __fsub_rn(c.y, piy[r]);
            float dz = __fsub_rn(c.z, piz[r]);
            if (WRAP) {
                dz = wrap_axis(dz, p.g.lo[2], p.g.hi[2], p.g.L[2]);
                dy = wrap_axis(dy, p.g.lo[1], p.g.hi[1], p.g.L[1]);
                dx = wrap_axis(dx, p.g.lo[0], p.g.hi[0], p.g.L[0]);
            }
            float rsq = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            bool hit = valid && !(rsq > p.rc2) && (vbase + t != vself[r]);
            if (p.map_type_start >= 0)                       // warp-uniform
                hit = hit && (((int)c.w >= p.map_type_start) == ((int)pit[r] >= p.map_type_start));
            const unsigned mask = __ballot_sync(HTF_FULL, hit);
            if (mask) {                                      // warp-uniform
                const int nh = __popc(mask);
                const int rank = __popc(mask & lt);
                // slot wraps modulo K like htf/TensorflowCompute.cc:370; when one chunk holds
                // more than K hits only the last writer of a slot may store.
                if (hit && (rank + K >= nh)) {
                    int q = cnt[r] + rank;
                    if (q >= K) q %= K;
                    rowbuf[r * K + q] = make_float4(dx, dy, dz, c.w);
                    if (WITH_IDX) rowidx[r * K + q] = cj;
                }
                cnt[r] += nh;
            }
        }
    }
}

template <bool WITH_IDX>
__global__ void __launch_bounds__(128) nlist_build_kernel(const NlistParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int cell = blockIdx.x * wpb + warp;
    if (cell >= p.g.ncell) return;

    const int K = p.K;
    // per-warp carve-up: cand[cap] f4 | rowbuf[R*K] f4 | (candidx[cap] i32 | rowidx[R*K] i32)
    const size_t per_warp = (size_t)(p.cap + ROWS_PER_PASS * K) * (WITH_IDX ? 20 : 16);
    unsigned char *base = smem_raw + per_warp * warp;
    float4 *cand = reinterpret_cast<float4 *>(base);
    float4 *rowbuf = cand + p.cap;
    int *candidx = reinterpret_cast<int *>(rowbuf + ROWS_PER_PASS * K);
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
            for (int t = lo + lane; t < hi; t += 32) {
                const int s = qb + (t - qo);
                cand[t - w0] = __ldg(p.spos + s);
                if (WITH_IDX) candidx[t - w0] = __ldg(p.sorted_idx + s);
            }
        }
        __syncwarp();
        return w1 - w0;
    };

    int mstage = 0;
    if (npass == 1) mstage = stage(0);

    for (int s0 = b; s0 < e; s0 += ROWS_PER_PASS) {
        float pix[ROWS_PER_PASS], piy[ROWS_PER_PASS], piz[ROWS_PER_PASS], pit[ROWS_PER_PASS];
        int vself[ROWS_PER_PASS], orig[ROWS_PER_PASS], cnt[ROWS_PER_PASS];
        bool rvalid[ROWS_PER_PASS];
        bool anyrow = false;
#pragma unroll
        for (int r = 0; r < ROWS_PER_PASS; r++) {
            const int s = s0 + r;
            rvalid[r] = s < e;
            orig[r] = rvalid[r] ? __ldg(p.sorted_idx + s) : -1;
            rvalid[r] = rvalid[r] && orig[r] >= p.row_lo && orig[r] < p.row_hi;
            float4 pi = rvalid[r] ? __ldg(p.spos + s) : make_float4(0.f, 0.f, 0.f, 0.f);
            pix[r] = pi.x; piy[r] = pi.y; piz[r] = pi.z; pit[r] = pi.w;
            vself[r] = self_base + s;
            cnt[r] = 0;
            anyrow |= rvalid[r];
        }
        if (!anyrow) continue;

        for (int pass = 0; pass < npass; pass++) {
            if (npass > 1) { __syncwarp(); mstage = stage(pass); }
            if (wrap)
                test_chunks<true, WITH_IDX>(p, cand, candidx, mstage, pass * cap, pix, piy, piz, pit, vself,
                                            rvalid, cnt, rowbuf, rowidx, lane);
            else
                test_chunks<false, WITH_IDX>(p, cand, candidx, mstage, pass * cap, pix, piy, piz, pit, vself,
                                             rvalid, cnt, rowbuf, rowidx, lane);
        }
        __syncwarp();

        // ---- flush: one contiguous K*16 B store per row, zero padded ----
#pragma unroll
        for (int r = 0; r < ROWS_PER_PASS; r++) {
            if (!rvalid[r]) continue;
            const size_t row = (size_t)(orig[r] - p.row_lo);
            float4 *dst = p.out + row * K;
            for (int s = lane; s < K; s += 32) {
                float4 v = (s < cnt[r]) ? rowbuf[r * K + s] : make_float4(0.f, 0.f, 0.f, 0.f);
                dst[s] = v;
                if (WITH_IDX) p.idx_out[row * K + s] = (s < cnt[r]) ? rowidx[r * K + s] : -1;
            }
            if (lane == 0) {
                if (p.count_out) p.count_out[row] = cnt[r];
                if (p.overflow && cnt[r] >= K) atomicMax(p.overflow, cnt[r]);
            }
        }
        __syncwarp();
    }
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
    const int bytes_per = with_idx ? 20 : 16;
    int wpb = 4;
    const size_t smem_max = 200 * 1024;
    // keep the per-block footprint within the opt-in limit; shrink the staging window first
    // (the kernel re-stages in passes), then the block
    while ((size_t)(cap + ROWS_PER_PASS * p.K) * bytes_per * wpb > smem_max) {
        if (cap > 64) cap = max(64, cap / 2 / 32 * 32);
        else if (wpb > 1) wpb /= 2;
        else return cudaErrorInvalidValue;
    }
    p.cap = cap;
    const size_t smem = (size_t)(cap + ROWS_PER_PASS * p.K) * bytes_per * wpb;
    const int grid = (g.ncell + wpb - 1) / wpb;
    cudaError_t e;
    if (with_idx) {
        e = cudaFuncSetAttribute(nlist_build_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        nlist_build_kernel<true><<<grid, wpb * 32, smem, st>>>(p);
    } else {
        e = cudaFuncSetAttribute(nlist_build_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        nlist_build_kernel<false><<<grid, wpb * 32, smem, st>>>(p);
    }
    ctx->launches += 1;
    return cudaGetLastError();
}
