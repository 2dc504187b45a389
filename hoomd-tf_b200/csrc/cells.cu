// Cell binning for the neighbor build: replaces the candidate search of HOOMD's
// NeighborList::compute (called at /root/reference htf/TensorflowCompute.cc:163).
//
// Pipeline (all HBM-bound, ~56 B/particle):
//   count   : cell id per particle, atomic count per cell
//   scan    : exclusive prefix sum of the counts (3 small kernels)
//   scatter : particle index -> slot of its cell (count-down atomics, no second zeroing)
//   order+gather : one warp per cell: (deterministic mode) bitonic sort of the cell's indices,
//             then the cell-sorted copy of the positions, the only array the build kernel reads
#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int cell_coord(float p, float lo, float inv_w, int n)
{
    float u = __fmul_rn(__fsub_rn(p, lo), inv_w);
    int c = (int)floorf(u);
    return max(0, min(n - 1, c));      // particles on/over the box faces go to the edge cells
}

__global__ void __launch_bounds__(256) cell_count_kernel(const float4 *__restrict__ pos, int n, CellGrid g,
                                                         int *__restrict__ cell_of, int *__restrict__ cell_cnt)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = __ldg(pos + i);
    {   // region of interest (sharded builds): particles that cannot be within r_cut of a local row are skipped
        const float q[3] = {p.x, p.y, p.z};
        bool in = true;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            if (g.roi_h[a] >= 0.0f) {
                float d = q[a] - g.roi_c[a];
                d = d >= g.half[a] ? d - g.L[a] : (d < -g.half[a] ? d + g.L[a] : d);
                in = in && fabsf(d) <= g.roi_h[a];
            }
        }
        if (!in) { cell_of[i] = -1; return; }
    }
    int cx = cell_coord(p.x, g.lo[0], g.inv_w[0], g.n[0]);
    int cy = cell_coord(p.y, g.lo[1], g.inv_w[1], g.n[1]);
    int cz = cell_coord(p.z, g.lo[2], g.inv_w[2], g.n[2]);
    int c = (cz * g.n[1] + cy) * g.n[0] + cx;
    cell_of[i] = c;
    atomicAdd(cell_cnt + c, 1);
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(HTF_FULL, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// exclusive scan of one tile held as SCAN_ITEMS per thread; returns the tile total
__device__ __forceinline__ int block_excl_scan(int (&v)[SCAN_ITEMS], int *warp_tot /* smem[8] */)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) s += v[k];
    int incl = warp_incl_scan(s, lane);
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int q = 0; q < SCAN_THREADS / 32; q++) {
        int t = warp_tot[q];
        if (q < w) base += t;
        total += t;
    }
    int run = base + incl - s;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { int t = v[k]; v[k] = run; run += t; }
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums_kernel(const int *__restrict__ in, int n,
                                                                      int *__restrict__ tile_sums)
{
    __shared__ int warp_tot[SCAN_THREADS / 32];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) v[k] = (base + k < n) ? in[base + k] : 0;
    int total = block_excl_scan(v, warp_tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// one block: exclusive scan of the tile sums in place, total -> *grand_total
__global__ void __launch_bounds__(SCAN_THREADS) scan_top_kernel(int *__restrict__ tile_sums, int ntiles,
                                                                int *__restrict__ grand_total)
{
    __shared__ int warp_tot[SCAN_THREADS / 32];
    int carry = 0;
    for (int t0 = 0; t0 < ntiles; t0 += SCAN_TILE) {
        const int base = t0 + threadIdx.x * SCAN_ITEMS;
        int v[SCAN_ITEMS];
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) v[k] = (base + k < ntiles) ? tile_sums[base + k] : 0;
        int total = block_excl_scan(v, warp_tot);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++)
            if (base + k < ntiles) tile_sums[base + k] = v[k] + carry;
        carry += total;
    }
    if (threadIdx.x == 0) *grand_total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const int *__restrict__ in, int n,
                                                                  const int *__restrict__ tile_offs,
                                                                  int *__restrict__ out)
{
    __shared__ int warp_tot[SCAN_THREADS / 32];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) v[k] = (base + k < n) ? in[base + k] : 0;
    block_excl_scan(v, warp_tot);
    const int off = tile_offs[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n) out[base + k] = v[k] + off;
}

// count-down scatter: cell_cnt[c] still holds the population from the count pass
__global__ void __launch_bounds__(256) cell_scatter_kernel(const int *__restrict__ cell_of, int n,
                                                           const int *__restrict__ cell_start,
                                                           int *__restrict__ cell_cnt, int *__restrict__ sorted_idx)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = cell_of[i];
    if (c < 0) return;                      // outside the region of interest
    int k = atomicSub(cell_cnt + c, 1) - 1;
    sorted_idx[cell_start[c] + k] = i;
}

// One warp per cell: (deterministic mode) bitonic-sort the cell's particle indices across the
// lanes, then gather the positions into the cell-sorted copy.  The count-down scatter leaves the
// indices of a cell in atomic order; sorting them makes the row-internal slot order of the
// neighbor tensor -- and therefore every fp32 force sum -- reproducible run to run.
template <bool SORT, int LPC>
__global__ void __launch_bounds__(256) cell_order_gather_kernel(const float4 *__restrict__ pos,
                                                                const int *__restrict__ cell_start, int ncell_win,
                                                                int layer, int z0, int nz,
                                                                const int *__restrict__ scattered,
                                                                int *__restrict__ sorted_idx,
                                                                float4 *__restrict__ spos)
{
    // LPC lanes per cell (a warp serves 32 / LPC cells).  Out of place: `scattered` holds the cell's members in
    // the order the atomics of the scatter produced; the rank of a member is the number of members with a
    // smaller index (indices are distinct), counted with broadcast loads that hit L1 -- no shuffles, no
    // intra-warp hand-shake, any cell population.
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int cw = gt / LPC, sub = gt % LPC;                                // cell inside the z-window
    if (cw >= ncell_win) return;
    const int lz = cw / layer;
    const int c = ((z0 + lz) % nz) * layer + (cw - lz * layer);
    const int b = __ldg(cell_start + c), n = __ldg(cell_start + c + 1) - b;
    for (int k = sub; k < n; k += LPC) {
        const int v = __ldg(scattered + b + k);
        int dst = k;
        if (SORT) {
            dst = 0;
            for (int j = 0; j < n; j++) dst += (__ldg(scattered + b + j) < v) ? 1 : 0;
        }
        sorted_idx[b + dst] = v;
        spos[b + dst] = __ldg(pos + v);
    }
}

// ---- cell population statistics (calibrates the staging capacities of the build kernels) ----
// stats[0] = particles binned, stats[1] = occupied cells, stats[2] = largest cell population
__global__ void __launch_bounds__(256) cell_stats_kernel(const int *__restrict__ cell_start, int ncell,
                                                         int *__restrict__ stats)
{
    int occ = 0, mx = 0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += gridDim.x * blockDim.x) {
        const int n = cell_start[c + 1] - cell_start[c];
        occ += n > 0;
        mx = max(mx, n);
    }
    occ = __reduce_add_sync(HTF_FULL, occ);
    mx = __reduce_max_sync(HTF_FULL, mx);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(stats + 1, occ);
        atomicMax(stats + 2, mx);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) stats[0] = cell_start[ncell];
}

// ---- stable selection of the particles beyond a plane (halo packing for the slab exchange) ----
// out[k] = k-th particle (in index order) with pos[axis] < thr (LESS) or > thr; the rest of out[0..cap)
// is filled with a far-away sentinel that the region-of-interest test of the binning rejects.
template <bool LESS>
__device__ __forceinline__ bool beyond(const float4 &p, int axis, float thr)
{
    const float v = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
    return LESS ? v < thr : v > thr;
}

template <bool LESS>
__global__ void __launch_bounds__(256) select_count_kernel(const float4 *__restrict__ pos, int n, int axis, float thr,
                                                           int *__restrict__ block_cnt)
{
    __shared__ int wsum[8];
    const int i = blockIdx.x * 256 + threadIdx.x;
    const bool sel = i < n && beyond<LESS>(__ldg(pos + i), axis, thr);
    const unsigned m = __ballot_sync(HTF_FULL, sel);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) t += wsum[w];
        block_cnt[blockIdx.x] = t;
    }
}

template <bool LESS>
__global__ void __launch_bounds__(256) select_scatter_kernel(const float4 *__restrict__ pos, int n, int axis, float thr,
                                                             const int *__restrict__ block_off, float4 *__restrict__ out,
                                                             int cap, int *__restrict__ overflow)
{
    __shared__ int wsum[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i = blockIdx.x * 256 + threadIdx.x;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) p = __ldg(pos + i);
    const bool sel = i < n && beyond<LESS>(p, axis, thr);
    const unsigned m = __ballot_sync(HTF_FULL, sel);
    if (lane == 0) wsum[w] = __popc(m);
    __syncthreads();
    int base = block_off[blockIdx.x];
    for (int q = 0; q < w; q++) base += wsum[q];
    const int k = base + __popc(m & ((1u << lane) - 1u));
    if (sel) {
        if (k < cap) out[k] = p;
        else if (overflow) atomicMax(overflow, k + 1);
    }
}

// ---- both slab faces in one pass: block_cnt[b] = particles of block b below thr_lo, block_cnt[nb + b] = above thr_hi ----
__global__ void __launch_bounds__(256) select_count2_kernel(const float4 *__restrict__ pos, int n, int axis, float thr_lo,
                                                            float thr_hi, int nb, int *__restrict__ block_cnt)
{
    __shared__ int wsum[16];
    const int i = blockIdx.x * 256 + threadIdx.x;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) p = __ldg(pos + i);
    const unsigned ml = __ballot_sync(HTF_FULL, i < n && beyond<true>(p, axis, thr_lo));
    const unsigned mh = __ballot_sync(HTF_FULL, i < n && beyond<false>(p, axis, thr_hi));
    if ((threadIdx.x & 31) == 0) { wsum[threadIdx.x >> 5] = __popc(ml); wsum[8 + (threadIdx.x >> 5)] = __popc(mh); }
    __syncthreads();
    if (threadIdx.x < 2) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) t += wsum[8 * threadIdx.x + w];
        block_cnt[threadIdx.x * nb + blockIdx.x] = t;
    }
    if (blockIdx.x == 0 && threadIdx.x == 2) block_cnt[2 * nb] = 0;      // the extra entry that makes off[2 nb] exist
}

// exclusive scan of a SMALL array (n <= 1024 * SCAN1_ITEMS) by one block, in[] -> out[], grand total -> *total:
// one launch instead of three for the halo compaction's per-block counts
constexpr int SCAN1_ITEMS = 16;
__global__ void __launch_bounds__(1024) scan_one_block_kernel(const int *__restrict__ in, int n, int *__restrict__ out,
                                                              int *__restrict__ total)
{
    __shared__ int wtot[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int per = (n + 1023) / 1024;                        // contiguous items per thread
    const int b = threadIdx.x * per;
    int v[SCAN1_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN1_ITEMS; k++) {
        v[k] = (k < per && b + k < n) ? in[b + k] : 0;
        s += v[k];
    }
    int incl = warp_incl_scan(s, lane);
    if (lane == 31) wtot[w] = incl;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int q = 0; q < 32; q++) {
        const int t = wtot[q];
        if (q < w) base += t;
        tot += t;
    }
    int run = base + incl - s;
#pragma unroll
    for (int k = 0; k < SCAN1_ITEMS; k++) {
        if (k < per && b + k < n) out[b + k] = run;
        run += v[k];
    }
    if (threadIdx.x == 0) *total = tot;
}

// block_off = exclusive scan of block_cnt[2 nb]; *total = its grand total.  Threads past n only pad: every entry of
// the two buffers beyond its face's count gets the sentinel, so no separate fill pass is needed.
__global__ void __launch_bounds__(256) select_scatter2_kernel(const float4 *__restrict__ pos, int n, int axis, float thr_lo,
                                                              float thr_hi, int nb, const int *__restrict__ block_off,
                                                              const int *__restrict__ total, float4 *__restrict__ out_lo,
                                                              float4 *__restrict__ out_hi, int cap, int *__restrict__ counts,
                                                              int *__restrict__ overflow, const HtfHaloDst *dst)
{
    __shared__ int wsum[16];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (dst) {
        // fused pack + send: the two faces go straight into the neighbours' receive buffers (peer memory over
        // NVLink), double buffered by the parity of the exchange epoch, which lives in device memory so that a
        // captured graph advances it by itself
        const int par = (int)(*reinterpret_cast<const volatile unsigned long long *>(&dst->epoch) & 1ull);
        out_lo = dst->lo[par];
        out_hi = dst->hi[par];
    }
    const int tot_lo = block_off[nb], tot_hi = *total - tot_lo;
    if (i == 0 && counts) { counts[0] = tot_lo; counts[1] = tot_hi; }
    if (i < cap) {
        const float4 far = make_float4(1e30f, 1e30f, 1e30f, 0.f);
        if (i >= tot_lo) out_lo[i] = far;
        if (i >= tot_hi) out_hi[i] = far;
    }
    if (blockIdx.x >= nb) return;                               // padding-only blocks (cap > n)
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) p = __ldg(pos + i);
    const bool sl = i < n && beyond<true>(p, axis, thr_lo), sh = i < n && beyond<false>(p, axis, thr_hi);
    const unsigned ml = __ballot_sync(HTF_FULL, sl), mh = __ballot_sync(HTF_FULL, sh);
    if (lane == 0) { wsum[w] = __popc(ml); wsum[8 + w] = __popc(mh); }
    __syncthreads();
    int bl = block_off[blockIdx.x], bh = block_off[nb + blockIdx.x] - tot_lo;
    for (int q = 0; q < w; q++) { bl += wsum[q]; bh += wsum[8 + q]; }
    const unsigned below = (1u << lane) - 1u;
    if (sl) {
        const int k = bl + __popc(ml & below);
        if (k < cap) out_lo[k] = p;
        else if (overflow) atomicMax(overflow, k + 1);
    }
    if (sh) {
        const int k = bh + __popc(mh & below);
        if (k < cap) out_hi[k] = p;
        else if (overflow) atomicMax(overflow, k + 1);
    }
}

// ---- the same compaction of both faces as ONE launch (instead of count + scan + scatter) ----
// A grid of 2 blocks per SM, all co-resident.  Warp w of block b owns the contiguous particles
// [(8 b + w) * items * 32, +items * 32) and keeps them IN REGISTERS (items <= IMAX, all loads of a thread in flight at
// once; a first version that re-read the selected particles iteration by iteration took 17 us because the blocks at
// the two ends of a spatially sorted slab hold nothing but face particles).  A block publishes its two counts as ONE
// 64-bit word tagged with the launch epoch (so the slots never have to be cleared), its threads read the words of all
// blocks (spinning on the tag: the launch is COOPERATIVE, so every block of the grid is resident, and a block publishes
// before it waits), and the block stores its selected particles at the prefix.  Order = index order, exactly as the
// three-kernel form.
template <int IMAX>
__global__ void __launch_bounds__(256, 2) select_fused2_kernel(const float4 *__restrict__ pos, int n, int axis, float thr_lo,
                                                               float thr_hi, int items, unsigned long long *slots,
                                                               unsigned *ctl /* [0] epoch, [1] blocks done */,
                                                               float4 *__restrict__ out_lo, float4 *__restrict__ out_hi, int cap,
                                                               int *__restrict__ counts, int *__restrict__ overflow,
                                                               const HtfHaloDst *dst)
{
    __shared__ int wsum[16];
    __shared__ int red[4];
    __shared__ int red4[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned e = *reinterpret_cast<volatile unsigned *>(&ctl[0]);
    const unsigned tag = e + 1u;
    if (dst) {
        const int par = (int)(*reinterpret_cast<const volatile unsigned long long *>(&dst->epoch) & 1ull);
        out_lo = dst->lo[par];
        out_hi = dst->hi[par];
    }
    const long long base = ((long long)blockIdx.x * 8 + w) * items * 32 + lane;
    float4 p[IMAX];
#pragma unroll
    for (int it = 0; it < IMAX; it++) {
        const long long i = base + it * 32;
        // a particle that is in neither face (coordinate 0 lies between the thresholds only by luck: use a flag value)
        p[it] = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7fc00001));
        if (it < items && i < n) p[it] = __ldg(pos + i);
    }
    int cl = 0, ch = 0;
#pragma unroll
    for (int it = 0; it < IMAX; it++) {
        const bool valid = __float_as_int(p[it].w) != 0x7fc00001;
        const float v = axis == 0 ? p[it].x : (axis == 1 ? p[it].y : p[it].z);
        cl += __popc(__ballot_sync(HTF_FULL, valid && v < thr_lo));
        ch += __popc(__ballot_sync(HTF_FULL, valid && v > thr_hi));
    }
    if (lane == 0) { wsum[w] = cl; wsum[8 + w] = ch; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int bl = 0, bh = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) { bl += wsum[q]; bh += wsum[8 + q]; }
        *reinterpret_cast<volatile unsigned long long *>(&slots[blockIdx.x]) =
            ((unsigned long long)tag << 32) | ((unsigned long long)(unsigned)bh << 16) | (unsigned long long)(unsigned)bl;
    }
    // counts of every block: prefix over the lower blocks and the grand totals.  Every thread polls at most
    // ceil(grid / 256) slots, so the block pays one or two L2 round trips (one warp walking all slots: ten)
    {
        int pl = 0, ph = 0, tl = 0, th = 0;
        for (int t = threadIdx.x; t < (int)gridDim.x; t += 256) {
            unsigned long long v;
            do {
                v = *reinterpret_cast<volatile unsigned long long *>(&slots[t]);
            } while ((unsigned)(v >> 32) != tag);
            const int a = (int)(v & 0xffffull), b = (int)((v >> 16) & 0xffffull);
            tl += a; th += b;
            if (t < (int)blockIdx.x) { pl += a; ph += b; }
        }
        pl = __reduce_add_sync(HTF_FULL, pl); ph = __reduce_add_sync(HTF_FULL, ph);
        tl = __reduce_add_sync(HTF_FULL, tl); th = __reduce_add_sync(HTF_FULL, th);
        if (lane == 0) { red4[4 * w] = pl; red4[4 * w + 1] = ph; red4[4 * w + 2] = tl; red4[4 * w + 3] = th; }
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        int t = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) t += red4[4 * q + threadIdx.x];
        red[threadIdx.x] = t;
    }
    __syncthreads();
    const int tl = red[2], th = red[3];
    int ol = red[0], oh = red[1];
    if (threadIdx.x == 0) {
        // every block reads the epoch before it gets here: the last one advances it for the next launch
        __threadfence();
        if (atomicAdd(&ctl[1], 1u) == gridDim.x - 1u) {
            ctl[1] = 0u;
            __threadfence();
            ctl[0] = tag;
        }
        if (blockIdx.x == 0 && counts) { counts[0] = tl; counts[1] = th; }
    }
    for (int q = 0; q < w; q++) { ol += wsum[q]; oh += wsum[8 + q]; }
    const unsigned below = (1u << lane) - 1u;
#pragma unroll
    for (int it = 0; it < IMAX; it++) {
        const bool valid = __float_as_int(p[it].w) != 0x7fc00001;
        const float v = axis == 0 ? p[it].x : (axis == 1 ? p[it].y : p[it].z);
        const bool sl = valid && v < thr_lo, sh = valid && v > thr_hi;
        const unsigned ml = __ballot_sync(HTF_FULL, sl), mh = __ballot_sync(HTF_FULL, sh);
        if (sl) {
            const int k = ol + __popc(ml & below);
            if (k < cap) out_lo[k] = p[it];
            else if (overflow) atomicMax(overflow, k + 1);
        }
        if (sh) {
            const int k = oh + __popc(mh & below);
            if (k < cap) out_hi[k] = p[it];
            else if (overflow) atomicMax(overflow, k + 1);
        }
        ol += __popc(ml);
        oh += __popc(mh);
    }
    // sentinel padding behind the two faces
    const float4 far = make_float4(1e30f, 1e30f, 1e30f, 0.f);
    const int gt = blockIdx.x * 256 + threadIdx.x, gs = gridDim.x * 256;
    for (int i = tl + gt; i < cap; i += gs) out_lo[i] = far;
    for (int i = th + gt; i < cap; i += gs) out_hi[i] = far;
}

__global__ void __launch_bounds__(256) fill_sentinel_kernel(float4 *__restrict__ out, int cap)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < cap) out[i] = make_float4(1e30f, 1e30f, 1e30f, 0.f);
}

}  // namespace

cudaError_t htf_launch_binning(htf_ctx *ctx, const float4 *pos, int64_t n64, cudaStream_t st)
{
    const int n = (int)n64;
    const CellGrid &g = ctx->grid;
    const int ncell = g.ncell;
    cudaError_t e = cudaMemsetAsync(ctx->d_cell_cnt, 0, sizeof(int) * (size_t)ncell, st);
    if (e != cudaSuccess) return e;
    if (n == 0) {
        e = cudaMemsetAsync(ctx->d_cell_start, 0, sizeof(int) * ((size_t)ncell + 1), st);
        return e;
    }
    const int pb = (n + 255) / 256;
    cell_count_kernel<<<pb, 256, 0, st>>>(pos, n, g, ctx->d_cell_of, ctx->d_cell_cnt);
    int scan_launches = 3;
    if (ncell <= 1024 * SCAN1_ITEMS) {
        // small grids (launch-bound systems): one block scans all the counts
        scan_one_block_kernel<<<1, 1024, 0, st>>>(ctx->d_cell_cnt, ncell, ctx->d_cell_start, ctx->d_cell_start + ncell);
        scan_launches = 1;
    } else {
        const int ntiles = (ncell + SCAN_TILE - 1) / SCAN_TILE;
        scan_tile_sums_kernel<<<ntiles, SCAN_THREADS, 0, st>>>(ctx->d_cell_cnt, ncell, ctx->d_block_sums);
        scan_top_kernel<<<1, SCAN_THREADS, 0, st>>>(ctx->d_block_sums, ntiles, ctx->d_cell_start + ncell);
        scan_apply_kernel<<<ntiles, SCAN_THREADS, 0, st>>>(ctx->d_cell_cnt, ncell, ctx->d_block_sums, ctx->d_cell_start);
    }
    cell_scatter_kernel<<<pb, 256, 0, st>>>(ctx->d_cell_of, n, ctx->d_cell_start, ctx->d_cell_cnt, ctx->d_scattered);
    const int layer = g.n[0] * g.n[1];
    const int ncell_win = layer * g.zcount;                     // only the cell layers the region of interest touches
#ifndef HTF_LPC
#define HTF_LPC 8
#endif
    constexpr int LPC = HTF_LPC;
    const int cb = (int)(((long long)ncell_win * LPC + 255) / 256);
    if (ctx->flags & 1 /* HTF_FLAG_DETERMINISTIC */)
        cell_order_gather_kernel<true, LPC><<<cb, 256, 0, st>>>(pos, ctx->d_cell_start, ncell_win, layer, g.z0, g.n[2],
                                                                ctx->d_scattered, ctx->d_sorted_idx, ctx->d_spos);
    else
        cell_order_gather_kernel<false, LPC><<<cb, 256, 0, st>>>(pos, ctx->d_cell_start, ncell_win, layer, g.z0, g.n[2],
                                                                 ctx->d_scattered, ctx->d_sorted_idx, ctx->d_spos);
    ctx->launches += 3 + scan_launches;
    return cudaGetLastError();
}


cudaError_t htf_launch_select(htf_ctx *ctx, const float4 *pos, int64_t n64, int axis, float thr, bool less,
                              float4 *out, int cap, int *d_count, int *d_overflow, cudaStream_t st)
{
    const int n = (int)n64;
    fill_sentinel_kernel<<<(cap + 255) / 256, 256, 0, st>>>(out, cap);
    ctx->launches += 1;
    if (n == 0) return d_count ? cudaMemsetAsync(d_count, 0, sizeof(int), st) : cudaGetLastError();
    const int nb = (n + 255) / 256;
    int *cnt = ctx->d_sel_cnt, *off = ctx->d_sel_off, *sums = ctx->d_sel_sums;
    if (less) select_count_kernel<true><<<nb, 256, 0, st>>>(pos, n, axis, thr, cnt);
    else select_count_kernel<false><<<nb, 256, 0, st>>>(pos, n, axis, thr, cnt);
    const int ntiles = (nb + SCAN_TILE - 1) / SCAN_TILE;
    scan_tile_sums_kernel<<<ntiles, SCAN_THREADS, 0, st>>>(cnt, nb, sums);
    scan_top_kernel<<<1, SCAN_THREADS, 0, st>>>(sums, ntiles, d_count ? d_count : sums + ntiles + 1);
    scan_apply_kernel<<<ntiles, SCAN_THREADS, 0, st>>>(cnt, nb, sums, off);
    if (less) select_scatter_kernel<true><<<nb, 256, 0, st>>>(pos, n, axis, thr, off, out, cap, d_overflow);
    else select_scatter_kernel<false><<<nb, 256, 0, st>>>(pos, n, axis, thr, off, out, cap, d_overflow);
    ctx->launches += 5;
    return cudaGetLastError();
}


cudaError_t htf_launch_select_pair(htf_ctx *ctx, const float4 *pos, int64_t n64, int axis, float thr_lo, float thr_hi,
                                   float4 *out_lo, float4 *out_hi, int cap, int *d_counts, int *d_overflow, cudaStream_t st,
                                   const HtfHaloDst *dst)
{
    const int n = (int)n64;
    const int nb = (n + 255) / 256;
    {
        // One launch instead of three (exchange 45.5 -> 35.0 us per step at 2 x 1 M rows).  Its blocks wait for each other,
        // so it goes out as a COOPERATIVE launch: the driver either places all blocks together or refuses, and a refusal
        // (SM-limited context, another resident kernel) falls back to the three-kernel form for good.
        // HTF_SELECT_FUSED=0 selects the three-kernel form.
        const char *fe = getenv("HTF_SELECT_FUSED");
        const bool fused_on = !(fe && fe[0] == '0') && !ctx->sel_fused_refused;
        const int g = max(1, min(min(2 * ctx->sm_count, HTF_SEL_MAX_BLOCKS), nb));
        const int items = (int)((n64 + (int64_t)g * 256 - 1) / ((int64_t)g * 256));
        if (fused_on && ctx->d_sel_slots && items <= 24) {
            unsigned *ctl = reinterpret_cast<unsigned *>(ctx->d_sel_slots + HTF_SEL_MAX_BLOCKS);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)g); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeCooperative;
            attr[0].val.cooperative = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            unsigned long long *slots = ctx->d_sel_slots;
            cudaError_t e;
            if (items <= 8)
                e = cudaLaunchKernelEx(&cfg, select_fused2_kernel<8>, pos, n, axis, thr_lo, thr_hi, items, slots, ctl, out_lo,
                                       out_hi, cap, d_counts, d_overflow, dst);
            else if (items <= 16)
                e = cudaLaunchKernelEx(&cfg, select_fused2_kernel<16>, pos, n, axis, thr_lo, thr_hi, items, slots, ctl, out_lo,
                                       out_hi, cap, d_counts, d_overflow, dst);
            else
                e = cudaLaunchKernelEx(&cfg, select_fused2_kernel<24>, pos, n, axis, thr_lo, thr_hi, items, slots, ctl, out_lo,
                                       out_hi, cap, d_counts, d_overflow, dst);
            if (e == cudaSuccess) {
                ctx->launches += 1;
                return cudaSuccess;
            }
            cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
            cudaStreamIsCapturing(st, &cs);
            if (cs != cudaStreamCaptureStatusNone) return e;          // a failed launch has already broken the capture
            (void)cudaGetLastError();
            ctx->sel_fused_refused = true;
        }
    }
    int *cnt = ctx->d_sel_cnt, *off = ctx->d_sel_off, *sums = ctx->d_sel_sums;
    if (nb > 0) select_count2_kernel<<<nb, 256, 0, st>>>(pos, n, axis, thr_lo, thr_hi, nb, cnt);
    const int m = 2 * nb + 1;                                   // one extra (zero) entry so that off[2 nb] exists
    if (nb == 0) {
        cudaError_t e = cudaMemsetAsync(cnt, 0, sizeof(int), st);
        if (e != cudaSuccess) return e;
    }
    const int ntiles = (m + SCAN_TILE - 1) / SCAN_TILE;
    int *total = sums + ntiles + 1;
    if (m <= 1024 * SCAN1_ITEMS) {
        scan_one_block_kernel<<<1, 1024, 0, st>>>(cnt, m, off, total);
        ctx->launches += 1;
    } else {
        scan_tile_sums_kernel<<<ntiles, SCAN_THREADS, 0, st>>>(cnt, m, sums);
        scan_top_kernel<<<1, SCAN_THREADS, 0, st>>>(sums, ntiles, total);
        scan_apply_kernel<<<ntiles, SCAN_THREADS, 0, st>>>(cnt, m, sums, off);
        ctx->launches += 3;
    }
    const int gb = max(nb, (cap + 255) / 256);
    select_scatter2_kernel<<<gb, 256, 0, st>>>(pos, n, axis, thr_lo, thr_hi, nb, off, total, out_lo, out_hi, cap, d_counts,
                                               d_overflow, dst);
    ctx->launches += 2;
    return cudaGetLastError();
}

// Synchronising (one cudaMemcpy): called by the nlist launcher only when the context has no valid
// calibration for the current box / cutoff / region of interest / particle count.
cudaError_t htf_cell_stats(htf_ctx *ctx, int h_stats[3], cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(ctx->d_stats, 0, 3 * sizeof(int), st);
    if (e != cudaSuccess) return e;
    const int ncell = ctx->grid.ncell;
    cell_stats_kernel<<<min((ncell + 255) / 256, 592), 256, 0, st>>>(ctx->d_cell_start, ncell, ctx->d_stats);
    ctx->launches += 1;
    e = cudaMemcpyAsync(h_stats, ctx->d_stats, 3 * sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(st);
}
