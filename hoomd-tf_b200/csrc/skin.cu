// Buffered ("skin") neighbor lists: the per-step half of the reference's own structure.
//
// In hoomd-tf the expensive search is HOOMD's NeighborList::compute, which runs only when particles have moved
// more than r_buff/2 (called at /root/reference htf/TensorflowCompute.cc:163); what runs EVERY
// step is prepareNeighbors (htf/TensorflowCompute.cc:304-374): walk the row's candidate indices, d = p_k - p_i,
// minimum image, skip if rsq > rc^2, write (dx, dy, dz, type) into the next slot, wrap the slot index modulo K.
// This file is that per-step pass on top of a candidate list built by the cell-list kernels with cutoff
// r_cut + skin (htf_skin_rebuild): one warp per row, lanes = candidates, positions gathered from L2 by original
// index, exact oracle arithmetic, ballot compaction into a shared row buffer, coalesced 16-byte stores with the
// zero padding.  HBM traffic per row: 4*KC bytes of candidate indices + 16 + 16K written (the position gathers hit
// the L2-resident 16 B/particle table); the kernel is HBM bound.
//
// A particle that has moved more than skin/2 since the rebuild makes the list unsafe; the kernel records it in a
// violation counter (htf_skin_status), the way HOOMD counts "dangerous builds".
#include "common.cuh"

namespace {

__device__ __forceinline__ float wrap_axis_s(float d, float lo, float hi, float L)
{
    const float adj = (d >= hi) ? L : ((d < lo) ? -L : 0.0f);
    return __fsub_rn(d, adj);
}

struct FilterParams {
    const float4 *pos;      // all particles (original order)
    const float4 *ref;      // positions at the last rebuild
    const int *cand;        // [rows][KC] candidate indices, -1 padded
    const int *cand_count;  // [rows] candidates found by the rebuild (may exceed KC: overflow of the candidate list)
    int row_lo, rows, KC, K;
    float rc2, half_skin2;
    float half[3], L[3];
    int map_type_start;
    float4 *out;
    int *idx_out;
    int *count_out;
    int *overflow;
    int *status;            // [0] rows whose own particle moved more than skin/2, [1] rows whose candidate list overflowed
};

template <bool WITH_IDX, bool MAPPED>
__global__ void __launch_bounds__(256) nlist_filter_kernel(const FilterParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = p.K;
    float4 *rowbuf = reinterpret_cast<float4 *>(smem_raw) + (size_t)warp * K;
    int *idxbuf = reinterpret_cast<int *>(reinterpret_cast<float4 *>(smem_raw) + (size_t)8 * K) + (size_t)warp * K;
    const unsigned lt = (1u << lane) - 1u;
    for (int row = blockIdx.x * 8 + warp; row < p.rows; row += gridDim.x * 8) {
        const int i = p.row_lo + row;
        const float4 pi = __ldg(p.pos + i);
        const int nc_all = __ldg(p.cand_count + row);
        const int nc = min(nc_all, p.KC);
        if (lane == 0) {
            const float4 r0 = __ldg(p.ref + i);
            float ex = wrap_axis_s(__fsub_rn(pi.x, r0.x), -p.half[0], p.half[0], p.L[0]);
            float ey = wrap_axis_s(__fsub_rn(pi.y, r0.y), -p.half[1], p.half[1], p.L[1]);
            float ez = wrap_axis_s(__fsub_rn(pi.z, r0.z), -p.half[2], p.half[2], p.L[2]);
            if (ex * ex + ey * ey + ez * ez > p.half_skin2) atomicAdd(p.status, 1);
            if (nc_all > p.KC) atomicAdd(p.status + 1, 1);
        }
        const int *crow = p.cand + (size_t)row * p.KC;
        int running = 0;
        for (int t0 = 0; t0 < nc; t0 += 32) {
            const int c = t0 + lane;
            const bool valid = c < nc;
            const int j = valid ? __ldg(crow + c) : i;
            const float4 pj = __ldg(p.pos + j);
            float dx = __fsub_rn(pj.x, pi.x), dy = __fsub_rn(pj.y, pi.y), dz = __fsub_rn(pj.z, pi.z);
            dz = wrap_axis_s(dz, -p.half[2], p.half[2], p.L[2]);
            dy = wrap_axis_s(dy, -p.half[1], p.half[1], p.L[1]);
            dx = wrap_axis_s(dx, -p.half[0], p.half[0], p.L[0]);
            const float rsq = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            bool keep = valid && (rsq <= p.rc2);
            if (MAPPED) keep = keep && (((int)pj.w >= p.map_type_start) == ((int)pi.w >= p.map_type_start));
            const unsigned m = __ballot_sync(HTF_FULL, keep);
            if (keep) {
                int slot = running + __popc(m & lt);
                if (slot >= K) slot %= K;                       // htf/TensorflowCompute.cc:370, the last writer of a slot wins
                rowbuf[slot] = make_float4(dx, dy, dz, pj.w);
                if (WITH_IDX) idxbuf[slot] = j;
            }
            running += __popc(m);
        }
        __syncwarp();
        const int nvalid = min(running, K);
        float4 *dst = p.out + (size_t)row * K;
        for (int sl = lane; sl < K; sl += 32) {
            dst[sl] = sl < nvalid ? rowbuf[sl] : make_float4(0.f, 0.f, 0.f, 0.f);
            if (WITH_IDX) p.idx_out[(size_t)row * K + sl] = sl < nvalid ? idxbuf[sl] : -1;
        }
        if (lane == 0) {
            if (p.count_out) p.count_out[row] = running;
            if (running >= K && p.overflow) atomicMax(p.overflow, running);
        }
        __syncwarp();
    }
}

}  // namespace

cudaError_t htf_launch_skin_filter(htf_ctx *ctx, const float4 *pos, int64_t row_lo, int64_t row_hi, float4 *out,
                                   int32_t *idx_out, int32_t *count_out, int32_t *overflow, cudaStream_t st)
{
    const int rows = (int)(row_hi - row_lo);
    if (rows <= 0) return cudaSuccess;
    FilterParams p;
    p.pos = pos; p.ref = ctx->d_skin_ref; p.cand = ctx->d_skin_cand; p.cand_count = ctx->d_skin_count;
    p.row_lo = (int)row_lo; p.rows = rows; p.KC = ctx->skin_kc; p.K = ctx->K;
    p.rc2 = ctx->r_cut * ctx->r_cut;
    p.half_skin2 = 0.25f * ctx->skin * ctx->skin;
    for (int a = 0; a < 3; a++) { p.half[a] = ctx->grid.half[a]; p.L[a] = ctx->grid.L[a]; }
    p.map_type_start = ctx->map_type_start;
    p.out = out; p.idx_out = idx_out; p.count_out = count_out; p.overflow = overflow;
    p.status = ctx->d_stats + 6;
    const bool with_idx = idx_out != nullptr, mapped = ctx->map_type_start >= 0;
    const size_t smem = (size_t)8 * p.K * 16 + (with_idx ? (size_t)8 * p.K * 4 : 0);
    if (smem > 48 * 1024) return cudaErrorInvalidValue;        // K <= 384 (307 with indices)
    int grid = (rows + 7) / 8;
    const int gmax = 32 * ctx->sm_count;
    if (grid > gmax) grid = gmax;
    if (with_idx) {
        if (mapped) nlist_filter_kernel<true, true><<<grid, 256, smem, st>>>(p);
        else nlist_filter_kernel<true, false><<<grid, 256, smem, st>>>(p);
    } else {
        if (mapped) nlist_filter_kernel<false, true><<<grid, 256, smem, st>>>(p);
        else nlist_filter_kernel<false, false><<<grid, 256, smem, st>>>(p);
    }
    ctx->launches += 1;
    return cudaGetLastError();
}
