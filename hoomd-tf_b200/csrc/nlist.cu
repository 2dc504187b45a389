// Padded neighbor tensor [rows, K, 4] from the cell-sorted positions.
//
// Replaces prepareNeighbors (/root/reference htf/TensorflowCompute.cc:304-374, CPU rule
// "skip if rsq > rc^2"), its GPU twin htf_gpu_reshape_nlist_kernel
// (htf/TensorflowCompute.cu:80-151: one thread per row, 16-byte stores strided by 16*K)
// and the cudaMemset before it (.cu:180).
//
// Work decomposition: one warp per cell.
//   stage : the positions of the 3x3x3 stencil (9 contiguous runs of the cell-sorted array
//           when the x-neighbours do not wrap) are copied to shared memory with cp.async;
//   test  : the rows of the cell are taken four at a time; every lane holds one candidate
//           per 32-candidate chunk, tests it against the four rows and, on a hit, appends the
//           candidate's 16-bit index to its own (lane-private) list of that row -- one
//           predicated store and one predicated add, no ballot / popc / branch;
//   emit  : per row, an exclusive scan of the lane counts gives every lane its slot range; the
//           lanes publish "slot -> candidate" in a small shared map, then lane s re-derives d
//           for slot s, s+32, ... and the row leaves the SM as contiguous coalesced 16-byte
//           stores, zero padding included -- no memset pass, no partial-sector traffic.
//
// Arithmetic is the oracle's, bit for bit: d = p_j - p_i, compare-and-shift minimum image
// (HOOMD BoxDim::minImage CPU branch), rsq = (dx*dx + dy*dy) + dz*dz with round-to-nearest
// mul/add that are never contracted to FMA, keep iff rsq <= rc^2.  The subtractions and the
// squares use Blackwell's packed fp32 pipe (add/mul .f32x2: two rows per instruction, each
// half IEEE-rounded like the scalar op); the sums stay scalar because ptxas contracts a
// packed mul feeding a packed add into FFMA2 even with .rn.
//
// Slot order inside a row is lane-major (lane 0's hits in candidate order, then lane 1's, ...):
// deterministic, and as unspecified as the reference's HOOMD-internal order.  When a row has
// more than K neighbors the slot index wraps modulo K and the last writer wins, exactly like
// htf/TensorflowCompute.cc:370 (with this kernel's hit order).
#include "common.cuh"

#include <math_constants.h>
#ifdef HTF_DEBUG_FLAGS
#include <cstdio>
#include <vector>
#endif

namespace {

#ifndef HTF_RPP
#define HTF_RPP 2
#endif
#ifndef HTF_TILE
#define HTF_TILE 4
#endif
constexpr int RPP = HTF_RPP;      // rows tested against each staged candidate chunk (even)

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
    int cap;             // per-warp window capacity (candidates), multiple of 32, <= 32768
    int cap_tile;        // tile kernel: candidates staged per block
    unsigned char *tile_flag;   // [tiles]: 1 = the tile kernel left this tile to the per-cell kernel
    int *flag_count;            // tiles flagged by this launch's tile kernel (nullptr: unknown, always scan)
    int *flag_count_next;       // the other parity's counter, zeroed by the per-cell kernel for the next launch
    int use_flags;       // per-cell kernel: process only cells of flagged tiles
    float4 *out;
    int *idx_out;
    int *count_out;
    int *overflow;
};

#ifdef HTF_EXP_COLD_NOINLINE
#define HTF_EMIT_INLINE __noinline__
#else
#define HTF_EMIT_INLINE __forceinline__
#endif

typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
// same, but pinned: used for loop-invariant operands so that the pair is built once, not per iteration
__device__ __forceinline__ f32x2 pack2_pinned(float lo, float hi)
{
    f32x2 r;
#ifdef HTF_EXP_NOPIN
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
#else
    asm volatile("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
#endif
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi)
{
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

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
__device__ __forceinline__ void sts_u16(unsigned addr, unsigned v)
{
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ unsigned lds_u16(unsigned addr)
{
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
    return v;
}

__device__ __forceinline__ int lds_i32(unsigned addr)
{
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float4 lds_f4(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

// same accesses without the compiler barrier, for read-only staged data / write-only lists inside one phase
__device__ __forceinline__ float4 lds_f4_ro(unsigned addr)
{
    float4 v;
    asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u16_nb(unsigned addr, unsigned v)
{
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v));
}

struct RowState {
    f32x2 x[RPP / 2], y[RPP / 2], z[RPP / 2];   // row pairs packed for the f32x2 pipe
    float t[RPP];                               // row types (mapped-nlist rule)
    int self_rel[RPP];      // index of the row's own particle inside the staged window (or -1)
    unsigned lp[RPP];       // shared-space byte address of this lane's next list entry for row r
};

// ---- test: append the window-relative index of every hit to the lane-private lists ----
// One 32-candidate chunk against the RPP rows: 1 LDS.128, 6 packed subs, 6 packed squares, 8 adds,
// 4 compares (+ the "is this the row's own particle" compare) and the predicated list appends.
// MASKED adds the "candidate lies past the end of the window" test (only tiny grids need it).
template <bool WRAP, bool MAPPED, bool MASKED>
__device__ __forceinline__ void test_chunk(const NlistParams &p, const float4 *cand, int t0, int mlen, RowState &rs,
                                           int lane)
{
    const int tl = t0 + lane;
    const float4 c = cand[tl];
    const bool pv = !MASKED || tl < mlen;
    const f32x2 cx = pack2(c.x, c.x), cy = pack2(c.y, c.y), cz = pack2(c.z, c.z);
#pragma unroll
    for (int h = 0; h < RPP / 2; h++) {
        float dx[2], dy[2], dz[2], xx[2], yy[2], zz[2];
        const f32x2 dx2 = sub2(cx, rs.x[h]), dy2 = sub2(cy, rs.y[h]), dz2 = sub2(cz, rs.z[h]);
        if (WRAP) {
            unpack2(dx2, dx[0], dx[1]); unpack2(dy2, dy[0], dy[1]); unpack2(dz2, dz[0], dz[1]);
#pragma unroll
            for (int u = 0; u < 2; u++) {
                dz[u] = wrap_axis(dz[u], -p.g.half[2], p.g.half[2], p.g.L[2]);
                dy[u] = wrap_axis(dy[u], -p.g.half[1], p.g.half[1], p.g.L[1]);
                dx[u] = wrap_axis(dx[u], -p.g.half[0], p.g.half[0], p.g.L[0]);
                xx[u] = __fmul_rn(dx[u], dx[u]); yy[u] = __fmul_rn(dy[u], dy[u]); zz[u] = __fmul_rn(dz[u], dz[u]);
            }
        } else {
            unpack2(mul2(dx2, dx2), xx[0], xx[1]);
            unpack2(mul2(dy2, dy2), yy[0], yy[1]);
            unpack2(mul2(dz2, dz2), zz[0], zz[1]);
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int r = 2 * h + u;
            const float rsq = __fadd_rn(__fadd_rn(xx[u], yy[u]), zz[u]);
            // rsq <= rc2 is !(rsq > rc2) for every non-NaN rsq; +inf rows / sentinels give inf or NaN -> no hit
            bool hit = (rsq <= p.rc2) & (tl != rs.self_rel[r]);
            if (MASKED) hit = hit & pv;
            if (MAPPED) hit = hit && (((int)c.w >= p.map_type_start) == ((int)rs.t[r] >= p.map_type_start));
            if (hit) {
                sts_u16(rs.lp[r], (unsigned)tl);
                rs.lp[r] += 64u;                    // lists are [k][lane] u16: next k is 32 entries on
            }
        }
    }
}

template <bool WRAP, bool MAPPED, bool MASKED>
__device__ __forceinline__ void test_window(const NlistParams &p, const float4 *cand, int mround, int mlen,
                                            RowState &rs, int lane)
{
#pragma unroll 1
    for (int t0 = 0; t0 < mround; t0 += 32) test_chunk<WRAP, MAPPED, MASKED>(p, cand, t0, mlen, rs, lane);
}

// ---- emit one row whose hits all come from the single staged window ----
// Scan of the lane counts -> slot ranges; "slot -> candidate" published in a small shared map; then lane s
// re-derives d for slots s, s+32, ... and stores them coalesced, zero padding included.
template <bool WITH_IDX>
__device__ HTF_EMIT_INLINE void emit_single_window(const NlistParams &p, unsigned cand_s, const int *candidx,
                                                   unsigned slotmap_s, unsigned list_s, int c_l, bool wrap,
                                                   const float4 &pi, int orig, int lane)
{
    const int K = p.K;
    int incl_c = c_l;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(HTF_FULL, incl_c, o);
        if (lane >= o) incl_c += t;
    }
    const int total = __shfl_sync(HTF_FULL, incl_c, 31);
    const size_t row = (size_t)(orig - p.row_lo);
    float4 *grow = p.out + row * K;
    if (total <= K) {
        // slots [0,total) are exactly the hits, lane-major
        const unsigned qa = slotmap_s + (unsigned)(incl_c - c_l) * 2u;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k < c_l) sts_u16(qa + 2u * k, lds_u16(list_s + 64u * k));
        if (__any_sync(HTF_FULL, c_l > 4))
            for (int k = 4; k < c_l; k++) sts_u16(qa + 2u * k, lds_u16(list_s + 64u * k));
    } else {
        // htf/TensorflowCompute.cc:370: slot = q mod K, the last writer of a slot wins -> only the last K hits
        const int first = total - K;
        int q = incl_c - c_l;
        for (int k = 0; k < c_l; k++, q++)
            if (q >= first) sts_u16(slotmap_s + 2u * (unsigned)(q % K), lds_u16(list_s + 64u * k));
    }
    __syncwarp();
    const int nvalid = min(total, K);
    for (int sl = lane; sl < K; sl += 32) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        int vi = -1;
        if (sl < nvalid) {
            const unsigned ci = lds_u16(slotmap_s + 2u * sl);
            const float4 cd = lds_f4(cand_s + ci * 16u);
            float dx = __fsub_rn(cd.x, pi.x), dy = __fsub_rn(cd.y, pi.y), dz = __fsub_rn(cd.z, pi.z);
            if (wrap) {
                dz = wrap_axis(dz, -p.g.half[2], p.g.half[2], p.g.L[2]);
                dy = wrap_axis(dy, -p.g.half[1], p.g.half[1], p.g.L[1]);
                dx = wrap_axis(dx, -p.g.half[0], p.g.half[0], p.g.L[0]);
            }
            v = make_float4(dx, dy, dz, cd.w);
            if (WITH_IDX) vi = candidx[ci];
        }
        if (!WITH_IDX || p.out) grow[sl] = v;                   // idx-only builds (candidate lists) pass out == nullptr
        if (WITH_IDX) p.idx_out[row * K + sl] = vi;
    }
    if (lane == 0) {
        if (p.count_out) p.count_out[row] = total;
        if (total >= K && p.overflow) atomicMax(p.overflow, total);
    }
    __syncwarp();
}

// Slot phase of the emit: lane s derives (d, type) for slots s, s+32, ... of one row and stores them
// coalesced.  KCH = K/32 when that is a compile-time-friendly value (fully unrolled, immediate offsets),
// 0 = generic K.
template <bool WITH_IDX, int KCH>
__device__ __forceinline__ void emit_slots(const NlistParams &p, unsigned cand_ws, const int *candidx_w,
                                           unsigned slotmap_s, const float4 &pi, bool wrap, int total, size_t row,
                                           int lane)
{
    const int K = KCH ? KCH * 32 : p.K;
    float4 *dst = p.out + row * K + lane;
    int *idst = WITH_IDX ? p.idx_out + row * K + lane : nullptr;
    const unsigned sa = slotmap_s + 2u * lane;
    if (KCH) {
#pragma unroll
        for (int i = 0; i < (KCH ? KCH : 1); i++) {
            const int sl = lane + 32 * i;
            const bool valid = sl < total;
            const unsigned ci = valid ? lds_u16(sa + 64u * i) : 0u;      // stale map entries are never dereferenced
            const float4 cd = lds_f4(cand_ws + ci * 16u);
            float dx = __fsub_rn(cd.x, pi.x), dy = __fsub_rn(cd.y, pi.y), dz = __fsub_rn(cd.z, pi.z);
            if (wrap) {
                dz = wrap_axis(dz, -p.g.half[2], p.g.half[2], p.g.L[2]);
                dy = wrap_axis(dy, -p.g.half[1], p.g.half[1], p.g.L[1]);
                dx = wrap_axis(dx, -p.g.half[0], p.g.half[0], p.g.L[0]);
            }
            if (!WITH_IDX || p.out) dst[32 * i] = valid ? make_float4(dx, dy, dz, cd.w) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (WITH_IDX) idst[32 * i] = valid ? candidx_w[ci] : -1;
        }
    } else {
        unsigned sb = sa;
        for (int sl = lane; sl < K; sl += 32, dst += 32, sb += 64u) {
            const bool valid = sl < total;
            const unsigned ci = valid ? lds_u16(sb) : 0u;
            const float4 cd = lds_f4(cand_ws + ci * 16u);
            float dx = __fsub_rn(cd.x, pi.x), dy = __fsub_rn(cd.y, pi.y), dz = __fsub_rn(cd.z, pi.z);
            if (wrap) {
                dz = wrap_axis(dz, -p.g.half[2], p.g.half[2], p.g.L[2]);
                dy = wrap_axis(dy, -p.g.half[1], p.g.half[1], p.g.L[1]);
                dx = wrap_axis(dx, -p.g.half[0], p.g.half[0], p.g.L[0]);
            }
            if (!WITH_IDX || p.out) *dst = valid ? make_float4(dx, dy, dz, cd.w) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (WITH_IDX) { *idst = valid ? candidx_w[ci] : -1; idst += 32; }
        }
    }
}

constexpr int TILE = HTF_TILE;    // cells per block along x in the tile kernel (= warps per block)
#ifdef HTF_EXP_ROWS
constexpr int NPMAX = 192;        // piece table capacity: (TILE + 2) * 9 <= NPMAX; rows kernel: (cells + 2) * 9 <= NPMAX
#else
constexpr int NPMAX = 128;        // piece table capacity: (TILE + 2) * 9 <= NPMAX  ->  TILE <= 12
#endif
constexpr int TILE_HDR = (2 * NPMAX + 32) * 4;   // bytes: piece table end[NPMAX] + adj[NPMAX] + colstart[24] + warp totals[8]
static_assert((TILE + 2) * 9 <= NPMAX && TILE * 32 >= (TILE + 2) * 9 && TILE + 3 <= 24, "tile size");

// One warp builds all rows of one cell with its own staging buffer (any density: candidates are
// re-staged in windows when they do not fit).
template <bool WITH_IDX, bool MAPPED>
__device__ __forceinline__ void build_cell(const NlistParams &p, const int cell, unsigned char *smem_raw)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int K = p.K;
    const int cap = p.cap;
    // per-warp carve-up (see per_warp_bytes): cand[cap] f4 | rowstage[K] f4 | lists[RPP][cap] u16 |
    //   runtab[64] i32 | slotmap[K] u16 (padded to 16 B) | (candidx[cap] i32 | idxstage[K] i32)
    const size_t slotmap_bytes = ((size_t)K * 2 + 15) & ~(size_t)15;
    size_t per_warp = (size_t)cap * 16 + (size_t)K * 16 + (size_t)RPP * cap * 2 + 256 + slotmap_bytes;
    if (WITH_IDX) per_warp += (size_t)cap * 4 + (((size_t)K * 4 + 15) & ~(size_t)15);
    unsigned char *base = smem_raw + per_warp * warp;
    float4 *cand = reinterpret_cast<float4 *>(base);
    float4 *rowstage = cand + cap;
    unsigned char *lists = reinterpret_cast<unsigned char *>(rowstage + K);
    int *runtab = reinterpret_cast<int *>(lists + (size_t)RPP * cap * 2);     // [0..31] run end, [32..63] src - t
    unsigned short *slotmap = reinterpret_cast<unsigned short *>(runtab + 64);
    int *candidx = reinterpret_cast<int *>(reinterpret_cast<unsigned char *>(slotmap) + slotmap_bytes);
    int *idxstage = candidx + cap;
    const unsigned cand_s = (unsigned)__cvta_generic_to_shared(cand);
    const unsigned candidx_s = (unsigned)__cvta_generic_to_shared(candidx);
    const unsigned lists_s = (unsigned)__cvta_generic_to_shared(lists);

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
    const int m = __shfl_sync(HTF_FULL, incl, 31);
    // run table: candidate t belongs to the first run q with t < end[q]; its source slot is t + adj[q]
    runtab[lane] = incl;
    runtab[32 + lane] = rb - (incl - rl);
    // the run that holds this cell's own particles -> virtual index of "self" for each row
    const unsigned selfmask = __ballot_sync(HTF_FULL, rl > 0 && rb <= b && b < rb + rl);
    const int self_base = __shfl_sync(HTF_FULL, (incl - rl) - rb, __ffs(selfmask) - 1);
    __syncwarp();

    const int npass = (m + cap - 1) / cap;

    // stage window `pass` of the virtual candidate list; returns its length rounded up to 32
    // (the tail is filled with +inf sentinels, which can never be a hit)
    const unsigned runtab_s = (unsigned)__cvta_generic_to_shared(runtab);
    auto stage = [&](int pass) {
        const int w0 = pass * cap, w1 = min(m, w0 + cap);
        // each lane walks the run table with a private cursor (runs are ~one chunk long, so the cursor
        // advances 0-2 entries per chunk); the copies are asynchronous, one wait for the whole window
        unsigned qa = runtab_s;
        int qend = lds_i32(qa), qadj = lds_i32(qa + 128u);
        unsigned dsta = cand_s + (unsigned)lane * 16u;
        for (int t = w0 + lane; t < w1; t += 32, dsta += 512u) {
            while (t >= qend) { qa += 4u; qend = lds_i32(qa); qadj = lds_i32(qa + 128u); }
            cp_async16(dsta, p.spos + (t + qadj));
            if (WITH_IDX) cp_async4(candidx_s + (unsigned)(t - w0) * 4u, p.sorted_idx + (t + qadj));
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
    // a dense cell that needs several windows is walked one row at a time: that row's partial
    // result lives in the shared staging row between windows
    const int rstep = (npass == 1) ? RPP : 1;

    for (int s0 = b; s0 < e; s0 += rstep) {
        RowState rs;
        bool anyrow = false;
        {
            float px[RPP], py[RPP], pz[RPP];
#pragma unroll
            for (int r = 0; r < RPP; r++) {
                const int s = min(s0 + r, e - 1);                    // both loads are independent of each other
                const int o = __ldg(p.sorted_idx + s);
                float4 pi = __ldg(p.spos + s);
                const bool ok = (r < rstep) && (s0 + r < e) && o >= p.row_lo && o < p.row_hi;
                if (!ok) pi = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, 0.f);       // never hits
                px[r] = pi.x; py[r] = pi.y; pz[r] = pi.z; rs.t[r] = pi.w;
                rs.self_rel[r] = ok ? 0 : -1;                        // refined per window below
                anyrow |= ok;
            }
            if (!anyrow) continue;
#pragma unroll
            for (int h = 0; h < RPP / 2; h++) {
                rs.x[h] = pack2_pinned(px[2 * h], px[2 * h + 1]);
                rs.y[h] = pack2_pinned(py[2 * h], py[2 * h + 1]);
                rs.z[h] = pack2_pinned(pz[2 * h], pz[2 * h + 1]);
            }
        }
        bool rvalid[RPP];
#pragma unroll
        for (int r = 0; r < RPP; r++) rvalid[r] = rs.self_rel[r] == 0;

        int cnt_acc = 0;                                             // multi-window mode: hits of the row so far
        for (int pass = 0; pass < npass; pass++) {
            if (npass > 1) { __syncwarp(); mround = stage(pass); }
#pragma unroll
            for (int r = 0; r < RPP; r++) {
                const int rel = self_base + s0 + r - pass * cap;       // own particle inside this window?
                rs.self_rel[r] = (rvalid[r] && rel >= 0 && rel < mround) ? rel : -1;
                rs.lp[r] = lists_s + (unsigned)(r * cap + lane) * 2u;
            }
            if (wrap) test_window<true, MAPPED, false>(p, cand, mround, mround, rs, lane);
            else test_window<false, MAPPED, false>(p, cand, mround, mround, rs, lane);
            __syncwarp();

            // ---- emit, row by row (kept rolled: this code runs once per row, not once per pair) ----
            const bool last = pass == npass - 1;
            const unsigned slotmap_s = (unsigned)__cvta_generic_to_shared(slotmap);
#pragma unroll 1
            for (int r = 0; r < rstep; r++) {
                const int srow = s0 + r;
                if (srow >= e) break;
                const int orig = __ldg(p.sorted_idx + srow);
                if (orig < p.row_lo || orig >= p.row_hi) continue;       // warp-uniform
                const float4 pi = __ldg(p.spos + srow);
                unsigned lp_r = rs.lp[0];
#pragma unroll
                for (int q = 1; q < RPP; q++) if (r == q) lp_r = rs.lp[q];
                const unsigned list_s = lists_s + (unsigned)(r * cap + lane) * 2u;
                const int c_l = (int)((lp_r - list_s) >> 6);
                if (npass == 1) {
                    emit_single_window<WITH_IDX>(p, cand_s, candidx, slotmap_s, list_s, c_l, wrap, pi, orig, lane);
                    continue;
                }
                int incl_c = c_l;                                        // scan of the lane counts
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(HTF_FULL, incl_c, o);
                    if (lane >= o) incl_c += t;
                }
                const int wtotal = __shfl_sync(HTF_FULL, incl_c, 31);
                const int total = cnt_acc + wtotal;
                float4 *grow = p.out + (size_t)(orig - p.row_lo) * K;

                // ---- dense cell, several windows: the row accumulates in the shared staging row ----
                const int first = total - K;            // > 0: more than K hits so far, only the last K survive
                for (int sl = lane; sl < K; sl += 32) slotmap[sl] = 0xffffu;     // 0xffff = none from this window
                __syncwarp();
                int q = cnt_acc + incl_c - c_l;
                for (int k = 0; k < c_l; k++, q++) {
                    const unsigned ci = lds_u16(list_s + (unsigned)k * 64u);
                    // htf/TensorflowCompute.cc:370: slot = q mod K, the last writer of a slot wins
                    if (first <= 0) slotmap[q] = (unsigned short)ci;
                    else if (q >= first) slotmap[q % K] = (unsigned short)ci;
                }
                __syncwarp();
                const size_t row = (size_t)(orig - p.row_lo);
                const bool direct = false;
                float4 *dst = rowstage;
                int *idst = WITH_IDX ? idxstage : nullptr;
                for (int sl = lane; sl < K; sl += 32) {
                    const unsigned ci = slotmap[sl];
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    int vi = -1;
                    if (ci != 0xffffu) {
                        const float4 cd = cand[ci];
                        float dx = __fsub_rn(cd.x, pi.x), dy = __fsub_rn(cd.y, pi.y), dz = __fsub_rn(cd.z, pi.z);
                        dz = wrap_axis(dz, -p.g.half[2], p.g.half[2], p.g.L[2]);     // a no-op for interior cells
                        dy = wrap_axis(dy, -p.g.half[1], p.g.half[1], p.g.L[1]);
                        dx = wrap_axis(dx, -p.g.half[0], p.g.half[0], p.g.L[0]);
                        v = make_float4(dx, dy, dz, cd.w);
                        if (WITH_IDX) vi = candidx[ci];
                    }
                    if (direct || ci != 0xffffu || (pass == 0)) {        // multi-window: keep earlier windows' slots
                        dst[sl] = v;
                        if (WITH_IDX) idst[sl] = vi;
                    }
                }
                if (!direct) {
                    cnt_acc = total;
                    __syncwarp();
                    if (last) {
                        for (int sl = lane; sl < K; sl += 32) {
                            if (!WITH_IDX || p.out) grow[sl] = rowstage[sl];
                            if (WITH_IDX) p.idx_out[row * K + sl] = idxstage[sl];
                        }
                    }
                }
                if (last && lane == 0) {
                    if (p.count_out) p.count_out[row] = total;
                    if (p.overflow && total >= K) atomicMax(p.overflow, total);
                }
                __syncwarp();
            }
        }
    }
}

// Per-cell kernel.  Stand-alone it walks every cell (one warp each).  As the second pass behind the
// tile kernel it is a small persistent grid that scans the tile flags and builds only the cells of
// flagged tiles -- with no flagged tile it costs a few microseconds.
template <bool WITH_IDX, bool MAPPED>
__global__ void __launch_bounds__(128) nlist_build_kernel(const NlistParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    if (!p.use_flags) {
        const int cw = blockIdx.x * wpb + warp, layer = p.g.n[0] * p.g.n[1];
        if (cw < layer * p.g.zcount) {
            const int lz = cw / layer;
            build_cell<WITH_IDX, MAPPED>(p, ((p.g.z0 + lz) % p.g.n[2]) * layer + (cw - lz * layer), smem_raw);
        }
        return;
    }
    if (p.flag_count) {
        // nothing flagged by the tile kernel (the usual case): leave at once
        const int nflag = *reinterpret_cast<volatile const int *>(p.flag_count);
        if (blockIdx.x == 0 && threadIdx.x == 0) *p.flag_count_next = 0;
        if (nflag == 0) return;
    }
    const int nx = p.g.n[0], tiles_x = (nx + TILE - 1) / TILE;
    const int per_layer = tiles_x * p.g.n[1];
    const int ntiles_win = per_layer * p.g.zcount;                  // tiles of the z-window only
    for (int tw = blockIdx.x * wpb + warp; tw < ntiles_win; tw += gridDim.x * wpb) {
        const int lz = tw / per_layer;
        const int tile = ((p.g.z0 + lz) % p.g.n[2]) * per_layer + (tw - lz * per_layer);
        if (!p.tile_flag[tile]) continue;
        const int tx = tile % tiles_x, row = tile / tiles_x;        // row = cz * ny + cy
        for (int c = 0; c < TILE && tx * TILE + c < nx; c++) build_cell<WITH_IDX, MAPPED>(p, row * nx + tx * TILE + c, smem_raw);
    }
}

// ------------------------------------------------------------------------------------------------
// Tile kernel: one block = TILE x-adjacent cells (one warp each) of one (y,z) cell row.  The block
// stages the (TILE+2) x 3 x 3 cell neighbourhood ONCE, column-major, so that each warp's 3x3x3
// stencil is one contiguous window of the shared buffer: per-cell staging traffic drops from 3 to
// (TILE+2)/TILE columns, the stencil / prefix set-up is paid once per block, and the smaller
// shared-memory footprint per warp doubles the resident warps.  Tiles whose neighbourhood does
// not fit (dense clusters) are flagged and left to the per-cell kernel above.
template <bool WITH_IDX, bool MAPPED>
__global__ void __launch_bounds__(TILE * 32) nlist_tile_kernel(const NlistParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = p.K, capB = p.cap_tile, capW = p.cap;
    int *ptab_end = reinterpret_cast<int *>(smem_raw);          // [NPMAX] inclusive prefix of piece lengths
    int *ptab_adj = ptab_end + NPMAX;                           // [NPMAX] source slot - staged index
    int *colstart = ptab_adj + NPMAX;                           // [<= TILE+3] staged offset of each column
    int *wtot = colstart + 24;                                  // [4] piece-length totals of warps 0..3
    float4 *cand = reinterpret_cast<float4 *>(smem_raw + TILE_HDR);
    int *candidx = reinterpret_cast<int *>(cand + capB + 32);
    unsigned char *wbase = WITH_IDX ? reinterpret_cast<unsigned char *>(candidx + capB + 32)
                                    : reinterpret_cast<unsigned char *>(candidx);
    const size_t per_warp = (size_t)RPP * capW * 2 + (((size_t)K * 2 + 15) & ~(size_t)15);
    unsigned char *lists = wbase + per_warp * warp;
    const unsigned lists_s = (unsigned)__cvta_generic_to_shared(lists);
    const unsigned slotmap_s = lists_s + (unsigned)(RPP * capW * 2);
    const unsigned cand_s = (unsigned)__cvta_generic_to_shared(cand);
    const unsigned candidx_s = (unsigned)__cvta_generic_to_shared(candidx);

    const int nx = p.g.n[0], ny = p.g.n[1], nz = p.g.n[2];
    const int tiles_x = gridDim.x;                              // = ceil(nx / TILE)
    const int tx = blockIdx.x, cy = blockIdx.y;
    const int cz = (p.g.z0 + (int)blockIdx.z) % nz;            // grid.z = cell layers of the z-window
    const int bid = (cz * ny + cy) * tiles_x + tx;
    const int cx0 = tx * TILE;
    const int nact = min(TILE, nx - cx0);                       // cells of this tile
    const int nly = min(ny, 3), nlz = min(nz, 3), nyz = nly * nlz;
    const bool xs = nx >= 3;                                    // x stencil = {c-1, c, c+1}; else every x cell
    const int ncol = xs ? nact + 2 : nx;
    const int npieces = ncol * nyz;                             // <= 6 * 9

    // ---- piece table: piece (col, j) = one stencil cell; prefix sums give the staged layout ----
    int pl = 0, pb = 0;
    if (tid < npieces) {
        const int col = tid / nyz, j = tid - col * nyz, jy = j % nly, jz = j / nly;
        int sx = xs ? cx0 - 1 + col : col;
        sx = sx < 0 ? sx + nx : (sx >= nx ? sx - nx : sx);
        int sy = ny <= 3 ? jy : cy + jy - 1;
        sy = sy < 0 ? sy + ny : (sy >= ny ? sy - ny : sy);
        int sz = nz <= 3 ? jz : cz + jz - 1;
        sz = sz < 0 ? sz + nz : (sz >= nz ? sz - nz : sz);
        const int c0 = (sz * ny + sy) * nx + sx;
        pb = __ldg(p.cell_start + c0);
        pl = __ldg(p.cell_start + c0 + 1) - pb;
    }
    int incl = pl;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(HTF_FULL, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31 && warp < 4) wtot[warp] = incl;              // totals of pieces [32w, 32w+32)
    const unsigned stage_bar = (unsigned)__cvta_generic_to_shared(wtot + 4);   // mbarrier of the bulk copies (8-byte aligned)
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(stage_bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    for (int w = 0; w < warp && w < 4; w++) incl += wtot[w];
    if (tid < NPMAX) {
        ptab_end[tid] = incl;
        ptab_adj[tid] = pb - (incl - pl);
        if (tid < npieces && tid % nyz == 0) colstart[tid / nyz] = incl - pl;
        if (tid == npieces - 1) colstart[ncol] = incl;
    }
    __syncthreads();
    const int mblock = colstart[ncol];
    bool fits = mblock <= capB;
    for (int w = 0; w < nact; w++) {
        const int wl = xs ? colstart[w + 3] - colstart[w] : mblock;
        fits = fits && wl <= capW;
    }
    if (tid == 0) {
        p.tile_flag[bid] = fits ? 0 : 1;
        if (!fits && p.flag_count) atomicAdd(p.flag_count, 1);
    }
    if (!fits) return;                                          // block-uniform
    const bool full = (p.row_lo == 0 && p.row_hi == p.n_all);
    if (!full) {
        // sharded build: skip the tile (before staging anything) when none of its cells holds a local row
        bool any = false;
        if (warp < nact) {
            const int c = (cz * ny + cy) * nx + cx0 + warp;
            const int cb_ = __ldg(p.cell_start + c), ce_ = __ldg(p.cell_start + c + 1);
            for (int s = cb_ + lane; s < ce_; s += 32) {
                const int o = __ldg(p.sorted_idx + s);
                any |= (o >= p.row_lo && o < p.row_hi);
            }
        }
        if (!__syncthreads_or(any)) return;
    }

    // ---- stage the whole neighbourhood once: one TMA bulk copy per piece (a piece = one stencil cell = one
    //      contiguous run of the cell-sorted positions), issued by the thread that owns the piece's table entry;
    //      completion is counted in bytes on one mbarrier ----
    if (tid == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(stage_bar), "r"((unsigned)mblock * 16u) : "memory");
    if (tid < npieces && pl > 0)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(cand_s + (unsigned)(incl - pl) * 16u), "l"(p.spos + pb), "r"((unsigned)pl * 16u), "r"(stage_bar) : "memory");
    if (WITH_IDX) {
        // original indices ride along as 4-byte cp.async (bulk copies need 16-byte granules): warp w takes pieces w, w+TILE, ...
        for (int q = warp; q < npieces; q += TILE) {
            const int qend = ptab_end[q], qadj = ptab_adj[q];
            const int qbeg = q == 0 ? 0 : ptab_end[q - 1];
            for (int t = qbeg + lane; t < qend; t += 32) cp_async4(candidx_s + (unsigned)t * 4u, p.sorted_idx + (t + qadj));
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    // 32 sentinels behind the staged data: the last chunk of the last window may read past its end
    if (warp == 0) cand[mblock + lane] = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, 0.f);
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(stage_bar) : "memory");
    __syncthreads();

    // ---- one warp per cell of the tile ----
    if (warp >= nact) return;
    const int cx = cx0 + warp;
    const int cell = (cz * ny + cy) * nx + cx;
    const int b = __ldg(p.cell_start + cell), e = __ldg(p.cell_start + cell + 1);
    if (e == b) return;
    if (!full) {
        bool any = false;
        for (int s = b + lane; s < e; s += 32) {
            const int o = __ldg(p.sorted_idx + s);
            any |= (o >= p.row_lo && o < p.row_hi);
        }
        if (!__any_sync(HTF_FULL, any)) return;
    }
    const bool wrap = !((nx >= 5 && cx >= 1 && cx <= nx - 2) && (ny >= 5 && cy >= 1 && cy <= ny - 2) &&
                        (nz >= 5 && cz >= 1 && cz <= nz - 2));
    const int ws = xs ? colstart[warp] : 0;
    const int mlen_true = (xs ? colstart[warp + 3] : mblock) - ws;
    const int mround = (mlen_true + 31) & ~31;
    // With >= 4 cells in x, whatever follows the window in the buffer is the x-column two cells away (or the
    // sentinels): farther than r_cut in x by construction of the grid, so it can never pass the cutoff test
    // and the "past the end of the window" mask is unnecessary.  Tiny grids keep the mask.
    // (test_window<..., MASKED = true> is used for nx < 4 only)
    // staged position of this cell's own particles: piece (own column, own (y,z))
    const int ps = (xs ? warp + 1 : cx) * nyz + (nz <= 3 ? cz : 1) * nly + (ny <= 3 ? cy : 1);
    const int self_base = (ps == 0 ? 0 : ptab_end[ps - 1]) - ws - b;
    const float4 *cand_w = cand + ws;
    const unsigned cand_ws = cand_s + (unsigned)ws * 16u;
#ifdef HTF_EXP_NOUNROLL
    const int kch = 0;
#elif defined(HTF_EXP_UNROLL3)
    const int kch = (K == 64) ? 2 : (K == 96 ? 3 : 0);
#else
    const int kch = (K == 64) ? 2 : 0;          // unrolled slot phase for K = 64; more copies cost more in I-cache
#endif
    const int *candidx_w = candidx + ws;

    for (int s0 = b; s0 < e; s0 += RPP) {
        RowState rs;
        int orig[RPP];
        bool anyrow = false;
        {
            float px[RPP], py[RPP], pz[RPP];
#pragma unroll
            for (int r = 0; r < RPP; r++) {
                const int s = min(s0 + r, e - 1);
                const int rel = self_base + s;                       // the row's own particle in the window
                float4 pi = cand_w[rel];
                const int o = __ldg(p.sorted_idx + s);               // needed for the row address anyway
                const bool ok = (s0 + r < e) && o >= p.row_lo && o < p.row_hi;
                orig[r] = ok ? o : -1;
                if (!ok) pi = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, 0.f);       // never hits
                px[r] = pi.x; py[r] = pi.y; pz[r] = pi.z; rs.t[r] = pi.w;
                rs.self_rel[r] = ok ? rel : -1;
                rs.lp[r] = lists_s + (unsigned)(r * capW + lane) * 2u;
                anyrow |= ok;
            }
            if (!anyrow) continue;
#pragma unroll
            for (int h = 0; h < RPP / 2; h++) {
                rs.x[h] = pack2_pinned(px[2 * h], px[2 * h + 1]);
                rs.y[h] = pack2_pinned(py[2 * h], py[2 * h + 1]);
                rs.z[h] = pack2_pinned(pz[2 * h], pz[2 * h + 1]);
            }
        }
#ifdef HTF_EXP_MASKALL
        if (!wrap) test_window<false, MAPPED, true>(p, cand_w, mround, mlen_true, rs, lane);
        else test_window<true, MAPPED, true>(p, cand_w, mround, mlen_true, rs, lane);
#else
        if (!wrap) test_window<false, MAPPED, false>(p, cand_w, mround, mround, rs, lane);
        else if (nx >= 4) test_window<true, MAPPED, false>(p, cand_w, mround, mround, rs, lane);
        else test_window<true, MAPPED, true>(p, cand_w, mround, mlen_true, rs, lane);
#endif
        __syncwarp();

        // ---- emit the batch.  Lane counts of two rows share one 32-bit scan (16 bits each). ----
        int cl[RPP], excl[RPP], tot[RPP];
#pragma unroll
        for (int r = 0; r < RPP; r++) cl[r] = (int)((rs.lp[r] - (lists_s + (unsigned)(r * capW + lane) * 2u)) >> 6);
#pragma unroll
        for (int h = 0; h < RPP / 2; h++) {
            const unsigned packed = (unsigned)cl[2 * h] | ((unsigned)cl[2 * h + 1] << 16);
            unsigned inc = packed;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(HTF_FULL, inc, o);
                if (lane >= o) inc += t;
            }
            const unsigned all = __shfl_sync(HTF_FULL, inc, 31);
            const unsigned ex = inc - packed;
            excl[2 * h] = (int)(ex & 0xffffu); excl[2 * h + 1] = (int)(ex >> 16);
            tot[2 * h] = (int)(all & 0xffffu); tot[2 * h + 1] = (int)(all >> 16);
        }
#pragma unroll
        for (int r = 0; r < RPP; r++) {
            if (orig[r] < 0) continue;                                   // warp-uniform
            const unsigned list_s = lists_s + (unsigned)(r * capW + lane) * 2u;
            const int total = tot[r];
            const size_t row = (size_t)(orig[r] - p.row_lo);
            if (total > K) {
                // overflowing row: modulo-K rule of htf/TensorflowCompute.cc:370 (cold path)
                const float4 pi = cand_w[self_base + s0 + r];
                emit_single_window<WITH_IDX>(p, cand_ws, candidx_w, slotmap_s, list_s, cl[r], wrap, pi, orig[r], lane);
                continue;
            }
            // slot -> candidate map: slots [0,total) are exactly the hits, lane-major
            const unsigned qa = slotmap_s + (unsigned)excl[r] * 2u;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (k < cl[r]) sts_u16(qa + 2u * k, lds_u16(list_s + 64u * k));
            if (__any_sync(HTF_FULL, cl[r] > 4))
                for (int k = 4; k < cl[r]; k++) sts_u16(qa + 2u * k, lds_u16(list_s + 64u * k));
            __syncwarp();
            const float4 pi = lds_f4(cand_ws + (unsigned)(self_base + s0 + r) * 16u);
            if (kch == 2) emit_slots<WITH_IDX, 2>(p, cand_ws, candidx_w, slotmap_s, pi, wrap, total, row, lane);
#ifdef HTF_EXP_UNROLL3
            else if (kch == 3) emit_slots<WITH_IDX, 3>(p, cand_ws, candidx_w, slotmap_s, pi, wrap, total, row, lane);
#endif
            else emit_slots<WITH_IDX, 0>(p, cand_ws, candidx_w, slotmap_s, pi, wrap, total, row, lane);
            if (lane == 0 && (p.count_out != nullptr || total == K)) {
                if (p.count_out) p.count_out[row] = total;
                if (total == K && p.overflow) atomicMax(p.overflow, total);
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Row-per-lane kernel -- an evaluated alternative, compiled only with -DHTF_EXP_ROWS (parity-green, but
// 0.59 ms against the tile kernel's 0.43 ms at 1M x 64: 324 instead of 388 warp-instructions per row, yet the
// per-block staging chain and 30 KB of shared memory per 64 rows hold it at 12 warps per SM and IPC 2.1).  A block builds up to 32*ROWS_NW CONSECUTIVE rows of the cell-sorted
// order, all from one x-row of cells; a lane owns one row.  The stencil of the cells the block touches is staged
// once (column-major, TMA bulk copies, as in the tile kernel), so a lane's 27-cell stencil is one contiguous
// window of the shared buffer:
//   test : every lane walks ITS window candidate by candidate (one LDS.128 per lane and candidate; lanes of
//          the same cell read the same address) and appends the window-relative index of each hit to its own
//          row list -- the list IS the row in slot order, so there is no ballot, no scan and no slot map;
//   emit : the warp then writes its rows one after the other, lane s deriving (d, type) for slots s, s+32, ...
//          from the row's list: contiguous coalesced 16-byte stores, zero padding included.
// Row lists are [slot][33] u16 (stride 33 keeps both the lane-wise appends and the slot-wise reads free of
// bank conflicts).  A list saturates at K entries; rows that reach K, segments whose stencil does not fit the
// buffer and x-rows with more rows than the grid has segments flag their tiles and are rebuilt exactly (modulo-K
// rule included) by the per-cell kernel, like tiles the tile kernel gives up on.
#ifndef HTF_ROWS_NW
#define HTF_ROWS_NW 2
#endif
constexpr int ROWS_NW = HTF_ROWS_NW;
constexpr int ROWS_PB = 32 * ROWS_NW;              // rows per block
constexpr int ROWS_MAXCELLS = NPMAX / 9 - 2;       // cells a segment may span: (cells + 2) * 9 pieces <= NPMAX
constexpr int ROWS_SLACK = 128;                    // readable sentinels behind the staged data: lanes run to the warp's longest window
constexpr int ROWS_HDR = (2 * NPMAX + 24 + 8 + 24) * 4;   // piece table | colstart[24] | misc[8] (mbarrier at +4) | cs[24]
static_assert(ROWS_MAXCELLS + 3 <= 24 && ROWS_HDR % 16 == 0 && NPMAX % 32 == 0, "rows kernel header");

__device__ __forceinline__ size_t rows_list_bytes(int K) { return (((size_t)(K + 1) * 66) + 15) & ~(size_t)15; }

// one candidate against this lane's row; appends to the lane's list on a hit
template <bool WRAP, bool MAPPED>
__device__ __forceinline__ void rows_test_one(const NlistParams &p, const float4 &q, int c, int wlen, int self_rel,
                                              const float4 &pi, f32x2 pxy, unsigned &ha, unsigned ha_dump)
{
    const f32x2 dxy = sub2(pack2(q.x, q.y), pxy);
    float dz = __fsub_rn(q.z, pi.z);
    float xx, yy, zz;
    if (WRAP) {
        float dx, dy;
        unpack2(dxy, dx, dy);
        dz = wrap_axis(dz, -p.g.half[2], p.g.half[2], p.g.L[2]);
        dy = wrap_axis(dy, -p.g.half[1], p.g.half[1], p.g.L[1]);
        dx = wrap_axis(dx, -p.g.half[0], p.g.half[0], p.g.L[0]);
        xx = __fmul_rn(dx, dx); yy = __fmul_rn(dy, dy);
    } else {
        unpack2(mul2(dxy, dxy), xx, yy);
    }
    zz = __fmul_rn(dz, dz);
    const float rsq = __fadd_rn(__fadd_rn(xx, yy), zz);
    bool hit = (rsq <= p.rc2) & (c != self_rel) & (c < wlen);
    if (MAPPED) hit = hit && (((int)q.w >= p.map_type_start) == ((int)pi.w >= p.map_type_start));
    if (hit) {
        sts_u16_nb(ha, (unsigned)c);
        ha = min(ha + 66u, ha_dump);
    }
}

// The window walk, four candidates per step.  ptxas cannot move a shared load above an earlier shared store (it has
// no alias information), so the next group's loads are issued by hand before this group's list appends.
template <bool WRAP, bool MAPPED>
__device__ __forceinline__ unsigned rows_test(const NlistParams &p, unsigned cbase, int maxlen, int wlen, int self_rel,
                                              const float4 &pi, unsigned ha, unsigned ha_dump)
{
    const f32x2 pxy = pack2_pinned(pi.x, pi.y);
    constexpr int G = 4;
    float4 q[G], n[G];
#pragma unroll
    for (int u = 0; u < G; u++) q[u] = lds_f4_ro(cbase + (unsigned)u * 16u);        // sentinels make any over-read harmless
#pragma unroll 1
    for (int c = 0; c < maxlen; c += G) {
#pragma unroll
        for (int u = 0; u < G; u++) n[u] = lds_f4_ro(cbase + (unsigned)(c + G + u) * 16u);
#pragma unroll
        for (int u = 0; u < G; u++) rows_test_one<WRAP, MAPPED>(p, q[u], c + u, wlen, self_rel, pi, pxy, ha, ha_dump);
#pragma unroll
        for (int u = 0; u < G; u++) q[u] = n[u];
    }
    return ha;
}

template <bool WITH_IDX, bool MAPPED>
__global__ void __launch_bounds__(ROWS_PB) nlist_rows_kernel(const NlistParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = p.K, capB = p.cap_tile;
    int *ptab_end = reinterpret_cast<int *>(smem_raw);          // [NPMAX] inclusive prefix of piece lengths
    int *ptab_adj = ptab_end + NPMAX;                           // [NPMAX] source slot - staged index
    int *colstart = ptab_adj + NPMAX;                           // [<= ROWS_MAXCELLS + 3] staged offset of each column
    int *misc = colstart + 24;                                  // [0] first cell, [1] last cell, [2] fits, [4..5] mbarrier
    int *cs = misc + 8;                                         // [<= ROWS_MAXCELLS + 1] sorted-order start of the segment's cells
    float4 *cand = reinterpret_cast<float4 *>(smem_raw + ROWS_HDR);
    int *candidx = reinterpret_cast<int *>(cand + capB + ROWS_SLACK);
    unsigned char *lists = WITH_IDX ? reinterpret_cast<unsigned char *>(candidx + capB + ROWS_SLACK)
                                    : reinterpret_cast<unsigned char *>(candidx);
    const unsigned cand_s = (unsigned)__cvta_generic_to_shared(cand);
    const unsigned candidx_s = (unsigned)__cvta_generic_to_shared(candidx);
    const unsigned hl_s = (unsigned)__cvta_generic_to_shared(lists) + (unsigned)(rows_list_bytes(K) * warp);
    const unsigned stage_bar = (unsigned)__cvta_generic_to_shared(misc + 4);

    const int nx = p.g.n[0], ny = p.g.n[1], nz = p.g.n[2];
    const int tiles_x = (nx + TILE - 1) / TILE;                 // flags use the tile kernel's tiling
    const int cy = blockIdx.y, cz = (p.g.z0 + (int)blockIdx.z) % nz;
    const int c_base = (cz * ny + cy) * nx;
    unsigned char *flag_row = p.tile_flag + (size_t)(cz * ny + cy) * tiles_x;
    const int R0 = __ldg(p.cell_start + c_base), R1 = __ldg(p.cell_start + c_base + nx);
    const int nrows = R1 - R0;
    if (nrows == 0) return;
    const int nseg = (nrows + ROWS_PB - 1) / ROWS_PB;
    if (nseg > (int)gridDim.x) {                                // denser than the launch was sized for: whole x-row to the fallback
#ifdef HTF_DEBUG_FLAGS
        if (tid == 0 && blockIdx.x == 0 && cy < 2) printf("flag A: nseg %d grid %d nrows %d cy %d cz %d\n", nseg, gridDim.x, nrows, cy, cz);
#endif
        if (blockIdx.x == 0)
            for (int t = tid; t < tiles_x; t += ROWS_PB) flag_row[t] = 1;
        return;
    }
    const int seg = blockIdx.x;
    if (seg >= nseg) return;
    const int rb = R0 + (int)((long long)seg * nrows / nseg), re = R0 + (int)((long long)(seg + 1) * nrows / nseg);

    if (!(p.row_lo == 0 && p.row_hi == p.n_all)) {
        // sharded build: skip the segment (before staging anything) when none of its rows is local
        bool any = false;
        for (int r = rb + tid; r < re; r += ROWS_PB) {
            const int o = __ldg(p.sorted_idx + r);
            any |= (o >= p.row_lo && o < p.row_hi);
        }
        if (!__syncthreads_or(any)) return;
    }

    // ---- which cells does the segment touch ----
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(stage_bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int c = tid; c < nx; c += ROWS_PB) {
        const int a = __ldg(p.cell_start + c_base + c), b = __ldg(p.cell_start + c_base + c + 1);
        if (a <= rb && rb < b) misc[0] = c;
        if (a < re && re <= b) misc[1] = c;
    }
    __syncthreads();
    const int cx_first = misc[0], cx_last = misc[1];
    const int ncs = cx_last - cx_first + 1;
    if (ncs > ROWS_MAXCELLS) {                                  // sparse x-row: too many cells for the piece table
#ifdef HTF_DEBUG_FLAGS
        if (tid == 0 && cy < 2) printf("flag B: ncs %d first %d last %d rb %d re %d\n", ncs, cx_first, cx_last, rb, re);
#endif
        for (int t = cx_first / TILE + tid; t <= cx_last / TILE; t += ROWS_PB) flag_row[t] = 1;
        return;
    }
    const int nly = min(ny, 3), nlz = min(nz, 3), nyz = nly * nlz;
    const int ncol = ncs + 2, npieces = ncol * nyz;
    if (tid <= ncs) cs[tid] = __ldg(p.cell_start + c_base + cx_first + tid);

    // ---- warp 0: piece table (piece (col, j) = one stencil cell), fit check, one TMA bulk copy per piece ----
    if (warp == 0) {
        int pl[NPMAX / 32], pb[NPMAX / 32], incl[NPMAX / 32];
        int carry = 0;
#pragma unroll
        for (int k = 0; k < NPMAX / 32; k++) {
            const int q = lane + 32 * k;
            pl[k] = 0; pb[k] = 0;
            if (q < npieces) {
                const int col = q / nyz, j = q - col * nyz, jy = j % nly, jz = j / nly;
                int sx = cx_first - 1 + col;
                sx = sx < 0 ? sx + nx : (sx >= nx ? sx - nx : sx);
                int sy = ny <= 3 ? jy : cy + jy - 1;
                sy = sy < 0 ? sy + ny : (sy >= ny ? sy - ny : sy);
                int sz = nz <= 3 ? jz : cz + jz - 1;
                sz = sz < 0 ? sz + nz : (sz >= nz ? sz - nz : sz);
                const int c0 = (sz * ny + sy) * nx + sx;
                pb[k] = __ldg(p.cell_start + c0);
                pl[k] = __ldg(p.cell_start + c0 + 1) - pb[k];
            }
        }
#pragma unroll
        for (int k = 0; k < NPMAX / 32; k++) {
            const int q = lane + 32 * k;
            int v = pl[k];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(HTF_FULL, v, o);
                if (lane >= o) v += t;
            }
            v += carry;
            incl[k] = v;
            carry = __shfl_sync(HTF_FULL, v, 31);
            ptab_end[q] = v;
            ptab_adj[q] = pb[k] - (v - pl[k]);
            if (q < npieces && q % nyz == 0) colstart[q / nyz] = v - pl[k];
        }
        const int mblock = carry;
        const bool fits = mblock <= capB;
        if (lane == 0) { colstart[ncol] = mblock; misc[2] = fits ? 1 : 0; }
        if (fits) {
            if (lane == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(stage_bar), "r"((unsigned)mblock * 16u) : "memory");
#pragma unroll
            for (int k = 0; k < NPMAX / 32; k++)
                if (pl[k] > 0)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(cand_s + (unsigned)(incl[k] - pl[k]) * 16u), "l"(p.spos + pb[k]), "r"((unsigned)pl[k] * 16u),
                                   "r"(stage_bar) : "memory");
        }
    }
    __syncthreads();
    if (!misc[2]) {                                             // stencil larger than the buffer: dense cluster -> fallback
#ifdef HTF_DEBUG_FLAGS
        if (tid == 0 && cy < 2) printf("flag C: mblock %d capB %d ncs %d\n", colstart[ncol], capB, ncs);
#endif
        for (int t = cx_first / TILE + tid; t <= cx_last / TILE; t += ROWS_PB) flag_row[t] = 1;
        return;
    }
    const int mblock = colstart[ncol];
    if (WITH_IDX) {
        for (int q = warp; q < npieces; q += ROWS_NW) {
            const int qend = ptab_end[q], qadj = ptab_adj[q];
            const int qbeg = q == 0 ? 0 : ptab_end[q - 1];
            for (int t = qbeg + lane; t < qend; t += 32) cp_async4(candidx_s + (unsigned)t * 4u, p.sorted_idx + (t + qadj));
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    for (int t = tid; t < ROWS_SLACK; t += ROWS_PB) cand[mblock + t] = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, 0.f);

    // ---- this lane's row, its cell and its window (overlaps the TMA latency) ----
    const int row = rb + warp * 32 + lane;
    bool active = row < re;
    int orig = -1;
    if (active) {
        orig = __ldg(p.sorted_idx + row);
        active = orig >= p.row_lo && orig < p.row_hi;
    }
    int w = -1;
    for (int j = 0; j < ncs; j++) w += (row >= cs[j]) ? 1 : 0;
    w = max(0, min(w, ncs - 1));
    const int ws = colstart[w], wlen = colstart[w + 3] - ws;
    const int ps = (w + 1) * nyz + (nz <= 3 ? cz : 1) * nly + (ny <= 3 ? cy : 1);      // the row's own cell: piece (w + 1, centre)
    const int self_rel = active ? ptab_end[ps - 1] + (row - cs[w]) - ws : -1;
    const int cx = cx_first + w;
    const bool wrapflag = !((nx >= 5 && cx >= 1 && cx <= nx - 2) && (ny >= 5 && cy >= 1 && cy <= ny - 2) &&
                            (nz >= 5 && cz >= 1 && cz <= nz - 2));
    int maxlen = active ? wlen : 0, maxend = active ? ws + wlen : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(HTF_FULL, maxlen, o));
    maxend = active ? ws + maxlen : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxend = max(maxend, __shfl_xor_sync(HTF_FULL, maxend, o));
    const bool anywrap = __any_sync(HTF_FULL, active && wrapflag);

    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(stage_bar) : "memory");
    __syncthreads();
    if (maxlen == 0) return;                                    // no row of this warp belongs to the shard
    if (maxend + 8 > mblock + ROWS_SLACK) {                     // a lane would read past the sentinels (wildly uneven cells)
#ifdef HTF_DEBUG_FLAGS
        if (lane == 0 && cy < 2) printf("flag D: maxend %d mblock %d maxlen %d\n", maxend, mblock, maxlen);
#endif
        if (active) flag_row[cx / TILE] = 1;
        return;
    }

    // ---- test ----
    const unsigned cbase = cand_s + (unsigned)ws * 16u;
    float4 pi = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, 0.f);                   // inactive lanes never hit
    if (active) pi = lds_f4(cbase + (unsigned)self_rel * 16u);
    const unsigned ha0 = hl_s + (unsigned)lane * 2u, ha_dump = ha0 + (unsigned)K * 66u;
    unsigned ha;
    if (!anywrap) ha = rows_test<false, MAPPED>(p, cbase, maxlen, wlen, self_rel, pi, ha0, ha_dump);
    else ha = rows_test<true, MAPPED>(p, cbase, maxlen, wlen, self_rel, pi, ha0, ha_dump);
    const int cnt = (int)((ha - ha0) / 66u);
#ifdef HTF_DEBUG_FLAGS
    if (active && cnt >= K && cy < 2) printf("flag E: cnt %d row %d w %d ws %d wlen %d self %d cx %d cy %d cz %d ncs %d\n", cnt, row, w, ws, wlen, self_rel, cx, cy, cz, ncs);
#endif
    if (active && cnt >= K) {                                   // full or overflowing row: exact modulo-K rule in the fallback
        flag_row[cx / TILE] = 1;
        active = false;
    }
    if (active && p.count_out) p.count_out[orig - p.row_lo] = cnt;
    __syncwarp();

    // ---- emit: one row after the other, lanes = slots ----
    unsigned todo = __ballot_sync(HTF_FULL, active);
    while (todo) {
        const int r = __ffs(todo) - 1;
        todo &= todo - 1;
        const int cnt_r = __shfl_sync(HTF_FULL, cnt, r);
        const unsigned cb_r = __shfl_sync(HTF_FULL, cbase, r);
        const int self_r = __shfl_sync(HTF_FULL, self_rel, r);
        const int orig_r = __shfl_sync(HTF_FULL, orig, r);
        const bool wrap_r = __shfl_sync(HTF_FULL, wrapflag ? 1 : 0, r) != 0;
        const float4 pr = lds_f4(cb_r + (unsigned)self_r * 16u);
        const size_t orow = (size_t)(orig_r - p.row_lo);
        float4 *dst = p.out + orow * K + lane;
        unsigned la = hl_s + (unsigned)r * 2u + (unsigned)lane * 66u;
        // two slots per step, both index loads and both candidate loads issued before either is used
        for (int sl = lane; sl < K; sl += 64, dst += 64, la += 64u * 66u) {
            const bool v0 = sl < cnt_r, v1 = sl + 32 < cnt_r, in1 = sl + 32 < K;
            const unsigned c0 = v0 ? lds_u16(la) : 0u;
            const unsigned c1 = v1 ? lds_u16(la + 32u * 66u) : 0u;
            const float4 q0 = lds_f4(cb_r + c0 * 16u);
            const float4 q1 = lds_f4(cb_r + c1 * 16u);
            float dx0 = __fsub_rn(q0.x, pr.x), dy0 = __fsub_rn(q0.y, pr.y), dz0 = __fsub_rn(q0.z, pr.z);
            float dx1 = __fsub_rn(q1.x, pr.x), dy1 = __fsub_rn(q1.y, pr.y), dz1 = __fsub_rn(q1.z, pr.z);
            if (wrap_r) {
                dz0 = wrap_axis(dz0, -p.g.half[2], p.g.half[2], p.g.L[2]);
                dy0 = wrap_axis(dy0, -p.g.half[1], p.g.half[1], p.g.L[1]);
                dx0 = wrap_axis(dx0, -p.g.half[0], p.g.half[0], p.g.L[0]);
                dz1 = wrap_axis(dz1, -p.g.half[2], p.g.half[2], p.g.L[2]);
                dy1 = wrap_axis(dy1, -p.g.half[1], p.g.half[1], p.g.L[1]);
                dx1 = wrap_axis(dx1, -p.g.half[0], p.g.half[0], p.g.L[0]);
            }
            dst[0] = v0 ? make_float4(dx0, dy0, dz0, q0.w) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (in1) dst[32] = v1 ? make_float4(dx1, dy1, dz1, q1.w) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (WITH_IDX) {
                const int *ci = candidx + (cb_r - cand_s) / 16u;
                p.idx_out[orow * K + sl] = v0 ? ci[c0] : -1;
                if (in1) p.idx_out[orow * K + sl + 32] = v1 ? ci[c1] : -1;
            }
        }
    }
}

template <bool WITH_IDX, bool MAPPED>
cudaError_t launch_rows_variant(const NlistParams &p, dim3 grid, size_t smem, cudaStream_t st)
{
    static size_t configured_dev[HTF_MAX_DEVICES] = {0};   // the attribute is per device
    size_t &configured = configured_dev[htf_current_device_slot()];
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(nlist_rows_kernel<WITH_IDX, MAPPED>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    nlist_rows_kernel<WITH_IDX, MAPPED><<<grid, ROWS_PB, smem, st>>>(p);
    return cudaGetLastError();
}

size_t rows_block_bytes(int capB, int K, bool with_idx)
{
    size_t b = ROWS_HDR + (size_t)(capB + ROWS_SLACK) * 16;
    if (with_idx) b += (size_t)(capB + ROWS_SLACK) * 4;
    b += (size_t)ROWS_NW * ((((size_t)(K + 1) * 66) + 15) & ~(size_t)15);
    return b;
}

template <bool WITH_IDX, bool MAPPED>
cudaError_t launch_tile_variant(const NlistParams &p, dim3 grid, size_t smem, cudaStream_t st)
{
    static size_t configured_dev[HTF_MAX_DEVICES] = {0};   // the attribute is per device
    size_t &configured = configured_dev[htf_current_device_slot()];
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(nlist_tile_kernel<WITH_IDX, MAPPED>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    nlist_tile_kernel<WITH_IDX, MAPPED><<<grid, TILE * 32, smem, st>>>(p);
    return cudaGetLastError();
}

size_t tile_block_bytes(int capB, int capW, int K, bool with_idx)
{
    size_t b = TILE_HDR + (size_t)(capB + 32) * 16;
    if (with_idx) b += (size_t)(capB + 32) * 4;
    b += (size_t)TILE * ((size_t)RPP * capW * 2 + (((size_t)K * 2 + 15) & ~(size_t)15));
    return b;
}

template <bool WITH_IDX, bool MAPPED>
cudaError_t launch_variant(const NlistParams &p, int grid, int wpb, size_t smem, cudaStream_t st)
{
    static size_t configured_dev[HTF_MAX_DEVICES] = {0};   // per instantiation and device: largest dynamic smem opted in so far
    size_t &configured = configured_dev[htf_current_device_slot()];
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(nlist_build_kernel<WITH_IDX, MAPPED>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    nlist_build_kernel<WITH_IDX, MAPPED><<<grid, wpb * 32, smem, st>>>(p);
    return cudaGetLastError();
}

size_t per_warp_bytes(int cap, int K, bool with_idx)
{
    size_t b = (size_t)cap * 16 + (size_t)K * 16 + (size_t)RPP * cap * 2 + 256 + (((size_t)K * 2 + 15) & ~(size_t)15);
    if (with_idx) b += (size_t)cap * 4 + (((size_t)K * 4 + 15) & ~(size_t)15);
    return b;
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
    // particles per occupied cell.  With a region of interest (sharded builds) only part of the grid is
    // populated, so the density is measured once per configuration (one small kernel + a synchronising
    // 12-byte copy) instead of being derived from n / ncell; unrestricted builds need no calibration.
    double cell_mean = (double)ctx->n_binned / (double)g.ncell;
    const bool restricted = g.roi_h[0] >= 0.f || g.roi_h[1] >= 0.f || g.roi_h[2] >= 0.f;
    if (restricted) {
        const double drift = ctx->calib_n > 0 ? fabs((double)ctx->n_binned - (double)ctx->calib_n) / (double)ctx->calib_n : 1.0;
        if (!ctx->calib_valid || drift > 0.05) {
            int st3[3] = {0, 0, 0};
            cudaError_t ce = htf_cell_stats(ctx, st3, st);
            if (ce != cudaSuccess) return ce;
            ctx->calib_cell_mean = st3[1] > 0 ? (double)st3[0] / (double)st3[1] : 1.0;
            ctx->calib_n = ctx->n_binned;
            ctx->calib_valid = true;
        }
        cell_mean = ctx->calib_cell_mean;
    }
    const double mean = (double)stencil * cell_mean;
    int cap = (int)(mean + 5.0 * sqrt(mean > 1.0 ? mean : 1.0)) + 32;
    cap = (cap + 31) / 32 * 32;
    if (cap > 32768) cap = 32768;        // candidate indices are 16 bit
    const bool with_idx = idx_out != nullptr;
    const bool mapped = ctx->map_type_start >= 0;
    cudaError_t e;

    // ---- pass 1: row-per-lane kernel (grids with >= 4 cells in x), else the tile kernel ----
    const int tiles_x = (g.n[0] + TILE - 1) / TILE;
    const int ntiles = tiles_x * g.n[1] * g.n[2];
    p.use_flags = 0;
    p.tile_flag = nullptr;
    p.flag_count = nullptr;
    p.flag_count_next = nullptr;
    bool tiled = false;
#ifdef HTF_EXP_ROWS
    if (g.n[0] >= 4 && g.n[1] <= 65535 && g.n[2] <= 65535 && cap <= 32768) {
        const int nyz = min(g.n[1], 3) * min(g.n[2], 3);
        const double cm = cell_mean > 0.05 ? cell_mean : 0.05;
        int ncells = (int)ceil((double)ROWS_PB / cm) + 2;                       // cells a full segment can span (two partial ones)
        if (ncells > ROWS_MAXCELLS) ncells = ROWS_MAXCELLS;
        const double bmean = (double)(ncells + 2) * nyz * cm;
        int capB = (int)(bmean + 5.0 * sqrt(bmean > 1.0 ? bmean : 1.0)) + 32;
        capB = (capB + 31) / 32 * 32;
        // rows per x-row of cells: a lattice-like fluid puts 2x2 or 3x3 lattice lines into a cell row, so allow twice
        // the mean (blocks beyond an x-row's own segment count exit at once; a denser x-row goes to the fallback)
        const double xrow = 2.0 * (double)g.n[0] * cm;
        const int segs = (int)((xrow + 5.0 * sqrt(xrow > 1.0 ? xrow : 1.0)) / ROWS_PB) + 1;
        const size_t bytes = rows_block_bytes(capB, p.K, with_idx);
        if (capB <= 32768 && bytes <= 100 * 1024 && segs <= 65535) {
            if ((e = htf_ensure_tile_flags(ctx, ntiles)) != cudaSuccess) return e;
            if ((e = cudaMemsetAsync(ctx->d_tile_flag, 0, (size_t)ntiles, st)) != cudaSuccess) return e;
            p.cap = cap;
            p.cap_tile = capB;
            p.tile_flag = ctx->d_tile_flag;
            ctx->launches += 1;
            const dim3 rg((unsigned)segs, (unsigned)g.n[1], (unsigned)g.zcount);
            e = with_idx ? (mapped ? launch_rows_variant<true, true>(p, rg, bytes, st)
                                   : launch_rows_variant<true, false>(p, rg, bytes, st))
                         : (mapped ? launch_rows_variant<false, true>(p, rg, bytes, st)
                                   : launch_rows_variant<false, false>(p, rg, bytes, st));
            if (e != cudaSuccess) return e;
            tiled = true;
#ifdef HTF_DEBUG_FLAGS
            {
                std::vector<unsigned char> h(ntiles);
                cudaMemcpyAsync(h.data(), ctx->d_tile_flag, ntiles, cudaMemcpyDeviceToHost, st);
                cudaStreamSynchronize(st);
                int nf = 0, firstf = -1;
                for (int i = 0; i < ntiles; i++) if (h[i]) { nf++; if (firstf < 0) firstf = i; }
                static int printed = 0;
                if (printed++ < 3) printf("[rows] flagged tiles %d of %d (first %d: tx %d cy %d cz %d) capB %d segs %d\n", nf, ntiles, firstf,
                                          firstf % tiles_x, (firstf / tiles_x) % g.n[1], firstf / tiles_x / g.n[1], capB, segs);
            }
#endif
        }
    }
#endif
    if (!tiled) {
        const int ncol = g.n[0] >= 3 ? min(TILE, g.n[0]) + 2 : g.n[0];
        const double bmean = (double)ncol * min(g.n[1], 3) * min(g.n[2], 3) * cell_mean;
        int capB = (int)(bmean + 5.0 * sqrt(bmean > 1.0 ? bmean : 1.0)) + 32;
        capB = (capB + 31) / 32 * 32;
        const size_t bytes = tile_block_bytes(capB, cap, p.K, with_idx);
        if (capB <= 32768 && bytes <= 100 * 1024 && g.n[1] <= 65535 && g.n[2] <= 65535) {           // keep >= 2 blocks per SM, else per-cell only
            if ((e = htf_ensure_tile_flags(ctx, ntiles)) != cudaSuccess) return e;
            p.cap = cap;
            p.cap_tile = capB;
            p.tile_flag = ctx->d_tile_flag;
            // two counters in turn: this launch counts into one, its per-cell pass zeroes the other for the next launch
            p.flag_count = ctx->d_flag_count + (ctx->flag_parity & 1);
            p.flag_count_next = ctx->d_flag_count + ((ctx->flag_parity + 1) & 1);
            ctx->flag_parity ^= 1;
            ctx->launches += 1;
            const dim3 tg((unsigned)tiles_x, (unsigned)g.n[1], (unsigned)g.zcount);
            e = with_idx ? (mapped ? launch_tile_variant<true, true>(p, tg, bytes, st)
                                   : launch_tile_variant<true, false>(p, tg, bytes, st))
                         : (mapped ? launch_tile_variant<false, true>(p, tg, bytes, st)
                                   : launch_tile_variant<false, false>(p, tg, bytes, st));
            if (e != cudaSuccess) return e;
            tiled = true;
        }
    }

    // ---- pass 2: per-cell kernel; after the tile kernel it only walks the tiles that were flagged ----
    p.use_flags = tiled ? 1 : 0;
    int wpb = 4;
    const size_t smem_max = 200 * 1024;
    // keep the per-block footprint within the opt-in limit; shrink the staging window first
    // (the kernel re-stages in passes), then the block
    while (per_warp_bytes(cap, p.K, with_idx) * wpb > smem_max) {
        if (cap > 64) cap = max(64, cap / 2 / 32 * 32);
        else if (wpb > 1) wpb /= 2;
        else return cudaErrorInvalidValue;
    }
    p.cap = cap;
    const size_t smem = per_warp_bytes(cap, p.K, with_idx) * wpb;
    int grid = (g.n[0] * g.n[1] * g.zcount + wpb - 1) / wpb;
    if (tiled) grid = min((tiles_x * g.n[1] * g.zcount + wpb - 1) / wpb, 2 * ctx->sm_count);   // persistent flag scan
    ctx->launches += 1;
    if (with_idx) return mapped ? launch_variant<true, true>(p, grid, wpb, smem, st)
                                : launch_variant<true, false>(p, grid, wpb, smem, st);
    return mapped ? launch_variant<false, true>(p, grid, wpb, smem, st)
                  : launch_variant<false, false>(p, grid, wpb, smem, st);
}
