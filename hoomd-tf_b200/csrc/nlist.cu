// Padded neighbor tensor [rows, K, 4] from the cell-sorted positions.
//
// Replaces prepareNeighbors (/root/reference htf/TensorflowCompute.cc:304-374, CPU rule
// "skip if rsq > rc^2"), its GPU twin htf_gpu_reshape_nlist_kernel
// (htf/TensorflowCompute.cu:80-151: one thread per row, 16-byte stores strided by 16*K)
// and the cudaMemset before it (.cu:180).
//
// Two kernels.  nlist_tile2_kernel (further down) does the work: a block takes TILE y-adjacent
// cells, one warp each, stages their common neighbourhood once with TMA bulk copies, tests every
// candidate chunk against two or four rows on the packed fp32 pipe (hits become bits of per-lane
// masks in registers) and emits each row through a per-warp stage as contiguous 16-byte stores,
// zero padding included -- no memset pass, no partial-sector traffic.  nlist_build_kernel is the
// general form (one warp per cell with its own staging buffer, windows for cells of any density,
// lane-private hit lists in shared memory): it runs as a second, persistent pass over the tiles
// the tile kernel flagged as too dense for its buffers, and alone when the tile kernel does not
// apply.
//
// Arithmetic is the oracle's, bit for bit: d = p_j - p_i, compare-and-shift minimum image
// (HOOMD BoxDim::minImage CPU branch), rsq = (dx*dx + dy*dy) + dz*dz with round-to-nearest
// mul/add that are never contracted to FMA, keep iff rsq <= rc^2.  The subtractions and the
// squares use Blackwell's packed fp32 pipe (add/mul .f32x2: two rows per instruction, each
// half IEEE-rounded like the scalar op); the sums stay scalar because ptxas contracts a
// packed mul feeding a packed add into FFMA2 even with .rn.
//
// Slot order inside a row is lane-major (lane 0's hits, then lane 1's, ...; within a lane in candidate
// order in the per-cell kernel, from the last chunk to the first in the tile kernel):
// deterministic, and as unspecified as the reference's HOOMD-internal order.  When a row has
// more than K neighbors the slot index wraps modulo K and the last writer wins, exactly like
// htf/TensorflowCompute.cc:370 (with this kernel's hit order).
#include "common.cuh"

#include <math_constants.h>
#include <cstdlib>
#include <cstring>
#ifdef HTF_DEBUG_FLAGS
#include <cstdio>
#include <vector>
#endif

namespace {

#ifndef HTF_RPP
#define HTF_RPP 2
#endif
#ifndef HTF_T2_WARPS2
#define HTF_T2_WARPS2 32      // resident warps per SM the two-pair form of the tile kernel is compiled for (63 registers)
#endif
#ifndef HTF_T2_STQ
#define HTF_T2_STQ ".cs"       // cache operator of the tensor stores: streaming (A/B in the step at 1 M x 64: "" 0.5296, ".cs" 0.5266, ".cg" 0.5302, ".wt" 0.5305 ms)
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
    unsigned long long one2;   // (1.0f, 1.0f): an OPAQUE packed one (see add2_exact)
    unsigned long long inv_l2[3], neg_l2[3], magic2;   // packed (1/L, 1/L), (-L, -L) per axis (0 where n < 5), (1.5 * 2^23) x 2
    int fast_wrap;             // every dimension has >= 5 cells: minimum image by rounding (see test_window_bits)
    int map_type_start;
    int cap;             // per-warp window capacity (candidates), multiple of 32, <= 4064
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

#define HTF_EMIT_INLINE __forceinline__

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
    asm volatile("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
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

// a + b for both halves, IEEE-rounded, on the packed pipe.  ptxas contracts mul.rn.f32x2 feeding add.rn.f32x2 into
// FFMA2 (even with .rn and -fmad=false), which would change rsq in the last bit; fma(a, 1, b) with a one it cannot
// see through (a kernel parameter) is the same rounded sum and cannot absorb the multiply that produced a.
__device__ __forceinline__ f32x2 add2_exact(f32x2 a, f32x2 b, f32x2 one)
{
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(one), "l"(b));
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
// low 16 bits of a 32-bit register, no 16-bit arithmetic on the way
__device__ __forceinline__ void sts_u16_r(unsigned addr, unsigned v)
{
    asm volatile("{\n\t.reg .b16 t;\n\tcvt.u16.u32 t, %1;\n\tst.shared.b16 [%0], t;\n\t}" ::"r"(addr), "r"(v));
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

struct RowState {
    f32x2 x[RPP / 2], y[RPP / 2], z[RPP / 2];   // row pairs packed for the f32x2 pipe
    float t[RPP];                               // row types (mapped-nlist rule)
    unsigned self_addr[RPP];// shared-memory byte address of the row's own particle (or 0xffffffff)
    unsigned lp[RPP];       // shared-space byte address of this lane's next list entry for row r
};

// ---- test: append every hit to the lane-private lists ----
// List entries are 16 bits: the candidate's byte address in shared memory minus `base`.  The tile kernel keeps its
// whole block below 64 KB of shared memory and uses base = 0, so the loop variable `addr` is at once the loop
// counter, the address of the candidate and the value stored on a hit; the per-cell kernel (larger blocks) passes the
// start of the warp's window (a window is at most 4064 candidates = 65,024 bytes).
// One 32-candidate chunk against the RPP rows of the hot variant: 1 LDS.128, 3 packed subs, 3 packed squares,
// 2 packed (exact) sums, 2 compares, 2 predicated appends (store + pointer bump), loop add / compare / branch.
// SELF adds the "is this the row's own particle" compare: only the one or two chunks that hold the rows' own
// particles run that variant.  MASKED adds the "candidate lies past the end of the window" test (tiny grids).
template <bool WRAP, bool MAPPED, bool MASKED, bool SELF, bool ABS>
__device__ __forceinline__ void test_one(const NlistParams &p, const float4 c, unsigned addr, unsigned mlen_addr,
                                         unsigned base, RowState &rs, const f32x2 one)
{
    const bool pv = !MASKED || addr < mlen_addr;
    const f32x2 cx = pack2(c.x, c.x), cy = pack2(c.y, c.y), cz = pack2(c.z, c.z);
#pragma unroll
    for (int h = 0; h < RPP / 2; h++) {
        float q[2];
        const f32x2 dx2 = sub2(cx, rs.x[h]), dy2 = sub2(cy, rs.y[h]), dz2 = sub2(cz, rs.z[h]);
        if (WRAP) {
            float dx[2], dy[2], dz[2];
            unpack2(dx2, dx[0], dx[1]); unpack2(dy2, dy[0], dy[1]); unpack2(dz2, dz[0], dz[1]);
#pragma unroll
            for (int u = 0; u < 2; u++) {
                dz[u] = wrap_axis(dz[u], -p.g.half[2], p.g.half[2], p.g.L[2]);
                dy[u] = wrap_axis(dy[u], -p.g.half[1], p.g.half[1], p.g.L[1]);
                dx[u] = wrap_axis(dx[u], -p.g.half[0], p.g.half[0], p.g.L[0]);
                q[u] = __fadd_rn(__fadd_rn(__fmul_rn(dx[u], dx[u]), __fmul_rn(dy[u], dy[u])), __fmul_rn(dz[u], dz[u]));
            }
        } else {
            // (dx*dx + dy*dy) + dz*dz, every operation rounded on its own
            unpack2(add2_exact(add2_exact(mul2(dx2, dx2), mul2(dy2, dy2), one), mul2(dz2, dz2), one), q[0], q[1]);
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int r = 2 * h + u;
            // rsq <= rc2 is !(rsq > rc2) for every non-NaN rsq; +inf rows / sentinels give inf or NaN -> no hit
            bool hit = q[u] <= p.rc2;
            if (SELF) hit = hit & (addr != rs.self_addr[r]);
            if (MASKED) hit = hit & pv;
            if (MAPPED) hit = hit && (((int)c.w >= p.map_type_start) == ((int)rs.t[r] >= p.map_type_start));
            if (hit) {
                sts_u16_r(rs.lp[r], ABS ? addr : addr - base);
                rs.lp[r] += 64u;                    // lists are [k][lane] u16: next k is 32 entries on
            }
        }
    }
}

template <bool WRAP, bool MAPPED, bool MASKED, bool SELF, bool ABS>
__device__ __forceinline__ void test_range(const NlistParams &p, unsigned addr, unsigned addr_end, unsigned mlen_addr,
                                           unsigned base, RowState &rs)
{
    const f32x2 one = p.one2;
    // keep `addr` ONE induction variable: hidden from the constant-offset splitting of the front end, which otherwise
    // carries (addr - header) and re-adds the header size three times per iteration
    asm volatile("mov.u32 %0, %0;" : "+r"(addr));
#pragma unroll 1
    for (; addr < addr_end; addr += 512u) {
        const float4 c = lds_f4_ro(addr);
        test_one<WRAP, MAPPED, MASKED, SELF, ABS>(p, c, addr, mlen_addr, base, rs, one);
    }
}

// the whole window [0, mround) (candidates, starting at shared address cand_ws) for this lane; the chunks that hold
// the rows' own particles run the SELF variant, all others skip that compare
template <bool WRAP, bool MAPPED, bool MASKED, bool ABS>
__device__ __forceinline__ void test_window(const NlistParams &p, unsigned cand_ws, int mround, int mlen, RowState &rs,
                                            int lane)
{
    unsigned smin = 0xffffffffu, smax = 0u;
    bool any = false;
#pragma unroll
    for (int r = 0; r < RPP; r++)
        if (rs.self_addr[r] != 0xffffffffu) { smin = min(smin, rs.self_addr[r]); smax = max(smax, rs.self_addr[r]); any = true; }
    const unsigned l16 = (unsigned)lane * 16u, end = cand_ws + (unsigned)mround * 16u, mlen_addr = cand_ws + (unsigned)mlen * 16u;
    unsigned a = end, b = end;                          // chunks [a, b) (addresses of chunk starts) hold a row's own particle
    if (any) { a = cand_ws + ((smin - cand_ws) & ~511u); b = cand_ws + ((smax - cand_ws) & ~511u) + 512u; }
    test_range<WRAP, MAPPED, MASKED, false, ABS>(p, cand_ws + l16, a, mlen_addr, cand_ws, rs);
    test_range<WRAP, MAPPED, MASKED, true, ABS>(p, a + l16, min(b, end), mlen_addr, cand_ws, rs);
    test_range<WRAP, MAPPED, MASKED, false, ABS>(p, b + l16, end, mlen_addr, cand_ws, rs);
}

// ---- emit one row whose hits all come from the single staged window ----
// Scan of the lane counts -> slot ranges; "slot -> candidate" published in a small shared map; then lane s
// re-derives d for slots s, s+32, ... and stores them coalesced, zero padding included.
template <bool WITH_IDX>
__device__ HTF_EMIT_INLINE void emit_single_window(const NlistParams &p, unsigned cand_s, unsigned idx_base,
                                                   const int *candidx, unsigned slotmap_s, unsigned list_s, int c_l,
                                                   bool wrap, const float4 &pi, int orig, int lane)
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
            const unsigned ci = lds_u16(slotmap_s + 2u * sl);       // byte offset into the window
            const float4 cd = lds_f4(cand_s + ci);
            float dx = __fsub_rn(cd.x, pi.x), dy = __fsub_rn(cd.y, pi.y), dz = __fsub_rn(cd.z, pi.z);
            if (wrap) {
                dz = wrap_axis(dz, -p.g.half[2], p.g.half[2], p.g.L[2]);
                dy = wrap_axis(dy, -p.g.half[1], p.g.half[1], p.g.L[1]);
                dx = wrap_axis(dx, -p.g.half[0], p.g.half[0], p.g.L[0]);
            }
            v = make_float4(dx, dy, dz, cd.w);
            if (WITH_IDX) vi = candidx[(ci - idx_base) >> 4];
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

constexpr int TILE = HTF_TILE;    // cells per block along x in the tile kernel (= warps per block)
static_assert(TILE * 32 >= (TILE + 2) * 9 && TILE + 3 <= 24, "tile size");

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
                rs.self_addr[r] = ok ? 0u : 0xffffffffu;             // refined per window below
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
        for (int r = 0; r < RPP; r++) rvalid[r] = rs.self_addr[r] == 0u;

        int cnt_acc = 0;                                             // multi-window mode: hits of the row so far
        for (int pass = 0; pass < npass; pass++) {
            if (npass > 1) { __syncwarp(); mround = stage(pass); }
#pragma unroll
            for (int r = 0; r < RPP; r++) {
                const int rel = self_base + s0 + r - pass * cap;       // own particle inside this window?
                rs.self_addr[r] = (rvalid[r] && rel >= 0 && rel < mround) ? cand_s + (unsigned)rel * 16u : 0xffffffffu;
                rs.lp[r] = lists_s + (unsigned)(r * cap + lane) * 2u;
            }
            if (wrap) test_window<true, MAPPED, false, false>(p, cand_s, mround, mround, rs, lane);
            else test_window<false, MAPPED, false, false>(p, cand_s, mround, mround, rs, lane);
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
                    emit_single_window<WITH_IDX>(p, cand_s, 0u, candidx, slotmap_s, list_s, c_l, wrap, pi, orig, lane);
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
                        const float4 cd = cand[ci >> 4];
                        float dx = __fsub_rn(cd.x, pi.x), dy = __fsub_rn(cd.y, pi.y), dz = __fsub_rn(cd.z, pi.z);
                        dz = wrap_axis(dz, -p.g.half[2], p.g.half[2], p.g.L[2]);     // a no-op for interior cells
                        dy = wrap_axis(dy, -p.g.half[1], p.g.half[1], p.g.L[1]);
                        dx = wrap_axis(dx, -p.g.half[0], p.g.half[0], p.g.L[0]);
                        v = make_float4(dx, dy, dz, cd.w);
                        if (WITH_IDX) vi = candidx[ci >> 4];
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
    const int nx = p.g.n[0], ny = p.g.n[1];
    const int tiles_t = (ny + TILE - 1) / TILE;
    const int per_layer = tiles_t * nx;
    const int ntiles_win = per_layer * p.g.zcount;                  // tiles of the z-window only
    for (int tw = blockIdx.x * wpb + warp; tw < ntiles_win; tw += gridDim.x * wpb) {
        const int lz = tw / per_layer;
        const int cz = (p.g.z0 + lz) % p.g.n[2];
        const int tile = cz * per_layer + (tw - lz * per_layer);
        if (!p.tile_flag[tile]) continue;
        // tile = (cz * tiles_y + ty) * nx + cx: cells (cx, ty * TILE + c, cz)
        const int cx = tile % nx, ty = (tile / nx) % tiles_t;
        for (int c = 0; c < TILE && ty * TILE + c < ny; c++)
            build_cell<WITH_IDX, MAPPED>(p, (cz * ny + ty * TILE + c) * nx + cx, smem_raw);
    }
}

// ------------------------------------------------------------------------------------------------
// Tile kernel (second form; the first one -- tiles along x, one bulk copy per stencil cell, hits appended to
// lane-private lists in shared memory -- needed 339 warp instructions per row against 255).  One block = TILE
// y-adjacent cells (one warp each) whose shared neighbourhood is staged once by TMA:
//   * the tile runs along y, so that the three x cells of a stencil row are ONE contiguous run of the
//     cell-sorted array: (TILE+2) x 3 bulk copies per block instead of (TILE+2) x 9;
//   * hits are not appended to lists in shared memory: every lane keeps one 32-bit mask per row, bit k =
//     "my candidate of chunk k is a neighbor" (a window is at most 32 chunks = 1024 candidates, larger
//     ones go to the per-cell kernel).  The test loop has no self test and no shared-memory store; the
//     row's own particle is one bit cleared afterwards;
//   * emit: after the scan of the lane counts, every lane walks its own bits, derives d for each of its
//     hits (conflict-free: lane l only reads column l of the staged chunks) and writes the finished
//     (dx,dy,dz,type) to its slots of a per-warp row stage; the row then leaves as coalesced 16-byte
//     stores, zero padding included.  No slot map, no gather with bank conflicts.
// Slot order inside a row: lane-major, within a lane from the last chunk to the first.
constexpr int NP2 = ((TILE + 2) * 9 + 31) / 32 * 32;   // piece table capacity of the second form (64 for TILE = 4)
constexpr int TILE2_HDR = (2 * NP2 + 32) * 4;
static_assert((TILE + 2) * 9 <= NP2, "tile size");
// NPB = row pairs tested per candidate load.  2 halves the shared-memory traffic and the loop overhead of the test
// phase per row but needs 63 registers (8 blocks per SM instead of 10): measured 1.90 vs 2.14 ms at 4 M x 96 (windows
// of 16 chunks), 0.373 vs 0.377 ms at 1 M x 64 (10 chunks).

__device__ __forceinline__ float4 lds_f4_v(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f4_v(unsigned addr, float x, float y, float z, float w)
{
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void sts_i32_v(unsigned addr, int v)
{
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned bfind_u32(unsigned m)
{
    unsigned k;
    asm("bfind.u32 %0, %1;" : "=r"(k) : "r"(m));
    return k;
}

// m |= bit under a predicate (one predicated LOP3; the plain C form compiles to SEL + LOP3)
__device__ __forceinline__ void or_if(unsigned &m, const bool h, const unsigned bit)
{
    asm("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %1, 0;\n\t@p or.b32 %0, %0, %2;\n\t}" : "+r"(m) : "r"((unsigned)h), "r"(bit));
}

struct RowPair {
    f32x2 x, y, z;          // the two rows packed for the f32x2 pipe
    float t[2];             // row types (mapped-nlist rule)
};

// One lane's share of the window [addr, end): candidate chunk k sets bit k of m[2h] / m[2h+1] when it is within the
// cutoff of the rows of pair h.  Per chunk and pair: 3 packed subs, 3 packed squares, 2 packed (exact) sums,
// 2 compares, 2 predicated ORs; per chunk: 1 LDS.128, shift, add, compare, branch.  NP = 2 tests four rows per
// candidate load: half the shared-memory traffic per row.
template <bool WRAP, bool MAPPED, bool MASKED, int NP>
__device__ __forceinline__ void test_window_bits(const NlistParams &p, unsigned addr, const unsigned end,
                                                 const unsigned mlen_addr, const RowPair (&rp)[NP], const f32x2 one,
                                                 const float rc2, unsigned (&m)[2 * NP])
{
    unsigned bit = 1u;
    asm volatile("mov.u32 %0, %0;" : "+r"(addr));
#pragma unroll 1
    for (; addr < end; addr += 512u, bit <<= 1) {
        const float4 c = lds_f4_ro(addr);
        const f32x2 cx = pack2(c.x, c.x), cy = pack2(c.y, c.y), cz = pack2(c.z, c.z);
#pragma unroll
        for (int h = 0; h < NP; h++) {
            const f32x2 dx2 = sub2(cx, rp[h].x), dy2 = sub2(cy, rp[h].y), dz2 = sub2(cz, rp[h].z);
            float q[2];
            if (WRAP && p.fast_wrap) {
                // With >= 5 cells per dimension a stencil candidate is either less than 2 cell edges away (no shift)
                // or, reached through the periodic boundary, at least n - 2 >= 3 edges away (one shift) -- never near
                // L/2 = n/2 edges.  Then d - L * rint(d / L) is bit-identical to HOOMD's compare-and-shift, and it
                // runs on the packed pipe: k = (d/L + 1.5 * 2^23) - 1.5 * 2^23, d' = fma(k, -L, d) (k L is exact, so
                // the fma rounds once, like the subtraction).
                f32x2 dd[3] = {dx2, dy2, dz2};
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    f32x2 k = add2_exact(mul2(dd[a], p.inv_l2[a]), p.magic2, one);
                    k = sub2(k, p.magic2);
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd[a]) : "l"(k), "l"(p.neg_l2[a]));
                }
                unpack2(add2_exact(add2_exact(mul2(dd[0], dd[0]), mul2(dd[1], dd[1]), one), mul2(dd[2], dd[2]), one), q[0], q[1]);
            } else if (WRAP) {
                float dx[2], dy[2], dz[2];
                unpack2(dx2, dx[0], dx[1]); unpack2(dy2, dy[0], dy[1]); unpack2(dz2, dz[0], dz[1]);
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    dz[u] = wrap_axis(dz[u], -p.g.half[2], p.g.half[2], p.g.L[2]);
                    dy[u] = wrap_axis(dy[u], -p.g.half[1], p.g.half[1], p.g.L[1]);
                    dx[u] = wrap_axis(dx[u], -p.g.half[0], p.g.half[0], p.g.L[0]);
                    q[u] = __fadd_rn(__fadd_rn(__fmul_rn(dx[u], dx[u]), __fmul_rn(dy[u], dy[u])), __fmul_rn(dz[u], dz[u]));
                }
            } else {
                unpack2(add2_exact(add2_exact(mul2(dx2, dx2), mul2(dy2, dy2), one), mul2(dz2, dz2), one), q[0], q[1]);
            }
            // rsq <= rc2 is !(rsq > rc2) for every non-NaN rsq; +inf rows / sentinels give inf or NaN -> no hit
            bool h0 = q[0] <= rc2, h1 = q[1] <= rc2;
            if (MASKED) { const bool pv = addr < mlen_addr; h0 = h0 & pv; h1 = h1 & pv; }
            if (MAPPED) {
                const bool cm = (int)c.w >= p.map_type_start;
                h0 = h0 && (cm == ((int)rp[h].t[0] >= p.map_type_start));
                h1 = h1 && (cm == ((int)rp[h].t[1] >= p.map_type_start));
            }
            or_if(m[2 * h], h0, bit);
            or_if(m[2 * h + 1], h1, bit);
        }
    }
}

// This lane's hits of one row -> (d, type) into the lane's slots of the row stage, from the highest set bit of m
// down.  Branch-free steps: a lane that has run out of bits computes on a harmless address (bfind(0) = -1 puts it
// 512 bytes below the lane's first candidate, inside the block's header) and only its store is predicated off.
template <bool WITH_IDX, bool WRAP>
__device__ __forceinline__ void emit_own_hits(const NlistParams &p, unsigned m, unsigned qa, unsigned qi,
                                              const unsigned lane_cand, const unsigned cand_s, const unsigned candidx_s,
                                              const float px, const float py, const float pz)
{
    auto step = [&]() {
        const bool act = m != 0u;
        const unsigned k = bfind_u32(m);
        unsigned bitk;
        asm("shl.b32 %0, 1, %1;" : "=r"(bitk) : "r"(k));        // 0 when k = 0xffffffff (shift counts clamp at 32)
        m ^= bitk;
        const unsigned a = lane_cand + (k << 9);
        const float4 cd = lds_f4_ro(a);
        float dx = __fsub_rn(cd.x, px), dy = __fsub_rn(cd.y, py), dz = __fsub_rn(cd.z, pz);
        if (WRAP) {
            dz = wrap_axis(dz, -p.g.half[2], p.g.half[2], p.g.L[2]);
            dy = wrap_axis(dy, -p.g.half[1], p.g.half[1], p.g.L[1]);
            dx = wrap_axis(dx, -p.g.half[0], p.g.half[0], p.g.L[0]);
        }
        if (act) sts_f4_v(qa, dx, dy, dz, cd.w);
        if (WITH_IDX) {
            if (act) sts_i32_v(qi, lds_i32(candidx_s + ((a - cand_s) >> 2)));
            qi += 4u;
        }
        qa += 16u;
    };
    if (!WITH_IDX && !WRAP) {
        // the same four steps spelled out: in the C form the predicated store becomes a branch per step and the
        // slot address is re-derived inside each of them
#pragma unroll
        for (int j = 0; j < 4; j++)
            asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 k, b, a;\n\t.reg .f32 x, y, z, w;\n\t"
                         "setp.ne.u32 p, %0, 0;\n\t"
                         "bfind.u32 k, %0;\n\t"
                         "shl.b32 b, 1, k;\n\t"
                         "xor.b32 %0, %0, b;\n\t"
                         "shl.b32 a, k, 9;\n\t"
                         "add.u32 a, a, %2;\n\t"
                         "ld.shared.v4.f32 {x, y, z, w}, [a];\n\t"
                         "sub.rn.f32 x, x, %3;\n\t"
                         "sub.rn.f32 y, y, %4;\n\t"
                         "sub.rn.f32 z, z, %5;\n\t"
                         "@p st.shared.v4.f32 [%1], {x, y, z, w};\n\t"
                         "add.u32 %1, %1, 16;\n\t}"
                         : "+r"(m), "+r"(qa) : "r"(lane_cand), "f"(px), "f"(py), "f"(pz) : "memory");
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) step();
    }
    while (__any_sync(HTF_FULL, m != 0u)) step();
}

// A row with more than K hits: hit number q lands in slot q mod K and only the last K hits are written
// (htf/TensorflowCompute.cc:370, "the last writer of a slot wins").  Cold path.
template <bool WITH_IDX>
__device__ __forceinline__ void emit_own_hits_over(const unsigned K, const float3 half, const float3 L, unsigned m, unsigned q, const unsigned first,
                                                const unsigned stage_s, const unsigned istage_s, const unsigned lane_cand,
                                                const unsigned cand_s, const unsigned candidx_s, const float px,
                                                const float py, const float pz)
{
    while (m) {
        const unsigned k = bfind_u32(m);
        m ^= 1u << k;
        const unsigned a = lane_cand + (k << 9);
        const float4 cd = lds_f4_ro(a);
        float dx = __fsub_rn(cd.x, px), dy = __fsub_rn(cd.y, py), dz = __fsub_rn(cd.z, pz);
        dz = wrap_axis(dz, -half.z, half.z, L.z);     // a no-op for interior cells
        dy = wrap_axis(dy, -half.y, half.y, L.y);
        dx = wrap_axis(dx, -half.x, half.x, L.x);
        if (q >= first) {
            const unsigned slot = q % K;
            sts_f4_v(stage_s + slot * 16u, dx, dy, dz, cd.w);
            if (WITH_IDX) sts_i32_v(istage_s + slot * 4u, lds_i32(candidx_s + ((a - cand_s) >> 2)));
        }
        q++;
    }
}

template <bool WITH_IDX, bool MAPPED, int KC, int NPB>
__global__ void __launch_bounds__(TILE * 32, (NPB == 2 ? HTF_T2_WARPS2 : 40) / TILE) nlist_tile2_kernel(const NlistParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int capB = p.cap_tile;
    int *ptab_end = reinterpret_cast<int *>(smem_raw);          // [NP2] inclusive prefix of piece lengths
    int *ptab_adj = ptab_end + NP2;                           // [NP2] source slot - staged index
    int *colstart = ptab_adj + NP2;                           // [<= TILE+3] staged offset of each column
    int *wtot = colstart + 24;                                  // [4] piece-length totals of warps 0..3
    float4 *cand = reinterpret_cast<float4 *>(smem_raw + TILE2_HDR);
    // rows in the per-warp stage.  One: a two-row stage (both rows of a pair emitted between one pair of warp barriers)
    // is no faster at K = 64 (0.363 vs 0.361 ms) and at K = 96 (3 KB per warp) it held the kernel at 6 blocks per SM
    constexpr unsigned RS = 1u;
    const unsigned istage_bytes = WITH_IDX ? (unsigned)(((size_t)RS * (KC ? KC : p.K) * 4 + 15) & ~(size_t)15) : 0u;
    const unsigned per_warp = RS * (unsigned)(KC ? KC : p.K) * 16u + istage_bytes;
    // one opaque base register: otherwise every shared address below is re-derived from SR_CgaCtaId where it is used
    unsigned smem_s = (unsigned)__cvta_generic_to_shared(smem_raw);
    asm volatile("mov.u32 %0, %0;" : "+r"(smem_s));
    const unsigned cand_s = smem_s + (unsigned)TILE2_HDR;
    const unsigned candidx_s = cand_s + (unsigned)(capB + 32) * 16u;
    unsigned stage_s = (WITH_IDX ? candidx_s + (unsigned)(capB + 32) * 4u : candidx_s) + per_warp * (unsigned)warp;   // [RS][K] float4
    asm volatile("mov.u32 %0, %0;" : "+r"(stage_s));
    const unsigned istage_s = stage_s + RS * (unsigned)(KC ? KC : p.K) * 16u;                                      // [RS][K] int

    const int nx = p.g.n[0], ny = p.g.n[1], nz = p.g.n[2];
    const int tiles_y = gridDim.y;                              // = ceil(ny / TILE)
    const int cx = blockIdx.x, ty = blockIdx.y;
    const int cz = (p.g.z0 + (int)blockIdx.z) % nz;            // grid.z = cell layers of the z-window
    const int bid = (cz * tiles_y + ty) * nx + cx;
    const int cy0 = ty * TILE;
    const int nact = min(TILE, ny - cy0);                       // cells of this tile
    const int nlx = min(nx, 3), nlz = min(nz, 3);
    const bool ys = ny >= 3;                                    // y stencil = {c-1, c, c+1}; else every y cell
    const int ncol = ys ? nact + 2 : ny;
    const bool merge = nx >= 3 && cx >= 1 && cx <= nx - 2;      // cells cx-1..cx+1 are one contiguous run
    const int ppc = merge ? nlz : nlz * nlx;                    // pieces per column
    const int npieces = ncol * ppc;                             // <= (TILE + 2) * 9

    // ---- piece table: prefix sums of the piece lengths give the staged (column-major) layout ----
    int pl = 0, pb = 0;
    if (tid < npieces) {
        const int col = tid / ppc, j = tid - col * ppc;
        const int jz = merge ? j : j / nlx, jx = merge ? 0 : j - jz * nlx;
        int sy = ys ? cy0 - 1 + col : col;
        sy = sy < 0 ? sy + ny : (sy >= ny ? sy - ny : sy);
        int sz = nz <= 3 ? jz : cz + jz - 1;
        sz = sz < 0 ? sz + nz : (sz >= nz ? sz - nz : sz);
        int sx = merge ? cx - 1 : (nx <= 3 ? jx : cx + jx - 1);
        sx = sx < 0 ? sx + nx : (sx >= nx ? sx - nx : sx);
        const int c0 = (sz * ny + sy) * nx + sx;
        pb = __ldg(p.cell_start + c0);
        pl = __ldg(p.cell_start + c0 + (merge ? 3 : 1)) - pb;
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
    if (tid < NP2) {
        ptab_end[tid] = incl;
        ptab_adj[tid] = pb - (incl - pl);
        if (tid < npieces && tid % ppc == 0) colstart[tid / ppc] = incl - pl;
        if (tid == npieces - 1) colstart[ncol] = incl;
    }
    __syncthreads();
    const int mblock = colstart[ncol];
    bool fits = mblock <= capB;
    for (int w = 0; w < nact; w++) {
        const int wl = ys ? colstart[w + 3] - colstart[w] : mblock;
        fits = fits && wl <= 1024;                              // 32 chunks: one mask bit each
    }
    if (tid == 0) {
        p.tile_flag[bid] = fits ? 0 : 1;
        if (!fits && p.flag_count) atomicAdd(p.flag_count, 1);
    }
    if (!fits) return;                                          // block-uniform
    const bool full = (p.row_lo == 0 && p.row_hi == p.n_all);
    const int cy = cy0 + warp;
    const int cell = (cz * ny + min(cy, ny - 1)) * nx + cx;
    int b = 0, e = 0;
    if (warp < nact) { b = __ldg(p.cell_start + cell); e = __ldg(p.cell_start + cell + 1); }
    // the rows of this warp's cell: lane l holds the particle index of row b + l (-1: not a row of this launch)
    int my_o = -1;
    if (b + lane < e) {
        my_o = __ldg(p.sorted_idx + b + lane);
        if (!full && (my_o < p.row_lo || my_o >= p.row_hi)) my_o = -1;
    }
    if (!full) {
        // sharded build: skip the tile (before staging anything) when none of its cells holds a local row
        bool any = my_o >= 0;
        for (int s = b + 32 + lane; s < e; s += 32) {
            const int o = __ldg(p.sorted_idx + s);
            any |= (o >= p.row_lo && o < p.row_hi);
        }
        if (!__syncthreads_or(any)) return;
    }

    // ---- stage the whole neighbourhood once: one TMA bulk copy per piece, issued by the thread that owns the
    //      piece's table entry; completion is counted in bytes on one mbarrier ----
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
    if (warp == 0) {
        cand[mblock + lane] = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, 0.f);
        // one warp polls the barrier, the others sleep in the block barrier below
        asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
                     "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(stage_bar) : "memory");
    }
    __syncthreads();
    if (warp != 0)      // completed by now: observing the phase makes the bulk copies visible to this thread too
        asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
                     "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(stage_bar) : "memory");

    // ---- one warp per cell of the tile ----
    if (warp >= nact || e == b) return;
    if (!full) {
        bool any = my_o >= 0;
        for (int s = b + 32 + lane; s < e; s += 32) {
            const int o = __ldg(p.sorted_idx + s);
            any |= (o >= p.row_lo && o < p.row_hi);
        }
        if (!__any_sync(HTF_FULL, any)) return;
    }
    const bool wrap = !((nx >= 5 && cx >= 1 && cx <= nx - 2) && (ny >= 5 && cy >= 1 && cy <= ny - 2) &&
                        (nz >= 5 && cz >= 1 && cz <= nz - 2));
    const int ws = ys ? colstart[warp] : 0;
    const int mlen_true = (ys ? colstart[warp + 3] : mblock) - ws;
    const int mround = (mlen_true + 31) & ~31;
    // With >= TILE + 2 cells along the tile axis, whatever follows the window in the buffer is the column two
    // cells away (or the sentinels): farther than r_cut by construction of the grid, so the "past the end of the
    // window" mask is unnecessary.  Smaller grids alias the periodic image of the warp's own stencil there.
    const bool masked = ny < TILE + 2;
    // staged position of this cell's own particles
    const int ps = (ys ? warp + 1 : cy) * ppc + (nz <= 3 ? cz : 1) * (merge ? 1 : nlx) + (merge ? 0 : (nx <= 3 ? cx : 1));
    const int self_base = -ptab_adj[ps] - ws;                   // row slot s sits at window index self_base + s
    const unsigned cand_ws = cand_s + (unsigned)ws * 16u;
    unsigned lane_cand = cand_ws + (unsigned)lane * 16u;
    asm volatile("mov.u32 %0, %0;" : "+r"(lane_cand));
    const unsigned wend = cand_ws + (unsigned)mround * 16u, mlen_addr = cand_ws + (unsigned)mlen_true * 16u;

    // the test loop's two constants as per-thread registers (through a shuffle, so that ptxas cannot prove them
    // uniform): as uniform-register operands they are re-loaded from the constant bank in every iteration
    f32x2 one_v = p.one2;
    float rc2_v = p.rc2;
#ifndef HTF_T2_UCONST
    {
        unsigned lo = (unsigned)one_v, hi = (unsigned)(one_v >> 32);
        lo = __shfl_sync(HTF_FULL, lo, 0); hi = __shfl_sync(HTF_FULL, hi, 0);
        one_v = (f32x2)lo | ((f32x2)hi << 32);
        rc2_v = __shfl_sync(HTF_FULL, rc2_v, 0);
    }
#endif

    const int K = KC ? KC : p.K;
    unsigned st_lane = stage_s + (unsigned)lane * 16u;           // this lane's slot of row 0 in the row stage
    const unsigned ist_lane = istage_s + (unsigned)lane * 4u;
    const unsigned rowbytes = (unsigned)K * 16u;
    const bool need_count = p.count_out != nullptr;
    unsigned long long out_lane_a = (unsigned long long)__cvta_generic_to_global(p.out + lane);
    // opaque copies: under the register cap ptxas otherwise re-derives these from tid / the parameters at every use
    asm volatile("mov.u32 %0, %0;" : "+r"(st_lane));
    asm volatile("mov.u64 %0, %0;" : "+l"(out_lane_a));

    for (int sb = b; sb < e; sb += 32) {
        if (sb != b) {                                          // cells with more than 32 rows (rare): next 32 indices
            my_o = -1;
            if (sb + lane < e) {
                my_o = __ldg(p.sorted_idx + sb + lane);
                if (!full && (my_o < p.row_lo || my_o >= p.row_hi)) my_o = -1;
            }
        }
        const int nrow = min(32, e - sb);
        int my_cnt = 0;                                         // neighbor count of row sb + lane
        unsigned pa = cand_ws + (unsigned)(self_base + sb) * 16u;        // shared address of row sb's own particle
        for (int r0 = 0; r0 < nrow; r0 += 2 * NPB, pa += 32u * NPB) {
            int o[2 * NPB];
            int oall = -1;
#pragma unroll
            for (int j = 0; j < 2 * NPB; j++) { o[j] = __shfl_sync(HTF_FULL, my_o, r0 + j); oall &= o[j]; }   // r0 + j <= 31
            if (oall < 0) continue;                             // no row of this batch belongs to this launch
            RowPair rp[NPB];
            unsigned clr[2 * NPB], m[2 * NPB];
#pragma unroll
            for (int h = 0; h < NPB; h++) {
                // rows past the end of the cell repeat the batch's first row; they test as +inf like every row that is
                // not emitted.  The selects double as the moves that pair the rows' coordinates in adjacent
                // registers; plain moves get re-materialised inside the test loop.
                const unsigned pa0 = pa + ((r0 + 2 * h < nrow) ? 32u * h : 0u);
                const unsigned pa1 = pa + ((r0 + 2 * h + 1 < nrow) ? 32u * h + 16u : 0u);
                const float4 pi0 = lds_f4_ro(pa0), pi1 = lds_f4_ro(pa1);
                const bool v0 = o[2 * h] >= 0, v1 = o[2 * h + 1] >= 0;
                rp[h].x = pack2_pinned(v0 ? pi0.x : CUDART_INF_F, v1 ? pi1.x : CUDART_INF_F);
                rp[h].y = pack2_pinned(v0 ? pi0.y : CUDART_INF_F, v1 ? pi1.y : CUDART_INF_F);
                rp[h].z = pack2_pinned(v0 ? pi0.z : CUDART_INF_F, v1 ? pi1.z : CUDART_INF_F);
                rp[h].t[0] = pi0.w; rp[h].t[1] = pi1.w;
                // The rows' own particles pass the test (d = 0): their bits are cleared afterwards.  Lane l holds the
                // candidates at lane_cand + 512 k; pa - lane_cand rotated right by 9 is the chunk number when the
                // difference is a multiple of 512 and a number >= 2^27 otherwise -- and shl clamps shift counts at
                // 32, so only the owning lane gets a bit.
                asm("{\n\t.reg .b32 d, s;\n\tsub.u32 d, %2, %4;\n\tshf.r.wrap.b32 s, d, d, 9;\n\tshl.b32 %0, 1, s;\n\t"
                    "sub.u32 d, %3, %4;\n\tshf.r.wrap.b32 s, d, d, 9;\n\tshl.b32 %1, 1, s;\n\t}"
                    : "=r"(clr[2 * h]), "=r"(clr[2 * h + 1]) : "r"(pa0), "r"(pa1), "r"(lane_cand));
                m[2 * h] = 0u; m[2 * h + 1] = 0u;
            }
            if (NPB == 2 && r0 + 2 >= nrow) {
                // only the first pair holds rows: test it alone
                RowPair rp1[1] = {rp[0]};
                unsigned m1[2] = {0u, 0u};
                if (!wrap) test_window_bits<false, MAPPED, false, 1>(p, lane_cand, wend, mlen_addr, rp1, one_v, rc2_v, m1);
                else if (!masked) test_window_bits<true, MAPPED, false, 1>(p, lane_cand, wend, mlen_addr, rp1, one_v, rc2_v, m1);
                else test_window_bits<true, MAPPED, true, 1>(p, lane_cand, wend, mlen_addr, rp1, one_v, rc2_v, m1);
                m[0] = m1[0]; m[1] = m1[1];
            } else {
                if (!wrap) test_window_bits<false, MAPPED, false, NPB>(p, lane_cand, wend, mlen_addr, rp, one_v, rc2_v, m);
                else if (!masked) test_window_bits<true, MAPPED, false, NPB>(p, lane_cand, wend, mlen_addr, rp, one_v, rc2_v, m);
                else test_window_bits<true, MAPPED, true, NPB>(p, lane_cand, wend, mlen_addr, rp, one_v, rc2_v, m);
            }

            // ---- emit, pair by pair (the row stage holds two rows).  Lane counts of a pair share one 32-bit scan. ----
            unsigned ex[NPB], all[NPB];
#pragma unroll
            for (int h = 0; h < NPB; h++) {
                m[2 * h] &= ~clr[2 * h]; m[2 * h + 1] &= ~clr[2 * h + 1];
                const unsigned packed = (unsigned)__popc(m[2 * h]) | ((unsigned)__popc(m[2 * h + 1]) << 16);
                unsigned inc = packed;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1)
                    asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tshfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n\t@p add.u32 %0, %0, t;\n\t}" : "+r"(inc) : "r"(d));
                all[h] = __shfl_sync(HTF_FULL, inc, 31);
                ex[h] = inc - packed;
            }
#pragma unroll
            for (int h = 0; h < NPB; h++) {
                if ((o[2 * h] & o[2 * h + 1]) < 0) continue;    // warp-uniform
                float px[2], py[2], pz[2];                      // the register halves of the packed rows: no moves
                unpack2(rp[h].x, px[0], px[1]); unpack2(rp[h].y, py[0], py[1]); unpack2(rp[h].z, pz[0], pz[1]);
                const unsigned tot[2] = {all[h] & 0xffffu, all[h] >> 16};
#pragma unroll
                for (int g = 0; g < 2 / (int)RS; g++) {         // RS = 2: both rows at once; RS = 1: row by row
                __syncwarp();                                   // the previous slot phase has left the row stage
#pragma unroll
                for (int rr = 0; rr < (int)RS; rr++) {
                    const int r = g * (int)RS + rr;
                    if (o[2 * h + r] < 0) continue;             // warp-uniform
                    const unsigned q = r ? (ex[h] >> 16) : (ex[h] & 0xffffu);
                    const unsigned st_r = stage_s + (rr ? rowbytes : 0u), ist_r = istage_s + (rr ? (unsigned)K * 4u : 0u);
                    if (tot[r] <= (unsigned)K) {
                        if (!wrap) emit_own_hits<WITH_IDX, false>(p, m[2 * h + r], st_r + q * 16u, ist_r + q * 4u, lane_cand, cand_s, candidx_s, px[r], py[r], pz[r]);
                        else emit_own_hits<WITH_IDX, true>(p, m[2 * h + r], st_r + q * 16u, ist_r + q * 4u, lane_cand, cand_s, candidx_s, px[r], py[r], pz[r]);
                    } else {
                        emit_own_hits_over<WITH_IDX>((unsigned)K, make_float3(p.g.half[0], p.g.half[1], p.g.half[2]),
                                                     make_float3(p.g.L[0], p.g.L[1], p.g.L[2]), m[2 * h + r], q, tot[r] - (unsigned)K,
                                                     st_r, ist_r, lane_cand, cand_s, candidx_s, px[r], py[r], pz[r]);
                    }
                }
                __syncwarp();
#pragma unroll
                for (int rr = 0; rr < (int)RS; rr++) {
                    const int r = g * (int)RS + rr;
                    const int orig = o[2 * h + r];
                    if (orig < 0) continue;                      // warp-uniform
                    const int total = (int)tot[r];
                    const int nvalid = min(total, K);
                    const unsigned row = (unsigned)(orig - p.row_lo);
                    const unsigned st_r = st_lane + (rr ? rowbytes : 0u), ist_r = ist_lane + (rr ? (unsigned)K * 4u : 0u);
                    if (KC && !WITH_IDX) {
                        const unsigned long long dsta = out_lane_a + (unsigned long long)row * (unsigned long long)(KC * 16);
#pragma unroll
                        for (int i = 0; i < (KC ? KC / 32 : 1); i++)
                            asm volatile("{\n\t.reg .pred p;\n\t.reg .f32 x, y, z, w;\n\t"
                                         "setp.lt.s32 p, %2, %3;\n\t"
                                         "mov.f32 x, 0f00000000;\n\tmov.f32 y, 0f00000000;\n\tmov.f32 z, 0f00000000;\n\tmov.f32 w, 0f00000000;\n\t"
                                         "@p ld.shared.v4.f32 {x, y, z, w}, [%1];\n\t"
                                         "st.global" HTF_T2_STQ ".v4.f32 [%0], {x, y, z, w};\n\t}"
                                         :: "l"(dsta + 512ull * i), "r"(st_r + 512u * i), "r"(lane + 32 * i), "r"(nvalid) : "memory");
                    } else {
                        float4 *dst = p.out + (size_t)row * (size_t)K + lane;
                        int *idst = WITH_IDX ? p.idx_out + (size_t)row * (size_t)K + lane : nullptr;
                        unsigned sa = st_r, ia = ist_r;
                        for (int sl = lane; sl < K; sl += 32, sa += 512u, ia += 128u, dst += 32) {
                            const bool valid = sl < nvalid;
                            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (valid) v = lds_f4_v(sa);
                            if (!WITH_IDX || p.out) *dst = v;
                            if (WITH_IDX) { *idst = valid ? lds_i32(ia) : -1; idst += 32; }
                        }
                    }
                    if (need_count && lane == r0 + 2 * h + r) my_cnt = total;      // stored once per 32 rows of the cell
                    if (total >= K && lane == 0 && p.overflow) atomicMax(p.overflow, total);
                }
                }
            }
        }
        // the counts of up to 32 rows of the cell in one store instruction
        if (need_count && my_o >= 0) p.count_out[my_o - p.row_lo] = my_cnt;
    }
}

template <bool WITH_IDX, bool MAPPED, int KC, int NPB>
cudaError_t launch_tile2_kc(const NlistParams &p, dim3 grid, size_t smem, cudaStream_t st)
{
    static size_t configured_dev[HTF_MAX_DEVICES] = {0};   // the attribute is per device
    size_t &configured = configured_dev[htf_current_device_slot()];
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(nlist_tile2_kernel<WITH_IDX, MAPPED, KC, NPB>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    nlist_tile2_kernel<WITH_IDX, MAPPED, KC, NPB><<<grid, TILE * 32, smem, st>>>(p);
    return cudaGetLastError();
}

// the slot phase is unrolled for the two cutoffs of the benchmark configurations (K = 64, 96); four rows per
// candidate load where cells hold enough rows for it (pairs2)
template <bool WITH_IDX, bool MAPPED>
cudaError_t launch_tile2_variant(const NlistParams &p, dim3 grid, size_t smem, cudaStream_t st, bool pairs2)
{
    if (!WITH_IDX && p.K == 64) return pairs2 ? launch_tile2_kc<WITH_IDX, MAPPED, 64, 2>(p, grid, smem, st)
                                              : launch_tile2_kc<WITH_IDX, MAPPED, 64, 1>(p, grid, smem, st);
    if (!WITH_IDX && p.K == 96) return pairs2 ? launch_tile2_kc<WITH_IDX, MAPPED, 96, 2>(p, grid, smem, st)
                                              : launch_tile2_kc<WITH_IDX, MAPPED, 96, 1>(p, grid, smem, st);
    return launch_tile2_kc<WITH_IDX, MAPPED, 0, 1>(p, grid, smem, st);
}

size_t tile2_block_bytes(int capB, int K, bool with_idx)
{
    size_t b = TILE2_HDR + (size_t)(capB + 32) * 16;
    if (with_idx) b += (size_t)(capB + 32) * 4;
    const size_t rs = 1;                                        // rows in the per-warp stage (see the kernel)
    b += (size_t)TILE * (rs * K * 16 + (with_idx ? ((rs * K * 4 + 15) & ~(size_t)15) : 0));
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
                             int32_t *count_out, int32_t *overflow, cudaStream_t st, int zoff, int zcnt, int lane)
{
    if (row_hi <= row_lo) return cudaSuccess;
    NlistParams p;
    p.g = ctx->grid;
    if (zcnt >= 0) {
        // slab of the z-window: cell layers (z0 + zoff + i) % nz, i < zcnt (the pipelined step builds slab by slab)
        if (zcnt == 0) return cudaSuccess;
        p.g.z0 = (ctx->grid.z0 + zoff) % ctx->grid.n[2];
        p.g.zcount = zcnt;
    }
    p.cell_start = ctx->d_cell_start;
    p.sorted_idx = ctx->d_sorted_idx;
    p.spos = ctx->d_spos;
    p.n_all = (int)ctx->n_binned;
    p.row_lo = (int)row_lo;
    p.row_hi = (int)row_hi;
    p.K = ctx->K;
    p.rc2 = ctx->r_cut * ctx->r_cut;
    p.one2 = 0x3f8000003f800000ull;
    {
        auto pack = [](float v) { unsigned u; memcpy(&u, &v, 4); return (unsigned long long)u | ((unsigned long long)u << 32); };
        p.fast_wrap = 1;
        for (int a = 0; a < 3; a++) {
            if (ctx->grid.n[a] < 5) p.fast_wrap = 0;
            p.inv_l2[a] = pack(1.0f / ctx->grid.L[a]);
            p.neg_l2[a] = pack(-ctx->grid.L[a]);
        }
        p.magic2 = pack(12582912.0f);
        static const bool exact_env = [] { const char *e2 = getenv("HTF_EXACT_WRAP"); return e2 && atoi(e2) != 0; }();
        if (exact_env) p.fast_wrap = 0;
    }
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
    if (cap > 4064) cap = 4064;          // list entries are 16-bit byte offsets into the window (16 B per candidate)
    const bool with_idx = idx_out != nullptr;
    const bool mapped = ctx->map_type_start >= 0;
    cudaError_t e;

    // ---- pass 1: the tile kernel ----
    p.use_flags = 0;
    p.tile_flag = nullptr;
    p.flag_count = nullptr;
    p.flag_count_next = nullptr;
    bool tiled = false;
    static const bool tile_off = [] { const char *e = getenv("HTF_TILE_KERNEL"); return e && atoi(e) == 0; }();   // 0: per-cell kernel only
    if (!tile_off) {
        // tiles along y, hit masks in registers (see nlist_tile2_kernel)
        const int tiles_y = (g.n[1] + TILE - 1) / TILE;
        const int ncol = g.n[1] >= 3 ? min(TILE, g.n[1]) + 2 : g.n[1];
        const double bmean = (double)ncol * min(g.n[0], 3) * min(g.n[2], 3) * cell_mean;
        // mean + 4.5 sigma (Poisson): about 3e-6 of the tiles of a homogeneous fluid go to the per-cell kernel
        int capB = (int)(bmean + 4.5 * sqrt(bmean > 1.0 ? bmean : 1.0)) + 32;
        capB = (capB + 31) / 32 * 32;
        const size_t bytes = tile2_block_bytes(capB, p.K, with_idx);
        if (bytes <= 100 * 1024 && tiles_y <= 65535 && g.n[2] <= 65535) {
            const int ntiles2 = tiles_y * g.n[0] * g.n[2];
            if ((e = htf_ensure_tile_flags(ctx, ntiles2)) != cudaSuccess) return e;
            p.cap = cap;
            p.cap_tile = capB;
            p.tile_flag = ctx->d_tile_flag;
            lane &= 1;
            p.flag_count = ctx->d_flag_count + 2 * lane + (ctx->flag_parity[lane] & 1);
            p.flag_count_next = ctx->d_flag_count + 2 * lane + ((ctx->flag_parity[lane] + 1) & 1);
            ctx->flag_parity[lane] ^= 1;
            ctx->launches += 1;
            const dim3 tg((unsigned)g.n[0], (unsigned)tiles_y, (unsigned)p.g.zcount);
            // four rows per candidate load when the cells hold enough rows to fill the batches
            static const int pairs_env = [] { const char *e2 = getenv("HTF_TILE_PAIRS"); return e2 ? atoi(e2) : 0; }();
            const bool pairs2 = pairs_env ? pairs_env == 2 : cell_mean >= 6.0;
            e = with_idx ? (mapped ? launch_tile2_variant<true, true>(p, tg, bytes, st, pairs2)
                                   : launch_tile2_variant<true, false>(p, tg, bytes, st, pairs2))
                         : (mapped ? launch_tile2_variant<false, true>(p, tg, bytes, st, pairs2)
                                   : launch_tile2_variant<false, false>(p, tg, bytes, st, pairs2));
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
    int grid = (g.n[0] * g.n[1] * p.g.zcount + wpb - 1) / wpb;
    if (tiled) {                                                                                 // persistent flag scan
        const int ntw = (g.n[1] + TILE - 1) / TILE * g.n[0] * p.g.zcount;
        grid = min((ntw + wpb - 1) / wpb, 2 * ctx->sm_count);
    }
    ctx->launches += 1;
    if (with_idx) return mapped ? launch_variant<true, true>(p, grid, wpb, smem, st)
                                : launch_variant<true, false>(p, grid, wpb, smem, st);
    return mapped ? launch_variant<false, true>(p, grid, wpb, smem, st)
                  : launch_variant<false, false>(p, grid, wpb, smem, st);
}
