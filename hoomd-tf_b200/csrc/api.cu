// extern "C" boundary of libhtf_b200.so (see include/htf_b200.h for the contract and the
// reference interfaces each entry point replaces).  Orchestration only: argument checks,
// scratch ownership, kernel sequencing on the caller's stream.  No device synchronisation.
#include "common.cuh"
#include "../../include/htf_b200.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <algorithm>
#include <vector>

namespace {

thread_local char g_create_err[512] = "";

constexpr int HTF_DEFAULT_PIPE_SLABS = 0;   // measured on B200: the back-to-back sequence wins at cfg3

void set_err(htf_ctx *ctx, const char *fmt, ...)
{
    char *dst = ctx ? ctx->err : g_create_err;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
}

struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

#define HTF_CUDA(ctx, call)                                                                  \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            set_err(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return e_ == cudaErrorMemoryAllocation ? HTF_ENOMEM : HTF_ECUDA;                 \
        }                                                                                    \
    } while (0)

template <typename T>
int dev_realloc(htf_ctx *ctx, T **ptr, size_t count)
{
    if (*ptr) { cudaFree(*ptr); *ptr = nullptr; }
    if (count == 0) count = 1;
    HTF_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(ptr), sizeof(T) * count));
    return HTF_OK;
}

int ensure_particles(htf_ctx *ctx, int64_t n)
{
    if (n <= ctx->n_cap) return HTF_OK;
    int rc;
    if ((rc = dev_realloc(ctx, &ctx->d_cell_of, (size_t)n))) return rc;
    if ((rc = dev_realloc(ctx, &ctx->d_sorted_idx, (size_t)n))) return rc;
    if ((rc = dev_realloc(ctx, &ctx->d_scattered, (size_t)n))) return rc;
    if ((rc = dev_realloc(ctx, &ctx->d_spos, (size_t)n))) return rc;
    ctx->n_cap = n;
    return HTF_OK;
}

int ensure_cells(htf_ctx *ctx, int ncell)
{
    if (ncell <= ctx->ncell_cap) return HTF_OK;
    int rc;
    if ((rc = dev_realloc(ctx, &ctx->d_cell_cnt, (size_t)ncell))) return rc;
    if ((rc = dev_realloc(ctx, &ctx->d_cell_start, (size_t)ncell + 1))) return rc;
    if ((rc = dev_realloc(ctx, &ctx->d_block_sums, (size_t)ncell / 1024 + 2))) return rc;
    ctx->ncell_cap = ncell;
    return HTF_OK;
}

// Cell layers in z that the region of interest can touch (+ one layer of slack on both sides).
void update_z_window(htf_ctx *ctx)
{
    CellGrid &g = ctx->grid;
    g.z0 = 0;
    g.zcount = g.n[2];
    if (!(g.roi_h[2] >= 0.0f) || g.n[2] < 1) return;
    const double w = (double)g.L[2] / (double)g.n[2];
    const double a = (double)g.roi_c[2] - (double)g.roi_h[2] - (double)g.lo[2];
    const double b = (double)g.roi_c[2] + (double)g.roi_h[2] - (double)g.lo[2];
    long long la = (long long)std::floor(a / w) - 1, lb = (long long)std::floor(b / w) + 1;
    long long cnt = lb - la + 1;
    if (cnt >= g.n[2]) return;
    long long z0 = la % g.n[2];
    if (z0 < 0) z0 += g.n[2];
    g.z0 = (int)z0;
    g.zcount = (int)cnt;
}

// Cell grid for the current box and cutoff.  Any cell edge >= r_cut is correct; the edge
// is kept 1e-4 above r_cut (see CellGrid) and the cell count is capped so that a huge,
// nearly empty box (e.g. compute_pairwise's 1e10 box) cannot exhaust memory.
int make_grid(htf_ctx *ctx)
{
    CellGrid &g = ctx->grid;
    double budget = 2.0 * (double)(ctx->n_max > 2048 ? ctx->n_max : 2048);
    double want[3];
    for (int a = 0; a < 3; a++) {
        double L = (double)g.L[a];
        double n = std::floor(L / ((double)ctx->r_cut * (1.0 + 1e-4)));
        if (!(n >= 1.0)) n = 1.0;
        if (n > 1024.0) n = 1024.0;
        want[a] = n;
    }
    while (want[0] * want[1] * want[2] > budget) {
        int big = 0;
        for (int a = 1; a < 3; a++) if (want[a] > want[big]) big = a;
        if (want[big] <= 1.0) break;
        want[big] = std::floor(want[big] * 0.8);
        if (want[big] < 1.0) want[big] = 1.0;
    }
    g.ncell = 1;
    for (int a = 0; a < 3; a++) {
        g.n[a] = (int)want[a];
        g.inv_w[a] = (float)((double)g.n[a] / (double)g.L[a]);
        g.ncell *= g.n[a];
    }
    ctx->binned = false;
    ctx->calib_valid = false;
    update_z_window(ctx);
    return ensure_cells(ctx, g.ncell);
}

int check_ctx(htf_ctx *ctx)
{
    if (!ctx) { set_err(nullptr, "null context"); return HTF_EINVAL; }
    return HTF_OK;
}

int upload_rdf_table(htf_ctx *ctx, float r_lo, float r_hi, int nbins, cudaStream_t st)
{
    if (nbins < 1 || !(r_hi > r_lo)) { set_err(ctx, "rdf: need nbins >= 1 and r_hi > r_lo"); return HTF_EINVAL; }
    if (ctx->d_rdf_thr && ctx->rdf_lo == r_lo && ctx->rdf_hi == r_hi && ctx->rdf_nbins == nbins) return HTF_OK;
    const int nb = nbins + 2;
    std::vector<float> thr((size_t)nb + 1);
    htf_rdf_thresholds(r_lo, r_hi, nbins, thr.data());
    if (ctx->rdf_nbins != nbins || !ctx->d_rdf_thr) {
        int rc = dev_realloc(ctx, &ctx->d_rdf_thr, (size_t)nb + 1);
        if (rc) return rc;
    }
    // pageable source: the runtime stages it before returning, so `thr` may go out of scope
    HTF_CUDA(ctx, cudaMemcpyAsync(ctx->d_rdf_thr, thr.data(), sizeof(float) * ((size_t)nb + 1),
                                  cudaMemcpyHostToDevice, st));
    ctx->rdf_lo = r_lo; ctx->rdf_hi = r_hi; ctx->rdf_nbins = nbins;
    return HTF_OK;
}

// ---- pipelined step -------------------------------------------------------------------------------------------
// The build kernel is instruction-issue bound (it uses ~40 % of the HBM bandwidth), the pair pass is HBM bound (it
// uses few issue slots).  Run back to back they add up; run side by side they share the SMs.  The z-window of cell
// layers is cut into slabs: slab i is built on the caller's stream, its pair pass (walking the slab's cell-sorted
// slots) runs on the context's auxiliary stream as soon as the slab is complete -- while slab i+1 is being built,
// and while most of slab i's rows are still in the 126 MB L2.
struct PassSpec {
    bool cv;
    float4 *fe; float *virial; int vcomp;
    const float *thr; int nb; unsigned long long *bins;
    float r0; float4 *cv_row; double *cv_sum;
};

int ensure_pipeline(htf_ctx *ctx, int nevents)
{
    if (!ctx->aux_stream) {
        int lo = 0, hi = 0;
        HTF_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        HTF_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, hi));   // pair pass first
        HTF_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->build2_stream, cudaStreamNonBlocking));
    }
    if (nevents > ctx->pipe_events_n) {
        cudaEvent_t *ev = static_cast<cudaEvent_t *>(realloc(ctx->pipe_events, sizeof(cudaEvent_t) * (size_t)nevents));
        if (!ev) { set_err(ctx, "host allocation failed"); return HTF_ENOMEM; }
        ctx->pipe_events = ev;
        for (int i = ctx->pipe_events_n; i < nevents; i++) {
            HTF_CUDA(ctx, cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
            ctx->pipe_events_n = i + 1;
        }
    }
    return HTF_OK;
}

cudaError_t launch_pass(htf_ctx *ctx, const PassSpec &ps, const float4 *nl, int64_t rows, const HtfSlab *slab, cudaStream_t st,
                        const int32_t *row_count)
{
    if (ps.cv)
        return htf_launch_lj_cv(ctx, nl, rows, ctx->K, ps.fe, ps.virial, ps.vcomp, ps.r0, ps.cv_row, ps.cv_sum, ps.thr, ps.nb,
                                ps.bins, st, slab, row_count);
    return htf_launch_lj(ctx, nl, rows, ctx->K, ps.fe, ps.virial, ps.vcomp, ps.thr, ps.nb, nullptr, 0, -1, -1, ps.bins, st, slab,
                         row_count);
}

// build rows [row_lo, row_hi) of the binned particles into nl, then the pair pass; pipelined when it pays
int build_and_pass(htf_ctx *ctx, int64_t row_lo, int64_t row_hi, float4 *nl, int32_t *d_overflow, const PassSpec &ps,
                   cudaStream_t st)
{
    const int64_t rows = row_hi - row_lo;
    const CellGrid &g = ctx->grid;
    // The builder's per-row neighbor counts ride along to the pair pass, which then reads (and computes on) only the
    // valid slots of each row: 27 % less DRAM traffic at liquid density and, with 4 lanes per row, 0.149 instead of
    // 0.180 ms at 1 M x 64 -- for 10 us more in the build.  HTF_ROW_COUNTS=0 turns it off.
    static const bool use_counts = [] { const char *e = getenv("HTF_ROW_COUNTS"); return !(e && atoi(e) == 0); }();
    int32_t *cnt = nullptr;
    if (use_counts) {
        if (rows > ctx->row_count_cap) {
            int rc = dev_realloc(ctx, &ctx->d_row_count, (size_t)rows);
            if (rc) return rc;
            ctx->row_count_cap = rows;
        }
        cnt = ctx->d_row_count;
    }
    int S = ctx->pipe_slabs;
    if (S > g.zcount / 2) S = g.zcount / 2;                       // at least two cell layers per slab
    if (rows < 131072 || S < 2) {                                 // small systems are launch bound: plain sequence
        HTF_CUDA(ctx, htf_launch_nlist(ctx, row_lo, row_hi, nl, nullptr, cnt, d_overflow, st));
        HTF_CUDA(ctx, launch_pass(ctx, ps, nl, rows, nullptr, st, cnt));
        return HTF_OK;
    }
    // slab boundaries in window-relative layers; a slab never crosses the periodic wrap of the window
    std::vector<int> cut;
    for (int i = 0; i <= S; i++) cut.push_back((int)((int64_t)i * g.zcount / S));
    const int lw = g.n[2] - g.z0;
    if (lw > 0 && lw < g.zcount) { cut.push_back(lw); std::sort(cut.begin(), cut.end()); cut.erase(std::unique(cut.begin(), cut.end()), cut.end()); }
    const int nslab = (int)cut.size() - 1;
    int rc = ensure_pipeline(ctx, nslab + 3);
    if (rc) return rc;
    cudaEvent_t ev_binned = ctx->pipe_events[nslab], ev_aux = ctx->pipe_events[nslab + 1], ev_b2 = ctx->pipe_events[nslab + 2];
    const bool two = ctx->pipe_build_streams > 1;
    if (two) {
        HTF_CUDA(ctx, cudaEventRecord(ev_binned, st));
        HTF_CUDA(ctx, cudaStreamWaitEvent(ctx->build2_stream, ev_binned, 0));
    }
    const int layer = g.n[0] * g.n[1];
    for (int i = 0; i < nslab; i++) {
        const int la = cut[i], nlay = cut[i + 1] - cut[i];
        const int lane = (two && (i & 1)) ? 1 : 0;
        cudaStream_t bs = lane ? ctx->build2_stream : st;
        HTF_CUDA(ctx, htf_launch_nlist(ctx, row_lo, row_hi, nl, nullptr, cnt, d_overflow, bs, la, nlay, lane));
        HTF_CUDA(ctx, cudaEventRecord(ctx->pipe_events[i], bs));
        HTF_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->pipe_events[i], 0));
        const int za = (g.z0 + la) % g.n[2];
        HtfSlab slab;
        slab.sorted_idx = ctx->d_sorted_idx;
        slab.slot_lo = ctx->d_cell_start + (size_t)za * layer;
        slab.slot_hi = ctx->d_cell_start + (size_t)(za + nlay) * layer;
        slab.row_lo = row_lo; slab.row_hi = row_hi;
        slab.blocks_per_sm = (i + 1 < nslab) ? ctx->pipe_pass_bps : 0;      // the last pass has the machine to itself
        const int64_t expect = (int64_t)((double)rows * nlay / g.zcount * 1.25) + 1024;
        HTF_CUDA(ctx, launch_pass(ctx, ps, nl, expect, &slab, ctx->aux_stream, cnt));
    }
    HTF_CUDA(ctx, cudaEventRecord(ev_aux, ctx->aux_stream));
    HTF_CUDA(ctx, cudaStreamWaitEvent(st, ev_aux, 0));
    if (two) {
        HTF_CUDA(ctx, cudaEventRecord(ev_b2, ctx->build2_stream));
        HTF_CUDA(ctx, cudaStreamWaitEvent(st, ev_b2, 0));
    }
    return HTF_OK;
}

}  // namespace

cudaError_t htf_ensure_tile_flags(htf_ctx *ctx, int ntiles)
{
    if (ntiles <= ctx->tile_flag_cap) return cudaSuccess;
    if (ctx->d_tile_flag) cudaFree(ctx->d_tile_flag);
    ctx->d_tile_flag = nullptr;
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&ctx->d_tile_flag), (size_t)ntiles);
    if (e == cudaSuccess) ctx->tile_flag_cap = ntiles;
    return e;
}

// bin(q) of the reference for a squared distance q: r = fp32 sqrt, TF CPU
// histogram_fixed_width rule (double step, truncation).  Monotone in q, so the kernel only
// needs the nb-1 switch points; they are found by bisection over fp32 bit patterns.
static int rdf_bin_of_q(float q, float r_lo, double step, int nb)
{
    float r = sqrtf(q);
    float v = r > r_lo ? r : r_lo;
    double t = (double)(float)(v - r_lo) / step;
    double last = (double)(nb - 1);
    if (t > last) t = last;
    return (int)t;
}

void htf_rdf_thresholds(float r_lo, float r_hi, int nbins, float *thr)
{
    const int nb = nbins + 2;
    const double step = (double)(float)(r_hi - r_lo) / (double)nb;
    thr[0] = 0.0f;
    thr[nb] = INFINITY;
    for (int b = 1; b < nb; b++) {
        // smallest non-negative float q with bin(q) >= b; bit patterns of non-negative floats are ordered
        uint32_t lo = 0u, hi = 0x7f800000u;     // [+0, +inf]
        float fhi;
        memcpy(&fhi, &hi, 4);
        if (rdf_bin_of_q(fhi, r_lo, step, nb) < b) { thr[b] = INFINITY; continue; }
        while (lo < hi) {
            uint32_t mid = lo + (hi - lo) / 2;
            float fm;
            memcpy(&fm, &mid, 4);
            if (rdf_bin_of_q(fm, r_lo, step, nb) >= b) hi = mid; else lo = mid + 1;
        }
        memcpy(&thr[b], &lo, 4);
    }
}

extern "C" {

int htf_abi_version(void) { return HTF_ABI_VERSION; }

int htf_create(htf_ctx **out, int device, int64_t n_max, int k, float r_cut, int flags)
{
    if (!out) { set_err(nullptr, "htf_create: out is NULL"); return HTF_EINVAL; }
    *out = nullptr;
    if (n_max < 0 || n_max > 2000000000LL) { set_err(nullptr, "htf_create: n_max out of range"); return HTF_EINVAL; }
    if (k < 1) { set_err(nullptr, "htf_create: nneighbor_cutoff must be >= 1"); return HTF_EINVAL; }
    if (!(r_cut > 0.0f)) { set_err(nullptr, "htf_create: r_cut must be > 0"); return HTF_EINVAL; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) {
        set_err(nullptr, "htf_create: no CUDA device %d (%s)", device, cudaGetErrorString(e));
        return HTF_ECUDA;
    }
    int major = 0, sms = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (major != 10 && !(flags & HTF_FLAG_ANY_ARCH)) {
        set_err(nullptr, "htf_create: device %d is sm_%d0, this library carries sm_100a code only", device, major);
        return HTF_EARCH;
    }
    htf_ctx *ctx = new (std::nothrow) htf_ctx();
    if (!ctx) { set_err(nullptr, "htf_create: host allocation failed"); return HTF_ENOMEM; }
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device; ctx->sm_count = sms; ctx->flags = flags; ctx->n_max = n_max; ctx->K = k;
    ctx->r_cut = r_cut; ctx->map_type_start = -1;
    ctx->pipe_slabs = HTF_DEFAULT_PIPE_SLABS;
    ctx->pipe_pass_bps = 2;
    ctx->pipe_build_streams = 2;
    if (const char *e = getenv("HTF_PIPE_SLABS")) ctx->pipe_slabs = atoi(e);
    if (const char *e = getenv("HTF_PIPE_PASS_BPS")) ctx->pipe_pass_bps = atoi(e);
    if (const char *e = getenv("HTF_PIPE_BUILD_STREAMS")) ctx->pipe_build_streams = atoi(e);
    for (int a = 0; a < 3; a++) ctx->grid.roi_h[a] = -1.0f;
    DeviceGuard guard(device);
    int rc = ensure_particles(ctx, n_max > 0 ? n_max : 1);
    if (!rc) rc = dev_realloc(ctx, &ctx->d_stats, 16);         // [0..2] cell statistics, [6..7] skin status, [8..11] flagged-tile counters
    if (!rc && cudaMemset(ctx->d_stats, 0, 16 * sizeof(int)) != cudaSuccess) rc = HTF_ECUDA;
    if (!rc) ctx->d_flag_count = ctx->d_stats + 8;
    if (!rc) rc = dev_realloc(ctx, &ctx->d_sel_slots, (size_t)HTF_SEL_MAX_BLOCKS + 2);
    if (!rc && cudaMemset(ctx->d_sel_slots, 0, (HTF_SEL_MAX_BLOCKS + 2) * sizeof(unsigned long long)) != cudaSuccess) rc = HTF_ECUDA;
    if (rc) { memcpy(g_create_err, ctx->err, sizeof(g_create_err)); htf_destroy(ctx); return rc; }
    *out = ctx;
    return HTF_OK;
}

void htf_destroy(htf_ctx *ctx)
{
    if (!ctx) return;
    DeviceGuard guard(ctx->device);
    if (ctx->skin_ctx) { htf_destroy(ctx->skin_ctx); ctx->skin_ctx = nullptr; }
    void *ptrs[] = {ctx->d_skin_cand, ctx->d_skin_count, ctx->d_skin_ref, ctx->d_cell_cnt, ctx->d_cell_start, ctx->d_block_sums, ctx->d_cell_of, ctx->d_sorted_idx, ctx->d_scattered,
                    ctx->d_spos, ctx->d_nlist_scratch, ctx->d_row_count, ctx->d_rdf_thr, ctx->d_tile_flag, ctx->d_stats,
                    ctx->d_sel_cnt, ctx->d_sel_off, ctx->d_sel_sums, ctx->d_sel_slots, ctx->d_train_packed, ctx->d_train_pred,
                    ctx->d_train_partial, ctx->d_train_loss_partial, ctx->d_mlp_pairs, ctx->d_mlp_blk};
    for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); i++) {
        if (!ptrs[i]) continue;
        cudaError_t e = cudaFree(ptrs[i]);
        if (e != cudaSuccess && getenv("HTF_DEBUG"))
            fprintf(stderr, "htf_destroy: cudaFree #%zu %p: %s\n", i, ptrs[i], cudaGetErrorString(e));
    }
    htf_comm_free(ctx);
    for (int i = 0; i < ctx->pipe_events_n; i++) cudaEventDestroy(ctx->pipe_events[i]);
    free(ctx->pipe_events);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->build2_stream) cudaStreamDestroy(ctx->build2_stream);
    (void)cudaGetLastError();       // never leave a sticky error behind for the next context
    delete ctx;
}

const char *htf_last_error(const htf_ctx *ctx) { return ctx ? ctx->err : g_create_err; }

int htf_set_box(htf_ctx *ctx, const float h_lo[3], const float h_hi[3], const float h_tilt[3])
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!h_lo || !h_hi) { set_err(ctx, "htf_set_box: null box"); return HTF_EINVAL; }
    if (h_tilt) {
        // the reference asserts sum(box[2]) < 1e-4 (htf/simmodel.py:195); use |.| so negative tilt is caught too
        float s = fabsf(h_tilt[0]) + fabsf(h_tilt[1]) + fabsf(h_tilt[2]);
        if (!(s < 1e-4f)) { set_err(ctx, "box is skewed"); return HTF_ESKEW; }
    }
    for (int a = 0; a < 3; a++) {
        if (!(h_hi[a] > h_lo[a])) { set_err(ctx, "htf_set_box: hi <= lo on axis %d", a); return HTF_EINVAL; }
        ctx->grid.lo[a] = h_lo[a]; ctx->grid.hi[a] = h_hi[a]; ctx->grid.L[a] = h_hi[a] - h_lo[a];
        ctx->grid.half[a] = 0.5f * ctx->grid.L[a];
    }
    ctx->box_set = true;
    DeviceGuard guard(ctx->device);
    return make_grid(ctx);
}

int htf_set_roi(htf_ctx *ctx, const float h_center[3], const float h_half_width[3])
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    for (int a = 0; a < 3; a++) {
        ctx->grid.roi_c[a] = (h_center && h_half_width) ? h_center[a] : 0.0f;
        ctx->grid.roi_h[a] = (h_center && h_half_width) ? h_half_width[a] : -1.0f;
    }
    ctx->binned = false;
    ctx->calib_valid = false;
    if (ctx->box_set) update_z_window(ctx);
    return HTF_OK;
}

int htf_pack_halo(htf_ctx *ctx, const float *d_pos, int64_t n, int axis, float threshold, int below,
                  float *d_out, int64_t capacity, int32_t *d_count, int32_t *d_overflow, void *stream)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (n < 0 || n > 2000000000LL || capacity < 1 || capacity > 2000000000LL || axis < 0 || axis > 2 || !d_out ||
        (n > 0 && !d_pos)) {
        set_err(ctx, "htf_pack_halo: bad arguments"); return HTF_EINVAL;
    }
    DeviceGuard guard(ctx->device);
    const int64_t nb = (n + 255) / 256 + 1;
    if (nb > ctx->sel_cap) {
        if ((rc = dev_realloc(ctx, &ctx->d_sel_cnt, (size_t)nb))) return rc;
        if ((rc = dev_realloc(ctx, &ctx->d_sel_off, (size_t)nb))) return rc;
        if ((rc = dev_realloc(ctx, &ctx->d_sel_sums, (size_t)nb / 1024 + 4))) return rc;
        ctx->sel_cap = nb;
    }
    HTF_CUDA(ctx, htf_launch_select(ctx, reinterpret_cast<const float4 *>(d_pos), n, axis, threshold, below != 0,
                                    reinterpret_cast<float4 *>(d_out), (int)capacity, d_count, d_overflow,
                                    (cudaStream_t)stream));
    return HTF_OK;
}

int htf_pack_halo_pair(htf_ctx *ctx, const float *d_pos, int64_t n, int axis, float threshold_lo, float threshold_hi,
                       float *d_out_lo, float *d_out_hi, int64_t capacity, int32_t *d_counts, int32_t *d_overflow,
                       void *stream)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (n < 0 || n > 1000000000LL || capacity < 1 || capacity > 2000000000LL || axis < 0 || axis > 2 || !d_out_lo ||
        !d_out_hi || (n > 0 && !d_pos)) {
        set_err(ctx, "htf_pack_halo_pair: bad arguments"); return HTF_EINVAL;
    }
    DeviceGuard guard(ctx->device);
    const int64_t need = 2 * ((n + 255) / 256) + 2;
    if (need > ctx->sel_cap) {
        if ((rc = dev_realloc(ctx, &ctx->d_sel_cnt, (size_t)need))) return rc;
        if ((rc = dev_realloc(ctx, &ctx->d_sel_off, (size_t)need))) return rc;
        if ((rc = dev_realloc(ctx, &ctx->d_sel_sums, (size_t)need / 1024 + 4))) return rc;
        ctx->sel_cap = need;
    }
    HTF_CUDA(ctx, htf_launch_select_pair(ctx, reinterpret_cast<const float4 *>(d_pos), n, axis, threshold_lo, threshold_hi,
                                         reinterpret_cast<float4 *>(d_out_lo), reinterpret_cast<float4 *>(d_out_hi),
                                         (int)capacity, d_counts, d_overflow, (cudaStream_t)stream));
    return HTF_OK;
}

int htf_eds_step(htf_ctx *ctx, const float *d_cv, const float *d_set_point, float *d_mean, float *d_ssd, int32_t *d_n,
                 float *d_alpha, float *d_adam_m, float *d_adam_v, float *d_adam_t, int period, float learning_rate,
                 float cv_scale, void *stream)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!d_cv || !d_set_point || !d_mean || !d_ssd || !d_n || !d_alpha || !d_adam_m || !d_adam_v || !d_adam_t || period < 1 ||
        !(cv_scale != 0.0f)) {
        set_err(ctx, "htf_eds_step: bad arguments"); return HTF_EINVAL;
    }
    DeviceGuard guard(ctx->device);
    HTF_CUDA(ctx, htf_launch_eds_step(ctx, d_cv, d_set_point, d_mean, d_ssd, d_n, d_alpha, d_adam_m, d_adam_v, d_adam_t, period,
                                      learning_rate, cv_scale, (cudaStream_t)stream));
    return HTF_OK;
}

int htf_integrate_half(htf_ctx *ctx, int half, float *d_pos, float *d_vel, const float *d_force, int64_t n, float dt,
                       float gamma, float kT, int flat, uint64_t seed, uint64_t timestep, void *stream)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!ctx->box_set) { set_err(ctx, "htf_integrate_half: call htf_set_box first"); return HTF_ESTATE; }
    if (n < 0 || n > 2000000000LL || (half != 0 && half != 1) || !(dt > 0.f) || gamma < 0.f || kT < 0.f ||
        (n > 0 && (!d_vel || !d_force || (half == 0 && !d_pos)))) {
        set_err(ctx, "htf_integrate_half: bad arguments"); return HTF_EINVAL;
    }
    DeviceGuard guard(ctx->device);
    HTF_CUDA(ctx, htf_launch_integrate(ctx, half, reinterpret_cast<float4 *>(d_pos), d_vel,
                                       reinterpret_cast<const float4 *>(d_force), n, dt, gamma, kT, flat,
                                       (unsigned long long)seed, (unsigned long long)timestep, (cudaStream_t)stream));
    return HTF_OK;
}

int htf_unstuff4(htf_ctx *ctx, const float *d_pos_hoomd, float *d_pos_out, int64_t n, void *stream)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (n < 0 || n > 2000000000LL || (n > 0 && (!d_pos_hoomd || !d_pos_out))) { set_err(ctx, "htf_unstuff4: bad arguments"); return HTF_EINVAL; }
    DeviceGuard guard(ctx->device);
    HTF_CUDA(ctx, htf_launch_unstuff4(ctx, reinterpret_cast<const float4 *>(d_pos_hoomd), reinterpret_cast<float4 *>(d_pos_out), n,
                                      (cudaStream_t)stream));
    return HTF_OK;
}

int htf_skin_configure(htf_ctx *ctx, float skin, int k_candidates)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!(skin > 0.0f) || k_candidates < 0) { set_err(ctx, "htf_skin_configure: need skin > 0"); return HTF_EINVAL; }
    if (ctx->K < 32) { set_err(ctx, "htf_skin_configure: buffered lists need nneighbor_cutoff >= 32"); return HTF_EINVAL; }
    if (k_candidates == 0) {
        const double g = (double)(ctx->r_cut + skin) / (double)ctx->r_cut;
        k_candidates = ((int)ceil(ctx->K * g * g * g) + 31) / 32 * 32;
    }
    if (k_candidates < ctx->K) k_candidates = ctx->K;
    DeviceGuard guard(ctx->device);
    if (ctx->skin_ctx) { htf_destroy(ctx->skin_ctx); ctx->skin_ctx = nullptr; }
    ctx->skin = skin; ctx->skin_kc = k_candidates;
    ctx->skin_row_lo = ctx->skin_row_hi = ctx->skin_n = -1;
    rc = htf_create(&ctx->skin_ctx, ctx->device, ctx->n_max, k_candidates, ctx->r_cut + skin, ctx->flags);
    if (rc) { set_err(ctx, "htf_skin_configure: %s", htf_last_error(nullptr)); return rc; }
    return HTF_OK;
}

int htf_skin_rebuild(htf_ctx *ctx, const float *d_pos_all, int64_t n_all, int64_t row_lo, int64_t row_hi, void *stream)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!ctx->skin_ctx) { set_err(ctx, "htf_skin_rebuild: call htf_skin_configure first"); return HTF_ESTATE; }
    if (!ctx->box_set) { set_err(ctx, "htf_skin_rebuild: call htf_set_box first"); return HTF_ESTATE; }
    if (n_all < 0 || row_lo < 0 || row_hi > n_all || row_lo > row_hi || (n_all > 0 && !d_pos_all)) {
        set_err(ctx, "htf_skin_rebuild: bad arguments"); return HTF_EINVAL;
    }
    DeviceGuard guard(ctx->device);
    htf_ctx *sk = ctx->skin_ctx;
    // same box, mapped rule and region of interest (widened by the skin) as the outer context
    if (!sk->box_set || memcmp(sk->grid.lo, ctx->grid.lo, sizeof(float) * 3) || memcmp(sk->grid.hi, ctx->grid.hi, sizeof(float) * 3)) {
        const float tilt[3] = {0.f, 0.f, 0.f};
        if ((rc = htf_set_box(sk, ctx->grid.lo, ctx->grid.hi, tilt))) { set_err(ctx, "htf_skin_rebuild: %s", htf_last_error(sk)); return rc; }
    }
    {
        float hw[3];
        for (int a = 0; a < 3; a++) hw[a] = ctx->grid.roi_h[a] >= 0.f ? ctx->grid.roi_h[a] + ctx->skin : -1.0f;
        if (memcmp(hw, sk->grid.roi_h, sizeof(hw)) || memcmp(ctx->grid.roi_c, sk->grid.roi_c, sizeof(float) * 3))
            if ((rc = htf_set_roi(sk, ctx->grid.roi_c, hw))) { set_err(ctx, "htf_skin_rebuild: %s", htf_last_error(sk)); return rc; }
    }
    sk->map_type_start = ctx->map_type_start;
    const int64_t rows = row_hi - row_lo;
    if (rows > ctx->skin_rows_cap) {
        if ((rc = dev_realloc(ctx, &ctx->d_skin_cand, (size_t)rows * ctx->skin_kc))) return rc;
        if ((rc = dev_realloc(ctx, &ctx->d_skin_count, (size_t)rows))) return rc;
        ctx->skin_rows_cap = rows;
    }
    if (n_all > ctx->skin_n_cap) {
        if ((rc = dev_realloc(ctx, &ctx->d_skin_ref, (size_t)n_all))) return rc;
        ctx->skin_n_cap = n_all;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = htf_bin_particles(sk, d_pos_all, n_all, stream))) { set_err(ctx, "htf_skin_rebuild: %s", htf_last_error(sk)); return rc; }
    HTF_CUDA(ctx, htf_launch_nlist(sk, row_lo, row_hi, nullptr, ctx->d_skin_cand, ctx->d_skin_count, nullptr, st));
    HTF_CUDA(ctx, cudaMemcpyAsync(ctx->d_skin_ref, d_pos_all, sizeof(float4) * (size_t)n_all, cudaMemcpyDeviceToDevice, st));
    ctx->launches += sk->launches; sk->launches = 0;
    ctx->skin_row_lo = row_lo; ctx->skin_row_hi = row_hi; ctx->skin_n = n_all;
    return HTF_OK;
}

int htf_skin_nlist(htf_ctx *ctx, const float *d_pos_all, int64_t n_all, int64_t row_lo, int64_t row_hi,
                   float *d_nlist_out, int32_t *d_idx_out, int32_t *d_count_out, int32_t *d_overflow, void *stream)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!ctx->skin_ctx || ctx->skin_n != n_all || ctx->skin_row_lo != row_lo || ctx->skin_row_hi != row_hi) {
        set_err(ctx, "htf_skin_nlist: no candidate lists for these rows (call htf_skin_rebuild first)"); return HTF_ESTATE;
    }
    if (row_hi > row_lo && (!d_nlist_out || !d_pos_all)) { set_err(ctx, "htf_skin_nlist: null argument"); return HTF_EINVAL; }
    DeviceGuard guard(ctx->device);
    HTF_CUDA(ctx, htf_launch_skin_filter(ctx, reinterpret_cast<const float4 *>(d_pos_all), row_lo, row_hi,
                                         reinterpret_cast<float4 *>(d_nlist_out), d_idx_out, d_count_out, d_overflow,
                                         (cudaStream_t)stream));
    return HTF_OK;
}

int htf_skin_status(htf_ctx *ctx, int32_t h_status[2], int reset, void *stream)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!h_status) return HTF_EINVAL;
    DeviceGuard guard(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    HTF_CUDA(ctx, cudaMemcpyAsync(h_status, ctx->d_stats + 6, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    HTF_CUDA(ctx, cudaStreamSynchronize(st));
    if (reset) HTF_CUDA(ctx, cudaMemsetAsync(ctx->d_stats + 6, 0, 2 * sizeof(int), st));
    return HTF_OK;
}

int htf_set_mapped_nlist(htf_ctx *ctx, int map_type_start)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    ctx->map_type_start = map_type_start < 0 ? -1 : map_type_start;
    return HTF_OK;
}

int htf_set_cutoff(htf_ctx *ctx, float r_cut, int k)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (k < 1 || !(r_cut > 0.0f)) { set_err(ctx, "htf_set_cutoff: need r_cut > 0 and k >= 1"); return HTF_EINVAL; }
    ctx->r_cut = r_cut; ctx->K = k;
    if (!ctx->box_set) return HTF_OK;
    DeviceGuard guard(ctx->device);
    return make_grid(ctx);
}

int htf_get_cell_grid(const htf_ctx *ctx, int h_ncell[3])
{
    if (!ctx || !h_ncell) return HTF_EINVAL;
    for (int a = 0; a < 3; a++) h_ncell[a] = ctx->grid.n[a];
    return ctx->box_set ? HTF_OK : HTF_ESTATE;
}

int64_t htf_launch_count(const htf_ctx *ctx) { return ctx ? ctx->launches : 0; }

int htf_bin_particles(htf_ctx *ctx, const float *d_pos_all, int64_t n_all, void *stream)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!ctx->box_set) { set_err(ctx, "htf_bin_particles: box not set"); return HTF_ESTATE; }
    if (n_all < 0 || n_all > 2000000000LL || (n_all > 0 && !d_pos_all)) {
        set_err(ctx, "htf_bin_particles: bad positions"); return HTF_EINVAL;
    }
    DeviceGuard guard(ctx->device);
    if ((rc = ensure_particles(ctx, n_all))) return rc;
    if (n_all > ctx->n_max) {                 // the cell budget follows the particle count
        ctx->n_max = n_all;
        if ((rc = make_grid(ctx))) return rc;
    }
    HTF_CUDA(ctx, htf_launch_binning(ctx, reinterpret_cast<const float4 *>(d_pos_all), n_all, (cudaStream_t)stream));
    ctx->binned = true;
    ctx->n_binned = n_all;
    return HTF_OK;
}

int htf_build_nlist(htf_ctx *ctx, const float *d_pos_all, int64_t n_all, int64_t row_lo, int64_t row_hi,
                    float *d_nlist_out, int32_t *d_idx_out, int32_t *d_count_out, int32_t *d_overflow,
                    void *stream)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    (void)d_pos_all;
    if (!ctx->binned || ctx->n_binned != n_all) {
        set_err(ctx, "htf_build_nlist: call htf_bin_particles on these %lld particles first", (long long)n_all);
        return HTF_ESTATE;
    }
    if (row_lo < 0 || row_hi > n_all || row_lo > row_hi) { set_err(ctx, "htf_build_nlist: bad row range"); return HTF_EINVAL; }
    if (row_hi > row_lo && !d_nlist_out) { set_err(ctx, "htf_build_nlist: null output"); return HTF_EINVAL; }
    DeviceGuard guard(ctx->device);
    HTF_CUDA(ctx, htf_launch_nlist(ctx, row_lo, row_hi, reinterpret_cast<float4 *>(d_nlist_out), d_idx_out,
                                   d_count_out, d_overflow, (cudaStream_t)stream));
    return HTF_OK;
}

int htf_lj_forces(htf_ctx *ctx, const float *d_nlist, int64_t rows, int k, const int32_t *d_row_count, float *d_force_energy,
                  float *d_virial, int virial_components, void *stream)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (rows < 0 || k < 1 || (rows > 0 && (!d_nlist || !d_force_energy))) { set_err(ctx, "htf_lj_forces: bad arguments"); return HTF_EINVAL; }
    if (d_virial && virial_components != 6 && virial_components != 9) {
        set_err(ctx, "htf_lj_forces: virial_components must be 6 or 9"); return HTF_EINVAL;
    }
    DeviceGuard guard(ctx->device);
    HTF_CUDA(ctx, htf_launch_lj(ctx, reinterpret_cast<const float4 *>(d_nlist), rows, k,
                                reinterpret_cast<float4 *>(d_force_energy), d_virial, virial_components,
                                nullptr, 0, nullptr, 0, -1, -1, nullptr, (cudaStream_t)stream, nullptr, d_row_count));
    return HTF_OK;
}

int htf_lj_forces_rdf(htf_ctx *ctx, const float *d_nlist, int64_t rows, int k, const int32_t *d_row_count, float *d_force_energy,
                      float *d_virial, int virial_components, int64_t *d_bins, float r_lo, float r_hi, int nbins, void *stream)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (rows < 0 || k < 1 || !d_bins || (rows > 0 && (!d_nlist || !d_force_energy))) {
        set_err(ctx, "htf_lj_forces_rdf: bad arguments"); return HTF_EINVAL;
    }
    if (d_virial && virial_components != 6 && virial_components != 9) {
        set_err(ctx, "htf_lj_forces_rdf: virial_components must be 6 or 9"); return HTF_EINVAL;
    }
    if (nbins + 2 > 1024) { set_err(ctx, "htf_lj_forces_rdf: nbins must be <= 1022"); return HTF_EINVAL; }
    DeviceGuard guard(ctx->device);
    if ((rc = upload_rdf_table(ctx, r_lo, r_hi, nbins, (cudaStream_t)stream))) return rc;
    HTF_CUDA(ctx, htf_launch_lj(ctx, reinterpret_cast<const float4 *>(d_nlist), rows, k,
                                reinterpret_cast<float4 *>(d_force_energy), d_virial, virial_components,
                                ctx->d_rdf_thr, nbins + 2, nullptr, 0, -1, -1,
                                reinterpret_cast<unsigned long long *>(d_bins), (cudaStream_t)stream, nullptr, d_row_count));
    return HTF_OK;
}

int htf_lj_cv_forces(htf_ctx *ctx, const float *d_nlist, int64_t rows, int k, const int32_t *d_row_count, float r0,
                     float *d_force_energy, float *d_virial, int virial_components, float *d_cv_row, double *d_cv_sum,
                     int64_t *d_bins, float r_lo, float r_hi, int nbins, void *stream)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (rows < 0 || k < 1 || !(r0 > 0.0f) || !d_cv_sum || (rows > 0 && (!d_nlist || !d_force_energy || !d_cv_row))) {
        set_err(ctx, "htf_lj_cv_forces: bad arguments"); return HTF_EINVAL;
    }
    if (d_virial && virial_components != 6 && virial_components != 9) {
        set_err(ctx, "htf_lj_cv_forces: virial_components must be 6 or 9"); return HTF_EINVAL;
    }
    DeviceGuard guard(ctx->device);
    const float *thr = nullptr;
    int nb = 0;
    if (d_bins) {
        if (nbins + 2 > 1024) { set_err(ctx, "htf_lj_cv_forces: nbins must be <= 1022"); return HTF_EINVAL; }
        if ((rc = upload_rdf_table(ctx, r_lo, r_hi, nbins, (cudaStream_t)stream))) return rc;
        thr = ctx->d_rdf_thr; nb = nbins + 2;
    }
    HTF_CUDA(ctx, htf_launch_lj_cv(ctx, reinterpret_cast<const float4 *>(d_nlist), rows, k,
                                   reinterpret_cast<float4 *>(d_force_energy), d_virial, virial_components, r0,
                                   reinterpret_cast<float4 *>(d_cv_row), d_cv_sum, thr, nb,
                                   reinterpret_cast<unsigned long long *>(d_bins), (cudaStream_t)stream, nullptr, d_row_count));
    return HTF_OK;
}

int htf_mlp_param_sizes(int *raw_count, int *packed_bytes)
{
    if (raw_count) *raw_count = htf_mlp_raw_count_host();
    if (packed_bytes) *packed_bytes = htf_mlp_packed_bytes_host();
    return HTF_OK;
}

int htf_mlp_pack(htf_ctx *ctx, const float *d_raw, void *d_packed, void *stream)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!d_raw || !d_packed) { set_err(ctx, "htf_mlp_pack: null argument"); return HTF_EINVAL; }
    DeviceGuard guard(ctx->device);
    HTF_CUDA(ctx, htf_launch_mlp_pack(ctx, d_raw, reinterpret_cast<unsigned char *>(d_packed), (cudaStream_t)stream));
    return HTF_OK;
}

int htf_mlp_forces(htf_ctx *ctx, const float *d_nlist, int64_t rows, int k, const int32_t *d_row_count, const void *d_packed,
                   float rbf_high, float *d_force_energy, void *stream)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (rows < 0 || k < 1 || !(rbf_high > 0.0f) || !d_packed || (rows > 0 && (!d_nlist || !d_force_energy))) {
        set_err(ctx, "htf_mlp_forces: bad arguments"); return HTF_EINVAL;
    }
    if ((reinterpret_cast<uintptr_t>(d_packed) & 15) != 0) { set_err(ctx, "htf_mlp_forces: packed blob must be 16-byte aligned"); return HTF_EINVAL; }
    DeviceGuard guard(ctx->device);
    HTF_CUDA(ctx, htf_launch_mlp(ctx, reinterpret_cast<const float4 *>(d_nlist), rows, k,
                                 reinterpret_cast<const unsigned char *>(d_packed), rbf_high,
                                 reinterpret_cast<float4 *>(d_force_energy), (cudaStream_t)stream, d_row_count));
    return HTF_OK;
}

int htf_mlp_train_grads(htf_ctx *ctx, const float *d_nlist, int64_t rows, int k, const float *d_raw, float rbf_high,
                        const float *d_labels, int64_t n_total, float *d_pred_out, float *d_grads, float *d_loss, void *stream)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (rows < 0 || k < 1 || !(rbf_high > 0.0f) || !d_raw || !d_grads || n_total < 1 || (rows > 0 && (!d_nlist || !d_labels))) {
        set_err(ctx, "htf_mlp_train_grads: bad arguments"); return HTF_EINVAL;
    }
    DeviceGuard guard(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (!ctx->d_train_packed) {
        if ((rc = dev_realloc(ctx, &ctx->d_train_packed, (size_t)htf_mlp_packed_bytes_host()))) return rc;
        if ((rc = dev_realloc(ctx, &ctx->d_train_partial, (size_t)htf_mlp_train_partial_floats(ctx->sm_count)))) return rc;
        if ((rc = dev_realloc(ctx, &ctx->d_train_loss_partial, (size_t)2 * ctx->sm_count))) return rc;
    }
    float *pred = d_pred_out;
    if (!pred) {
        if (rows > ctx->train_pred_rows) {
            if ((rc = dev_realloc(ctx, &ctx->d_train_pred, (size_t)rows * 4))) return rc;
            ctx->train_pred_rows = rows;
        }
        pred = ctx->d_train_pred;
    }
    // pass 1: predictions of the current parameters (tcgen05 inference kernel); pass 2: reverse sweep
    HTF_CUDA(ctx, htf_launch_mlp_pack(ctx, d_raw, ctx->d_train_packed, st));
    HTF_CUDA(ctx, htf_launch_mlp(ctx, reinterpret_cast<const float4 *>(d_nlist), rows, k, ctx->d_train_packed, rbf_high,
                                 reinterpret_cast<float4 *>(pred), st));
    HTF_CUDA(ctx, htf_launch_mlp_train(ctx, reinterpret_cast<const float4 *>(d_nlist), rows, k, d_raw, rbf_high,
                                       reinterpret_cast<const float4 *>(pred), reinterpret_cast<const float4 *>(d_labels), n_total,
                                       ctx->d_train_partial, ctx->d_train_loss_partial, d_grads, d_loss, st));
    return HTF_OK;
}

int htf_adam_step(htf_ctx *ctx, float *d_params, const float *d_grads, float *d_m, float *d_v, float *d_t, int64_t count,
                  float learning_rate, float beta1, float beta2, float epsilon, void *stream)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (count < 0 || count > 2000000000LL || (count > 0 && (!d_params || !d_grads || !d_m || !d_v || !d_t))) {
        set_err(ctx, "htf_adam_step: bad arguments"); return HTF_EINVAL;
    }
    DeviceGuard guard(ctx->device);
    HTF_CUDA(ctx, htf_launch_adam(ctx, d_params, d_grads, d_m, d_v, d_t, (int)count, learning_rate, beta1, beta2, epsilon,
                                  (cudaStream_t)stream));
    return HTF_OK;
}

int htf_rdf_hist(htf_ctx *ctx, const float *d_nlist, int64_t rows, int k, const float *d_row_type,
                 int64_t row_type_stride, float r_lo, float r_hi, int nbins, int type_i, int type_j,
                 int64_t *d_bins, void *stream)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (rows < 0 || k < 1 || (rows > 0 && !d_nlist) || !d_bins) { set_err(ctx, "htf_rdf_hist: bad arguments"); return HTF_EINVAL; }
    if (type_i >= 0 && !d_row_type) { set_err(ctx, "htf_rdf_hist: type_i needs the row types"); return HTF_EINVAL; }
    if (nbins + 2 > 1024) { set_err(ctx, "htf_rdf_hist: nbins must be <= 1022"); return HTF_EINVAL; }
    DeviceGuard guard(ctx->device);
    if ((rc = upload_rdf_table(ctx, r_lo, r_hi, nbins, (cudaStream_t)stream))) return rc;
    HTF_CUDA(ctx, htf_launch_rdf(ctx, reinterpret_cast<const float4 *>(d_nlist), rows, k,
                                 d_row_type, (long long)row_type_stride, ctx->d_rdf_thr, nbins + 2,
                                 type_i, type_j, reinterpret_cast<unsigned long long *>(d_bins),
                                 (cudaStream_t)stream));
    return HTF_OK;
}

static int lj_step_impl(htf_ctx *ctx, const float *d_pos_all, int64_t n_all, int64_t row_lo, int64_t row_hi,
                        float *d_nlist_out, float *d_force_energy, float *d_virial, int virial_components,
                        int32_t *d_overflow, int64_t *d_bins, float r_lo, float r_hi, int nbins, void *stream, bool rebin)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (row_lo < 0 || row_hi > n_all || row_lo > row_hi) { set_err(ctx, "htf_lj_step: bad row range"); return HTF_EINVAL; }
    const int64_t rows = row_hi - row_lo;
    if (rows > 0 && !d_force_energy) { set_err(ctx, "htf_lj_step: null force output"); return HTF_EINVAL; }
    if (d_virial && virial_components != 6 && virial_components != 9) {
        set_err(ctx, "htf_lj_step: virial_components must be 6 or 9"); return HTF_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    DeviceGuard guard(ctx->device);
    if (rebin) {
        if ((rc = htf_bin_particles(ctx, d_pos_all, n_all, stream))) return rc;
    } else if (!ctx->binned || ctx->n_binned != n_all) {
        set_err(ctx, "htf_lj_rows: call htf_bin_particles on these %lld particles first", (long long)n_all);
        return HTF_ESTATE;
    }
    float *nl = d_nlist_out;
    if (!nl) {
        const int64_t need = rows * ctx->K * 4;
        if (need > ctx->nlist_scratch_elems) {
            if ((rc = dev_realloc(ctx, &ctx->d_nlist_scratch, (size_t)need))) return rc;
            ctx->nlist_scratch_elems = need;
        }
        nl = ctx->d_nlist_scratch;
    }
    PassSpec ps = {};
    ps.cv = false;
    ps.fe = reinterpret_cast<float4 *>(d_force_energy); ps.virial = d_virial; ps.vcomp = virial_components;
    ps.bins = reinterpret_cast<unsigned long long *>(d_bins);
    if (d_bins) {
        if (nbins + 2 > 1024) { set_err(ctx, "htf_lj_step: nbins must be <= 1022"); return HTF_EINVAL; }
        if ((rc = upload_rdf_table(ctx, r_lo, r_hi, nbins, st))) return rc;
        ps.thr = ctx->d_rdf_thr; ps.nb = nbins + 2;
    }
    if (rows <= 0) return HTF_OK;
    return build_and_pass(ctx, row_lo, row_hi, reinterpret_cast<float4 *>(nl), d_overflow, ps, st);
}

int htf_lj_step(htf_ctx *ctx, const float *d_pos_all, int64_t n_all, int64_t row_lo, int64_t row_hi,
                float *d_nlist_out, float *d_force_energy, float *d_virial, int virial_components,
                int32_t *d_overflow, int64_t *d_bins, float r_lo, float r_hi, int nbins, void *stream)
{
    return lj_step_impl(ctx, d_pos_all, n_all, row_lo, row_hi, d_nlist_out, d_force_energy, d_virial, virial_components,
                        d_overflow, d_bins, r_lo, r_hi, nbins, stream, true);
}

int htf_lj_rows(htf_ctx *ctx, int64_t n_all, int64_t row_lo, int64_t row_hi, float *d_nlist_out, float *d_force_energy,
                float *d_virial, int virial_components, int32_t *d_overflow, int64_t *d_bins, float r_lo, float r_hi,
                int nbins, void *stream)
{
    return lj_step_impl(ctx, nullptr, n_all, row_lo, row_hi, d_nlist_out, d_force_energy, d_virial, virial_components,
                        d_overflow, d_bins, r_lo, r_hi, nbins, stream, false);
}

int htf_lj_cv_step(htf_ctx *ctx, const float *d_pos_all, int64_t n_all, int64_t row_lo, int64_t row_hi,
                   float *d_nlist_out, float r0, float *d_force_energy, float *d_virial, int virial_components,
                   float *d_cv_row, double *d_cv_sum, int32_t *d_overflow, int64_t *d_bins, float r_lo, float r_hi,
                   int nbins, void *stream)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (row_lo < 0 || row_hi > n_all || row_lo > row_hi) { set_err(ctx, "htf_lj_cv_step: bad row range"); return HTF_EINVAL; }
    const int64_t rows = row_hi - row_lo;
    if (!(r0 > 0.0f) || !d_cv_sum || (rows > 0 && (!d_force_energy || !d_cv_row))) {
        set_err(ctx, "htf_lj_cv_step: bad arguments"); return HTF_EINVAL;
    }
    if (d_virial && virial_components != 6 && virial_components != 9) {
        set_err(ctx, "htf_lj_cv_step: virial_components must be 6 or 9"); return HTF_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    DeviceGuard guard(ctx->device);
    if (d_pos_all) {
        if ((rc = htf_bin_particles(ctx, d_pos_all, n_all, stream))) return rc;
    } else if (!ctx->binned || ctx->n_binned != n_all) {      // NULL positions: the rows of the last htf_bin_particles
        set_err(ctx, "htf_lj_cv_step: call htf_bin_particles on these %lld particles first", (long long)n_all);
        return HTF_ESTATE;
    }
    float *nl = d_nlist_out;
    if (!nl) {
        const int64_t need = rows * ctx->K * 4;
        if (need > ctx->nlist_scratch_elems) {
            if ((rc = dev_realloc(ctx, &ctx->d_nlist_scratch, (size_t)need))) return rc;
            ctx->nlist_scratch_elems = need;
        }
        nl = ctx->d_nlist_scratch;
    }
    PassSpec ps = {};
    ps.cv = true;
    ps.fe = reinterpret_cast<float4 *>(d_force_energy); ps.virial = d_virial; ps.vcomp = virial_components;
    ps.r0 = r0; ps.cv_row = reinterpret_cast<float4 *>(d_cv_row); ps.cv_sum = d_cv_sum;
    ps.bins = reinterpret_cast<unsigned long long *>(d_bins);
    if (d_bins) {
        if (nbins + 2 > 1024) { set_err(ctx, "htf_lj_cv_step: nbins must be <= 1022"); return HTF_EINVAL; }
        if ((rc = upload_rdf_table(ctx, r_lo, r_hi, nbins, st))) return rc;
        ps.thr = ctx->d_rdf_thr; ps.nb = nbins + 2;
    }
    if (rows <= 0) return HTF_OK;
    return build_and_pass(ctx, row_lo, row_hi, reinterpret_cast<float4 *>(nl), d_overflow, ps, st);
}

int htf_set_pipeline(htf_ctx *ctx, int slabs)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (slabs < 0 || slabs > 256) { set_err(ctx, "htf_set_pipeline: slabs must be in [0, 256]"); return HTF_EINVAL; }
    ctx->pipe_slabs = slabs;
    return HTF_OK;
}

}  // extern "C"
