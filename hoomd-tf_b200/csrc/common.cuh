// Internal declarations shared by the kernels of libhtf_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define HTF_WARP 32
#define HTF_FULL 0xffffffffu

// Cell grid of the periodic orthorhombic box.  Cell edge >= r_cut*(1+1e-4) in every
// dimension, so all neighbors of a particle live in the 3x3x3 stencil of its cell even
// when the fp32 cell index of a particle on a cell face rounds to either side.
struct CellGrid {
    float lo[3], hi[3], L[3];
    float half[3];       // L / 2: minimum image is "d >= half -> d -= L; d < -half -> d += L" (HOOMD boxes are
                         // centred, so half == hi and -half == lo there; other origins work the same way)
    float inv_w[3];      // n / L
    int n[3];
    int ncell;
    // region of interest for binning (row sharding over GPUs): only particles whose minimum-image distance
    // from roi_c is <= roi_h on every axis are binned; roi_h < 0 disables the test on that axis
    float roi_c[3], roi_h[3];
    // z-window of cell layers that can hold binned particles: layers (z0 + i) % n[2], i < zcount
    // (the whole grid when the region of interest does not restrict z)
    int z0, zcount;
};

struct HtfComm;

constexpr int HTF_SEL_MAX_BLOCKS = 1024;     // grid limit of the one-launch halo selection (select_fused2_kernel)

struct htf_ctx {
    int device;
    int sm_count;
    int flags;
    int64_t n_max;
    int K;
    float r_cut;
    int map_type_start;
    bool box_set;
    bool binned;
    int64_t n_binned;
    CellGrid grid;
    // device scratch owned by the context
    int *d_cell_cnt;      // [ncell_cap]
    int *d_cell_start;    // [ncell_cap + 1]
    int *d_block_sums;    // [scan blocks]
    int ncell_cap;
    int *d_cell_of;       // [n_cap]
    int *d_sorted_idx;    // [n_cap] cell-sorted slot -> particle index
    int *d_scattered;     // [n_cap] the same before the in-cell ordering (output of the atomic scatter)
    float4 *d_spos;       // [n_cap] cell-sorted positions
    int64_t n_cap;
    int *d_stats;                 // [3] cell population statistics (see htf_cell_stats)
    bool calib_valid;             // staging capacities calibrated for the current configuration
    int64_t calib_n;              // n_binned at calibration time
    double calib_cell_mean;       // particles per occupied cell
    int *d_sel_cnt, *d_sel_off, *d_sel_sums;   // halo selection scratch (per 256-particle block)
    int64_t sel_cap;
    bool sel_fused_refused;            // the driver refused the cooperative launch of select_fused2_kernel once: three-kernel form from then on
    unsigned long long *d_sel_slots;   // [HTF_SEL_MAX_BLOCKS] tagged per-block counts of the one-launch selection + [1] epoch / done counter
    unsigned char *d_tile_flag;   // [tiles] written by the tile kernel, read by the per-cell kernel
    int tile_flag_cap;
    int *d_flag_count;            // per build lane two counters (inside d_stats) of tiles the tile kernel flagged, used in turn
    int flag_parity[2];           // lanes: builds that may run concurrently (pipelined step) use different counters
    // buffered ("skin") lists: candidates within r_cut + skin, rebuilt by htf_skin_rebuild, filtered every step
    float skin;
    int skin_kc;                  // candidate capacity per row
    htf_ctx *skin_ctx;            // inner context with cutoff r_cut + skin and K = skin_kc
    int *d_skin_cand, *d_skin_count;
    float4 *d_skin_ref;           // positions at the last rebuild
    int64_t skin_rows_cap, skin_n_cap;
    int64_t skin_row_lo, skin_row_hi, skin_n;   // what the current lists cover (-1: none)
    int *d_row_count;         // lazily sized [rows]: neighbors per row of the last fused build (the pair pass behind it
    int64_t row_count_cap;    //   skips the padded slots of every row instead of reading their zeros)
    float *d_nlist_scratch;   // lazily sized [rows][K][4] for htf_lj_step(d_nlist_out = NULL)
    int64_t nlist_scratch_elems;
    // RDF threshold table (device) and the key it was built for
    float *d_rdf_thr;
    float rdf_lo, rdf_hi;
    int rdf_nbins;
    // pipelined step: the force pass of cell-layer slab i runs on `aux` while slab i+1 is being built
    cudaStream_t aux_stream;      // pair passes
    cudaStream_t build2_stream;   // every other slab is built here, so that one build's tail overlaps the next one's start
    cudaEvent_t *pipe_events;     // [pipe_events_n]: slab built (i), then: binned, aux done, build2 done
    int pipe_events_n;
    int pipe_slabs;               // slabs per step (<= 1: no pipelining)
    int pipe_pass_bps;            // blocks per SM of a slab's pair pass
    int pipe_build_streams;       // 1 or 2
    HtfComm *comm;                // peer-memory exchange state (comm.cu), nullptr until htf_comm_create
    // compacted valid pairs of the MLP inference pass (one more copy of the tensor) and its block offsets
    float4 *d_mlp_pairs;
    int64_t mlp_pairs_cap;
    int *d_mlp_blk;
    int mlp_blk_cap;
    // scratch of the MLP training step: packed bf16 parameters, predictions, per-block partial gradients / loss sums
    unsigned char *d_train_packed;
    float *d_train_pred, *d_train_partial;
    double *d_train_loss_partial;
    int64_t train_pred_rows;
    int64_t launches;
    char err[512];
};

// ---- launchers (each returns a cudaError_t from the launch) ----
cudaError_t htf_launch_binning(htf_ctx *ctx, const float4 *pos, int64_t n, cudaStream_t st);

// zcnt >= 0 restricts the build to cell layers [zoff, zoff + zcnt) of the context's z-window
cudaError_t htf_launch_nlist(htf_ctx *ctx, int64_t row_lo, int64_t row_hi, float4 *out, int32_t *idx_out,
                             int32_t *count_out, int32_t *overflow, cudaStream_t st, int zoff = 0, int zcnt = -1,
                             int lane = 0);

// Slab mode of the pair pass (pipelined step): walk the cell-sorted slots [*slot_lo, *slot_hi) (device values,
// e.g. two entries of cell_start) and evaluate row sorted_idx[slot] - row_lo for particles inside [row_lo, row_hi).
// `rows` of the launcher is then only the expected number of rows (grid sizing).
struct HtfSlab {
    const int *sorted_idx;
    const int *slot_lo, *slot_hi;
    long long row_lo, row_hi;
    int blocks_per_sm;      // > 0: cap the grid (the pass then shares every SM with the concurrent build instead of
                            // taking the whole machine for a moment)
};

cudaError_t htf_launch_lj(htf_ctx *ctx, const float4 *nlist, int64_t rows, int K, float4 *fe, float *virial,
                          int vcomp, const float *rdf_thr, int nb, const float *row_type, long long row_type_stride,
                          int type_i, int type_j, unsigned long long *bins, cudaStream_t st,
                          const HtfSlab *slab = nullptr, const int32_t *row_count = nullptr);

cudaError_t htf_launch_rdf(htf_ctx *ctx, const float4 *nlist, int64_t rows, int K, const float *row_type,
                           long long row_type_stride, const float *thr, int nb, int type_i, int type_j,
                           unsigned long long *bins, cudaStream_t st);

cudaError_t htf_ensure_tile_flags(htf_ctx *ctx, int ntiles);
cudaError_t htf_cell_stats(htf_ctx *ctx, int h_stats[3], cudaStream_t st);
// cudaFuncSetAttribute is per device: launchers cache what they have opted into per device slot
constexpr int HTF_MAX_DEVICES = 64;
inline int htf_current_device_slot()
{
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0) d = 0;
    return d % HTF_MAX_DEVICES;
}

cudaError_t htf_launch_eds_step(htf_ctx *ctx, const float *cv, const float *set_point, float *mean, float *ssd, int *n,
                                float *alpha, float *adam_m, float *adam_v, float *adam_t, int period, float lr,
                                float cv_scale, cudaStream_t st);
cudaError_t htf_launch_integrate(htf_ctx *ctx, int half, float4 *pos, float *vel, const float4 *force, int64_t n, float dt,
                                 float gamma, float kT, int flat, unsigned long long seed, unsigned long long step,
                                 cudaStream_t st);
cudaError_t htf_launch_unstuff4(htf_ctx *ctx, const float4 *in, float4 *out, int64_t n, cudaStream_t st);
cudaError_t htf_launch_skin_filter(htf_ctx *ctx, const float4 *pos, int64_t row_lo, int64_t row_hi, float4 *out,
                                   int32_t *idx_out, int32_t *count_out, int32_t *overflow, cudaStream_t st);
// destination table of the fused pack + send (lives in the rank's own comm window; see comm.cu)
struct HtfHaloDst {
    float4 *lo[2], *hi[2];           // [parity]: where this rank's low / high face goes (peer memory)
    unsigned long long epoch;        // halo exchanges completed so far
};
cudaError_t htf_launch_select_pair(htf_ctx *ctx, const float4 *pos, int64_t n, int axis, float thr_lo, float thr_hi,
                                   float4 *out_lo, float4 *out_hi, int cap, int *d_counts, int *d_overflow, cudaStream_t st,
                                   const HtfHaloDst *dst = nullptr);
void htf_comm_free(htf_ctx *ctx);
cudaError_t htf_launch_select(htf_ctx *ctx, const float4 *pos, int64_t n, int axis, float thr, bool less,
                              float4 *out, int cap, int *d_count, int *d_overflow, cudaStream_t st);

cudaError_t htf_launch_lj_cv(htf_ctx *ctx, const float4 *nlist, int64_t rows, int K, float4 *fe, float *virial,
                             int vcomp, float r0, float4 *cv_row, double *cv_sum, const float *rdf_thr, int nb,
                             unsigned long long *bins, cudaStream_t st, const HtfSlab *slab = nullptr,
                             const int32_t *row_count = nullptr);

int htf_mlp_packed_bytes_host();
int htf_mlp_raw_count_host();
cudaError_t htf_launch_mlp_pack(htf_ctx *ctx, const float *raw, unsigned char *packed, cudaStream_t st);
cudaError_t htf_launch_mlp(htf_ctx *ctx, const float4 *nlist, int64_t rows, int K, const unsigned char *packed,
                           float rbf_high, float4 *fe, cudaStream_t st, const int32_t *row_count = nullptr);

// pairwise-MLP training step (mlp_train.cu)
int htf_mlp_train_partial_floats(int sm_count);
cudaError_t htf_launch_mlp_train(htf_ctx *ctx, const float4 *nlist, int64_t rows, int K, const float *raw, float rbf_high,
                                 const float4 *pred, const float4 *labels, int64_t n_total, float *partial, double *loss_partial,
                                 float *grads, float *loss, cudaStream_t st);
cudaError_t htf_launch_adam(htf_ctx *ctx, float *params, const float *grads, float *m, float *v, float *t, int n, float lr,
                            float beta1, float beta2, float eps, cudaStream_t st);

// host: thresholds q_b (b = 1..nb-1) in rsq space such that bin(q) = #{b : q >= q_b}
void htf_rdf_thresholds(float r_lo, float r_hi, int nbins, float *thr /* [nbins+1] */);
