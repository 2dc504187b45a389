/*
 * htf_b200.h -- C ABI of libhtf_b200.so, the sm_100a implementation of hoomd-tf's
 * nlist -> forces+virial hot path.
 *
 * Each entry point names the reference interface it replaces (paths relative to the
 * hoomd-tf source tree, v2.4.0).  The reference crosses this boundary three ways:
 *   - the pybind11 module `_htf` (htf/module.cc:16-28, htf/TensorflowCompute.cc:422-486),
 *   - the TensorFlow custom ops HoomdToTf / TfToHoomd that memcpy through a CommStruct*
 *     baked into the graph as an integer attr (htf/hoomd2tf_op/hoomd2tf.cc:15-34,64-89;
 *     htf/tf2hoomd_op/tf2hoomd.cc:17-24,48-59; htf/CommStruct.h:29-104),
 *   - the C++ -> Python callback _finish_update (htf/TensorflowCompute.cc:219-226).
 * Here the whole path is one shared library with plain pointers and sizes; tensors are
 * owned by the caller (torch / DLPack / cudaMalloc -- the library does not care) and are
 * borrowed for the duration of the enqueue.
 *
 * Conventions
 *   - every function returns 0 on success or a negative HTF_E* code; nothing throws
 *     across the ABI; htf_last_error() returns a human-readable message.
 *   - all device work is enqueued on the caller's stream (a cudaStream_t passed as
 *     void*; NULL = the legacy default stream) and is asynchronous: no implicit device
 *     synchronisation (the reference ends every batch with cudaDeviceSynchronize,
 *     htf/TensorflowCompute.cc:208-211).  The exceptions are one-off or explicit: scratch growth
 *     (cudaMalloc/cudaFree) the first time a size is seen, the 12-byte density read-back of the first
 *     htf_build_nlist after htf_set_roi (restricted binnings only), and htf_skin_status.
 *   - one context per device, not re-entrant per context (the reference is single
 *     threaded too: htf/tf2hoomd_op/tf2hoomd.cc:11-12).
 *   - "pos" is always float[n][4] = (x, y, z, type) with the type stored as a float
 *     VALUE (what the reference produces with its unstuff4 kernel, htf/TFArrayComm.cu:9-15).
 *   - "nlist" is always float[rows][K][4] = (dx, dy, dz, type_j), zero padded, dx = minimum
 *     image of pos[j]-pos[i] (htf/TensorflowCompute.cc:353-367).
 *   - the box is orthorhombic: lo[3], hi[3], tilt[3] must be 0 (htf/simmodel.py:195).
 *   - pointers named d_* are DEVICE pointers, h_* are HOST pointers.
 */
#ifndef HTF_B200_H
#define HTF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HTF_OK            0
#define HTF_EINVAL       -1   /* bad argument */
#define HTF_ECUDA        -2   /* a CUDA runtime call or kernel launch failed */
#define HTF_ENOMEM       -3   /* device allocation failed */
#define HTF_ESTATE       -4   /* call order (e.g. box not set) */
#define HTF_ESKEW        -5   /* box has non-zero tilt factors */
#define HTF_EARCH        -6   /* device is not sm_100 */

/* htf_create flags */
#define HTF_FLAG_DETERMINISTIC  1   /* sort particles inside each cell by index: row-internal order and
                                       therefore fp32 force sums are bit-reproducible run to run */
#define HTF_FLAG_ANY_ARCH       2   /* do not refuse devices other than sm_100 (kernels still need sm_100a SASS) */

typedef struct htf_ctx htf_ctx;

/* ABI version of this header; htf_abi_version() of the loaded library must match. */
#define HTF_ABI_VERSION 18
int htf_abi_version(void);

/*
 * Replaces the TensorflowCompute constructor + reallocate()
 * (htf/TensorflowCompute.cc:31-121, GPU subclass :493-537): sizes the scratch that belongs
 * to the path (cell list, cell-sorted positions) for up to n_max particles.
 * k = nneighbor_cutoff (>=1), r_cut > 0.
 */
int htf_create(htf_ctx **out, int device, int64_t n_max, int k, float r_cut, int flags);
void htf_destroy(htf_ctx *ctx);

/* Message of the last failing call on ctx (ctx == NULL: last failing htf_create). */
const char *htf_last_error(const htf_ctx *ctx);

/*
 * Replaces updateBox() (htf/TensorflowCompute.cc:272-282) and the skew assertion of
 * SimModel.compute_inputs (htf/simmodel.py:195).  Host pointers.  Returns HTF_ESKEW when
 * |tilt| sums to >= 1e-4.
 */
int htf_set_box(htf_ctx *ctx, const float h_lo[3], const float h_hi[3], const float h_tilt[3]);

/* Mapped-nlist pair rule (htf/tensorflowcompute.py:284-305, TensorflowCompute::setMappedNlist
 * htf/TensorflowCompute.cc:472-478): pairs are listed iff both types are < map_type_start or
 * both are >= map_type_start.  map_type_start < 0 switches the rule off (default). */
int htf_set_mapped_nlist(htf_ctx *ctx, int map_type_start);

/* Change r_cut / nneighbor_cutoff of an existing context (tfcompute.attach arguments,
 * htf/tensorflowcompute.py:38-39). */
int htf_set_cutoff(htf_ctx *ctx, float r_cut, int k);

/*
 * Row sharding over GPUs: restrict the binning to the particles that can matter for this rank's rows.
 * Only particles whose minimum-image distance from h_center is <= h_half_width on every axis are binned
 * (h_half_width[a] < 0: no restriction on axis a; pass NULL, NULL to switch the region off).  The caller
 * chooses the region as the bounding interval of its rows plus r_cut plus a skin, the role HOOMD's ghost
 * layer width plays under MPI domain decomposition (htf/test-py/test_mpi_tensorflow.py:59-80).
 */
int htf_set_roi(htf_ctx *ctx, const float h_center[3], const float h_half_width[3]);

/*
 * Slab (halo) exchange between row shards: copies, in index order, every particle of d_pos[n] whose
 * coordinate on `axis` is < threshold (below != 0) or > threshold (below == 0) into d_out[capacity][4];
 * unused entries are set to a far-away sentinel that the region of interest rejects, so the buffer can be
 * sent to the neighbouring rank and appended to its positions as is.  d_count (nullable) receives the number
 * selected, d_overflow (nullable) is max'ed with it when it exceeds capacity.  This is the role of HOOMD's
 * ghost-particle exchange under MPI (htf/test-py/test_mpi_tensorflow.py:59-80); the reference has no code for it.
 */
int htf_pack_halo(htf_ctx *ctx, const float *d_pos, int64_t n, int axis, float threshold, int below,
                  float *d_out, int64_t capacity, int32_t *d_count, int32_t *d_overflow, void *stream);

/*
 * Both faces of a slab in one pass (5 launches instead of 12): particles with coordinate < threshold_lo go to
 * d_out_lo, those with coordinate > threshold_hi to d_out_hi, each in index order, sentinel padded to `capacity`;
 * d_counts (nullable) receives the two counts.  Same role and semantics as two htf_pack_halo calls.
 */
int htf_pack_halo_pair(htf_ctx *ctx, const float *d_pos, int64_t n, int axis, float threshold_lo, float threshold_hi,
                       float *d_out_lo, float *d_out_hi, int64_t capacity, int32_t *d_counts, int32_t *d_overflow,
                       void *stream);

/*
 * EDSLayer.call (htf/layers.py:142-195) as one launch on device-resident scalars: Welford mean / ssd of the collective
 * variable over the second half of every `period` calls, one tf.compat.v1 Adam step (beta1 .9, beta2 .999, eps 1e-8,
 * lr_t = lr sqrt(1 - b2^t) / (1 - b1^t)) on alpha at call period - 1, n <- (n + 1) mod period.  d_cv is the current
 * collective variable (device scalar); all state is fp32 except the int32 counter d_n.  The bias energy alpha * CV and
 * its forces are applied by the caller (htf_lj_cv_forces gives the CV's gradient row sums).
 */
int htf_eds_step(htf_ctx *ctx, const float *d_cv, const float *d_set_point, float *d_mean, float *d_ssd, int32_t *d_n,
                 float *d_alpha, float *d_adam_m, float *d_adam_v, float *d_adam_t, int period, float learning_rate,
                 float cv_scale, void *stream);

/*
 * The step either side of the path (what HOOMD's integrator does around the reference's compute:
 * hoomd.md.integrate.nve / langevin, htf/test-py/benchmark.py:39-48), unit mass, so that a trajectory stays on the GPU:
 *   half == 0 : v += dt/2 a;  x += dt v;  x wrapped into the box of htf_set_box;  flat != 0 keeps z = 0 (2-D systems)
 *   half == 1 : v += dt/2 a
 *   a = F, or F - gamma v + sqrt(2 gamma kT / (dt/2)) N(0,1) when gamma > 0 and kT > 0: every half kick carries its own
 *   random impulse (Philox, keyed by (seed, particle, timestep, half)), so that <v^2> relaxes to kT.
 * d_pos float[n][4] (w untouched), d_vel float[n][3], d_force float[n][4] (w = energy, ignored).
 */
int htf_integrate_half(htf_ctx *ctx, int half, float *d_pos, float *d_vel, const float *d_force, int64_t n, float dt,
                       float gamma, float kT, int flat, uint64_t seed, uint64_t timestep, void *stream);

/*
 * Buffered ("skin") neighbor lists -- the reference's own split of the work: HOOMD's NeighborList::compute searches
 * within r_cut + r_buff and runs only when particles have moved more than r_buff/2 (htf/TensorflowCompute.cc:163),
 * prepareNeighbors (htf/TensorflowCompute.cc:304-374) filters the candidates by r_cut EVERY step.
 *   htf_skin_configure : enable with a skin (HOOMD's r_buff) and a candidate capacity per row (0 = derived from K).
 *   htf_skin_rebuild   : the search.  Bins the particles with cell edge >= r_cut + skin, stores the candidate indices
 *                        of rows [row_lo,row_hi) and the positions they were built from.
 *   htf_skin_nlist     : the per-step pass.  Same output and arithmetic as htf_build_nlist (neighbor sets, values and
 *                        the modulo-K rule; slot order = candidate order), valid while no particle has moved more than
 *                        skin/2 since the rebuild; particle order and count must not change between rebuilds.
 *   htf_skin_status    : synchronising.  h_status[0] = rows whose particle had moved more than skin/2 when a list was
 *                        used, h_status[1] = rows whose candidate list overflowed its capacity; both must be 0.
 */
int htf_skin_configure(htf_ctx *ctx, float skin, int k_candidates);
int htf_skin_rebuild(htf_ctx *ctx, const float *d_pos_all, int64_t n_all, int64_t row_lo, int64_t row_hi, void *stream);
int htf_skin_nlist(htf_ctx *ctx, const float *d_pos_all, int64_t n_all, int64_t row_lo, int64_t row_hi,
                   float *d_nlist_out, int32_t *d_idx_out, int32_t *d_count_out, int32_t *d_overflow, void *stream);
int htf_skin_status(htf_ctx *ctx, int32_t h_status[2], int reset, void *stream);

/*
 * Replaces HOOMD's NeighborList::compute (called at htf/TensorflowCompute.cc:163) for this
 * path: bins all n_all particles into a cell list (cell edge >= r_cut) and writes the
 * cell-sorted copy used by htf_build_nlist.  Must be called again whenever positions change.
 */
int htf_bin_particles(htf_ctx *ctx, const float *d_pos_all, int64_t n_all, void *stream);

/*
 * Replaces prepareNeighbors (CPU htf/TensorflowCompute.cc:304-374; GPU :546-586 ->
 * htf_gpu_reshape_nlist htf/TensorflowCompute.cu:80-209), the memset before it (.cu:180) and
 * the HoomdToTf copy of the result (htf/hoomd2tf_op/hoomd2tf.cc:64-89).
 * Builds rows [row_lo,row_hi) of the padded neighbor tensor from the binned particles
 * (row batching = the reference's batch_size/offset chunking, htf/TensorflowCompute.cc:143-150).
 *   d_nlist_out   float[rows][K][4]
 *   d_idx_out     nullable int32[rows][K]: neighbor particle index per slot, -1 padded
 *                 (validation only; the reference has no equivalent)
 *   d_count_out   nullable int32[rows]: number of neighbors found (may exceed K: slots then
 *                 wrap modulo K exactly like htf/TensorflowCompute.cc:370)
 *   d_overflow    nullable int32[1]: atomically max'ed with count for every row with
 *                 count >= K ("Neighbor list is full!", htf/simmodel.py:216-224); the caller
 *                 zeroes it.
 */
int htf_build_nlist(htf_ctx *ctx, const float *d_pos_all, int64_t n_all, int64_t row_lo, int64_t row_hi,
                    float *d_nlist_out, int32_t *d_idx_out, int32_t *d_count_out, int32_t *d_overflow,
                    void *stream);

/*
 * Replaces the LJModel graph (htf/test-py/build_examples.py:67-77 = benchmark.py:12-23):
 * nlist_rinv (htf/simmodel.py:618-635) -> pair energy -> compute_nlist_forces (:526-555) ->
 * _add_energy (:558-578) -> _compute_virial (:509-523), and the TfToHoomd copies plus the
 * 3x3 -> 6 virial scatter (htf/tf2hoomd_op/tf2hoomd.cc:48-59, htf/TensorflowCompute.cc:285-301,
 * htf/TensorflowCompute.cu:41-71).
 *   k               second dimension of d_nlist
 *   d_row_count     nullable int32[rows]: htf_build_nlist's d_count_out for these rows.  With it the pass reads only
 *                   the first min(count, k) slots of a row -- the rest is the builder's zero padding, which adds
 *                   nothing to any sum (and lands in bin 0 of the histogram) -- about 27 % less DRAM traffic at
 *                   liquid density, bit-identical results.  NULL: every slot is read.  (Measured on B200: the pass
 *                   is then bound by instruction issue and no faster, so the fused steps htf_lj_step / htf_lj_rows /
 *                   htf_lj_cv_step only do this when HTF_ROW_COUNTS=1.)
 *   d_force_energy  float[rows][4] = (Fx, Fy, Fz, e_i)
 *   d_virial        nullable; virial_components = 6: float[rows][6] (xx,xy,xz,yy,yz,zz);
 *                   = 9: float[rows][9] row-major 3x3 (what get_virial_array returns,
 *                   htf/tensorflowcompute.py:388-392)
 */
int htf_lj_forces(htf_ctx *ctx, const float *d_nlist, int64_t rows, int k, const int32_t *d_row_count,
                  float *d_force_energy, float *d_virial, int virial_components, void *stream);

/*
 * htf_lj_forces with the compute_rdf histogram (below) fused into the same read of the neighbor
 * tensor: the LJ + RDF model of BASELINE config 2 (htf/test-py/build_examples.py:297-314) in one
 * pass.  No type filter; d_bins as in htf_rdf_hist.
 */
int htf_lj_forces_rdf(htf_ctx *ctx, const float *d_nlist, int64_t rows, int k, const int32_t *d_row_count,
                      float *d_force_energy, float *d_virial, int virial_components, int64_t *d_bins, float r_lo,
                      float r_hi, int nbins, void *stream);

/*
 * EDS-biased model of BASELINE config 5 in one pass over the neighbor tensor: htf_lj_forces (+ optional
 * fused RDF histogram as in htf_lj_forces_rdf) plus a smooth coordination-number collective variable
 *   cn_i = sum_j s(r_ij),  s(r) = (1 - (r/r0)^6) / (1 - (r/r0)^12) = 1 / (1 + (r/r0)^6),  CV = mean_i cn_i
 * (the differentiable form of the `mean(rinv > 0)` coordination number of sphinx-docs/source/running.rst:100-105;
 * r uses nlist_rinv's safe norm, padded slots are excluded exactly like there).
 *   d_cv_row  float[rows][4] = (sum_j ds/dd_ij (x, y, z), cn_i): the bias force on row i is
 *             2 * alpha / N * (x, y, z)   (compute_nlist_forces convention, htf/simmodel.py:542-550)
 *   d_cv_sum  double[1], atomically += sum_i cn_i (caller zeroes; all-reduce it across row shards)
 * EDSLayer (htf/layers.py:101-195) turns the CV into alpha on the host side of the ABI.
 */
int htf_lj_cv_forces(htf_ctx *ctx, const float *d_nlist, int64_t rows, int k, const int32_t *d_row_count, float r0,
                     float *d_force_energy, float *d_virial, int virial_components, float *d_cv_row, double *d_cv_sum,
                     int64_t *d_bins, float r_lo, float r_hi, int nbins, void *stream);

/*
 * Pairwise-MLP neural force field over the neighbor tensor (BASELINE config 3): the per-pair analogue of the
 * reference's learned models (`RBF`, htf/test-py/build_examples.py:231-241 with RBFExpansion htf/layers.py:7-49, and
 * the Dense stack of examples/08): r -> 32 radial basis features (centres linspace(0, rbf_high, 32)) ->
 * 3 x Dense(64, tanh) -> Dense(1); e_i = 1/2 sum_j u_ij over the non-padded slots; forces by
 * compute_nlist_forces' rule (htf/simmodel.py:542-550).  Forward and input gradient are six GEMMs per 128 pairs
 * on the tensor cores (tcgen05, bf16 operands, fp32 accumulate in TMEM); replaces Keras Dense + tf.gradients.
 *   htf_mlp_param_sizes  number of fp32 values of the raw parameter blob / bytes of the packed blob
 *   htf_mlp_pack         raw blob, torch.nn.Linear layout  W1[64][32] b1[64] W2[64][64] b2[64] W3[64][64] b3[64]
 *                        w4[64] b4[1]  ->  packed bf16 operand layouts (call again whenever the weights change)
 *   htf_mlp_forces       d_force_energy float[rows][4] = (Fx, Fy, Fz, e_i), overwritten.  d_row_count: nullable
 *                        int32[rows], htf_build_nlist's d_count_out: the compaction of the valid pairs that precedes
 *                        the kernel on large tensors then reads 4 bytes per row instead of the whole tensor twice
 */
int htf_mlp_param_sizes(int *raw_count, int *packed_bytes);
int htf_mlp_pack(htf_ctx *ctx, const float *d_raw, void *d_packed, void *stream);
int htf_mlp_forces(htf_ctx *ctx, const float *d_nlist, int64_t rows, int k, const int32_t *d_row_count, const void *d_packed,
                   float rbf_high, float *d_force_energy, void *stream);

/*
 * Replaces compute_rdf's histogram (htf/simmodel.py:638-669: masked_nlist :672-693, tf.norm,
 * tf.histogram_fixed_width over nbins+2 bins).  Adds this call's counts to d_bins
 * (int64[nbins+2], caller zeroes; bins 0 and nbins+1 are the ones the reference drops).
 *   d_row_type  nullable: type of the particle of each row, read as d_row_type[row * row_type_stride]
 *               (pass positions + 3 with stride 4 for `positions[:, 3]`) -- needed when
 *               type_i >= 0;  type_i / type_j < 0 = None.
 *   k           second dimension of d_nlist (any k >= 1, independent of the context's cutoff).
 */
int htf_rdf_hist(htf_ctx *ctx, const float *d_nlist, int64_t rows, int k, const float *d_row_type,
                 int64_t row_type_stride, float r_lo, float r_hi, int nbins, int type_i, int type_j,
                 int64_t *d_bins, void *stream);

/*
 * One pass of TensorflowCompute::computeForces for the built-in LJ model
 * (htf/TensorflowCompute.cc:130-216): bin, build rows [row_lo,row_hi), forces, virial and --
 * when d_bins != NULL -- the RDF histogram fused into the force pass.
 * d_nlist_out may be NULL: the context then keeps the tensor in its own scratch.
 */
int htf_lj_step(htf_ctx *ctx, const float *d_pos_all, int64_t n_all, int64_t row_lo, int64_t row_hi,
                float *d_nlist_out, float *d_force_energy, float *d_virial, int virial_components,
                int32_t *d_overflow,
                int64_t *d_bins, float r_lo, float r_hi, int nbins,
                void *stream);

/*
 * HOOMD's Scalar4 positions carry the particle type as the bit pattern of an int in .w; this library takes the
 * type as a float value.  Replaces TFArrayComm::receiveArray(..., unstuff4) and its kernel
 * (htf/TFArrayComm.h:86-130, htf/TFArrayComm.cu:9-28): one pass over n positions, in place when
 * d_pos_hoomd == d_pos_out.  A HOOMD-side binding calls this once per step on its position array.
 */
int htf_unstuff4(htf_ctx *ctx, const float *d_pos_hoomd, float *d_pos_out, int64_t n, void *stream);

/*
 * The row-batch form of htf_lj_step: rows [row_lo,row_hi) of the particles binned by the last htf_bin_particles
 * (n_all of them), no re-binning -- the reference's batch loop builds the neighbor list once and then walks
 * row chunks (htf/TensorflowCompute.cc:143-150,162-163,188-194).
 */
int htf_lj_rows(htf_ctx *ctx, int64_t n_all, int64_t row_lo, int64_t row_hi,
                float *d_nlist_out, float *d_force_energy, float *d_virial, int virial_components,
                int32_t *d_overflow,
                int64_t *d_bins, float r_lo, float r_hi, int nbins,
                void *stream);

/*
 * The same step for the EDS-biased model of BASELINE config 5: bin, build, then the fused LJ + smooth
 * coordination CV (+ RDF) pass of htf_lj_cv_forces (replaces, per step, the reference's
 * computeForces -> Python -> TF graph round trip for an EDSLayer model, htf/TensorflowCompute.cc:130-216,
 * htf/layers.py:101-195).  d_cv_sum and d_bins are accumulated into (zero them first).
 * d_pos_all == NULL: no re-binning, the rows of the last htf_bin_particles (n_all particles).
 */
int htf_lj_cv_step(htf_ctx *ctx, const float *d_pos_all, int64_t n_all, int64_t row_lo, int64_t row_hi,
                   float *d_nlist_out, float r0, float *d_force_energy, float *d_virial, int virial_components,
                   float *d_cv_row, double *d_cv_sum, int32_t *d_overflow,
                   int64_t *d_bins, float r_lo, float r_hi, int nbins, void *stream);

/*
 * htf_lj_step / htf_lj_cv_step pipeline the build and the pair pass: the z-window of cell layers is cut into
 * `slabs` slabs, slab i is built on the caller's stream and its pair pass runs on a stream owned by the context
 * while slab i+1 is being built (the build is instruction-issue bound, the pair pass HBM bound: side by side they
 * overlap, and the pass finds most of the slab still in L2).  The caller's stream waits for the last pass, so the
 * step is stream-ordered as a whole and can be captured into a CUDA graph (call it once outside the capture first:
 * the auxiliary stream and events are created on first use).  slabs <= 1 switches pipelining off, which is the
 * default: measured on B200 the build needs the whole register file for its own occupancy, so a co-resident pass
 * slows it by as much as it hides (cfg3: 0.63 vs 0.61 ms; cfg5: 3.60 vs 3.73 ms with 8 slabs) -- see DESIGN.md.
 * Environment override HTF_PIPE_SLABS.  Systems below 131072 rows are never pipelined.
 * The reference runs every stage back to back and ends with cudaDeviceSynchronize (htf/TensorflowCompute.cc:208-211).
 */
int htf_set_pipeline(htf_ctx *ctx, int slabs);

/*
 * Training step of the pairwise-MLP force field (BASELINE config 4, online force matching): what Keras does behind
 * model.train_on_batch(x=inputs, y=labels) in the reference's label mode (htf/tensorflowcompute.py:346-370, labels from
 * htf/TensorflowCompute.cc:177-187,:251-269) -- the gradient of MSE(compute_nlist_forces output [N,4], labels [N,4])
 * with respect to the parameters, i.e. a backward pass THROUGH the force gradient -- as one reverse sweep on the
 * tensor cores.  d_raw is the fp32 blob of htf_mlp_pack (10,497 values), d_grads receives the gradient in the same
 * layout, d_loss (nullable) this rank's share of the loss: sum over its rows / (4 n_total).  n_total = rows of ALL
 * ranks (the mean of the loss runs over them); with several ranks the caller sums d_grads / d_loss over the ranks
 * (htf_comm_allreduce_f32 or NCCL).  d_pred_out (nullable, [rows,4]) receives the model's forces + energy.
 */
int htf_mlp_train_grads(htf_ctx *ctx, const float *d_nlist, int64_t rows, int k, const float *d_raw, float rbf_high,
                        const float *d_labels, int64_t n_total, float *d_pred_out, float *d_grads, float *d_loss,
                        void *stream);

/*
 * One Adam step on count parameters, fused (m, v, bias-corrected step, update; the step counter d_t is a device scalar
 * that the call increments): tf.keras.optimizers.Adam as used by train_on_batch -- lr_t = lr sqrt(1 - b2^t) / (1 - b1^t),
 * p -= lr_t m / (sqrt(v) + epsilon).  Keras defaults: lr 1e-3, beta 0.9 / 0.999, epsilon 1e-7.
 */
int htf_adam_step(htf_ctx *ctx, float *d_params, const float *d_grads, float *d_m, float *d_v, float *d_t, int64_t count,
                  float learning_rate, float beta1, float beta2, float epsilon, void *stream);

/*
 * ---- exchange step of the row-sharded path over peer memory (NVLink / NVSwitch) ----
 * One process per GPU of one node.  The reference shards through HOOMD's MPI domain decomposition, whose ghost-particle
 * exchange is host MPI traffic every step (htf/test-py/test_mpi_tensorflow.py:59-80); here every rank owns a window
 * of device memory that its peers map (CUDA IPC), and the per-step exchange is device code only -- no NCCL, no host.
 *
 *   htf_comm_create   allocates this rank's window (halo receive buffers for `halo_capacity` particles per face,
 *                     all-reduce mailboxes, flags) and returns its 64-byte IPC handle in h_handle_out.
 *   htf_comm_connect  h_handles = the handles of ALL ranks, [world][64] in rank order (exchanged by the caller over
 *                     whatever it has: MPI in a HOOMD plugin, torch.distributed in this repository's host code).
 *   htf_comm_exchange_halo   rows are slabs along `axis`: packs the own particles (d_local[0..n_own)) below
 *                     threshold_lo and above threshold_hi (stable order, sentinel padded) DIRECTLY into the previous /
 *                     next rank's window, signals them, waits for their faces and stores those at
 *                     d_local[n_own .. n_own + halo_capacity) (from the next rank) and
 *                     d_local[n_own + halo_capacity .. n_own + 2 halo_capacity) (from the previous rank); d_local
 *                     is then what htf_bin_particles takes (with htf_set_roi rejecting the sentinels).
 *                     d_overflow (nullable) receives max(count + 1) when a face exceeds the capacity.
 *   htf_comm_allreduce_i64 / _f64 / _f32   in-place sum over all ranks of count <= 16384 values (the RDF histogram,
 *                     the CV sum and particle count of EDS, the ~10.5k weight gradients of the training step):
 *                     contributions are summed in rank order, so every rank gets the same bits.
 *   htf_comm_status   0, or != 0 when a wait gave up after 20 s (a peer died); synchronises the stream.
 * All exchange calls are stream-ordered, asynchronous and CUDA-graph capturable (epochs live in device memory).
 */
#define HTF_COMM_HANDLE_BYTES 64
int htf_comm_create(htf_ctx *ctx, int rank, int world, int64_t halo_capacity, unsigned char *h_handle_out);
int htf_comm_connect(htf_ctx *ctx, const unsigned char *h_handles);
int htf_comm_exchange_halo(htf_ctx *ctx, float *d_local, int64_t n_own, int axis, float threshold_lo, float threshold_hi,
                           int32_t *d_overflow, void *stream);
int htf_comm_allreduce_i64(htf_ctx *ctx, int64_t *d_values, int count, void *stream);
int htf_comm_allreduce_f64(htf_ctx *ctx, double *d_values, int count, void *stream);
int htf_comm_allreduce_f32(htf_ctx *ctx, float *d_values, int count, void *stream);
int htf_comm_status(htf_ctx *ctx, int32_t *h_status, void *stream);
int htf_comm_destroy(htf_ctx *ctx);

/* Number of kernels the library has launched on this context since creation
 * (bench.py's gpu_launches claim). */
int64_t htf_launch_count(const htf_ctx *ctx);

/* Cell grid chosen by the last htf_set_box/htf_set_cutoff: h_ncell[3]. */
int htf_get_cell_grid(const htf_ctx *ctx, int h_ncell[3]);

#ifdef __cplusplus
}
#endif
#endif /* HTF_B200_H */
