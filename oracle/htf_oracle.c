/*
 * htf_oracle.c -- CPU restatement of hoomd-tf's nlist -> forces+virial path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and only as the checker / the CPU
 * baseline.  The product path (hoomd-tf_b200/) never imports or links it.
 *
 * What is restated (all citations relative to /root/reference):
 *   - htf_oracle_nlist*      htf/TensorflowCompute.cc:304-374 (prepareNeighbors, CPU rule)
 *   - htf_oracle_lj          htf/simmodel.py:618-635 (nlist_rinv), :581-594 (safe_norm),
 *                            htf/test-py/build_examples.py:67-77 (LJModel),
 *                            htf/simmodel.py:526-578 (compute_nlist_forces, _add_energy),
 *                            htf/simmodel.py:509-523 (_compute_virial),
 *                            htf/TensorflowCompute.cc:285-301 (3x3 -> 6 virial layout)
 *   - htf_oracle_rdf_hist    htf/simmodel.py:638-693 (compute_rdf, masked_nlist)
 *
 * Third-party arithmetic that is NOT under /root/reference and is restated from
 * its published behaviour (see DESIGN.md "Oracle"):
 *   - HOOMD-blue (>=2.6, CI pins 2.7.0/2.8.2/2.9.2) BoxDim::minImage, CPU branch:
 *     per axis "if (w >= hi) w -= L; else if (w < lo) w += L", single image.  HOOMD boxes are
 *     centred on the origin (lo = -L/2, hi = L/2); the rule is applied with +-L/2 so that boxes
 *     with another origin (iter_from_trajectory uses lo = 0, htf/utils.py:702) behave the same.
 *   - TensorFlow (>=2.3, CI pins 2.3.2/2.4.1) tf.norm = sqrt(sum(x*x)),
 *     tf.histogram_fixed_width CPU kernel: step = double(hi-lo)/nbins,
 *     bin = int32(min(double(max(v,lo)-lo)/step, nbins-1)).
 *
 * Pinning status.  The reference stores no golden vectors for this path and
 * neither hoomd nor tensorflow can be imported in the build container, so the
 * oracle is pinned only through the reference's known-answer-by-construction
 * tests restated in tests/test_oracle_reference_cases.py (square / bcc lattices,
 * LJ vs analytic LJ, nlist vs O(N^2), virial vs pair virial, typed-RDF symmetry).
 * The neighbor tensor, LJ energy/force and virial are pinned that way.
 * PARITY UNPINNED: the RDF bin arithmetic (TF histogram rule is recalled, the
 * reference tests only check sum>0 and symmetry) -- stated here and in DESIGN.md.
 *
 * All arithmetic is fp32, round-to-nearest, no FMA contraction (compile with
 * -ffp-contract=off), operations in the written order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float lo[3], hi[3], L[3], half[3]; } obox_t;

static void make_box(const float *lo, const float *hi, obox_t *b)
{
    for (int a = 0; a < 3; a++) {
        b->lo[a] = lo[a]; b->hi[a] = hi[a]; b->L[a] = hi[a] - lo[a];
        b->half[a] = 0.5f * b->L[a];       /* HOOMD boxes are centred: half == hi, -half == lo */
    }
}

/* HOOMD BoxDim::minImage, CPU branch (called at htf/TensorflowCompute.cc:358). */
static inline void min_image(const obox_t *b, float *d)
{
    for (int a = 2; a >= 0; a--) {          /* z, then y, then x */
        if (d[a] >= b->half[a]) d[a] -= b->L[a];
        else if (d[a] < -b->half[a]) d[a] += b->L[a];
    }
}

/* One (i, j) candidate of the row loop, htf/TensorflowCompute.cc:341-371.
 * Returns 1 and fills d[] when the pair is kept. */
static inline int pair_kept(const obox_t *b, const float *pi, const float *pj, float rc2, float *d)
{
    d[0] = pj[0] - pi[0]; d[1] = pj[1] - pi[1]; d[2] = pj[2] - pi[2];   /* :353-355 */
    min_image(b, d);                                                     /* :358 */
    float rsq = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];                /* :359 */
    return !(rsq > rc2);
}

static inline void row_append(float *out_row, int32_t *idx_row, int32_t *cnt, int K,
                              const float *d, float type_j, int32_t j)
{
    int slot = (int)(*cnt % K);                                          /* :370 (wraps) */
    out_row[4 * slot + 0] = d[0]; out_row[4 * slot + 1] = d[1];
    out_row[4 * slot + 2] = d[2]; out_row[4 * slot + 3] = type_j;        /* :361-367 */
    if (idx_row) idx_row[slot] = j;
    *cnt += 1;
}

/* mapped-nlist pair rule, htf/tensorflowcompute.py:297-304: a pair is listed iff both
 * particles are all-atom or both are mapped beads.  map_type_start < 0 disables it. */
static inline int pair_allowed(float ti, float tj, int map_type_start)
{
    if (map_type_start < 0) return 1;
    return ((int)ti >= map_type_start) == ((int)tj >= map_type_start);
}

/*
 * Neighbor tensor, brute force over all j != i in ascending j (any candidate superset
 * gives the same set because of the r_cut filter, htf/TensorflowCompute.cc:359).
 * pos: [n_all,4] (x,y,z,type-as-float); out: [rows,K,4] zero padded (:311);
 * idx: nullable [rows,K], -1 padded; count: nullable [rows] = number of kept pairs
 * (may exceed K: the reference then wraps modulo K, :370).
 */
int htf_oracle_nlist(const float *pos, int64_t n_all, const float *lo, const float *hi,
                     float r_cut, int K, int64_t row_lo, int64_t row_hi, int map_type_start,
                     float *out, int32_t *idx, int32_t *count)
{
    if (K < 1 || row_lo < 0 || row_hi > n_all || row_lo > row_hi) return -1;
    obox_t b; make_box(lo, hi, &b);
    const float rc2 = r_cut * r_cut;
    const int64_t rows = row_hi - row_lo;
    memset(out, 0, sizeof(float) * 4 * (size_t)K * (size_t)rows);
    if (idx) for (int64_t t = 0; t < rows * K; t++) idx[t] = -1;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = row_lo; i < row_hi; i++) {
        const int64_t r = i - row_lo;
        int32_t cnt = 0;
        for (int64_t j = 0; j < n_all; j++) {
            float d[3];
            if (j == i) continue;
            if (!pair_allowed(pos[4 * i + 3], pos[4 * j + 3], map_type_start)) continue;
            if (!pair_kept(&b, pos + 4 * i, pos + 4 * j, rc2, d)) continue;
            row_append(out + 4 * (size_t)K * r, idx ? idx + (size_t)K * r : NULL, &cnt, K,
                       d, pos[4 * j + 3], (int32_t)j);
        }
        if (count) count[r] = cnt;
    }
    return 0;
}

/* ---- cell-list candidate generation (stands in for HOOMD's NeighborList::compute,
 * called at htf/TensorflowCompute.cc:163).  Same result as the brute force version,
 * including slot order (hits are emitted in ascending j). ---- */
typedef struct { int n[3]; float w_inv[3]; int64_t *start; int32_t *items; } ocells_t;

static inline int cell_of(const obox_t *b, const ocells_t *c, const float *p, int a)
{
    int v = (int)floor(((double)p[a] - (double)b->lo[a]) * (double)c->w_inv[a]);
    if (v < 0) v = 0;
    if (v >= c->n[a]) v = c->n[a] - 1;
    return v;
}

static int cmp_i32(const void *a, const void *b)
{
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

int htf_oracle_nlist_cells(const float *pos, int64_t n_all, const float *lo, const float *hi,
                           float r_cut, int K, int64_t row_lo, int64_t row_hi, int map_type_start,
                           float *out, int32_t *idx, int32_t *count)
{
    if (K < 1 || row_lo < 0 || row_hi > n_all || row_lo > row_hi) return -1;
    obox_t b; make_box(lo, hi, &b);
    ocells_t c;
    int64_t ncell = 1;
    for (int a = 0; a < 3; a++) {
        /* cell edge >= 1.001 r_cut so that a rounding error in cell_of can never hide a pair */
        int n = (int)floor((double)b.L[a] / ((double)r_cut * 1.001));
        if (n < 1) n = 1;
        c.n[a] = n; c.w_inv[a] = (float)((double)n / (double)b.L[a]);
        ncell *= n;
    }
    c.start = (int64_t *)calloc((size_t)ncell + 1, sizeof(int64_t));
    c.items = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_all > 0 ? n_all : 1));
    int32_t *cid = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_all > 0 ? n_all : 1));
    if (!c.start || !c.items || !cid) return -2;
    for (int64_t i = 0; i < n_all; i++) {
        int cx = cell_of(&b, &c, pos + 4 * i, 0), cy = cell_of(&b, &c, pos + 4 * i, 1),
            cz = cell_of(&b, &c, pos + 4 * i, 2);
        cid[i] = (cz * c.n[1] + cy) * c.n[0] + cx;
        c.start[cid[i] + 1]++;
    }
    for (int64_t q = 0; q < ncell; q++) c.start[q + 1] += c.start[q];
    {
        int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * (size_t)ncell);
        memcpy(fill, c.start, sizeof(int64_t) * (size_t)ncell);
        for (int64_t i = 0; i < n_all; i++) c.items[fill[cid[i]]++] = (int32_t)i;  /* ascending i per cell */
        free(fill);
    }
    const float rc2 = r_cut * r_cut;
    const int64_t rows = row_hi - row_lo;
    memset(out, 0, sizeof(float) * 4 * (size_t)K * (size_t)rows);
    if (idx) for (int64_t t = 0; t < rows * K; t++) idx[t] = -1;
#pragma omp parallel
    {
        int32_t cap = 1024, *cand = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = row_lo; i < row_hi; i++) {
            const int64_t r = i - row_lo;
            int ci[3] = { cid[i] % c.n[0], (cid[i] / c.n[0]) % c.n[1], cid[i] / (c.n[0] * c.n[1]) };
            int lst[3][3], nl[3];
            for (int a = 0; a < 3; a++) {           /* unique cells of the stencil in each dimension */
                if (c.n[a] <= 3) { nl[a] = c.n[a]; for (int q = 0; q < c.n[a]; q++) lst[a][q] = q; }
                else {
                    nl[a] = 3;
                    lst[a][0] = (ci[a] + c.n[a] - 1) % c.n[a]; lst[a][1] = ci[a]; lst[a][2] = (ci[a] + 1) % c.n[a];
                }
            }
            int32_t nc = 0;
            for (int z = 0; z < nl[2]; z++) for (int y = 0; y < nl[1]; y++) for (int x = 0; x < nl[0]; x++) {
                int64_t q = ((int64_t)lst[2][z] * c.n[1] + lst[1][y]) * c.n[0] + lst[0][x];
                for (int64_t t = c.start[q]; t < c.start[q + 1]; t++) {
                    if (nc == cap) { cap *= 2; cand = (int32_t *)realloc(cand, sizeof(int32_t) * (size_t)cap); }
                    cand[nc++] = c.items[t];
                }
            }
            qsort(cand, (size_t)nc, sizeof(int32_t), cmp_i32);
            int32_t cnt = 0;
            for (int32_t t = 0; t < nc; t++) {
                const int64_t j = cand[t];
                float d[3];
                if (j == i) continue;
                if (!pair_allowed(pos[4 * i + 3], pos[4 * j + 3], map_type_start)) continue;
                if (!pair_kept(&b, pos + 4 * i, pos + 4 * j, rc2, d)) continue;
                row_append(out + 4 * (size_t)K * r, idx ? idx + (size_t)K * r : NULL, &cnt, K,
                           d, pos[4 * j + 3], (int32_t)j);
            }
            if (count) count[r] = cnt;
        }
        free(cand);
    }
    free(c.start); free(c.items); free(cid);
    return 0;
}

/*
 * LJModel forces, per-particle energy and virial from the neighbor tensor.
 * nlist [rows,K,4] -> force_energy [rows,4] = (Fx,Fy,Fz,e_i)  (htf/simmodel.py:574-577)
 *                     virial9 (nullable) [rows,9] row-major 3x3 (htf/simmodel.py:509-523)
 *                     virial6 (nullable) [rows,6] = xx,xy,xz,yy,yz,zz (htf/TensorflowCompute.cc:294-299)
 */
int htf_oracle_lj(const float *nlist, int64_t rows, int K, float *force_energy, float *virial9, float *virial6)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; r++) {
        float f[3] = { 0.f, 0.f, 0.f }, e = 0.f, v[9] = { 0 };
        for (int s = 0; s < K; s++) {
            const float *d = nlist + 4 * ((size_t)K * r + s);
            /* nlist_rinv: safe_norm(delta = 3e-6/3/10 = 1e-7), htf/simmodel.py:630-631 */
            float ax = d[0] + 1e-7f, ay = d[1] + 1e-7f, az = d[2] + 1e-7f;
            float rt = sqrtf(ax * ax + ay * ay + az * az);
            float si = (rt > 3e-6f) ? 1.0f / (rt + 3e-6f) : 0.0f;      /* :632-635 */
            float s2 = si * si, s6 = s2 * s2 * s2;                       /* rinv**6 */
            e += 2.0f * (s6 * s6 - s6);                                  /* build_examples.py:72-75 */
            if (si == 0.0f) continue;                                    /* zero gradient: padded slot */
            /* 2 * d(sum_i e_i)/d(nlist[i,s,:]) in closed form, htf/simmodel.py:542-548 */
            float coef = (24.0f * s6 * si - 48.0f * s6 * s6 * si) / rt;
            float fx = coef * ax, fy = coef * ay, fz = coef * az;
            f[0] += fx; f[1] += fy; f[2] += fz;                          /* :549-550 */
            if (virial9 || virial6) {
                float rm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);   /* :514-515, un-offset d */
                float fm = sqrtf(fx * fx + fy * fy + fz * fz);              /* :516-517, w-gradient is 0 */
                float w = (rm == 0.0f) ? 0.0f : fm / (2.0f * rm);           /* divide_no_nan :518 */
                for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) v[3 * k + l] += w * d[k] * d[l];
            }
        }
        force_energy[4 * r + 0] = f[0]; force_energy[4 * r + 1] = f[1];
        force_energy[4 * r + 2] = f[2]; force_energy[4 * r + 3] = e;
        if (virial9) for (int q = 0; q < 9; q++) virial9[9 * r + q] = -1.0f * v[q];   /* :521 */
        if (virial6) {
            static const int pick[6] = { 0, 1, 2, 4, 5, 8 };
            for (int q = 0; q < 6; q++) virial6[6 * r + q] = -1.0f * v[pick[q]];
        }
    }
    return 0;
}

/*
 * compute_rdf bin counts (htf/simmodel.py:638-669) with the masked_nlist type filter
 * (:672-693).  hist has nbins+2 entries; the reference drops the first and last (:668).
 * type_i / type_j < 0 mean "None".  row_type: [rows] particle type of each row
 * (positions[:,3]); may be NULL when type_i < 0.
 */
int htf_oracle_rdf_hist(const float *nlist, int64_t rows, int K, const float *row_type,
                        float r_lo, float r_hi, int nbins, int type_i, int type_j, int64_t *hist)
{
    const int nb = nbins + 2;
    for (int q = 0; q < nb; q++) hist[q] = 0;
    const double step = (double)(float)(r_hi - r_lo) / (double)nb;
    const double last = (double)(nb - 1);
    for (int64_t r = 0; r < rows; r++) {
        if (type_i >= 0 && !(row_type[r] == (float)type_i)) continue;      /* boolean_mask :686-688 */
        for (int s = 0; s < K; s++) {
            const float *d = nlist + 4 * ((size_t)K * r + s);
            float m = 1.0f;
            if (type_j >= 0) m = (d[3] == (float)type_j) ? 1.0f : 0.0f;    /* :689-692 */
            float x = d[0] * m, y = d[1] * m, z = d[2] * m;
            float rr = sqrtf(x * x + y * y + z * z);                        /* tf.norm :661 */
            float v = rr > r_lo ? rr : r_lo;
            double q = (double)(float)(v - r_lo) / step;
            if (q > last) q = last;
            hist[(int32_t)q] += 1;
        }
    }
    return 0;
}

/*
 * Smooth coordination-number CV of BASELINE config 5: cn_i = sum_j 1/(1 + (rt/r0)^6) over the
 * non-padded slots, rt = nlist_rinv's safe norm (htf/simmodel.py:630-631); grad[i] = sum_j d s/d d_ij.
 * The reference has no implementation of this CV: this is the specification the kernel is
 * held to, the differentiable form of `mean(rinv > 0)` (sphinx-docs/source/running.rst:100-105).
 */
int htf_oracle_cv(const float *nlist, int64_t rows, int K, float r0, float *coord, float *grad)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; r++) {
        float cn = 0.f, g[3] = { 0.f, 0.f, 0.f };
        for (int s = 0; s < K; s++) {
            const float *d = nlist + 4 * ((size_t)K * r + s);
            float ax = d[0] + 1e-7f, ay = d[1] + 1e-7f, az = d[2] + 1e-7f;
            float rt = sqrtf(ax * ax + ay * ay + az * az);
            if (!(rt > 3e-6f)) continue;
            float x = rt / r0, x2 = x * x, x6 = x2 * x2 * x2;
            float sw = 1.0f / (1.0f + x6);
            cn += sw;
            float cg = -6.0f * x6 * sw * sw / (rt * rt);
            g[0] += cg * ax; g[1] += cg * ay; g[2] += cg * az;
        }
        coord[r] = cn;
        grad[3 * r + 0] = g[0]; grad[3 * r + 1] = g[1]; grad[3 * r + 2] = g[2];
    }
    return 0;
}

int htf_oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Number of OpenMP threads for the following calls (bench.py's CPU legs: torchrun exports
 * OMP_NUM_THREADS=1, which would otherwise time one core).  n <= 0: all online processors. */
int htf_oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n <= 0) n = omp_get_num_procs();
    omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}
