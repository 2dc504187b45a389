// Exchange step of the row-sharded path over PEER MEMORY (NVLink / NVSwitch), behind the C ABI.
//
// The reference has no multi-GPU code of its own: it inherits HOOMD's MPI domain decomposition, whose ghost
// exchange goes through host MPI calls every step (/root/reference htf/test-py/test_mpi_tensorflow.py:59-80).
// Here one process drives one GPU; every rank owns a "window" (one cudaMalloc, exported with cudaIpcGetMemHandle)
// that its peers map with cudaIpcOpenMemHandle.  The per-step traffic never touches the host and needs no NCCL:
//
//   halo exchange   the stable compaction of the two slab faces (select_fused2_kernel, cells.cu) writes the faces
//                   STRAIGHT INTO the neighbours' windows (fused pack + send); the gather kernel then publishes the
//                   epoch to the neighbours' flags (release, system scope), waits for both own flags (acquire) and
//                   copies the received faces behind the rank's own rows, where the binning expects them.
//                   Two launches in all: the one-launch compaction (select_fused2_kernel, cooperative; count + one-block
//                   scan + scatter-to-peer when it cannot be used) and signal + wait + gather.
//   all-reduce      every rank stores its (small) vector into every peer's window, publishes the epoch, waits for all
//                   contributions and sums them in rank order -- the same bits on every rank, int64 or fp64.
//
// Buffers are double buffered by epoch parity.  That is enough without a "consumed" handshake: a rank can only be
// one exchange ahead of a neighbour (it waits for that neighbour's face of the current epoch before it proceeds), so
// when it writes parity p of epoch e+2 the neighbour has finished epoch e.  Epochs live in device memory, so a CUDA
// graph that captured the exchange advances them by itself on every replay.
#include "common.cuh"
#include "../../include/htf_b200.h"

#include <cstdio>
#include <cstring>
#include <new>

namespace {

constexpr int COMM_MAX_WORLD = 16;
constexpr int COMM_AR_MAX = 16384;                    // values per all-reduce (mailbox slots are 8 bytes wide)
constexpr unsigned long long COMM_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;

// header of a window; peers write the flags and the data areas, the owner reads them
struct CommHeader {
    unsigned long long halo_flag[2];                  // [0] face from the next rank, [1] from the previous: epochs done
    unsigned long long ar_flag[COMM_MAX_WORLD];       // all-reduce epochs done, per source rank
    unsigned long long ar_epoch;                      // all-reduces completed by the owner
    int status;                                       // != 0: a wait timed out (sticky)
    int blocks_done;                                  // last-block-done counter of the gather kernel
    HtfHaloDst dst;                                   // where the owner's faces go + the halo epoch
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// spin until *flag >= want; false on timeout
__device__ __forceinline__ bool wait_flag(const unsigned long long *flag, unsigned long long want)
{
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(flag) < want) {
        __nanosleep(64);
        if (global_ns() - t0 > COMM_TIMEOUT_NS) return false;
    }
    return true;
}

// After the pack kernel (stream order: this rank's faces are in the neighbours' windows): block 0 publishes the epoch
// to both neighbours; every block then waits for both faces of this epoch and copies its share of them behind the
// rank's own rows: local[n_own + {0, cap} ...]
__global__ void __launch_bounds__(256) halo_gather_kernel(CommHeader *own, unsigned long long *flag_at_prev,
                                                          unsigned long long *flag_at_next,
                                                          const float4 *recv /* [2 parity][2 side][cap] */,
                                                          float4 *local_halo /* [2][cap] */, int cap)
{
    const unsigned long long e = *reinterpret_cast<volatile unsigned long long *>(&own->dst.epoch);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        __threadfence_system();
        st_release_sys(flag_at_prev, e + 1);          // "the face from your next rank is complete"
        st_release_sys(flag_at_next, e + 1);          // "the face from your previous rank is complete"
    }
    if (threadIdx.x == 0) {
        const bool ok = wait_flag(&own->halo_flag[0], e + 1) && wait_flag(&own->halo_flag[1], e + 1);
        if (!ok) atomicExch(&own->status, 1);
    }
    __syncthreads();
    const float4 *src = recv + (size_t)(e & 1ull) * 2 * cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * cap; i += gridDim.x * blockDim.x) local_halo[i] = src[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&own->blocks_done, 1) == (int)gridDim.x - 1) {      // every block has read the epoch and copied
            own->blocks_done = 0;
            own->dst.epoch = e + 1;
        }
    }
}

// block p stores this rank's vector into peer p's window, then publishes the epoch there
template <typename T>
__global__ void __launch_bounds__(256) ar_push_kernel(const CommHeader *own, const T *values, int count, int rank,
                                                      T *const *peer_data /* [world]: base of ar_data in peer p */,
                                                      unsigned long long *const *peer_flag /* [world]: &ar_flag[rank] in peer p */)
{
    const unsigned long long e = own->ar_epoch;
    T *dst = peer_data[blockIdx.x] + ((size_t)(e & 1ull) * COMM_MAX_WORLD + rank) * COMM_AR_MAX;
    for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = values[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(peer_flag[blockIdx.x], e + 1);
}

template <typename T>
__global__ void __launch_bounds__(256) ar_sum_kernel(CommHeader *own, const T *data /* own ar_data */, T *values, int count,
                                                     int world)
{
    const unsigned long long e = own->ar_epoch;
    if (threadIdx.x < world) {
        if (!wait_flag(&own->ar_flag[threadIdx.x], e + 1)) atomicExch(&own->status, 2);
    }
    __syncthreads();
    const T *src = data + (size_t)(e & 1ull) * COMM_MAX_WORLD * COMM_AR_MAX;
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        T acc = 0;
        for (int r = 0; r < world; r++) acc += src[(size_t)r * COMM_AR_MAX + i];      // rank order: same bits everywhere
        values[i] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) own->ar_epoch = e + 1;
}

}  // namespace

struct HtfComm {
    int rank, world;
    int64_t cap;
    unsigned char *window;            // own window (device)
    size_t window_bytes;
    size_t off_ar, off_halo;          // byte offsets of the data areas
    unsigned char *peer[COMM_MAX_WORLD];      // mapped windows (peer[rank] == window)
    bool connected;
    void **d_peer_data;               // device tables for the all-reduce push
    unsigned long long **d_peer_flag;
};

namespace {
void set_comm_err(htf_ctx *ctx, const char *what, cudaError_t e)
{
    snprintf(ctx->err, sizeof(ctx->err), "%s: %s", what, cudaGetErrorString(e));
}
}  // namespace

void htf_comm_free(htf_ctx *ctx)
{
    HtfComm *c = ctx->comm;
    if (!c) return;
    for (int p = 0; p < c->world; p++)
        if (p != c->rank && c->peer[p]) cudaIpcCloseMemHandle(c->peer[p]);
    if (c->d_peer_data) cudaFree(c->d_peer_data);
    if (c->d_peer_flag) cudaFree(c->d_peer_flag);
    if (c->window) cudaFree(c->window);
    delete c;
    ctx->comm = nullptr;
}

extern "C" {

int htf_comm_create(htf_ctx *ctx, int rank, int world, int64_t halo_capacity, unsigned char *h_handle_out)
{
    if (!ctx || !h_handle_out) return HTF_EINVAL;
    if (world < 2 || world > COMM_MAX_WORLD || rank < 0 || rank >= world || halo_capacity < 1 || halo_capacity > 500000000LL) {
        snprintf(ctx->err, sizeof(ctx->err), "htf_comm_create: need 2 <= world <= %d, 0 <= rank < world, capacity >= 1", COMM_MAX_WORLD);
        return HTF_EINVAL;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == HTF_COMM_HANDLE_BYTES, "handle size");
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    cudaSetDevice(ctx->device);
    htf_comm_free(ctx);
    HtfComm *c = new (std::nothrow) HtfComm();
    if (!c) return HTF_ENOMEM;
    memset(c, 0, sizeof(*c));
    c->rank = rank; c->world = world; c->cap = halo_capacity;
    c->off_ar = (sizeof(CommHeader) + 1023) / 1024 * 1024;
    c->off_halo = c->off_ar + (size_t)2 * COMM_MAX_WORLD * COMM_AR_MAX * 8;
    c->window_bytes = c->off_halo + (size_t)4 * halo_capacity * sizeof(float4);
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&c->window), c->window_bytes);
    if (e == cudaSuccess) e = cudaMemset(c->window, 0, c->window_bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, c->window);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&c->d_peer_data), sizeof(void *) * COMM_MAX_WORLD);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&c->d_peer_flag), sizeof(void *) * COMM_MAX_WORLD);
    if (e != cudaSuccess) {
        set_comm_err(ctx, "htf_comm_create", e);
        ctx->comm = c;
        htf_comm_free(ctx);
        (void)cudaGetLastError();
        if (prev_dev >= 0) cudaSetDevice(prev_dev);
        return e == cudaErrorMemoryAllocation ? HTF_ENOMEM : HTF_ECUDA;
    }
    memcpy(h_handle_out, &h, sizeof(h));
    c->peer[rank] = c->window;
    ctx->comm = c;
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
    return HTF_OK;
}

int htf_comm_connect(htf_ctx *ctx, const unsigned char *h_handles)
{
    if (!ctx || !ctx->comm || !h_handles) { if (ctx) snprintf(ctx->err, sizeof(ctx->err), "htf_comm_connect: call htf_comm_create first"); return HTF_ESTATE; }
    HtfComm *c = ctx->comm;
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    cudaSetDevice(ctx->device);
    int rc = HTF_OK;
    for (int p = 0; p < c->world && rc == HTF_OK; p++) {
        if (p == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, h_handles + (size_t)p * HTF_COMM_HANDLE_BYTES, sizeof(h));
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { set_comm_err(ctx, "htf_comm_connect: cudaIpcOpenMemHandle", e); (void)cudaGetLastError(); rc = HTF_ECUDA; break; }
        c->peer[p] = static_cast<unsigned char *>(ptr);
    }
    if (rc == HTF_OK) {
        // where this rank's faces go: low face -> previous rank's "from next" buffer (side 0), high face -> next rank's
        // "from previous" buffer (side 1)
        const int prev = (c->rank + c->world - 1) % c->world, next = (c->rank + 1) % c->world;
        HtfHaloDst d;
        memset(&d, 0, sizeof(d));
        for (int par = 0; par < 2; par++) {
            d.lo[par] = reinterpret_cast<float4 *>(c->peer[prev] + c->off_halo) + ((size_t)par * 2 + 0) * c->cap;
            d.hi[par] = reinterpret_cast<float4 *>(c->peer[next] + c->off_halo) + ((size_t)par * 2 + 1) * c->cap;
        }
        d.epoch = 0;
        void *pd[COMM_MAX_WORLD];
        unsigned long long *pf[COMM_MAX_WORLD];
        for (int p = 0; p < COMM_MAX_WORLD; p++) {
            pd[p] = p < c->world ? c->peer[p] + c->off_ar : nullptr;
            pf[p] = p < c->world ? &reinterpret_cast<CommHeader *>(c->peer[p])->ar_flag[c->rank] : nullptr;
        }
        cudaError_t e = cudaMemcpy(&reinterpret_cast<CommHeader *>(c->window)->dst, &d, sizeof(d), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(c->d_peer_data, pd, sizeof(pd), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(c->d_peer_flag, pf, sizeof(pf), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { set_comm_err(ctx, "htf_comm_connect", e); rc = HTF_ECUDA; }
        else c->connected = true;
    }
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
    return rc;
}

int htf_comm_exchange_halo(htf_ctx *ctx, float *d_local, int64_t n_own, int axis, float threshold_lo, float threshold_hi,
                           int32_t *d_overflow, void *stream)
{
    if (!ctx) return HTF_EINVAL;
    HtfComm *c = ctx->comm;
    if (!c || !c->connected) { snprintf(ctx->err, sizeof(ctx->err), "htf_comm_exchange_halo: call htf_comm_create / htf_comm_connect first"); return HTF_ESTATE; }
    if (!d_local || n_own < 0 || n_own > 1000000000LL || axis < 0 || axis > 2) { snprintf(ctx->err, sizeof(ctx->err), "htf_comm_exchange_halo: bad arguments"); return HTF_EINVAL; }
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    cudaSetDevice(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = HTF_OK;
    // scratch of the compaction (per 256-particle block), as in htf_pack_halo_pair
    const int64_t need = 2 * ((n_own + 255) / 256) + 2;
    if (need > ctx->sel_cap) {
        cudaError_t e = cudaSuccess;
        if (ctx->d_sel_cnt) cudaFree(ctx->d_sel_cnt);
        if (ctx->d_sel_off) cudaFree(ctx->d_sel_off);
        if (ctx->d_sel_sums) cudaFree(ctx->d_sel_sums);
        ctx->d_sel_cnt = ctx->d_sel_off = ctx->d_sel_sums = nullptr;
        e = cudaMalloc(reinterpret_cast<void **>(&ctx->d_sel_cnt), sizeof(int) * (size_t)need);
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&ctx->d_sel_off), sizeof(int) * (size_t)need);
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&ctx->d_sel_sums), sizeof(int) * ((size_t)need / 1024 + 4));
        if (e != cudaSuccess) { set_comm_err(ctx, "htf_comm_exchange_halo", e); rc = HTF_ENOMEM; }
        else ctx->sel_cap = need;
    }
    if (rc == HTF_OK) {
        CommHeader *own = reinterpret_cast<CommHeader *>(c->window);
        const int prev = (c->rank + c->world - 1) % c->world, next = (c->rank + 1) % c->world;
        float4 *local = reinterpret_cast<float4 *>(d_local);
        // 1. fused pack + send: both faces, stable order, sentinel padded, written into the neighbours' windows
        cudaError_t e = htf_launch_select_pair(ctx, local, n_own, axis, threshold_lo, threshold_hi, nullptr, nullptr, (int)c->cap,
                                               nullptr, d_overflow, st, &own->dst);
        // 2. publish the epoch to both neighbours, wait for both of their faces, copy them behind the own rows
        if (e == cudaSuccess) {
            const int blocks = (int)((2 * c->cap + 255) / 256 < 64 ? (2 * c->cap + 255) / 256 : 64);
            halo_gather_kernel<<<blocks, 256, 0, st>>>(own, &reinterpret_cast<CommHeader *>(c->peer[prev])->halo_flag[0],
                                                       &reinterpret_cast<CommHeader *>(c->peer[next])->halo_flag[1],
                                                       reinterpret_cast<const float4 *>(c->window + c->off_halo),
                                                       local + n_own, (int)c->cap);
            ctx->launches += 1;
            e = cudaGetLastError();
        }
        if (e != cudaSuccess) { set_comm_err(ctx, "htf_comm_exchange_halo", e); rc = HTF_ECUDA; }
    }
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
    return rc;
}

static int comm_allreduce(htf_ctx *ctx, void *d_values, int count, int kind /* 0 i64, 1 f64, 2 f32 */, void *stream)
{
    if (!ctx) return HTF_EINVAL;
    HtfComm *c = ctx->comm;
    if (!c || !c->connected) { snprintf(ctx->err, sizeof(ctx->err), "htf_comm_allreduce: call htf_comm_create / htf_comm_connect first"); return HTF_ESTATE; }
    if (!d_values || count < 1 || count > COMM_AR_MAX) { snprintf(ctx->err, sizeof(ctx->err), "htf_comm_allreduce: 1 <= count <= %d", COMM_AR_MAX); return HTF_EINVAL; }
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    cudaSetDevice(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    CommHeader *own = reinterpret_cast<CommHeader *>(c->window);
    if (kind == 2) {
        ar_push_kernel<float><<<c->world, 256, 0, st>>>(own, static_cast<const float *>(d_values), count, c->rank,
                                                        reinterpret_cast<float *const *>(c->d_peer_data), c->d_peer_flag);
        ar_sum_kernel<float><<<1, 256, 0, st>>>(own, reinterpret_cast<const float *>(c->window + c->off_ar),
                                                static_cast<float *>(d_values), count, c->world);
    } else if (kind == 1) {
        ar_push_kernel<double><<<c->world, 256, 0, st>>>(own, static_cast<const double *>(d_values), count, c->rank,
                                                         reinterpret_cast<double *const *>(c->d_peer_data), c->d_peer_flag);
        ar_sum_kernel<double><<<1, 256, 0, st>>>(own, reinterpret_cast<const double *>(c->window + c->off_ar),
                                                 static_cast<double *>(d_values), count, c->world);
    } else {
        ar_push_kernel<long long><<<c->world, 256, 0, st>>>(own, static_cast<const long long *>(d_values), count, c->rank,
                                                            reinterpret_cast<long long *const *>(c->d_peer_data), c->d_peer_flag);
        ar_sum_kernel<long long><<<1, 256, 0, st>>>(own, reinterpret_cast<const long long *>(c->window + c->off_ar),
                                                    static_cast<long long *>(d_values), count, c->world);
    }
    ctx->launches += 2;
    cudaError_t e = cudaGetLastError();
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
    if (e != cudaSuccess) { set_comm_err(ctx, "htf_comm_allreduce", e); return HTF_ECUDA; }
    return HTF_OK;
}

int htf_comm_allreduce_i64(htf_ctx *ctx, int64_t *d_values, int count, void *stream)
{
    return comm_allreduce(ctx, d_values, count, 0, stream);
}

int htf_comm_allreduce_f64(htf_ctx *ctx, double *d_values, int count, void *stream)
{
    return comm_allreduce(ctx, d_values, count, 1, stream);
}

int htf_comm_allreduce_f32(htf_ctx *ctx, float *d_values, int count, void *stream)
{
    return comm_allreduce(ctx, d_values, count, 2, stream);
}

int htf_comm_status(htf_ctx *ctx, int32_t *h_status, void *stream)
{
    if (!ctx || !h_status) return HTF_EINVAL;
    *h_status = 0;
    HtfComm *c = ctx->comm;
    if (!c) return HTF_OK;
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    cudaSetDevice(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemcpyAsync(h_status, &reinterpret_cast<CommHeader *>(c->window)->status, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
    if (e != cudaSuccess) { set_comm_err(ctx, "htf_comm_status", e); return HTF_ECUDA; }
    return HTF_OK;
}

int htf_comm_destroy(htf_ctx *ctx)
{
    if (!ctx) return HTF_EINVAL;
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    cudaSetDevice(ctx->device);
    htf_comm_free(ctx);
    (void)cudaGetLastError();
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
    return HTF_OK;
}

}  // extern "C"
