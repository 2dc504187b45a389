// Minimal on-device integrators: the step either side of the nlist -> forces path, so that a trajectory
// (BASELINE config 1: 1000 steps) keeps positions resident on the GPU.  They stand in for what HOOMD-blue does
// around the reference's compute (hoomd.md.integrate.nve / langevin with mode_standard(dt), as used by
// /root/reference htf/test-py/benchmark.py:39-48 and the test suite); unit mass, like the reference's lattices.
//
//   first half  : v += dt/2 * a;  x += dt * v;  wrap x into [lo, hi)
//   second half : v += dt/2 * a
//   a = F for NVE;  a = F - gamma v + sqrt(2 gamma kT / (dt/2)) * N(0,1) for Langevin: each half kick lasts dt/2 and
//   carries its own, independent, random impulse (fluctuation-dissipation for a step of dt/2), so <v^2> -> kT.
// One thread per particle; positions are float4 (w = type, untouched), velocities float[N][3], forces float4.
#include "common.cuh"

#include <curand_kernel.h>

namespace {

struct IntegrateParams {
    float4 *pos;
    float *vel;
    const float4 *force;
    int n;
    float dt, gamma, noise;     // noise = sqrt(4 gamma kT / dt), 0 for NVE
    int half;
    float lo[3], L[3];
    int flat;                   // 2-D system: z stays 0
    unsigned long long seed, step;
};

__device__ __forceinline__ float wrap_into(float x, float lo, float L)
{
    x -= floorf((x - lo) / L) * L;
    return x >= lo + L ? x - L : x;                 // fp32 rounding can land exactly on hi
}

template <bool LANGEVIN>
__global__ void __launch_bounds__(256) integrate_first_kernel(const IntegrateParams p)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const float4 f = p.force[i];
    float vx = p.vel[3 * i], vy = p.vel[3 * i + 1], vz = p.vel[3 * i + 2];
    // friction acts whenever gamma > 0 (gamma = 0 makes it a no-op), the random impulse only when kT > 0:
    // a damped kT = 0 run gets the same friction in both half kicks
    float ax = f.x - p.gamma * vx, ay = f.y - p.gamma * vy, az = f.z - p.gamma * vz;
    if (LANGEVIN) {
        curandStatePhilox4_32_10_t st;
        curand_init(p.seed, (unsigned long long)i, p.step * 8ull, &st);
        const float4 g = curand_normal4(&st);
        ax += p.noise * g.x;
        ay += p.noise * g.y;
        az += p.noise * (p.flat ? 0.f : g.z);
    }
    const float h = 0.5f * p.dt;
    vx += h * ax; vy += h * ay; vz += h * az;
    float4 x = p.pos[i];
    x.x = wrap_into(x.x + p.dt * vx, p.lo[0], p.L[0]);
    x.y = wrap_into(x.y + p.dt * vy, p.lo[1], p.L[1]);
    x.z = p.flat ? 0.f : wrap_into(x.z + p.dt * vz, p.lo[2], p.L[2]);
    if (p.flat) vz = 0.f;
    p.pos[i] = x;
    p.vel[3 * i] = vx; p.vel[3 * i + 1] = vy; p.vel[3 * i + 2] = vz;
}

template <bool LANGEVIN>
__global__ void __launch_bounds__(256) integrate_second_kernel(const IntegrateParams p)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const float4 f = p.force[i];
    const float h = 0.5f * p.dt;
    float vx = p.vel[3 * i], vy = p.vel[3 * i + 1], vz = p.vel[3 * i + 2];
    float ax = f.x - p.gamma * vx, ay = f.y - p.gamma * vy, az = f.z - p.gamma * vz;
    if (LANGEVIN) {
        curandStatePhilox4_32_10_t st;
        curand_init(p.seed, (unsigned long long)i, p.step * 8ull + 4ull, &st);
        const float4 g = curand_normal4(&st);
        ax += p.noise * g.x; ay += p.noise * g.y; az += p.noise * g.z;
    }
    vx += h * ax; vy += h * ay; vz += h * az;
    if (p.flat) vz = 0.f;
    p.vel[3 * i] = vx; p.vel[3 * i + 1] = vy; p.vel[3 * i + 2] = vz;
}

// HOOMD keeps the particle type as the BIT PATTERN of an int in the w component of its Scalar4 positions; the path
// works with the type as a float VALUE.  Same conversion as the reference's htf_gpu_unstuff4 kernel
// (/root/reference htf/TFArrayComm.cu:9-28), as one coalesced float4 pass (in place when in == out).
__global__ void __launch_bounds__(256) unstuff4_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = in[i];
    v.w = (float)__float_as_int(v.w);
    out[i] = v;
}

}  // namespace

cudaError_t htf_launch_unstuff4(htf_ctx *ctx, const float4 *in, float4 *out, int64_t n, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    unstuff4_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, (int)n);
    ctx->launches += 1;
    return cudaGetLastError();
}

cudaError_t htf_launch_integrate(htf_ctx *ctx, int half, float4 *pos, float *vel, const float4 *force, int64_t n, float dt,
                                 float gamma, float kT, int flat, unsigned long long seed, unsigned long long step,
                                 cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    IntegrateParams p;
    p.pos = pos; p.vel = vel; p.force = force; p.n = (int)n; p.dt = dt; p.gamma = gamma;
    p.noise = (gamma > 0.f && kT > 0.f) ? sqrtf(4.0f * gamma * kT / dt) : 0.f;
    p.half = half;
    for (int a = 0; a < 3; a++) { p.lo[a] = ctx->grid.lo[a]; p.L[a] = ctx->grid.L[a]; }
    p.flat = flat; p.seed = seed; p.step = step;
    const int nb = (int)((n + 255) / 256);
    if (half == 0) {
        if (p.noise > 0.f) integrate_first_kernel<true><<<nb, 256, 0, st>>>(p);
        else integrate_first_kernel<false><<<nb, 256, 0, st>>>(p);
    } else {
        if (p.noise > 0.f) integrate_second_kernel<true><<<nb, 256, 0, st>>>(p);
        else integrate_second_kernel<false><<<nb, 256, 0, st>>>(p);
    }
    ctx->launches += 1;
    return cudaGetLastError();
}
