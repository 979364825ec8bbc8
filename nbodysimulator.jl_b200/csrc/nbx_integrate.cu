// nbx_integrate.cu -- O(N), HBM-bound kernels: AoS<->SoA at the boundary, the fused velocity-Verlet /
// thermostat updates that keep the state on the device across steps, temperature reductions, the
// counter-based RNG of the stochastic thermostats, and the roofline microbenchmarks.
//
// Reference pieces restated here:
//   md_temperature / berendsen_acceleration!   src/thermostats.jl:76-91
//   nosehoover_acceleration!                   src/thermostats.jl:121-128
//   Langevin drift + noise                     src/nbody_to_ode.jl:575-595
//   Andersen velocity resampling               src/nbody_simulation_result.jl:504-540
//   VelocityVerlet / EM update formulas        upstream OrdinaryDiffEqSymplecticRK / StochasticDiffEq
#include "nbx_internal.cuh"

namespace nbx {

constexpr int kRedThreads = 256;
constexpr int kRedBlocksMax = 592;

static int red_blocks(const nbx_ctx *c, int64_t n)
{
    int64_t b = (n + kRedThreads - 1) / kRedThreads;
    const int64_t cap = (int64_t)c->sm_count * 4 < kRedBlocksMax ? (int64_t)c->sm_count * 4 : kRedBlocksMax;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// ------------------------------------------------------------------------------------------------
// boundary transposes: Julia Matrix{Float64} 3 x ncols (AoS) <-> SoA rows of stride ld
// ------------------------------------------------------------------------------------------------
__global__ void aos_to_soa_kernel(const double *__restrict__ aos, double *__restrict__ soa, int64_t n, int64_t ld)
{
    // one warp-wide pass over 32 columns = 96 consecutive doubles staged through shared memory
    __shared__ double tile[8][96];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t col0 = ((int64_t)blockIdx.x * 8 + warp) * 32;
    if (col0 >= n) return;
    const int64_t ncol = n - col0 < 32 ? n - col0 : 32;
    for (int k = lane; k < 3 * ncol; k += 32) tile[warp][k] = aos[3 * col0 + k];
    __syncwarp();
    if (lane < ncol) {
        soa[col0 + lane] = tile[warp][3 * lane];
        soa[ld + col0 + lane] = tile[warp][3 * lane + 1];
        soa[2 * ld + col0 + lane] = tile[warp][3 * lane + 2];
    }
}

__global__ void soa_to_aos_kernel(const double *__restrict__ soa, double *__restrict__ aos, int64_t ld,
                                  int64_t ncols_total, int64_t lo, int64_t hi)
{
    __shared__ double tile[8][96];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t col0 = ((int64_t)blockIdx.x * 8 + warp) * 32;
    if (col0 >= ncols_total) return;
    const int64_t ncol = ncols_total - col0 < 32 ? ncols_total - col0 : 32;
    if (lane < ncol) {
        const int64_t i = col0 + lane;
        const bool live = i >= lo && i < hi;
        tile[warp][3 * lane] = live ? soa[i] : 0.0;
        tile[warp][3 * lane + 1] = live ? soa[ld + i] : 0.0;
        tile[warp][3 * lane + 2] = live ? soa[2 * ld + i] : 0.0;
    }
    __syncwarp();
    for (int k = lane; k < 3 * ncol; k += 32) aos[3 * col0 + k] = tile[warp][k];
}

int launch_aos_to_soa(nbx_ctx *c, const double *aos, double *soa, int64_t ncols_used)
{
    if (ncols_used <= 0) return NBX_OK;
    const int64_t blocks = (ncols_used + 255) / 256;
    timer_begin(c, NBX_T_TRANSPOSE);
    aos_to_soa_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(aos, soa, ncols_used, c->npad);
    timer_end(c, NBX_T_TRANSPOSE);
    NBX_CUDA(c, cudaGetLastError());
    return NBX_OK;
}

int launch_soa_to_aos(nbx_ctx *c, const double *soa, double *aos, int64_t n, int64_t ncols_total, int64_t lo,
                      int64_t hi)
{
    (void)n;
    if (ncols_total <= 0) return NBX_OK;
    const int64_t blocks = (ncols_total + 255) / 256;
    timer_begin(c, NBX_T_TRANSPOSE);
    soa_to_aos_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(soa, aos, c->npad, ncols_total, lo, hi);
    timer_end(c, NBX_T_TRANSPOSE);
    NBX_CUDA(c, cudaGetLastError());
    return NBX_OK;
}

__global__ void fill_kernel(double *p, double v, int64_t count)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = v;
}

int launch_fill(nbx_ctx *c, double *p, double v, int64_t count)
{
    if (count <= 0) return NBX_OK;
    int64_t blocks = (count + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    fill_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(p, v, count);
    NBX_CUDA(c, cudaGetLastError());
    return NBX_OK;
}

// ------------------------------------------------------------------------------------------------
// deterministic block reductions
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v)
{
    __shared__ double wsum[kRedThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) wsum[warp] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kRedThreads / 32; ++w) s += wsum[w];
    return s; // valid in thread 0
}

// out[0] = sum of partial[0..nb) in ascending order (single block)
__global__ void final_sum_kernel(const double *__restrict__ partial, int nb, double *__restrict__ out,
                                 double *__restrict__ out2 = nullptr)
{
    double v = 0.0;
    for (int i = threadIdx.x; i < nb; i += kRedThreads) v += partial[i];
    const double s = block_sum(v);
    if (threadIdx.x == 0) {
        out[0] = s;
        if (out2) out2[0] = s;
    }
}

// sum_i m_i |v_i|^2 over columns [lo,hi)   (md_temperature numerator, src/thermostats.jl:88)
__global__ void mv2_partial_kernel(const double *__restrict__ vel, const double *__restrict__ mass, int64_t ld,
                                   int64_t lo, int64_t hi, double *__restrict__ partial)
{
    double s = 0.0;
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
        const double vx = vel[i], vy = vel[ld + i], vz = vel[2 * ld + i];
        s += mass[i] * (vx * vx + vy * vy + vz * vz);
    }
    const double b = block_sum(s);
    if (threadIdx.x == 0) partial[blockIdx.x] = b;
}

int ensure_red(nbx_ctx *c)
{
    if (c->d_red) return NBX_OK;
    c->red_cap = kRedBlocksMax * 4;
    NBX_TRY(dev_alloc(c, &c->d_red, (size_t)c->red_cap));
    NBX_CUDA(c, cudaMemsetAsync(c->d_red, 0, sizeof(double) * (size_t)c->red_cap, c->stream)); // (the tail holds a block ticket)
    return NBX_OK;
}

int launch_sum_mv2(nbx_ctx *c, const double *vel, int64_t lo, int64_t hi)
{
    NBX_TRY(ensure_red(c));
    const int nb = red_blocks(c, hi - lo);
    timer_begin(c, NBX_T_INTEGRATE);
    mv2_partial_kernel<<<nb, kRedThreads, 0, c->stream>>>(vel, c->mass, c->npad, lo, hi, c->d_red);
    final_sum_kernel<<<1, kRedThreads, 0, c->stream>>>(c->d_red, nb, c->d_scal);
    timer_end(c, NBX_T_INTEGRATE);
    NBX_CUDA(c, cudaGetLastError());
    return NBX_OK;
}

// ------------------------------------------------------------------------------------------------
// RHS thermostats ("simultaneous accelerations", src/nbody_to_ode.jl:484-486)
// scal[0] = sum m v^2 (whole system), scal[1] = zeta, scal[2] = zeta_dot (written here)
// ------------------------------------------------------------------------------------------------
__global__ void berendsen_kernel(double *__restrict__ acc, const double *__restrict__ vel, int64_t ld, int64_t lo,
                                 int64_t hi, const double *__restrict__ scal, double kB, double ndf, double T0,
                                 double gamma)
{
    const double T = scal[0] / (kB * ndf);
    // `if inv(T) == Inf` branch of src/thermostats.jl:78-82
    const double s = (1.0 / T == INFINITY) ? gamma : gamma * (T0 / T - 1.0);
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int d = 0; d < 3; ++d) acc[d * ld + i] += s * vel[d * ld + i];
    }
}

__global__ void nosehoover_kernel(double *__restrict__ acc, const double *__restrict__ vel, int64_t ld, int64_t lo,
                                  int64_t hi, double *__restrict__ scal, double kB, double ndf, double T0,
                                  double inv_tau2)
{
    const double zeta = scal[1];
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int d = 0; d < 3; ++d) acc[d * ld + i] -= zeta * vel[d * ld + i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const double T = scal[0] / (kB * ndf);
        scal[2] = inv_tau2 * (T / T0 - (ndf + 1.0) / ndf); // src/thermostats.jl:126
    }
}

int launch_thermostat_rhs(nbx_ctx *c, double *acc, const double *vel)
{
    if (c->thermo != NBX_THERMO_BERENDSEN && c->thermo != NBX_THERMO_NOSEHOOVER) return NBX_OK;
    const int64_t lo = c->tgt_lo, hi = c->tgt_hi;
    const int nb = red_blocks(c, hi - lo);
    const double ndf = (double)(3 * c->thN - c->thNc);
    timer_begin(c, NBX_T_INTEGRATE);
    if (c->thermo == NBX_THERMO_BERENDSEN) {
        berendsen_kernel<<<nb, kRedThreads, 0, c->stream>>>(acc, vel, c->npad, lo, hi, c->d_scal, c->kB, ndf, c->T0,
                                                           0.5 / c->tparam);
    } else {
        const double it = 1.0 / c->tparam;
        nosehoover_kernel<<<nb, kRedThreads, 0, c->stream>>>(acc, vel, c->npad, lo, hi, c->d_scal, c->kB, ndf, c->T0,
                                                            it * it);
    }
    timer_end(c, NBX_T_INTEGRATE);
    NBX_CUDA(c, cudaGetLastError());
    return NBX_OK;
}

// ------------------------------------------------------------------------------------------------
// velocity Verlet halves (own shard only)
// ------------------------------------------------------------------------------------------------
__global__ void vv_pos_kernel(double *__restrict__ pos, const double *__restrict__ vel, const double *__restrict__ acc,
                              int64_t ld, int64_t lo, int64_t hi, double dt, double hdt2, double *__restrict__ scal,
                              int nose, const int *__restrict__ dyn)
{
    if (dyn) hi = min(hi, lo + (int64_t)dyn[0]);
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int64_t k = d * ld + i;
            pos[k] = fma(hdt2, acc[k], fma(dt, vel[k], pos[k]));
        }
    }
    // the extra Nose-Hoover column: "position" zeta, "velocity" zeta_dot, zero acceleration
    if (nose && blockIdx.x == 0 && threadIdx.x == 0) scal[1] = fma(dt, scal[2], scal[1]);
}

// v+ = v + dt/2 (a_old + a_new), and the block partials of sum m v+^2 for the next temperature.
// BEREND: the Berendsen RHS term (src/thermostats.jl:76-83) a_new += gamma (T0/T - 1) v(t) is applied on the
// fly (T from scal[0] = sum m v(t)^2, which only final_sum_kernel overwrites, after this kernel).
template <bool BEREND>
__global__ void vv_vel_kernel(double *__restrict__ vel, const double *__restrict__ a_old, double *__restrict__ a_new,
                              const double *__restrict__ mass, int64_t ld, int64_t lo, int64_t hi, double hdt,
                              double *__restrict__ partial, const double *scal, double kB, double ndf,
                              double T0, double gamma, const int *__restrict__ dyn, int *__restrict__ ticket,
                              double *out, double *out2) // (scal, out, out2 all point into the scalar block)
{
    __shared__ int last_block;
    if (dyn) hi = min(hi, lo + (int64_t)dyn[0]);
    double sc = 0.0;
    if (BEREND) {
        const double T = scal[0] / (kB * ndf);
        sc = (1.0 / T == INFINITY) ? gamma : gamma * (T0 / T - 1.0);
    }
    double s = 0.0;
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
        double v2 = 0.0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int64_t k = d * ld + i;
            const double v0 = vel[k];
            double an = a_new[k];
            if (BEREND) {
                an += sc * v0; // same expression as berendsen_kernel
                a_new[k] = an;
            }
            const double v = fma(hdt, a_old[k] + an, v0);
            vel[k] = v;
            v2 = fma(v, v, v2);
        }
        s = fma(mass[i], v2, s);
    }
    const double b = block_sum(s);
    // the last block to finish adds the partials exactly as final_sum_kernel does (one launch less per step)
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = b;
        __threadfence();
        last_block = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (!last_block) return;
    __threadfence();
    double t = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += kRedThreads) t += __ldcg(partial + i);
    const double total = block_sum(t);
    if (threadIdx.x == 0) {
        out[0] = total;
        if (out2) out2[0] = total;
        *ticket = 0;
    }
}

int launch_vv_pos(nbx_ctx *c, double dt)
{
    const int64_t lo = c->tgt_lo, hi = c->tgt_hi;
    const int nb = red_blocks(c, hi - lo);
    timer_begin(c, NBX_T_INTEGRATE);
    vv_pos_kernel<<<nb, kRedThreads, 0, c->stream>>>(c->pos, c->vel, c->acc, c->npad, lo, hi, dt, 0.5 * dt * dt,
                                                    c->d_scal, c->thermo == NBX_THERMO_NOSEHOOVER ? 1 : 0, c->dyn);
    timer_end(c, NBX_T_INTEGRATE);
    NBX_CUDA(c, cudaGetLastError());
    return NBX_OK;
}

// acc_old = a(t), acc = a(t+dt) on entry; with_thermostat: the RHS thermostat term is still to be added to acc
int launch_vv_vel(nbx_ctx *c, double dt, bool with_thermostat)
{
    NBX_TRY(ensure_red(c));
    const int64_t lo = c->tgt_lo, hi = c->tgt_hi;
    const int nb = red_blocks(c, hi - lo);
    const double ndf = (double)(3 * c->thN - c->thNc);
    bool fused = false;
    if (with_thermostat) {
        if (c->thermo == NBX_THERMO_BERENDSEN) fused = true;
        else NBX_TRY(launch_thermostat_rhs(c, c->acc, c->vel));
    }
    // (T_slot != 0: the distributed loops keep the all-reduced sum in its own slot; the local one goes to both)
    int *ticket = reinterpret_cast<int *>(c->d_red + c->red_cap - 1);
    double *out2 = c->T_slot ? c->d_scal + c->T_slot : nullptr;
    timer_begin(c, NBX_T_INTEGRATE);
    if (fused)
        vv_vel_kernel<true><<<nb, kRedThreads, 0, c->stream>>>(c->vel, c->acc_old, c->acc, c->mass, c->npad, lo, hi,
                                                              0.5 * dt, c->d_red, c->d_scal + c->T_slot, c->kB, ndf, c->T0,
                                                              0.5 / c->tparam, c->dyn, ticket, c->d_scal, out2);
    else
        vv_vel_kernel<false><<<nb, kRedThreads, 0, c->stream>>>(c->vel, c->acc_old, c->acc, c->mass, c->npad, lo, hi,
                                                               0.5 * dt, c->d_red, c->d_scal, 0.0, 1.0, 0.0, 0.0, c->dyn, ticket, c->d_scal, out2);
    timer_end(c, NBX_T_INTEGRATE);
    NBX_CUDA(c, cudaGetLastError());
    return NBX_OK;
}

// ------------------------------------------------------------------------------------------------
// counter-based RNG: Philox4x32-10, keyed by the context seed, counter = (atom, step, stream, 0)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32(uint32_t (&ctr)[4], uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr[0]), lo0 = 0xD2511F53u * ctr[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr[2]), lo1 = 0xCD9E8D57u * ctr[2];
        const uint32_t n0 = hi1 ^ ctr[1] ^ k0, n1 = lo1, n2 = hi0 ^ ctr[3] ^ k1, n3 = lo0;
        ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ double u01(uint32_t a, uint32_t b)
{
    // 53-bit uniform in (0,1)
    const uint64_t m = ((uint64_t)a << 21) ^ (uint64_t)(b >> 11);
    return ((double)m + 0.5) * (1.0 / 9007199254740992.0);
}
// three independent standard normals and one uniform for (atom i, step, stream)
__device__ __forceinline__ void normals3(uint64_t seed, uint64_t step, uint32_t stream, uint64_t i, double (&g)[3],
                                         double &u)
{
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const uint32_t hi = (uint32_t)(step >> 32) ^ (stream << 24);
    uint32_t c0[4] = {(uint32_t)i, (uint32_t)(i >> 32), (uint32_t)step, hi};
    uint32_t c1[4] = {(uint32_t)i, (uint32_t)(i >> 32), (uint32_t)step, hi ^ 0x00800000u};
    philox4x32(c0, k0, k1);
    philox4x32(c1, k0, k1);
    double sn, cs;
    const double ra = sqrt(-2.0 * log(u01(c0[0], c0[1])));
    sincospi(2.0 * u01(c0[2], c0[3]), &sn, &cs);
    g[0] = ra * cs;
    g[1] = ra * sn;
    uint32_t c2[4] = {(uint32_t)i, (uint32_t)(i >> 32), (uint32_t)step, hi ^ 0x00400000u};
    philox4x32(c2, k0, k1);
    const double rb = sqrt(-2.0 * log(u01(c1[0], c1[1])));
    g[2] = rb * cospi(2.0 * u01(c1[2], c1[3]));
    u = u01(c2[0], c2[1]);
}

// Euler-Maruyama on the Langevin SDE (src/nbody_to_ode.jl:575-595):
//   x+ = x + dt v ;  v+ = v + dt (a - gamma v) + sigma sqrt(dt) xi
// sigma is the reference's scalar sqrt(2 gamma kb T / m_1) for every atom (:592).
// Water (the SDEProblem of WaterSPCFw, :600-680), as written there: the oxygen column gets -(gamma v) / mO on top of the
// -gamma v that every column gets at the end (the same term subtracted from the hydrogen columns inside the oxygen loop is
// zeroed again by the hydrogen loop, :627-636), and the noise amplitudes are sqrt(2 gamma kb T) / mO and / mH (:668-669:
// divided by the mass, not by its root).
__global__ void em_kernel(double *__restrict__ pos, double *__restrict__ vel, const double *__restrict__ acc,
                          const double *__restrict__ mass, int64_t ld, int64_t lo, int64_t hi, double dt, double gamma,
                          double sig_sqdt, int water, uint64_t seed, uint64_t step, double *__restrict__ partial)
{
    double s = 0.0;
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
        double g[3], u;
        normals3(seed, step, 1u, (uint64_t)i, g, u);
        const double m = mass[i];
        const bool oxygen = water && (i % 3 == 0);
        const double amp = water ? sig_sqdt / m : sig_sqdt; // (water: sig_sqdt = sqrt(2 gamma kb T) sqrt(dt))
        double v2 = 0.0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int64_t k = d * ld + i;
            const double v = vel[k];
            pos[k] = fma(dt, v, pos[k]);
            double drift = acc[k];
            if (oxygen) drift -= gamma * v / m;
            drift -= gamma * v;
            const double vn = v + dt * drift + amp * g[d];
            vel[k] = vn;
            v2 = fma(vn, vn, v2);
        }
        s = fma(m, v2, s);
    }
    const double b = block_sum(s);
    if (threadIdx.x == 0) partial[blockIdx.x] = b;
}

int launch_em_step(nbx_ctx *c, double dt)
{
    NBX_TRY(ensure_red(c));
    const int64_t lo = c->tgt_lo, hi = c->tgt_hi;
    const int nb = red_blocks(c, hi - lo);
    const double sigma = c->water ? sqrt(2.0 * c->tparam * c->kB * c->T0) : sqrt(2.0 * c->tparam * c->kB * c->T0 / c->h_m1);
    timer_begin(c, NBX_T_INTEGRATE);
    em_kernel<<<nb, kRedThreads, 0, c->stream>>>(c->pos, c->vel, c->acc, c->mass, c->npad, lo, hi, dt, c->tparam,
                                                sigma * sqrt(dt), c->water ? 1 : 0, c->seed, c->rng_step, c->d_red);
    final_sum_kernel<<<1, kRedThreads, 0, c->stream>>>(c->d_red, nb, c->d_scal);
    timer_end(c, NBX_T_INTEGRATE);
    NBX_CUDA(c, cudaGetLastError());
    c->rng_step++;
    return NBX_OK;
}

// Andersen: after every step each body collides with probability nu*dt and gets a fresh
// Maxwell-Boltzmann velocity sqrt(kb T / m_1) randn(3); water: O and both H of molecule i with
// sqrt(kb T / mO), sqrt(kb T / mH) (src/nbody_simulation_result.jl:522-540).
__global__ void andersen_kernel(double *__restrict__ vel, const double *__restrict__ mass, int64_t ld, int64_t lo,
                                int64_t hi, int water, double prob, double kT, double m1, uint64_t seed, uint64_t step)
{
    const int64_t nunits = water ? (hi - lo) / 3 : (hi - lo);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nunits; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t first = water ? lo + 3 * t : lo + t;
        double g[3], u;
        normals3(seed, step, 2u, (uint64_t)first, g, u);
        if (u < prob) {
            const int na = water ? 3 : 1;
            for (int a = 0; a < na; ++a) {
                const int64_t i = first + a;
                if (a > 0) normals3(seed, step, 2u + a, (uint64_t)first, g, u);
                const double sd = sqrt(kT / (water ? mass[i] : m1));
#pragma unroll
                for (int d = 0; d < 3; ++d) vel[d * ld + i] = sd * g[d];
            }
        }
    }
}

int launch_andersen(nbx_ctx *c, double dt)
{
    const int64_t lo = c->tgt_lo, hi = c->tgt_hi;
    const int nb = red_blocks(c, hi - lo);
    timer_begin(c, NBX_T_INTEGRATE);
    andersen_kernel<<<nb, kRedThreads, 0, c->stream>>>(c->vel, c->mass, c->npad, lo, hi, c->water, c->tparam * dt,
                                                      c->kB * c->T0, c->h_m1, c->seed, c->rng_step);
    timer_end(c, NBX_T_INTEGRATE);
    NBX_CUDA(c, cudaGetLastError());
    c->rng_step++;
    return launch_sum_mv2(c, c->vel, lo, hi);
}

// ------------------------------------------------------------------------------------------------
// input validation: NaN/Inf coordinates (the reference's wrap loops would never terminate)
// ------------------------------------------------------------------------------------------------
__global__ void finite_kernel(const double *__restrict__ soa, int64_t ld, int64_t n, int *__restrict__ flag)
{
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        bad = bad || !isfinite(soa[i]) || !isfinite(soa[ld + i]) || !isfinite(soa[2 * ld + i]);
    if (bad) *flag = 1;
}

// flag lives in the int view of d_scal[15]; the caller reads it back with the results
int check_finite(nbx_ctx *c, const double *soa, int64_t n)
{
    int *flag = reinterpret_cast<int *>(c->d_scal + 15);
    NBX_CUDA(c, cudaMemsetAsync(flag, 0, sizeof(double), c->stream));
    const int nb = red_blocks(c, n);
    finite_kernel<<<nb, kRedThreads, 0, c->stream>>>(soa, c->npad, n, flag);
    NBX_CUDA(c, cudaGetLastError());
    return NBX_OK;
}

// ------------------------------------------------------------------------------------------------
// kinetic energy / temperature of the resident state
// kinetic_energy: sum |v|^2 m/2 (src/nbody_simulation_result.jl:209-212)
// ------------------------------------------------------------------------------------------------
__global__ void ekin_partial_kernel(const double *__restrict__ vel, const double *__restrict__ mass, int64_t ld,
                                    int64_t lo, int64_t hi, double *__restrict__ partial)
{
    double s = 0.0;
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
        const double vx = vel[i], vy = vel[ld + i], vz = vel[2 * ld + i];
        s += (vx * vx + vy * vy + vz * vz) * (mass[i] / 2);
    }
    const double b = block_sum(s);
    if (threadIdx.x == 0) partial[blockIdx.x] = b;
}

int reduce_kinetic(nbx_ctx *c, double *ekin, double *temp)
{
    NBX_TRY(ensure_red(c));
    const int64_t lo = c->tgt_lo, hi = c->tgt_hi;
    const int nb = red_blocks(c, hi - lo);
    double h[2] = {0, 0};
    ekin_partial_kernel<<<nb, kRedThreads, 0, c->stream>>>(c->vel, c->mass, c->npad, lo, hi, c->d_red);
    final_sum_kernel<<<1, kRedThreads, 0, c->stream>>>(c->d_red, nb, c->d_scal + 8);
    mv2_partial_kernel<<<nb, kRedThreads, 0, c->stream>>>(c->vel, c->mass, c->npad, lo, hi, c->d_red + kRedBlocksMax);
    final_sum_kernel<<<1, kRedThreads, 0, c->stream>>>(c->d_red + kRedBlocksMax, nb, c->d_scal + 9);
    NBX_CUDA(c, cudaGetLastError());
    NBX_CUDA(c, cudaMemcpyAsync(h, c->d_scal + 8, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    NBX_CUDA(c, cudaStreamSynchronize(c->stream));
    if (ekin) *ekin = h[0];
    if (temp) {
        const int64_t N = c->thN > 0 ? c->thN : c->n;
        const double kB = c->kB != 0.0 ? c->kB : 1.0;
        *temp = h[1] / (kB * (double)(3 * N - c->thNc));
    }
    return NBX_OK;
}

// ------------------------------------------------------------------------------------------------
// roofline microbenchmarks
// ------------------------------------------------------------------------------------------------
template <int CHAINS>
__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b)
{
    double x[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) x[k] = (double)(threadIdx.x + k) * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) x[k] = fma(x[k], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s += x[k];
    if (s == 123.456) out[0] = s; // never true; keeps the chains alive
}

int measure_fp64_peak(nbx_ctx *c, double *tflops, double *mhz)
{
    constexpr int CH = 8;
    const int blocks = c->sm_count * 8, threads = 256, iters = 1 << 15;
    cudaEvent_t e0, e1;
    NBX_CUDA(c, cudaEventCreate(&e0));
    NBX_CUDA(c, cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        NBX_CUDA(c, cudaEventRecord(e0, c->stream));
        dfma_kernel<CH><<<blocks, threads, 0, c->stream>>>(c->d_scal + 10, iters, 0.999999, 1e-9);
        NBX_CUDA(c, cudaEventRecord(e1, c->stream));
        NBX_CUDA(c, cudaEventSynchronize(e1));
        float ms = 0.f;
        NBX_CUDA(c, cudaEventElapsedTime(&ms, e0, e1));
        const double fma_count = (double)blocks * threads * (double)iters * CH;
        const double tf = 2.0 * fma_count / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (tflops) *tflops = best;
    // 64 DFMA lanes per SM per clock
    if (mhz) *mhz = best * 1e12 / 2.0 / (64.0 * c->sm_count) / 1e6;
    return NBX_OK;
}

__global__ void copy_kernel(const double4 *__restrict__ src, double4 *__restrict__ dst, int64_t n4)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

int measure_hbm_peak(nbx_ctx *c, double *gbs)
{
    const size_t bytes = (size_t)1 << 30; // 1 GiB each way: well beyond the 126 MB L2
    double4 *a = nullptr, *b = nullptr;
    NBX_CUDA(c, cudaMalloc((void **)&a, bytes));
    cudaError_t e = cudaMalloc((void **)&b, bytes);
    if (e != cudaSuccess) { cudaFree(a); return cuda_fail(c, e, "cudaMalloc(copy benchmark)"); }
    cudaMemsetAsync(a, 0, bytes, c->stream);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    const int64_t n4 = (int64_t)(bytes / sizeof(double4));
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0, c->stream);
        copy_kernel<<<c->sm_count * 16, 512, 0, c->stream>>>(a, b, n4);
        cudaEventRecord(e1, c->stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double g = 2.0 * (double)bytes / (ms * 1e-3) / 1e9;
        if (rep > 0 && g > best) best = g;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(a);
    cudaFree(b);
    NBX_CUDA(c, cudaGetLastError());
    if (gbs) *gbs = best;
    return NBX_OK;
}

// CUDA loads kernels lazily, at their first launch, and loading may synchronise the context: a kernel that is first
// launched while another member's kernel spins on a flag would deadlock the pair (CUDA programming guide, "Lazy
// Loading": concurrent execution).  Every kernel of the distributed loops is therefore loaded when a context is created.
void preload_integrate()
{
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, aos_to_soa_kernel);
    cudaFuncGetAttributes(&a, soa_to_aos_kernel);
    cudaFuncGetAttributes(&a, fill_kernel);
    cudaFuncGetAttributes(&a, final_sum_kernel);
    cudaFuncGetAttributes(&a, mv2_partial_kernel);
    cudaFuncGetAttributes(&a, berendsen_kernel);
    cudaFuncGetAttributes(&a, vv_pos_kernel);
    cudaFuncGetAttributes(&a, vv_vel_kernel<true>);
    cudaFuncGetAttributes(&a, vv_vel_kernel<false>);
    cudaFuncGetAttributes(&a, em_kernel);
    cudaFuncGetAttributes(&a, andersen_kernel);
    cudaFuncGetAttributes(&a, finite_kernel);
    cudaFuncGetAttributes(&a, ekin_partial_kernel);
    cudaGetLastError();
}

} // namespace nbx
