// nbx_energy.cu -- potential-energy reductions of the resident state (diagnostics for drift checks).
//
//   lennard_jones_potential            src/nbody_simulation_result.jl:293-319  (r2 clamped to R2 outside the cutoff)
//   electrostatic_potential            src/nbody_simulation_result.jl:321-351  (q_j / R outside the cutoff)
//   harmonic_bonds_potential           src/nbody_simulation_result.jl:353-372  (each bond from both ends, / 4)
//   valence_angle_harmonic_potential   src/nbody_simulation_result.jl:374-397
// The reference defines no gravitational or magnetostatic potential energy (:239-264); neither do we.
// Pair sums run over i < j with the reference's exact distance predicate; one thread per i, block
// partials summed in a fixed order (deterministic).
#include "nbx_internal.cuh"

namespace nbx {

constexpr int kEThreads = 128;

__device__ __forceinline__ double eblock_sum(double v)
{
    __shared__ double wsum[kEThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) wsum[warp] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kEThreads / 32; ++w) s += wsum[w];
    return s;
}

struct EBox { int kind; double b0, b1, b2, b3, b4, b5; };

__device__ __forceinline__ double pair_r2(const EBox &bx, double xi, double yi, double zi, double xj, double yj, double zj)
{
    double x = __dsub_rn(xi, xj), y = __dsub_rn(yi, yj), z = __dsub_rn(zi, zj);
    if (bx.kind == NBX_BC_CUBIC) {
        x = wrap_cubic(x, bx.b1, bx.b0); y = wrap_cubic(y, bx.b1, bx.b0); z = wrap_cubic(z, bx.b1, bx.b0);
    } else if (bx.kind == NBX_BC_PERIODIC) {
        x = wrap_range(x, bx.b0, bx.b1); y = wrap_range(y, bx.b2, bx.b3); z = wrap_range(z, bx.b4, bx.b5);
    }
    return r2_unfused(x, y, z);
}

// stride 1: all columns; stride 3: oxygen columns of water
__global__ void lj_energy_kernel(const double *__restrict__ px, int64_t ld, int n, int stride, EBox bx, double sigma2,
                                 double R2, double *__restrict__ partial)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = t * stride;
    double e = 0.0;
    if (i < n) {
        const double xi = px[i], yi = px[ld + i], zi = px[2 * ld + i];
        for (int j = i + stride; j < n; j += stride) {
            double r2 = pair_r2(bx, xi, yi, zi, px[j], px[ld + j], px[2 * ld + j]);
            if (!(r2 < R2)) r2 = R2;
            const double q = sigma2 / r2;
            const double s6 = q * q * q;
            e += s6 * s6 - s6;
        }
    }
    const double b = eblock_sum(e);
    if (threadIdx.x == 0) partial[blockIdx.x] = b;
}

__global__ void coulomb_energy_kernel(const double *__restrict__ px, int64_t ld, const double *__restrict__ q, int n,
                                      int water, EBox bx, double R, double R2, double *__restrict__ partial)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (i < n) {
        const double xi = px[i], yi = px[ld + i], zi = px[2 * ld + i];
        const int first = water ? 3 * (i / 3) + 3 : i + 1;
        double ei = 0.0;
        for (int j = first; j < n; ++j) {
            const double r2 = pair_r2(bx, xi, yi, zi, px[j], px[ld + j], px[2 * ld + j]);
            ei += (r2 < R2) ? q[j] / sqrt(r2) : q[j] / R;
        }
        e = ei * q[i];
    }
    const double b = eblock_sum(e);
    if (threadIdx.x == 0) partial[blockIdx.x] = b;
}

__global__ void bonded_energy_kernel(const double *__restrict__ px, int64_t ld, int nmol, double rOH, double kb,
                                     double aHOH0, double ka, double *__restrict__ partial_b,
                                     double *__restrict__ partial_a)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    double eb = 0.0, ea = 0.0;
    if (m < nmol) {
        const int64_t o = 3 * (int64_t)m;
        const double ax = px[o + 1] - px[o], ay = px[ld + o + 1] - px[ld + o], az = px[2 * ld + o + 1] - px[2 * ld + o];
        const double cx = px[o + 2] - px[o], cy = px[ld + o + 2] - px[ld + o], cz = px[2 * ld + o + 2] - px[2 * ld + o];
        const double na = sqrt(ax * ax + ay * ay + az * az), nc = sqrt(cx * cx + cy * cy + cz * cz);
        const double da = na - rOH, dc = nc - rOH;
        eb = 2.0 * (da * da * kb) + 2.0 * (dc * dc * kb); // each bond is visited from both of its ends
        const double ang = acos((ax * cx + ay * cy + az * cz) / (na * nc));
        const double d = ang - aHOH0;
        ea = ka * (d * d);
    }
    const double sb = eblock_sum(eb);
    __syncthreads();
    const double sa = eblock_sum(ea);
    if (threadIdx.x == 0) { partial_b[blockIdx.x] = sb; partial_a[blockIdx.x] = sa; }
}

__global__ void esum_kernel(const double *__restrict__ partial, int nb, double scale, double *__restrict__ out)
{
    double v = 0.0;
    for (int i = threadIdx.x; i < nb; i += kEThreads) v += partial[i];
    const double s = eblock_sum(v);
    if (threadIdx.x == 0) out[0] += scale * s;
}

int reduce_potential(nbx_ctx *c, double *epot)
{
    const int n = (int)c->n;
    const int nb_max = (n + kEThreads - 1) / kEThreads;
    double *partial = nullptr, *out = c->d_scal + 11;
    NBX_TRY(dev_alloc(c, &partial, (size_t)2 * nb_max));
    cudaMemsetAsync(out, 0, sizeof(double), c->stream);
    EBox bx{c->bc_kind, c->bc[0], c->bc[1], c->bc[2], c->bc[3], c->bc[4], c->bc[5]};
    if (c->bc_kind == NBX_BC_CUBIC) bx.b1 = 0.5 * c->bc[0];
    if (c->has_lj) {
        const int stride = c->water ? 3 : 1;
        const int nt = (n + stride - 1) / stride;
        const int nb = (nt + kEThreads - 1) / kEThreads;
        lj_energy_kernel<<<nb, kEThreads, 0, c->stream>>>(c->pos, c->npad, n, stride, bx, c->lj_sigma2, c->lj_R2, partial);
        esum_kernel<<<1, kEThreads, 0, c->stream>>>(partial, nb, 4.0 * c->lj_eps, out);
    }
    if (c->has_coul) {
        coulomb_energy_kernel<<<nb_max, kEThreads, 0, c->stream>>>(c->pos, c->npad, c->charge, n, c->water, bx, c->el_R,
                                                                  c->el_R2, partial);
        esum_kernel<<<1, kEThreads, 0, c->stream>>>(partial, nb_max, c->el_k, out);
    }
    if (c->has_spcfw) {
        const int nmol = n / 3;
        const int nb = (nmol + kEThreads - 1) / kEThreads;
        bonded_energy_kernel<<<nb, kEThreads, 0, c->stream>>>(c->pos, c->npad, nmol, c->rOH, c->k_bond, c->aHOH, c->k_angle,
                                                             partial, partial + nb_max);
        esum_kernel<<<1, kEThreads, 0, c->stream>>>(partial, nb, 0.25, out);
        esum_kernel<<<1, kEThreads, 0, c->stream>>>(partial + nb_max, nb, 0.5, out);
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(epot, out, sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(partial);
    if (e != cudaSuccess) return cuda_fail(c, e, "reduce_potential");
    return NBX_OK;
}

} // namespace nbx
