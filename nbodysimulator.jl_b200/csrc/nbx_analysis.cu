// nbx_analysis.cu -- the first caller-side hotspots after the step loop: rdf and msd of saved frames.
//
// rdf (src/nbody_simulation_result.jl:664-709) is O(frames x N^2) on the host and dominates as soon as the step loop is
// fast; msd (:730-783) is O(frames x N).  Both are restated per FRAME here; the caller loops over the frames and
// normalises exactly as the reference does (:695-707).
//
// rdf_kernel: pairs i < j of the Lennard-Jones index set (all columns, or every third = the oxygens of water), the
// reference's distance (rij = ri - rj on the unwrapped coordinates, the wrap loops of src/boundary_conditions.jl:143-160,
// un-fused r2), `r2 < (0.5 L)^2`, bin = ceil(sqrt(r2) / dr) with IEEE sqrt and division, `1 < bin <= maxbin` -> += 2.
// The histogram is integer, so the result is BIT-EXACT whatever the order of the pairs.  Tiles of 256 x 256 pairs,
// the j tile staged in shared memory, a per-block shared histogram flushed with 64-bit atomics.
#include "nbx_internal.cuh"

namespace nbx {

constexpr int kRdfTile = 256;

__global__ void __launch_bounds__(kRdfTile) rdf_kernel(const double *__restrict__ px, int64_t ld, int m, int stride, double L,
                                                       double radius, double lim, double dr, int maxbin,
                                                       unsigned long long *__restrict__ hist)
{
    const int ti = blockIdx.y, tj = blockIdx.x;
    if (tj < ti) return; // upper triangle of the tile pairs
    extern __shared__ int sh_hist[];                  // [maxbin]
    __shared__ double sx[kRdfTile], sy[kRdfTile], sz[kRdfTile];
    const int t = threadIdx.x;
    for (int b = t; b < maxbin; b += kRdfTile) sh_hist[b] = 0;
    const int jj = tj * kRdfTile + t;
    if (jj < m) {
        const int64_t j = (int64_t)jj * stride;
        sx[t] = px[j]; sy[t] = px[ld + j]; sz[t] = px[2 * ld + j];
    }
    __syncthreads();
    const int ii = ti * kRdfTile + t;
    if (ii < m) {
        const int64_t i = (int64_t)ii * stride;
        const double xi = px[i], yi = px[ld + i], zi = px[2 * ld + i];
        const int nj = min(kRdfTile, m - tj * kRdfTile);
        for (int s = (ti == tj ? t + 1 : 0); s < nj; ++s) { // j > i
            double rx = __dsub_rn(xi, sx[s]), ry = __dsub_rn(yi, sy[s]), rz = __dsub_rn(zi, sz[s]);
            rx = wrap_cubic(rx, radius, L);
            ry = wrap_cubic(ry, radius, L);
            rz = wrap_cubic(rz, radius, L);
            const double r2 = r2_unfused(rx, ry, rz);
            if (r2 < lim) {
                const double b = ceil(__ddiv_rn(__dsqrt_rn(r2), dr));
                if (b > 1.0 && b <= (double)maxbin) atomicAdd(&sh_hist[(int)b - 1], 2);
            }
        }
    }
    __syncthreads();
    for (int b = t; b < maxbin; b += kRdfTile)
        if (sh_hist[b]) atomicAdd(&hist[b], (unsigned long long)sh_hist[b]);
}

// block partials of sum |r - r0|^2 (atoms) or of the mass-weighted molecular displacement (water)
__global__ void msd_partial_kernel(const double *__restrict__ p, const double *__restrict__ p0, int64_t ld, int64_t n, int water,
                                   double mO, double mH, double *__restrict__ partial)
{
    __shared__ double wsum[8];
    double s = 0.0;
    const int64_t count = water ? n / 3 : n;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (int64_t)gridDim.x * blockDim.x) {
        if (!water) {
            const double d0 = p[k] - p0[k], d1 = p[ld + k] - p0[ld + k], d2 = p[2 * ld + k] - p0[2 * ld + k];
            s += d0 * d0 + d1 * d1 + d2 * d2;
        } else {
            double d[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double *q = p + c * ld + 3 * k, *q0 = p0 + c * ld + 3 * k;
                d[c] = ((q[0] - q0[0]) * mO + (q[1] - q0[1]) * mH + (q[2] - q0[2]) * mH) / (2 * mH + mO);
            }
            s += d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) b += wsum[w];
        partial[blockIdx.x] = b;
    }
}

__global__ void msd_final_kernel(const double *__restrict__ partial, int nb, double *__restrict__ out)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s = 0.0;
    for (int b = 0; b < nb; ++b) s += partial[b];
    out[0] = s;
}

static int analysis_frame(nbx_ctx *c, const double *u_host, double **rows)
{
    if (!u_host) { // the resident positions
        if (!c->resident) return fail(c, NBX_ERR_INVALID, "no resident state and no frame given");
        *rows = c->pos;
        return NBX_OK;
    }
    double *&dst = *rows;
    const size_t bytes = sizeof(double) * 3 * (size_t)c->ncols;
    NBX_CUDA(c, cudaMemcpyAsync(c->aos_dv, u_host, bytes, cudaMemcpyHostToDevice, c->stream)); // aos_dv: free between evaluations
    NBX_TRY(launch_aos_to_soa(c, c->aos_dv, dst, c->n));
    return NBX_OK;
}

int analysis_rdf_reset(nbx_ctx *c, int maxbin)
{
    if (maxbin < 2 || maxbin > 8192) return fail(c, NBX_ERR_INVALID, "nbx_rdf_reset: 2 <= maxbin <= 8192");
    if (c->an_bins != maxbin) {
        NBX_TRY(dev_alloc(c, &c->an_hist, (size_t)maxbin));
        c->an_bins = maxbin;
    }
    NBX_CUDA(c, cudaMemsetAsync(c->an_hist, 0, sizeof(unsigned long long) * (size_t)maxbin, c->stream));
    c->an_frames = 0;
    return NBX_OK;
}

int analysis_rdf_add(nbx_ctx *c, const double *u_host)
{
    if (c->bc_kind != NBX_BC_CUBIC) return fail(c, NBX_ERR_UNSUPPORTED, "nbx_rdf_add: rdf reads pbc.L (CubicPeriodicBoundaryConditions)");
    if (c->slab.on) return fail(c, NBX_ERR_INVALID, "nbx_rdf_add: slab-decomposed context");
    if (!c->an_hist) NBX_TRY(analysis_rdf_reset(c, 1000)); // maxbin = 1000, src/nbody_simulation_result.jl:671
    if (!c->an_pos) NBX_TRY(dev_alloc(c, &c->an_pos, (size_t)3 * (size_t)c->npad));
    double *rows = c->an_pos;
    NBX_TRY(analysis_frame(c, u_host, &rows));
    const int stride = c->water ? 3 : 1; // obtain_data_for_lennard_jones_interaction: oxygens of water, else everything
    const int m = (int)(c->n / stride);
    const double L = c->bc[0];
    const int nt = (m + kRdfTile - 1) / kRdfTile;
    const dim3 grid((unsigned)nt, (unsigned)nt);
    rdf_kernel<<<grid, kRdfTile, sizeof(int) * (size_t)c->an_bins, c->stream>>>(rows, c->npad, m, stride, L, 0.5 * L,
                                                                              (0.5 * L) * (0.5 * L), L / (double)c->an_bins,
                                                                              c->an_bins, c->an_hist);
    NBX_CUDA(c, cudaGetLastError());
    c->an_frames += 1;
    return NBX_OK;
}

int analysis_rdf_get(nbx_ctx *c, int64_t *hist, int64_t cap, int64_t *frames)
{
    if (!c->an_hist) return fail(c, NBX_ERR_INVALID, "nbx_rdf_get: no histogram (nbx_rdf_add first)");
    if (frames) *frames = c->an_frames;
    if (!hist) return NBX_OK;
    if (cap < c->an_bins) return fail(c, NBX_ERR_CAPACITY, "nbx_rdf_get: %d bins", c->an_bins);
    NBX_CUDA(c, cudaMemcpyAsync(hist, c->an_hist, sizeof(unsigned long long) * (size_t)c->an_bins, cudaMemcpyDeviceToHost, c->stream));
    NBX_CUDA(c, cudaStreamSynchronize(c->stream));
    return NBX_OK;
}

int analysis_msd(nbx_ctx *c, const double *u0_host, const double *u_host, double *out)
{
    if (!u0_host || !out) return fail(c, NBX_ERR_INVALID, "nbx_msd: u0 and out are required");
    if (c->slab.on) return fail(c, NBX_ERR_INVALID, "nbx_msd: slab-decomposed context");
    if (!c->an_pos) NBX_TRY(dev_alloc(c, &c->an_pos, (size_t)3 * (size_t)c->npad));
    if (!c->an_pos0) NBX_TRY(dev_alloc(c, &c->an_pos0, (size_t)3 * (size_t)c->npad));
    if (!c->an_red) NBX_TRY(dev_alloc(c, &c->an_red, (size_t)1024 + 1));
    double *r0 = c->an_pos0, *r = c->an_pos;
    NBX_TRY(analysis_frame(c, u0_host, &r0));
    NBX_TRY(analysis_frame(c, u_host, &r));
    const int64_t count = c->water ? c->n / 3 : c->n;
    int nb = (int)((count + 255) / 256);
    if (nb > 1024) nb = 1024;
    if (nb < 1) nb = 1;
    double mO = 0.0, mH = 0.0;
    if (c->water) { // masses of the first molecule (system.mO, system.mH)
        double hm[2];
        NBX_CUDA(c, cudaMemcpyAsync(hm, c->mass, sizeof hm, cudaMemcpyDeviceToHost, c->stream));
        NBX_CUDA(c, cudaStreamSynchronize(c->stream));
        mO = hm[0]; mH = hm[1];
    }
    msd_partial_kernel<<<nb, 256, 0, c->stream>>>(r, r0, c->npad, c->n, c->water ? 1 : 0, mO, mH, c->an_red);
    msd_final_kernel<<<1, 32, 0, c->stream>>>(c->an_red, nb, c->an_red + 1024);
    NBX_CUDA(c, cudaGetLastError());
    double s = 0.0;
    NBX_CUDA(c, cudaMemcpyAsync(&s, c->an_red + 1024, sizeof s, cudaMemcpyDeviceToHost, c->stream));
    NBX_CUDA(c, cudaStreamSynchronize(c->stream));
    *out = s / (double)count;
    return NBX_OK;
}

void analysis_free(nbx_ctx *c)
{
    cudaFree(c->an_hist); cudaFree(c->an_pos); cudaFree(c->an_pos0); cudaFree(c->an_red);
    c->an_hist = nullptr; c->an_pos = c->an_pos0 = c->an_red = nullptr;
    c->an_bins = 0; c->an_frames = 0;
}

} // namespace nbx
