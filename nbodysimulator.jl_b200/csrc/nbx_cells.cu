// nbx_cells.cu -- cell-list rebuild and cutoff pair kernels over the cell list (sm_100a).
//
// The reference has no spatial data structure: pairwise_lennard_jones_acceleration!
// (src/basic_potentials.jl:240-272) and pairwise_electrostatic_acceleration! (:274-304) test
// `r2 < R2` on every one of the N-1 partners of each particle.  This file produces the same
// in-cutoff pair set from a cell list: the predicate itself is evaluated exactly as the reference
// does (rij = ri - rj on the UNWRAPPED coordinates, the compare-and-subtract wrap loops of
// src/boundary_conditions.jl:138-165, un-fused r2), only the set of candidates it is applied to is
// pruned.  Cells have edge >= R (1 + 1e-6), so every pair the predicate can accept lies in the 27
// surrounding cells; binning uses its own wrapped copy x - L floor(x/L)
// (cf. src/nbody_simulation_result.jl:571), whose rounding (<= 1e-13 L) is far inside the margin.
//
// Rebuild: cell id + histogram (atomics) -> exclusive scan -> scatter -> per-cell sort by particle
// index (makes the order, hence every sum, deterministic) -> gather positions into cell order.
#include "nbx_internal.cuh"

namespace nbx {

// ------------------------------------------------------------------------------------------------
// planning
// ------------------------------------------------------------------------------------------------
// Only the cubic minimum-image box can be cell-listed; PeriodicBoundaryConditions is not a minimum
// image (src/boundary_conditions.jl:111-136) and InfiniteBox has no box.
int cells_plan(nbx_ctx *c, double R, int64_t n, CellGrid *g)
{
    g->valid = false;
    if (!c->opt_cell_list || c->bc_kind != NBX_BC_CUBIC) return NBX_OK;
    const double L = c->bc[0];
    if (!(L > 0.0) || !(R > 0.0) || !isfinite(R) || !isfinite(L)) return NBX_OK;
    double q = floor(L / (R * (1.0 + 1e-6)));
    if (q < 3.0) return NBX_OK;
    // sparse systems: no point in having many more cells than particles
    const double cap = ceil(cbrt(4.0 * (double)(n > 1 ? n : 1)));
    if (q > cap) q = cap < 3.0 ? 3.0 : cap;
    if (q > 1024.0) q = 1024.0;
    const int nc = (int)q;
    g->nc[0] = g->nc[1] = g->nc[2] = nc;
    g->ncell = (int64_t)nc * nc * nc;
    g->lo[0] = g->lo[1] = g->lo[2] = 0.0;
    g->len[0] = g->len[1] = g->len[2] = L;
    g->valid = true;
    return NBX_OK;
}

// ------------------------------------------------------------------------------------------------
// rebuild kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int cell_coord(double x, double L, int nc)
{
    double w = x - L * floor(x / L);
    if (w < 0.0) w += L;
    if (w >= L) w -= L;
    int cx = (int)(w * ((double)nc / L));
    return cx < 0 ? 0 : (cx >= nc ? nc - 1 : cx);
}

__global__ void cell_id_kernel(const double *__restrict__ px, int64_t ld, int n, double L, int nc,
                               int *__restrict__ cell_of, int *__restrict__ count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cx = cell_coord(px[i], L, nc), cy = cell_coord(px[ld + i], L, nc), cz = cell_coord(px[2 * ld + i], L, nc);
    const int cid = (cz * nc + cy) * nc + cx; // x fastest: the three x-neighbours of a cell are contiguous
    cell_of[i] = cid;
    atomicAdd(&count[cid], 1);
}

constexpr int kScanBlock = 1024;

// exclusive scan of in[0..n) by blocks of 1024; block totals to sums[]
__global__ void scan_block_kernel(const int *__restrict__ in, int *__restrict__ out, int n, int *__restrict__ sums)
{
    __shared__ int wsum[32];
    const int i = blockIdx.x * kScanBlock + threadIdx.x;
    const int v = i < n ? in[i] : 0;
    int s = v;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
    }
    if (lane == 31) wsum[warp] = s;
    __syncthreads();
    if (warp == 0) {
        int w = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        wsum[lane] = w;
    }
    __syncthreads();
    const int base = warp > 0 ? wsum[warp - 1] : 0;
    if (i < n) out[i] = base + s - v;
    if (threadIdx.x == kScanBlock - 1) sums[blockIdx.x] = base + s;
}

// serial-by-chunks exclusive scan of the block totals (single block), total -> sums[nb]
__global__ void scan_sums_kernel(int *__restrict__ sums, int nb)
{
    __shared__ int wsum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base0 = 0; base0 < nb; base0 += kScanBlock) {
        const int i = base0 + threadIdx.x;
        const int v = i < nb ? sums[i] : 0;
        int s = v;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane == 31) wsum[warp] = s;
        __syncthreads();
        if (warp == 0) {
            int w = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            wsum[lane] = w;
        }
        __syncthreads();
        const int b = (warp > 0 ? wsum[warp - 1] : 0) + carry;
        if (i < nb) sums[i] = b + s - v;
        __syncthreads();
        if (threadIdx.x == kScanBlock - 1) carry = b + s;
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[nb] = carry;
}

__global__ void scan_add_kernel(int *__restrict__ out, int n, const int *__restrict__ sums)
{
    const int i = blockIdx.x * kScanBlock + threadIdx.x;
    if (i < n) out[i] += sums[blockIdx.x];
    if (i == n - 1 || (n == 0 && i == 0)) { /* total written by the caller kernel below */ }
}

__global__ void scan_total_kernel(int *__restrict__ out, int n, const int *__restrict__ sums, int nb)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) out[n] = sums[nb];
}

__global__ void scatter_kernel(const int *__restrict__ cell_of, int n, const int *__restrict__ start,
                               int *__restrict__ fill, int *__restrict__ sorted_idx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cid = cell_of[i];
    const int slot = start[cid] + atomicAdd(&fill[cid], 1);
    sorted_idx[slot] = i;
}

// one thread per cell: insertion sort of the cell's particle indices (ascending)
__global__ void cell_sort_kernel(const int *__restrict__ start, int ncell, int *__restrict__ sorted_idx)
{
    const int cid = blockIdx.x * blockDim.x + threadIdx.x;
    if (cid >= ncell) return;
    const int b = start[cid], e = start[cid + 1];
    for (int k = b + 1; k < e; ++k) {
        const int v = sorted_idx[k];
        int m = k - 1;
        while (m >= b && sorted_idx[m] > v) { sorted_idx[m + 1] = sorted_idx[m]; --m; }
        sorted_idx[m + 1] = v;
    }
}

__global__ void gather_kernel(const double *__restrict__ px, int64_t ld, const double *__restrict__ w,
                              const int *__restrict__ sorted_idx, const int *__restrict__ cell_of, int n,
                              double *__restrict__ spos, int64_t sld, double *__restrict__ sw,
                              int *__restrict__ scell)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int i = sorted_idx[k];
    spos[k] = px[i];
    spos[sld + k] = px[ld + i];
    spos[2 * sld + k] = px[2 * ld + i];
    if (w) sw[k] = w[i];
    scell[k] = cell_of[i];
}

static int ensure_cells(nbx_ctx *c, CellList *cl, int64_t n, int64_t ncell)
{
    if (n > cl->cap_n) {
        const int64_t np = ((n + kPad - 1) / kPad) * kPad;
        NBX_TRY(dev_alloc(c, &cl->cell_of, (size_t)np));
        NBX_TRY(dev_alloc(c, &cl->sorted_idx, (size_t)np));
        NBX_TRY(dev_alloc(c, &cl->scell, (size_t)np));
        NBX_TRY(dev_alloc(c, &cl->spos, (size_t)3 * np));
        NBX_TRY(dev_alloc(c, &cl->sw, (size_t)np));
        cl->cap_n = np;
        cl->sld = np;
    }
    if (ncell > cl->cap_cells) {
        const int64_t nb = (ncell + kScanBlock - 1) / kScanBlock;
        NBX_TRY(dev_alloc(c, &cl->count, (size_t)ncell + 1));
        NBX_TRY(dev_alloc(c, &cl->start, (size_t)ncell + 1));
        NBX_TRY(dev_alloc(c, &cl->fill, (size_t)ncell + 1));
        NBX_TRY(dev_alloc(c, &cl->sums, (size_t)nb + 1));
        cl->cap_cells = ncell;
    }
    return NBX_OK;
}

// Rebuild cl for the n particles of the SoA rows px (stride ld); w = optional per-particle weight
// (charge) carried into cell order.
int cells_build(nbx_ctx *c, CellList *cl, const double *px, const double *w, int64_t n, int64_t ld)
{
    const CellGrid &g = cl->grid;
    if (!g.valid) return fail(c, NBX_ERR_INVALID, "cells_build without a valid grid");
    NBX_TRY(ensure_cells(c, cl, n, g.ncell));
    const int ncell = (int)g.ncell, ni = (int)n;
    const int nb = (ncell + kScanBlock - 1) / kScanBlock;
    timer_begin(c, NBX_T_CELL_BUILD);
    cudaMemsetAsync(cl->count, 0, sizeof(int) * (size_t)(ncell + 1), c->stream);
    cudaMemsetAsync(cl->fill, 0, sizeof(int) * (size_t)(ncell + 1), c->stream);
    cell_id_kernel<<<(ni + 255) / 256, 256, 0, c->stream>>>(px, ld, ni, g.len[0], g.nc[0], cl->cell_of, cl->count);
    scan_block_kernel<<<nb, kScanBlock, 0, c->stream>>>(cl->count, cl->start, ncell, cl->sums);
    scan_sums_kernel<<<1, kScanBlock, 0, c->stream>>>(cl->sums, nb);
    scan_add_kernel<<<nb, kScanBlock, 0, c->stream>>>(cl->start, ncell, cl->sums);
    scan_total_kernel<<<1, 32, 0, c->stream>>>(cl->start, ncell, cl->sums, nb);
    scatter_kernel<<<(ni + 255) / 256, 256, 0, c->stream>>>(cl->cell_of, ni, cl->start, cl->fill, cl->sorted_idx);
    cell_sort_kernel<<<(ncell + 127) / 128, 128, 0, c->stream>>>(cl->start, ncell, cl->sorted_idx);
    gather_kernel<<<(ni + 255) / 256, 256, 0, c->stream>>>(px, ld, w, cl->sorted_idx, cl->cell_of, ni, cl->spos,
                                                          cl->sld, cl->sw, cl->scell);
    timer_end(c, NBX_T_CELL_BUILD);
    NBX_CUDA(c, cudaGetLastError());
    cl->n = n;
    return NBX_OK;
}

// ------------------------------------------------------------------------------------------------
// pair kernel: one thread per target (cell order), 27 neighbour cells, exact reference predicate
// ------------------------------------------------------------------------------------------------
struct CellPairArgs {
    const double *sx, *sy, *sz, *sw;
    const int *sorted_idx, *scell, *start;
    int n, nc;
    double L, radius, R2, sigma2;
};

template <int POT, int EXCL, int MODE>
__device__ __forceinline__ void cell_pair_visit(const CellPairArgs &a, int k, double xi, double yi, double zi, int i,
                                                int cid, double &f0, double &f1, double &f2, int &cnt,
                                                int32_t *__restrict__ list)
{
    const int nc = a.nc;
    const int cx = cid % nc, cy = (cid / nc) % nc, cz = cid / (nc * nc);
    (void)k;
#pragma unroll 1
    for (int dz = -1; dz <= 1; ++dz) {
        int z = cz + dz; z = z < 0 ? z + nc : (z >= nc ? z - nc : z);
#pragma unroll 1
        for (int dy = -1; dy <= 1; ++dy) {
            int y = cy + dy; y = y < 0 ? y + nc : (y >= nc ? y - nc : y);
#pragma unroll 1
            for (int dx = -1; dx <= 1; ++dx) {
                int x = cx + dx; x = x < 0 ? x + nc : (x >= nc ? x - nc : x);
                const int cc = (z * nc + y) * nc + x;
                const int b = a.start[cc], e = a.start[cc + 1];
                for (int m = b; m < e; ++m) {
                    const int j = a.sorted_idx[m];
                    const bool excl = EXCL == 0 ? (j == i) : ((j / 3) == (i / 3));
                    if (excl) continue;
                    double rx = __dsub_rn(xi, a.sx[m]), ry = __dsub_rn(yi, a.sy[m]), rz = __dsub_rn(zi, a.sz[m]);
                    rx = wrap_cubic(rx, a.radius, a.L);
                    ry = wrap_cubic(ry, a.radius, a.L);
                    rz = wrap_cubic(rz, a.radius, a.L);
                    const double r2 = r2_unfused(rx, ry, rz);
                    if (r2 < a.R2) {
                        if (MODE == 0) {
                            double f;
                            if (POT == 0) {
                                const double inv = 1.0 / r2;
                                const double q = a.sigma2 * inv;
                                const double s6 = q * q * q;
                                const double s12 = s6 * s6;
                                f = (2.0 * s12 - s6) * inv;
                            } else {
                                f = w_rinv3(r2, a.sw[m]);
                            }
                            f0 = fma(f, rx, f0);
                            f1 = fma(f, ry, f1);
                            f2 = fma(f, rz, f2);
                        } else {
                            if (MODE == 2) list[cnt] = j;
                            ++cnt;
                        }
                    }
                }
            }
        }
    }
}

// scale kinds as in the all-pairs reduce kernel: 2 -> scale / m_i (LJ), 1 -> scale q_i / m_i (Coulomb)
template <int POT, int EXCL>
__global__ void __launch_bounds__(128) cell_force_kernel(const CellPairArgs a, double scale,
                                                         const double *__restrict__ mass, int mstride,
                                                         const double *__restrict__ charge, int lo, int hi,
                                                         double *__restrict__ acc, int64_t ld, int accumulate)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.n) return;
    const int i = a.sorted_idx[k];
    if (i < lo || i >= hi) return;
    double f0 = 0.0, f1 = 0.0, f2 = 0.0;
    int cnt = 0;
    cell_pair_visit<POT, EXCL, 0>(a, k, a.sx[k], a.sy[k], a.sz[k], i, a.scell[k], f0, f1, f2, cnt, nullptr);
    double coeff = scale / mass[(size_t)i * mstride];
    if (POT == 1) coeff *= charge[i];
    if (accumulate) {
        acc[i] += coeff * f0; acc[ld + i] += coeff * f1; acc[2 * ld + i] += coeff * f2;
    } else {
        acc[i] = coeff * f0; acc[ld + i] = coeff * f1; acc[2 * ld + i] = coeff * f2;
    }
}

template <int EXCL, int MODE>
__global__ void __launch_bounds__(128) cell_neigh_kernel(const CellPairArgs a, int *__restrict__ counts,
                                                         const int64_t *__restrict__ offsets,
                                                         int32_t *__restrict__ list)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.n) return;
    const int i = a.sorted_idx[k];
    double f0 = 0, f1 = 0, f2 = 0;
    int cnt = 0;
    int32_t *mine = MODE == 2 ? list + offsets[i] : nullptr;
    cell_pair_visit<0, EXCL, MODE>(a, k, a.sx[k], a.sy[k], a.sz[k], i, a.scell[k], f0, f1, f2, cnt, mine);
    if (MODE == 1) counts[i] = cnt;
    if (MODE == 2) { // ascending partner order, as the reference's j loop visits them
        for (int p = 1; p < cnt; ++p) {
            const int32_t v = mine[p];
            int m = p - 1;
            while (m >= 0 && mine[m] > v) { mine[m + 1] = mine[m]; --m; }
            mine[m + 1] = v;
        }
    }
}

static CellPairArgs make_args(const nbx_ctx *c, const CellList *cl, double R2)
{
    CellPairArgs a{};
    a.sx = cl->spos; a.sy = cl->spos + cl->sld; a.sz = cl->spos + 2 * cl->sld; a.sw = cl->sw;
    a.sorted_idx = cl->sorted_idx; a.scell = cl->scell; a.start = cl->start;
    a.n = (int)cl->n; a.nc = cl->grid.nc[0];
    a.L = c->bc[0]; a.radius = 0.5 * c->bc[0]; a.R2 = R2; a.sigma2 = c->lj_sigma2;
    return a;
}

// pot 0: LJ (self exclusion), 1: Coulomb (self exclusion), 2: Coulomb (own-molecule exclusion)
int launch_cells_force(nbx_ctx *c, CellList *cl, int pot, int64_t lo, int64_t hi, int mstride, double *acc_out,
                       int64_t ld_out, bool accumulate)
{
    const int n = (int)cl->n;
    if (n == 0) return NBX_OK;
    const int blocks = (n + 127) / 128;
    const int acc_flag = accumulate ? 1 : 0;
    timer_begin(c, NBX_T_PAIR_CELLS);
    if (pot == 0) {
        const CellPairArgs a = make_args(c, cl, c->lj_R2);
        cell_force_kernel<0, 0><<<blocks, 128, 0, c->stream>>>(a, 24.0 * c->lj_eps, c->mass, mstride, c->charge,
                                                             (int)lo, (int)hi, acc_out, ld_out, acc_flag);
    } else if (pot == 1) {
        const CellPairArgs a = make_args(c, cl, c->el_R2);
        cell_force_kernel<1, 0><<<blocks, 128, 0, c->stream>>>(a, c->el_k, c->mass, 1, c->charge, (int)lo, (int)hi,
                                                             acc_out, ld_out, acc_flag);
    } else {
        const CellPairArgs a = make_args(c, cl, c->el_R2);
        cell_force_kernel<1, 1><<<blocks, 128, 0, c->stream>>>(a, c->el_k, c->mass, 1, c->charge, (int)lo, (int)hi,
                                                             acc_out, ld_out, acc_flag);
    }
    timer_end(c, NBX_T_PAIR_CELLS);
    NBX_CUDA(c, cudaGetLastError());
    return NBX_OK;
}

// ------------------------------------------------------------------------------------------------
// neighbour lists (diagnostic / parity API): CSR over original particle indices
// ------------------------------------------------------------------------------------------------
// brute-force variant for boxes that cannot be cell-listed: one thread per target, all j ascending
__global__ void brute_neigh_kernel(const double *__restrict__ px, int64_t ld, int n, int bc_kind, double b0, double b1,
                                   double b2, double b3, double b4, double b5, double R2, int mode,
                                   int *__restrict__ counts, const int64_t *__restrict__ offsets,
                                   int32_t *__restrict__ list)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double xi = px[i], yi = px[ld + i], zi = px[2 * ld + i];
    int cnt = 0;
    int32_t *mine = mode == 2 ? list + offsets[i] : nullptr;
    for (int j = 0; j < n; ++j) {
        if (j == i) continue;
        double x = __dsub_rn(xi, px[j]), y = __dsub_rn(yi, px[ld + j]), z = __dsub_rn(zi, px[2 * ld + j]);
        if (bc_kind == NBX_BC_CUBIC) {
            x = wrap_cubic(x, b1, b0); y = wrap_cubic(y, b1, b0); z = wrap_cubic(z, b1, b0);
        } else if (bc_kind == NBX_BC_PERIODIC) {
            x = wrap_range(x, b0, b1); y = wrap_range(y, b2, b3); z = wrap_range(z, b4, b5);
        }
        if (r2_unfused(x, y, z) < R2) {
            if (mode == 2) mine[cnt] = j;
            ++cnt;
        }
    }
    if (mode == 1) counts[i] = cnt;
}

// The LJ predicate's in-cutoff pair set of the n particles in px (the resident positions, or the
// oxygen sub-system for water).  Host arrays: offsets[n+1], list[cap].
int cells_neighbors(nbx_ctx *c, CellList *cl, const double *px, int64_t n, int64_t ld, double R2, int64_t *offsets,
                    int32_t *list, int64_t cap)
{
    const int ni = (int)n;
    int *d_counts = nullptr;
    int64_t *d_off = nullptr;
    int32_t *d_list = nullptr;
    NBX_TRY(dev_alloc(c, &d_counts, (size_t)n + 1));
    std::vector<int> hc((size_t)n);
    const bool use_cells = cl->grid.valid;
    const int blocks = (ni + 127) / 128;
    CellPairArgs a{};
    double b[6] = {c->bc[0], c->bc[1], c->bc[2], c->bc[3], c->bc[4], c->bc[5]};
    if (c->bc_kind == NBX_BC_CUBIC) b[1] = 0.5 * c->bc[0];
    if (use_cells) {
        int rc = cells_build(c, cl, px, nullptr, n, ld);
        if (rc != NBX_OK) { cudaFree(d_counts); return rc; }
        a = make_args(c, cl, R2);
        cell_neigh_kernel<0, 1><<<blocks, 128, 0, c->stream>>>(a, d_counts, nullptr, nullptr);
    } else {
        brute_neigh_kernel<<<blocks, 128, 0, c->stream>>>(px, ld, ni, c->bc_kind, b[0], b[1], b[2], b[3], b[4], b[5],
                                                         R2, 1, d_counts, nullptr, nullptr);
    }
    cudaError_t e = cudaMemcpyAsync(hc.data(), d_counts, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { cudaFree(d_counts); return cuda_fail(c, e, "neighbour counts"); }
    offsets[0] = 0;
    for (int64_t i = 0; i < n; ++i) offsets[i + 1] = offsets[i] + hc[(size_t)i];
    const int64_t total = offsets[n];
    if (total > cap) {
        cudaFree(d_counts);
        return fail(c, NBX_ERR_CAPACITY, "nbx_neighbors: list needs %lld entries, capacity %lld", (long long)total,
                    (long long)cap);
    }
    int rc = NBX_OK;
    if (total > 0) {
        rc = dev_alloc(c, &d_off, (size_t)n + 1);
        if (rc == NBX_OK) rc = dev_alloc(c, &d_list, (size_t)total);
        if (rc == NBX_OK) {
            cudaMemcpyAsync(d_off, offsets, sizeof(int64_t) * (size_t)(n + 1), cudaMemcpyHostToDevice, c->stream);
            if (use_cells)
                cell_neigh_kernel<0, 2><<<blocks, 128, 0, c->stream>>>(a, d_counts, d_off, d_list);
            else
                brute_neigh_kernel<<<blocks, 128, 0, c->stream>>>(px, ld, ni, c->bc_kind, b[0], b[1], b[2], b[3], b[4],
                                                                 b[5], R2, 2, d_counts, d_off, d_list);
            e = cudaMemcpyAsync(list, d_list, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            if (e != cudaSuccess) rc = cuda_fail(c, e, "neighbour list");
        }
    }
    cudaFree(d_counts);
    if (d_off) cudaFree(d_off);
    if (d_list) cudaFree(d_list);
    return rc;
}

void cells_free(CellList *cl)
{
    cudaFree(cl->cell_of); cudaFree(cl->count); cudaFree(cl->start); cudaFree(cl->fill); cudaFree(cl->sums);
    cudaFree(cl->sorted_idx); cudaFree(cl->scell); cudaFree(cl->spos); cudaFree(cl->sw);
    *cl = CellList{};
}

} // namespace nbx
