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
// Rebuild: cell id + histogram (atomics) -> exclusive scan -> scatter -> rank inside the cell by particle id
// (makes the order, hence every sum, deterministic) fused with the gather of the records into cell order.
//
// Pair kernel (cell_pairs2_kernel): one lane per target.  Phase 1 scans the 27 surrounding cells
// (9 x-rows, each one contiguous slot range plus at most one periodic wrap-around cell) with an fp32
// distance test on wrapped coordinates in cell units; the threshold carries a margin that covers the
// fp32 error, so it can only ever over-accept.  Survivors (about 15 % of the candidates in a liquid)
// go to a per-lane queue in shared memory.  Phase 2 drains the queues in lock step: the reference's
// exact fp64 predicate decides, and the force is evaluated only there.
#include "nbx_internal.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

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
__global__ void cell_id_kernel(const double *__restrict__ px, int64_t ld, int n, double L, int nc,
                               int *__restrict__ cell_of, int *__restrict__ arrival, int *__restrict__ count,
                               const int *__restrict__ dyn, const int *__restrict__ cond)
{
    if (cond && !cond[0]) return; // Verlet mode: the list is still valid, nothing to rebuild
    n = dyn_loc(dyn, n);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int cx = cell_coord(px[i], L, nc), cy = cell_coord(px[ld + i], L, nc), cz = cell_coord(px[2 * ld + i], L, nc);
        const int cid = (cz * nc + cy) * nc + cx; // x fastest: the three x-neighbours of a cell are contiguous
        cell_of[i] = cid;
        arrival[i] = atomicAdd(&count[cid], 1);  // arbitrary but unique slot inside the cell (ordered later)
    }
}

constexpr int kScanBlock = 1024;

// exclusive scan of in[0..n) by blocks of 1024; block totals to sums[]
// (round_to > 1: every input is first rounded up to a multiple of it -- padded cell sizes of nbx_fused.cu)
__global__ void scan_block_kernel(const int *__restrict__ in, int *__restrict__ out, int n, int *__restrict__ sums,
                                  const int *__restrict__ cond, int round_to)
{
    __shared__ int wsum[32];
    if (cond && !cond[0]) return;
    const int i = blockIdx.x * kScanBlock + threadIdx.x;
    int v = i < n ? in[i] : 0;
    if (round_to > 1) v = (v + round_to - 1) / round_to * round_to;
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
__global__ void scan_sums_kernel(int *__restrict__ sums, int nb, const int *__restrict__ cond)
{
    __shared__ int wsum[32];
    __shared__ int carry;
    if (cond && !cond[0]) return;
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

__global__ void scan_add_kernel(int *__restrict__ out, int n, const int *__restrict__ sums, int nb,
                                const int *__restrict__ cond)
{
    if (cond && !cond[0]) return;
    const int i = blockIdx.x * kScanBlock + threadIdx.x;
    if (i < n) out[i] += sums[blockIdx.x];
    if (i == 0) out[n] = sums[nb]; // grand total closes the CSR
}

// exclusive scan out[0..n] of in[0..n) (inputs rounded up to multiples of round_to), out[n] = total; sums: scratch of
// ceil(n / 1024) + 1 ints.  The three launches return at once when cond && !cond[0].
int cells_scan(nbx_ctx *c, const int *in, int *out, int n, int *sums, int round_to, const int *cond)
{
    const int nb = (n + kScanBlock - 1) / kScanBlock;
    scan_block_kernel<<<nb, kScanBlock, 0, c->stream>>>(in, out, n, sums, cond, round_to);
    scan_sums_kernel<<<1, kScanBlock, 0, c->stream>>>(sums, nb, cond);
    scan_add_kernel<<<nb, kScanBlock, 0, c->stream>>>(out, n, sums, nb, cond);
    NBX_CUDA(c, cudaGetLastError());
    return NBX_OK;
}

// slot = cell start + arrival rank (no atomics); the ordering key travels with the index
__global__ void scatter_kernel(const int *__restrict__ cell_of, const int *__restrict__ arrival,
                               const int *__restrict__ gid, int n, const int *__restrict__ start,
                               int *__restrict__ tmp_idx, int *__restrict__ tmp_key, const int *__restrict__ dyn,
                               const int *__restrict__ cond)
{
    if (cond && !cond[0]) return;
    n = dyn_loc(dyn, n);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int slot = start[cell_of[i]] + arrival[i];
        tmp_idx[slot] = i;
        if (gid) tmp_key[slot] = gid[i];
    }
}

// Final slot of a particle = cell start + its rank among the cell's members by key (global particle id, or
// the local index when the context holds the whole system): the cell order, hence every force sum, is
// deterministic and independent of how a decomposition numbers its local particles.  The same thread then
// writes the particle's records in cell order: the exact coordinates + weight (one 32-byte record), and the
// fp32 prefilter record: wrapped coordinates in cell units [0, nc) + the exclusion key (particle id, or
// molecule id = id / 3 for the own-molecule exclusion of src/nbody_to_ode.jl:331-351).
__global__ void rank_gather_kernel(const double *__restrict__ px, int64_t ld, const double *__restrict__ w,
                                   const int *__restrict__ tmp_idx, const int *__restrict__ tmp_key,
                                   const int *__restrict__ cell_of, const int *__restrict__ start, int n, double L,
                                   int nc, int key_div, int *__restrict__ sorted_idx, double4 *__restrict__ sp4,
                                   float4 *__restrict__ sl4, int *__restrict__ scell, const int *__restrict__ dyn,
                                   const int *__restrict__ cond, int *__restrict__ slot_of)
{
    if (cond && !cond[0]) return;
    n = dyn_loc(dyn, n);
    const int *keys = tmp_key ? tmp_key : tmp_idx;
    const double s = (double)nc / L;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int i = tmp_idx[k];
        const int cid = cell_of[i];
        const int b = start[cid], e = start[cid + 1];
        const int mine = keys[k];
        int rank = 0;
        for (int m = b; m < e; ++m) rank += keys[m] < mine ? 1 : 0;
        const int dst = b + rank;
        const double x = px[i], y = px[ld + i], z = px[2 * ld + i];
        sorted_idx[dst] = i;
        slot_of[i] = dst;
        sp4[dst] = make_double4(x, y, z, w ? w[i] : 0.0);
        sl4[dst] = make_float4((float)(wrapped_coord(x, L) * s), (float)(wrapped_coord(y, L) * s),
                               (float)(wrapped_coord(z, L) * s), __int_as_float(mine / key_div));
        scell[dst] = cid;
    }
}

// grid of an element-wise (grid-stride) kernel: enough blocks to fill the machine, few enough that a launch whose
// condition is off costs next to nothing
static unsigned map_grid(const nbx_ctx *c, int n, int threads)
{
    const int full = (n + threads - 1) / threads;
    const int cap = c->sm_count * (2048 / threads);
    return (unsigned)(full < cap ? (full > 0 ? full : 1) : cap);
}

static int ensure_cells(nbx_ctx *c, CellList *cl, int64_t n, int64_t ncell)
{
    if (n > cl->cap_n) {
        const int64_t np = ((n + kPad - 1) / kPad) * kPad;
        NBX_TRY(dev_alloc(c, &cl->cell_of, (size_t)np));
        NBX_TRY(dev_alloc(c, &cl->arrival, (size_t)np));
        NBX_TRY(dev_alloc(c, &cl->tmp_idx, (size_t)np));
        NBX_TRY(dev_alloc(c, &cl->tmp_key, (size_t)np));
        NBX_TRY(dev_alloc(c, &cl->sorted_idx, (size_t)np));
        NBX_TRY(dev_alloc(c, &cl->slot_of, (size_t)np));
        NBX_TRY(dev_alloc(c, &cl->scell, (size_t)np));
        NBX_TRY(dev_alloc(c, &cl->sp4, (size_t)np));
        NBX_TRY(dev_alloc(c, &cl->sl4, (size_t)np));
        cl->cap_n = np;
    }
    if (ncell > cl->cap_cells) {
        const int64_t nb = (ncell + kScanBlock - 1) / kScanBlock;
        NBX_TRY(dev_alloc(c, &cl->count, (size_t)ncell + 1));
        NBX_TRY(dev_alloc(c, &cl->start, (size_t)ncell + 1));
        NBX_TRY(dev_alloc(c, &cl->sums, (size_t)nb + 1));
        cl->cap_cells = ncell;
    }
    return NBX_OK;
}

// Rebuild cl for the n particles of the SoA rows px (stride ld); w = optional per-particle weight
// (charge) carried into cell order.
int cells_build(nbx_ctx *c, CellList *cl, const double *px, const double *w, const int *gid, int64_t n, int64_t ld,
                int key_div, const int *cond)
{
    const CellGrid &g = cl->grid;
    if (!g.valid) return fail(c, NBX_ERR_INVALID, "cells_build without a valid grid");
    NBX_TRY(ensure_cells(c, cl, n, g.ncell));
    const int ncell = (int)g.ncell, ni = (int)n;
    const int nb = (ncell + kScanBlock - 1) / kScanBlock;
    timer_begin(c, NBX_T_CELL_BUILD);
    cudaMemsetAsync(cl->count, 0, sizeof(int) * (size_t)(ncell + 1), c->stream);
    cell_id_kernel<<<map_grid(c, ni, 256), 256, 0, c->stream>>>(px, ld, ni, g.len[0], g.nc[0], cl->cell_of, cl->arrival,
                                                           cl->count, c->dyn, cond);
    scan_block_kernel<<<nb, kScanBlock, 0, c->stream>>>(cl->count, cl->start, ncell, cl->sums, cond, 1);
    scan_sums_kernel<<<1, kScanBlock, 0, c->stream>>>(cl->sums, nb, cond);
    scan_add_kernel<<<nb, kScanBlock, 0, c->stream>>>(cl->start, ncell, cl->sums, nb, cond);
    scatter_kernel<<<map_grid(c, ni, 256), 256, 0, c->stream>>>(cl->cell_of, cl->arrival, gid, ni, cl->start, cl->tmp_idx,
                                                           cl->tmp_key, c->dyn, cond);
    rank_gather_kernel<<<map_grid(c, ni, 128), 128, 0, c->stream>>>(px, ld, w, cl->tmp_idx, gid ? cl->tmp_key : nullptr,
                                                               cl->cell_of, cl->start, ni, g.len[0], g.nc[0], key_div,
                                                               cl->sorted_idx, cl->sp4, cl->sl4, cl->scell, c->dyn, cond, cl->slot_of);
    timer_end(c, NBX_T_CELL_BUILD);
    NBX_CUDA(c, cudaGetLastError());
    cl->n = n;
    return NBX_OK;
}

// ------------------------------------------------------------------------------------------------
// pair kernel: one thread per target (cell order), 27 neighbour cells, exact reference predicate
// ------------------------------------------------------------------------------------------------
struct CellPairArgs {
    const double4 *sp4;
    const float4 *sl4;
    const int *sorted_idx, *scell, *start;
    int n, nc;
    double L, radius, R2, sigma2;
    float R2f;     // prefilter threshold on the fp32 squared distance in cell units (margin included)
    int hi_radius; // high word of `radius`: |x| with a smaller high word needs no wrapping
    int hi_far2;   // high word of fl(radius^2): an unwrapped r2 with a smaller high word has no component to wrap
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
                    const double4 pj = a.sp4[m];
                    double rx = __dsub_rn(xi, pj.x), ry = __dsub_rn(yi, pj.y), rz = __dsub_rn(zi, pj.z);
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
                                f = w_rinv3(r2, pj.w);
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
                                                         double *__restrict__ acc, int64_t ld, int accumulate,
                                                         const int *__restrict__ dyn)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= dyn_loc(dyn, a.n)) return;
    const int i = a.sorted_idx[k];
    if (i < lo || i >= (dyn ? min(hi, dyn[0]) : hi)) return;
    double f0 = 0.0, f1 = 0.0, f2 = 0.0;
    int cnt = 0;
    const double4 pi = a.sp4[k];
    cell_pair_visit<POT, EXCL, 0>(a, k, pi.x, pi.y, pi.z, i, a.scell[k], f0, f1, f2, cnt, nullptr);
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
    const double4 pi = a.sp4[k];
    cell_pair_visit<0, EXCL, MODE>(a, k, pi.x, pi.y, pi.z, i, a.scell[k], f0, f1, f2, cnt, mine);
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


// ------------------------------------------------------------------------------------------------
// pair kernel v2: fp32 prefilter -> per-lane survivor queue -> exact fp64 predicate + force
// ------------------------------------------------------------------------------------------------
constexpr int kQCap = 32; // survivors a lane can hold before the warp drains its queues

// MODE 0: accelerations; 1: in-cutoff partner counts; 2: partner lists (CSR via offsets)
// POT 0: Lennard-Jones (src/basic_potentials.jl:253-266); 1: Coulomb (:288-297).  The exclusion (self, or own
// molecule) is the key stored in sl4[].w.
template <int POT, int MODE>
__global__ void __launch_bounds__(128, 8) cell_pairs2_kernel(const CellPairArgs a, double scale,
                                                          const double *__restrict__ mass, int mstride,
                                                          const double *__restrict__ charge, int lo, int hi,
                                                          double *__restrict__ acc, int64_t ld, int accumulate,
                                                          int *__restrict__ counts, const int64_t *__restrict__ offsets,
                                                          int32_t *__restrict__ list, const int *__restrict__ dyn,
                                                          const int *__restrict__ cond)
{
    __shared__ int q[kQCap * 128];
    if (cond && !cond[0]) return; // Verlet mode: runs only as the fallback when a list overflowed
    const int n_loc = dyn_loc(dyn, a.n);
    if (dyn) hi = min(hi, dyn[0]);
    constexpr unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x;
    // block-stride over the 128-slot groups (the grid may be a bound, or deliberately small for the fallback launch)
    for (int blk = blockIdx.x; blk * 128 < n_loc; blk += gridDim.x) {
    const int k = blk * 128 + tid;
    const int kk = k < n_loc ? k : n_loc - 1;
    const float4 me = a.sl4[kk];
    const int key = __float_as_int(me.w);
    const int i = a.sorted_idx[kk];
    const bool live = k < n_loc && i >= lo && i < hi;
    const double4 pi = a.sp4[kk];
    const int cid = a.scell[kk];
    const int nc = a.nc;
    const int cx = cid % nc, cy = (cid / nc) % nc, cz = cid / (nc * nc);
    const float fnc = (float)nc;
    double f0 = 0.0, f1 = 0.0, f2 = 0.0;
    int cnt = 0, found = 0;
    int32_t *mine = (MODE == 2 && live) ? list + offsets[i] : nullptr;

    // phase 2: every lane drains its own queue; the reference's predicate decides
    auto pair = [&](const double4 pj, int m) {
        double rx = __dsub_rn(pi.x, pj.x), ry = __dsub_rn(pi.y, pj.y), rz = __dsub_rn(pi.z, pj.z);
        const int hx = __double2hiint(rx) & 0x7fffffff, hy = __double2hiint(ry) & 0x7fffffff,
                  hz = __double2hiint(rz) & 0x7fffffff;
        if (max(hx, max(hy, hz)) >= a.hi_radius) { // rare: the pair straddles a periodic face
            rx = wrap_cubic(rx, a.radius, a.L);
            ry = wrap_cubic(ry, a.radius, a.L);
            rz = wrap_cubic(rz, a.radius, a.L);
        }
        const double r2 = r2_unfused(rx, ry, rz);
        // r2 >= 0 and R2 > 0: IEEE order == order of the bit patterns (keeps the test off the FP64 pipe)
        if (__double_as_longlong(r2) < __double_as_longlong(a.R2)) {
            if (MODE == 0) {
                double f;
                if (POT == 0) {
                    const double inv = rcp_fast(r2);
                    const double qq = a.sigma2 * inv;
                    const double s6 = qq * qq * qq;
                    f = (s6 * inv) * fma(2.0, s6, -1.0); // (2 s12 - s6) / r2
                } else {
                    f = w_rinv3(r2, pj.w);
                }
                f0 = fma(f, rx, f0);
                f1 = fma(f, ry, f1);
                f2 = fma(f, rz, f2);
            } else {
                if (MODE == 2) mine[found] = a.sorted_idx[m];
                ++found;
            }
        }
    };
    auto flush = [&]() {
        const int wmax = __reduce_max_sync(FULL, cnt);
        for (int s = 0; s < wmax; s += 2) { // two gathers in flight per lane
            const bool v0 = s < cnt, v1 = s + 1 < cnt;
            const int m0 = v0 ? q[s * 128 + tid] : kk, m1 = v1 ? q[(s + 1) * 128 + tid] : kk;
            const double4 p0 = load_rec(a.sp4 + m0), p1 = load_rec(a.sp4 + m1);
            if (v0) pair(p0, m0);
            if (v1) pair(p1, m1);
        }
        cnt = 0;
    };

    // phase 1: fp32 test of the slots [b, e) against the (shifted) target
    auto test = [&](const float4 cj, int m, float tx, float ty, float tz) {
        const float dx = tx - cj.x, dy = ty - cj.y, dz = tz - cj.z;
        const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (r2 < a.R2f && __float_as_int(cj.w) != key) {
            q[cnt * 128 + tid] = m;
            ++cnt;
        }
    };
    auto scan = [&](int b, int e, float tx, float ty, float tz) {
        int m = b;
        for (;;) {
            // four records in flight per lane: the loop is bound by load latency, not by arithmetic
            while (m + 4 <= e && cnt <= kQCap - 4) {
                const float4 c0 = __ldg(&a.sl4[m]), c1 = __ldg(&a.sl4[m + 1]), c2 = __ldg(&a.sl4[m + 2]),
                             c3 = __ldg(&a.sl4[m + 3]);
                test(c0, m, tx, ty, tz);
                test(c1, m + 1, tx, ty, tz);
                test(c2, m + 2, tx, ty, tz);
                test(c3, m + 3, tx, ty, tz);
                m += 4;
            }
            while (m < e && e - m < 4 && cnt < kQCap) {
                test(__ldg(&a.sl4[m]), m, tx, ty, tz);
                ++m;
            }
            if (!__any_sync(FULL, m < e)) break;
            flush(); // some lane's queue is (nearly) full
        }
    };

#pragma unroll 1
    for (int r = 0; r < 9; ++r) {
        const int dz = r / 3 - 1, dy = r - (r / 3) * 3 - 1;
        int z = cz + dz, y = cy + dy;
        float tz = me.z, ty = me.y;
        // a neighbour cell reached through a periodic face holds the image w_j -+ nc of its particles
        if (z < 0) { z += nc; tz += fnc; } else if (z >= nc) { z -= nc; tz -= fnc; }
        if (y < 0) { y += nc; ty += fnc; } else if (y >= nc) { y -= nc; ty -= fnc; }
        const int row = (z * nc + y) * nc;
        const int xa = cx > 0 ? cx - 1 : 0, xb = cx < nc - 1 ? cx + 1 : nc - 1;
        int b = a.start[row + xa], e = a.start[row + xb + 1];
        if (!live) e = b;
        scan(b, e, me.x, ty, tz);
        int b2 = 0, e2 = 0;
        float tx = me.x;
        if (cx == 0) { b2 = a.start[row + nc - 1]; e2 = a.start[row + nc]; tx += fnc; }
        else if (cx == nc - 1) { b2 = a.start[row]; e2 = a.start[row + 1]; tx -= fnc; }
        if (!live) e2 = b2;
        if (__any_sync(FULL, e2 > b2)) scan(b2, e2, tx, ty, tz);
    }
    flush();

    if (!live) continue;
    if (MODE == 0) {
        double coeff = scale / mass[(size_t)i * mstride];
        if (POT == 1) coeff *= charge[i];
        if (accumulate) {
            acc[i] += coeff * f0; acc[ld + i] += coeff * f1; acc[2 * ld + i] += coeff * f2;
        } else {
            acc[i] = coeff * f0; acc[ld + i] = coeff * f1; acc[2 * ld + i] = coeff * f2;
        }
    } else if (MODE == 1) {
        counts[i] = found;
    } else { // ascending partner order, as the reference's j loop visits them
        for (int p = 1; p < found; ++p) {
            const int32_t v = mine[p];
            int m = p - 1;
            while (m >= 0 && mine[m] > v) { mine[m + 1] = mine[m]; --m; }
            mine[m + 1] = v;
        }
    }
    } // block-stride loop
}

// ------------------------------------------------------------------------------------------------
// Verlet lists: the survivor set of the fp32 scan, kept across steps
// ------------------------------------------------------------------------------------------------
// The scan of cell_pairs2_kernel with the threshold widened to R + skin produces, per slot, the list of slots
// that can come within R while no particle has moved more than skin/2 since the build (|d(t) - d(build)| <
// skin).  Every later evaluation applies the reference's exact predicate to the listed pairs only, so the
// in-cutoff pair set is unchanged; the list is rebuilt -- on the device's own decision, no host round trip --
// as soon as some particle's displacement from its build-time position exceeds skin/2.
// flags: [0] rebuild now, [1] a list overflowed its capacity (sticky: from then on every evaluation rebuilds
// the cells and takes the scan-per-step kernel instead).
__global__ void verlet_check_kernel(const double *__restrict__ px, int64_t ld, const double *__restrict__ ref, int64_t rld,
                                    int n, double lim2, int *__restrict__ flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && flags[1]) flags[0] = 1;
    if (i >= n) return;
    const double dx = px[i] - ref[i], dy = px[ld + i] - ref[rld + i], dz = px[2 * ld + i] - ref[2 * rld + i];
    const double d2 = dx * dx + dy * dy + dz * dz;
    if (!(d2 <= lim2)) flags[0] = 1; // also catches NaN
}

struct VerletArgs {
    int *list;   // [cap][stride]: entry e of slot k at e * stride + k
    int *nlist;  // [n]
    int *flags;
    int cap;
    int64_t stride;
    const int *dyn; // slab mode: [0] own, [1] ghosts on the device (launch sizes are bounds); else null
    int consume;    // the force kernel clears flags[0] (no refresh kernel ran: the position update refreshed the records)
    int banked;     // the lists are laid out in blocks of four by record position inside a 128-byte line (see verlet_build_slot)
    int branchfree; // force kernel: batches of four evaluated without branches (interleaved dependency chains)
};

// one lane per slot: the fp32 scan of cell_pairs2_kernel, survivors appended to the slot's list
__device__ __forceinline__ void verlet_build_slot(const CellPairArgs &a, const VerletArgs &v, int k);

__global__ void __launch_bounds__(128) verlet_build_kernel(const CellPairArgs a, const VerletArgs v)
{
    if (!v.flags[0] || v.flags[1]) return;
    const int n_loc = dyn_loc(v.dyn, a.n);
    for (int k = blockIdx.x * 128 + threadIdx.x; k < n_loc; k += gridDim.x * 128)
        verlet_build_slot(a, v, k);
}

// Banked layout (VerletArgs::banked).  A gathered record is one 32-byte sector; the L1 data stage serves a 256-bit warp
// load four lanes at a time and needs one pass per DISTINCT sector of the same position inside a 128-byte line (measured,
// profiles/micro/l1_gather.cu: 32 random records cost 20-21 cycles per warp gather, 9.5-12 when the four lanes of every
// group read the four different positions).  The position of a record is its slot number mod 4, so the list of a slot is
// written in blocks of four entries, entry p of a block being a record of position p, and the force kernel makes the four
// lanes of a group read four different p of their blocks at any time.  A slot's partners are not spread evenly over the
// four positions: entries of an over-full position that do not fit below the list's final block count go to places
// left open by the other positions (in scan order -- nothing depends on anything but the slot numbers), the last
// places still open hold the slot's own number as a sentinel (r2 = 0 exactly: the force kernel's predicate is
// 0 < r2 < R2).
__device__ __forceinline__ void verlet_build_slot(const CellPairArgs &a, const VerletArgs &v, int k)
{
    const float4 me = a.sl4[k];
    const int key = __float_as_int(me.w);
    const int cid = a.scell[k];
    const int nc = a.nc;
    const int cx = cid % nc, cy = (cid / nc) % nc, cz = cid / (nc * nc);
    const float fnc = (float)nc;
    auto sweep = [&](auto &&hit) {
        auto scan = [&](int b, int e, float tx, float ty, float tz) {
            for (int m = b; m < e; ++m) {
                const float4 cj = __ldg(&a.sl4[m]);
                const float dx = tx - cj.x, dy = ty - cj.y, dz = tz - cj.z;
                const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                if (r2 < a.R2f && __float_as_int(cj.w) != key) hit(m);
            }
        };
#pragma unroll 1
        for (int r = 0; r < 9; ++r) {
            const int dz = r / 3 - 1, dy = r - (r / 3) * 3 - 1;
            int z = cz + dz, y = cy + dy;
            float tz = me.z, ty = me.y;
            if (z < 0) { z += nc; tz += fnc; } else if (z >= nc) { z -= nc; tz -= fnc; }
            if (y < 0) { y += nc; ty += fnc; } else if (y >= nc) { y -= nc; ty -= fnc; }
            const int row = (z * nc + y) * nc;
            const int xa = cx > 0 ? cx - 1 : 0, xb = cx < nc - 1 ? cx + 1 : nc - 1;
            scan(a.start[row + xa], a.start[row + xb + 1], me.x, ty, tz);
            if (cx == 0) scan(a.start[row + nc - 1], a.start[row + nc], me.x + fnc, ty, tz);
            else if (cx == nc - 1) scan(a.start[row], a.start[row + 1], me.x - fnc, ty, tz);
        }
    };
    if (!v.banked) {
        int cnt = 0;
        sweep([&](int m) {
            if (cnt < v.cap) v.list[(size_t)cnt * v.stride + k] = m;
            ++cnt;
        });
        v.nlist[k] = cnt < v.cap ? cnt : v.cap;
        if (cnt > v.cap) v.flags[1] = 1;
        return;
    }
    // one scan: survivor number t of position p goes to row p of block t at once (the rows reach up to twice the
    // capacity: a position may hold more than a quarter of a list); the four counters live in one register pair
    unsigned long long have = 0ull;
    int cnt = 0;
    bool over = false;
    const int bmax = v.cap >> 1; // blocks that exist (2 x cap rows)
    sweep([&](int m) {
        const int p = m & 3;
        const int t = (int)((have >> (p * 16)) & 0xffffull);
        if (t < bmax) v.list[(size_t)(4 * t + ((p - k) & 3)) * v.stride + k] = m;
        else over = true;
        have += 1ull << (p * 16);
        ++cnt;
    });
    if (cnt > v.cap || over) { // (cap is a multiple of four)
        v.nlist[k] = 0;
        v.flags[1] = 1;
        return;
    }
    // the list ends after `blocks` blocks: entries beyond (of over-full positions) move to the places the other
    // positions left open below, blocks [count_of(q), blocks) of position q; what stays open holds the sentinel
    const int blocks = (cnt + 3) >> 2;
    auto count_of = [&](int p) { return (int)((have >> (p * 16)) & 0xffffull); };
    int hq = 0, hh = count_of(0);
    auto settle = [&]() { while (hq < 4 && hh >= blocks) { ++hq; hh = hq < 4 ? count_of(hq) : 0; } };
    settle();
    for (int p = 0; p < 4; ++p)
        for (int t = blocks; t < count_of(p); ++t) {
            const int m = v.list[(size_t)(4 * t + ((p - k) & 3)) * v.stride + k];
            v.list[(size_t)(4 * hh + ((hq - k) & 3)) * v.stride + k] = m;
            ++hh; settle();
        }
    while (hq < 4) { v.list[(size_t)(4 * hh + ((hq - k) & 3)) * v.stride + k] = k; ++hh; settle(); } // sentinel: the slot itself
    v.nlist[k] = 4 * blocks;
}

// after a rebuild: remember the positions the list was built from
__global__ void verlet_ref_kernel(const double *__restrict__ px, int64_t ld, double *__restrict__ ref, int64_t rld, int n,
                                  int *__restrict__ flags, const int *__restrict__ dyn)
{
    if (!flags[0]) return;
    n = dyn_loc(dyn, n);
    if (blockIdx.x == 0 && threadIdx.x == 0) flags[2] += 1; // rebuild counter (diagnostics)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        ref[i] = px[i]; ref[rld + i] = px[ld + i]; ref[2 * rld + i] = px[2 * ld + i];
    }
}

// every evaluation: current exact coordinates into the cell-order records; the rebuild request is consumed here
// (nothing after this kernel reads flags[0])
__global__ void verlet_refresh_kernel(const double *__restrict__ px, int64_t ld, const double *__restrict__ w,
                                      const int *__restrict__ sorted_idx, int n, double4 *__restrict__ sp4,
                                      int *__restrict__ flags, const int *__restrict__ dyn)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) flags[0] = 0;
    if (k >= dyn_loc(dyn, n)) return;
    const int i = sorted_idx[k];
    sp4[k] = make_double4(px[i], px[ld + i], px[2 * ld + i], w ? w[i] : 0.0);
}

// P lanes share one target (P = 1, 2, 4, 8): lane s of the group takes the entries s, s + P, ... and the partial sums
// meet in a butterfly (fixed order).  Small systems with long lists (32,768 water oxygens x 136 entries, 98,304
// charges x 424) otherwise leave most of the machine idle and walk every list serially.
// the targets of one tile of blockDim.x threads
template <int POT, int P>
__device__ __forceinline__ void verlet_force_tile(const CellPairArgs &a, const VerletArgs &v, double scale,
                                                  const double *__restrict__ mass, int mstride,
                                                  const double *__restrict__ charge, int lo, int hi,
                                                  double *__restrict__ acc, int64_t ld, int accumulate, int tile)
{
    const int t = tile * (int)blockDim.x + threadIdx.x;
    const int k = t / P, sub = t % P;
    const int n_loc = dyn_loc(v.dyn, a.n);
    if (v.dyn) hi = min(hi, v.dyn[0]); // slab mode: the own particles are the targets, the ghosts only sources
    const bool valid = k < n_loc;
    const int kk = valid ? k : max(n_loc - 1, 0);
    const int i = a.sorted_idx[kk];
    const bool live = valid && i >= lo && i < hi;
    if (P == 1 && !live) return;
    const double4 pi = a.sp4[kk];
    const int cnt = live ? v.nlist[kk] : 0;
    double f0 = 0.0, f1 = 0.0, f2 = 0.0;
    // The wrap loops of the reference change a component only when |component| >= L/2, and then the UNWRAPPED r2 is at
    // least fl((L/2)^2) > R2 (rounding is monotone).  So: r2 of the raw differences first; inside the cutoff means no
    // component needed wrapping and r2 is the reference's; outside, only an r2 whose high word reaches that of
    // fl((L/2)^2) can belong to a pair that straddles a periodic face (rare) and is wrapped and tested again.
    // Inside = 0 < r2 < R2 on the bit patterns (one unsigned compare of r2 - 1 ulp; r2 = 0 is the sentinel entry of a
    // banked list, the slot itself).
    const unsigned long long R2m1 = (unsigned long long)__double_as_longlong(a.R2) - 1ull;
    auto inside = [&](double r2) { return (unsigned long long)__double_as_longlong(r2) - 1ull < R2m1; };
    auto rewrap = [&](double &rx, double &ry, double &rz, double &r2) { // rare
        if (__double2hiint(r2) >= a.hi_far2) {
            rx = wrap_cubic(rx, a.radius, a.L);
            ry = wrap_cubic(ry, a.radius, a.L);
            rz = wrap_cubic(rz, a.radius, a.L);
            r2 = r2_unfused(rx, ry, rz);
        }
    };
    auto strength = [&](double r2, double w) {
        if (POT == 0) { // (2 s12 - s6) / r2 = 2 (s6 / r2) (s6 - 1/2): the exact factor 2 is applied with the scale
            const double inv = rcp_fast(r2);
            const double qq = a.sigma2 * inv;
            const double s6 = qq * qq * qq;
            return (s6 * inv) * (s6 - 0.5);
        }
        return w_rinv3(r2, w);
    };
    auto pair = [&](const double4 pj) {
        double rx = __dsub_rn(pi.x, pj.x), ry = __dsub_rn(pi.y, pj.y), rz = __dsub_rn(pi.z, pj.z);
        double r2 = r2_unfused(rx, ry, rz);
        if (!inside(r2)) {
            if (__double2hiint(r2) < a.hi_far2) return;
            rewrap(rx, ry, rz, r2);
            if (!inside(r2)) return;
        }
        const double f = strength(r2, pj.w);
        f0 = fma(f, rx, f0);
        f1 = fma(f, ry, f1);
        f2 = fma(f, rz, f2);
    };
    // a batch of four without branches: the four dependency chains (r2 -> reciprocal -> strength) interleave, which is
    // what the kernel lacks with 7 warps per scheduler; entries outside the cutoff (a quarter of the list) add an exact 0.
    // Same operations in the same order as four calls of pair(): bit-identical sums.
    auto pair4 = [&](const double4 q0, const double4 q1, const double4 q2, const double4 q3) {
        double x0 = __dsub_rn(pi.x, q0.x), y0 = __dsub_rn(pi.y, q0.y), z0 = __dsub_rn(pi.z, q0.z);
        double x1 = __dsub_rn(pi.x, q1.x), y1 = __dsub_rn(pi.y, q1.y), z1 = __dsub_rn(pi.z, q1.z);
        double x2 = __dsub_rn(pi.x, q2.x), y2 = __dsub_rn(pi.y, q2.y), z2 = __dsub_rn(pi.z, q2.z);
        double x3 = __dsub_rn(pi.x, q3.x), y3 = __dsub_rn(pi.y, q3.y), z3 = __dsub_rn(pi.z, q3.z);
        double s0 = r2_unfused(x0, y0, z0), s1 = r2_unfused(x1, y1, z1), s2 = r2_unfused(x2, y2, z2), s3 = r2_unfused(x3, y3, z3);
        if (max(max(__double2hiint(s0), __double2hiint(s1)), max(__double2hiint(s2), __double2hiint(s3))) >= a.hi_far2) {
            rewrap(x0, y0, z0, s0); rewrap(x1, y1, z1, s1); rewrap(x2, y2, z2, s2); rewrap(x3, y3, z3, s3);
        }
        double g0 = strength(s0, q0.w), g1 = strength(s1, q1.w), g2 = strength(s2, q2.w), g3 = strength(s3, q3.w);
        g0 = inside(s0) ? g0 : 0.0; g1 = inside(s1) ? g1 : 0.0; g2 = inside(s2) ? g2 : 0.0; g3 = inside(s3) ? g3 : 0.0;
        f0 = fma(g0, x0, f0); f1 = fma(g0, y0, f1); f2 = fma(g0, z0, f2);
        f0 = fma(g1, x1, f0); f1 = fma(g1, y1, f1); f2 = fma(g1, z1, f2);
        f0 = fma(g2, x2, f0); f1 = fma(g2, y2, f1); f2 = fma(g2, z2, f2);
        f0 = fma(g3, x3, f0); f1 = fma(g3, y3, f1); f2 = fma(g3, z3, f2);
    };
    const int *lp = v.list + kk;
    int e = sub;
    // banked lists (verlet_build_slot): the four lanes of a group must read four different record positions at any time.
    // Row r of a block of slot k holds position (r + k) mod 4.  P = 1 (four targets per group, rows walked in step) and
    // P >= 4 (one target per group, lane sub reads row sub) satisfy that as they are; P = 2 (two targets per group: rows
    // {0, 1} then {2, 3}) needs the odd slot one row ahead.
    const int ahead = (v.banked && P == 2) ? (kk & 1) : 0;
    auto at = [&](int q) -> size_t { return (size_t)(P == 2 ? ((q & ~3) | ((q + ahead) & 3)) : q) * v.stride; };
    // four gathers in flight per lane; the list entries of the NEXT batch are fetched before the current one is evaluated,
    // so that a batch costs one memory round trip (the gathers), not two in a row (entries, then gathers).  Measured and
    // dropped in r02 (1,048,576 atoms, ms per step against 0.244 for this loop): the next batch's gathers issued ahead
    // into a second set of landing registers (86 registers) 0.309; eight gathers per batch (96 registers, 5 blocks per
    // SM) 0.283; the next batch's records prefetched into L1 (CCTL.PF1, no registers) 0.293; list rows prefetched into
    // L2 two to eight batches ahead: no change; 10 or 12 blocks per SM (48 / 40 registers, spills) 0.333 / 0.339;
    // persistent CTAs that keep an SM on one contiguous range of tiles (record misses of L1 -58 %) 0.283 -- per-SM
    // time did not move, the static ranges only added imbalance.
    bool have = e + 3 * P < cnt;
    int m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    auto entry = [&](int q) { return ld_stream(lp + at(q)); };
    if (have) { m0 = entry(e); m1 = entry(e + P); m2 = entry(e + 2 * P); m3 = entry(e + 3 * P); }
    while (have) {
        const double4 p0 = load_rec(a.sp4 + m0), p1 = load_rec(a.sp4 + m1), p2 = load_rec(a.sp4 + m2), p3 = load_rec(a.sp4 + m3);
        e += 4 * P;
        have = e + 3 * P < cnt;
        if (have) { m0 = entry(e); m1 = entry(e + P); m2 = entry(e + 2 * P); m3 = entry(e + 3 * P); }
        if (v.branchfree) pair4(p0, p1, p2, p3);
        else { pair(p0); pair(p1); pair(p2); pair(p3); }
    }
    for (; e < cnt; e += P) pair(load_rec(a.sp4 + entry(e)));
    if (P > 1) {
#pragma unroll
        for (int o = 1; o < P; o <<= 1) {
            f0 += __shfl_xor_sync(0xffffffffu, f0, o);
            f1 += __shfl_xor_sync(0xffffffffu, f1, o);
            f2 += __shfl_xor_sync(0xffffffffu, f2, o);
        }
        if (!live || sub != 0) return;
    }
    double coeff = (POT == 0 ? 2.0 * scale : scale) / mass[(size_t)i * mstride];
    if (POT == 1) coeff *= charge[i];
    if (accumulate) {
        acc[i] += coeff * f0; acc[ld + i] += coeff * f1; acc[2 * ld + i] += coeff * f2;
    } else {
        acc[i] = coeff * f0; acc[ld + i] = coeff * f1; acc[2 * ld + i] = coeff * f2;
    }
}

// (The kernel has no barrier and no shared memory: a block is only the unit whose registers are freed together.  Blocks of
// 64 or 32 threads, which let a finished warp make room before the longest list of 128 targets is done, measured 0.2364 /
// 0.2400 ms per step against 0.2402 at 1,048,576 atoms: within the noise.)
template <int POT, int P>
__global__ void __launch_bounds__(128, 8) verlet_force_kernel(const CellPairArgs a, const VerletArgs v, double scale,
                                                              const double *__restrict__ mass, int mstride,
                                                              const double *__restrict__ charge, int lo, int hi,
                                                              double *__restrict__ acc, int64_t ld, int accumulate)
{
    if (v.consume && blockIdx.x == 0 && threadIdx.x == 0) v.flags[0] = 0; // the rebuild request was consumed by the chain before
    if (v.flags[1]) return; // overflow: the scan-per-step kernel takes over
    verlet_force_tile<POT, P>(a, v, scale, mass, mstride, charge, lo, hi, acc, ld, accumulate, blockIdx.x);
}

template <int POT>
static void launch_verlet_force(nbx_ctx *c, const CellPairArgs &a, const VerletArgs &v, double scale, int mstride, int lo,
                                int hi, double *acc_out, int64_t ld_out, int acc_flag)
{
    // lanes per target: enough threads to fill the machine (148 SMs x 2048) when the system is small
    const int64_t want = (int64_t)c->sm_count * 2048;
    // (a slab launches for a capacity bound: the targets are about its share of the box)
    const int64_t ntgt = c->slab.on ? std::max<int64_t>(1, c->slab.n_total / c->slab.nranks) : (int64_t)a.n;
    int P = 1;
    while (P < 8 && ntgt * P < want) P <<= 1;
    if (c->opt_verlet_lanes > 0) P = c->opt_verlet_lanes;
    const unsigned blocks = (unsigned)(((int64_t)a.n * P + 127) / 128);
#define NBX_VF(PP) verlet_force_kernel<POT, PP><<<blocks, 128, 0, c->stream>>>(a, v, scale, c->mass, mstride, c->charge, lo, hi, \
                                                                              acc_out, ld_out, acc_flag)
    if (P == 1) NBX_VF(1);
    else if (P == 2) NBX_VF(2);
    else if (P == 4) NBX_VF(4);
    else NBX_VF(8);
#undef NBX_VF
}

// Position update of velocity Verlet fused with the two O(N) passes the lists need on every evaluation: the
// displacement check against the build-time positions (-> rebuild request) and the refresh of the cell-order record
// (through the inverse permutation slot_of).  Same arithmetic as vv_pos_kernel / verlet_check_kernel /
// verlet_refresh_kernel; saves two launches and re-reading the positions twice.
__global__ void vv_pos_lists_kernel(double *__restrict__ pos, const double *__restrict__ vel, const double *__restrict__ acc,
                                    const double *__restrict__ w, int64_t ld, int n, double dt, double hdt2,
                                    const double *__restrict__ ref, int64_t rld, double lim2, const int *__restrict__ slot_of,
                                    double4 *__restrict__ sp4, int *__restrict__ flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && flags[1]) flags[0] = 1;
    if (i >= n) return;
    const double x = fma(hdt2, acc[i], fma(dt, vel[i], pos[i]));
    const double y = fma(hdt2, acc[ld + i], fma(dt, vel[ld + i], pos[ld + i]));
    const double z = fma(hdt2, acc[2 * ld + i], fma(dt, vel[2 * ld + i], pos[2 * ld + i]));
    pos[i] = x; pos[ld + i] = y; pos[2 * ld + i] = z;
    const double dx = x - ref[i], dy = y - ref[rld + i], dz = z - ref[2 * rld + i];
    const double d2 = dx * dx + dy * dy + dz * dz;
    if (!(d2 <= lim2)) flags[0] = 1; // also catches NaN
    sp4[slot_of[i]] = make_double4(x, y, z, w ? w[i] : 0.0);
}

// true when the next position update of nbx_step_vv can run as vv_pos_lists_kernel for this list
bool lists_can_fuse_update(const nbx_ctx *c, const CellList *cl, const double *px)
{
    return c->opt_fuse_update && cl->v_valid && cl->v_px == px && cl->v_n == c->n && cl->slot_of && cl->v_ref && !c->slab.on &&
           c->dyn == nullptr && c->thermo != NBX_THERMO_NOSEHOOVER;
}

int launch_vv_pos_lists(nbx_ctx *c, CellList *cl, const double *w, double dt)
{
    const int n = (int)c->n;
    const double lim = 0.5 * cl->v_skin * (1.0 - 1e-9);
    timer_begin(c, NBX_T_INTEGRATE);
    vv_pos_lists_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->pos, c->vel, c->acc, w, c->npad, n, dt, 0.5 * dt * dt, cl->v_ref,
                                                               cl->cap_n, lim * lim, cl->slot_of, cl->sp4, cl->v_flags);
    timer_end(c, NBX_T_INTEGRATE);
    NBX_CUDA(c, cudaGetLastError());
    cl->prechecked = true;
    return NBX_OK;
}

// ------------------------------------------------------------------------------------------------
// the rebuild chain as the body of a graph IF node (while nbx_step_vv captures)
// ------------------------------------------------------------------------------------------------
__global__ void cond_set_kernel(cudaGraphConditionalHandle h, const int *__restrict__ flag)
{
    cudaGraphSetConditional(h, flag[0] != 0 ? 1u : 0u);
}

// While a graph with IF nodes is being captured: a conditional handle for a kernel of the caller's to set
// (cudaGraphSetConditional) before cond_scope_begin(..., pre) adds the node.  *ok = false otherwise.
int cond_handle_create(nbx_ctx *c, cudaGraphConditionalHandle *h, bool *ok)
{
    *ok = false;
    if (!c->cond_capture || c->cond_fail || !c->aux_stream) return NBX_OK;
    cudaStreamCaptureStatus st;
    unsigned long long id;
    cudaGraph_t g = nullptr;
    const cudaGraphNode_t *deps = nullptr;
    size_t nd = 0;
    if (cudaStreamGetCaptureInfo_v2(c->stream, &st, &id, &g, &deps, &nd) != cudaSuccess || st != cudaStreamCaptureStatusActive || !g)
        return NBX_OK;
    if (cudaGraphConditionalHandleCreate(h, g, 0, cudaGraphCondAssignDefault) != cudaSuccess) { c->cond_fail = true; cudaGetLastError(); return NBX_OK; }
    *ok = true;
    return NBX_OK;
}

// From here to cond_scope_end the launches on c->stream land in the IF node's body graph.  No-op outside a capture.
// pre: a handle from cond_handle_create that a kernel of the caller's has set already (else a one-thread kernel sets it
// from flag[0] here).
int cond_scope_begin(nbx_ctx *c, const int *flag, CondScope *sc, const cudaGraphConditionalHandle *pre)
{
    sc->active = false;
    if (!c->cond_capture || c->cond_fail || !c->aux_stream) return NBX_OK;
    cudaStreamCaptureStatus st;
    unsigned long long id;
    cudaGraph_t g = nullptr;
    const cudaGraphNode_t *deps = nullptr;
    size_t nd = 0;
    if (cudaStreamGetCaptureInfo_v2(c->stream, &st, &id, &g, &deps, &nd) != cudaSuccess || st != cudaStreamCaptureStatusActive || !g)
        return NBX_OK;
    cudaGraphConditionalHandle h;
    if (pre) h = *pre;
    else {
        if (cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault) != cudaSuccess) { c->cond_fail = true; cudaGetLastError(); return NBX_OK; }
        cond_set_kernel<<<1, 1, 0, c->stream>>>(h, flag);
    }
    if (cudaStreamGetCaptureInfo_v2(c->stream, &st, &id, &g, &deps, &nd) != cudaSuccess) { c->cond_fail = true; return fail(c, NBX_ERR_CUDA, "graph IF node: capture info"); }
    cudaGraphNodeParams p = {};
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = h;
    p.conditional.type = cudaGraphCondTypeIf;
    p.conditional.size = 1;
    cudaGraphNode_t node;
    if (cudaGraphAddNode(&node, g, deps, nd, &p) != cudaSuccess) { c->cond_fail = true; cudaGetLastError(); return fail(c, NBX_ERR_CUDA, "graph IF node: cudaGraphAddNode"); }
    if (cudaStreamUpdateCaptureDependencies(c->stream, &node, 1, cudaStreamSetCaptureDependencies) != cudaSuccess) {
        c->cond_fail = true; cudaGetLastError();
        return fail(c, NBX_ERR_CUDA, "graph IF node: cudaStreamUpdateCaptureDependencies");
    }
    if (cudaStreamBeginCaptureToGraph(c->aux_stream, p.conditional.phGraph_out[0], nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        c->cond_fail = true; cudaGetLastError();
        return fail(c, NBX_ERR_CUDA, "graph IF node: cudaStreamBeginCaptureToGraph");
    }
    sc->saved = c->stream;
    c->stream = c->aux_stream;
    sc->active = true;
    return NBX_OK;
}

int cond_scope_end(nbx_ctx *c, CondScope *sc)
{
    if (!sc->active) return NBX_OK;
    sc->active = false;
    cudaGraph_t body = nullptr;
    const cudaError_t e = cudaStreamEndCapture(c->stream, &body);
    c->stream = sc->saved;
    if (e != cudaSuccess) { c->cond_fail = true; cudaGetLastError(); return fail(c, NBX_ERR_CUDA, "graph IF node: body capture"); }
    return NBX_OK;
}

static CellPairArgs make_args(const nbx_ctx *c, const CellList *cl, double R2)
{
    CellPairArgs a{};
    a.sp4 = cl->sp4; a.sl4 = cl->sl4;
    a.sorted_idx = cl->sorted_idx; a.scell = cl->scell; a.start = cl->start;
    a.n = (int)cl->n; a.nc = cl->grid.nc[0];
    a.L = c->bc[0]; a.radius = 0.5 * c->bc[0]; a.R2 = R2; a.sigma2 = c->lj_sigma2;
    // fp32 prefilter: coordinates in cell units lie in [0, nc] (|rounding| <= nc 2^-25 each); a component of
    // the displacement is off by < 4 nc 2^-24, so for r ~ R <= 1 cell the squared distance is off by
    // < 2 sqrt(3) * 4 nc 2^-24 + O(2^-22) relative.  The margin doubles that.
    const double cell = a.L / (double)a.nc;
    const double margin = 32.0 * (double)a.nc * 5.9604644775390625e-8 + 1e-5;
    a.R2f = nextafterf((float)(R2 / (cell * cell) * (1.0 + margin)), INFINITY);
    int64_t bits;
    memcpy(&bits, &a.radius, sizeof bits);
    a.hi_radius = (int)(bits >> 32);
    const double far2 = a.radius * a.radius;
    memcpy(&bits, &far2, sizeof bits);
    a.hi_far2 = (int)(bits >> 32);
    return a;
}

// pot 0: LJ (self exclusion), 1: Coulomb (self exclusion), 2: Coulomb (own-molecule exclusion)
int launch_cells_force(nbx_ctx *c, CellList *cl, int pot, int64_t lo, int64_t hi, int mstride, double *acc_out,
                       int64_t ld_out, bool accumulate, const int *cond)
{
    const int n = (int)cl->n;
    if (n == 0) return NBX_OK;
    int blocks = (n + 127) / 128;
    if (cond && blocks > c->sm_count * 8) blocks = c->sm_count * 8; // conditional (fallback) launch: block-stride, cheap when off
    const int acc_flag = accumulate ? 1 : 0;
    timer_begin(c, NBX_T_PAIR_CELLS);
    if (c->opt_prefilter) {
        if (pot == 0) {
            const CellPairArgs a = make_args(c, cl, c->lj_R2);
            cell_pairs2_kernel<0, 0><<<blocks, 128, 0, c->stream>>>(a, 24.0 * c->lj_eps, c->mass, mstride, c->charge,
                                                                  (int)lo, (int)hi, acc_out, ld_out, acc_flag, nullptr,
                                                                  nullptr, nullptr, c->dyn, cond);
        } else { // the exclusion (self / own molecule) was fixed when the list was built (key_div)
            const CellPairArgs a = make_args(c, cl, c->el_R2);
            cell_pairs2_kernel<1, 0><<<blocks, 128, 0, c->stream>>>(a, c->el_k, c->mass, 1, c->charge, (int)lo, (int)hi,
                                                                  acc_out, ld_out, acc_flag, nullptr, nullptr, nullptr, c->dyn, cond);
        }
    } else if (pot == 0) {
        const CellPairArgs a = make_args(c, cl, c->lj_R2);
        cell_force_kernel<0, 0><<<blocks, 128, 0, c->stream>>>(a, 24.0 * c->lj_eps, c->mass, mstride, c->charge,
                                                             (int)lo, (int)hi, acc_out, ld_out, acc_flag, c->dyn);
    } else if (pot == 1) {
        const CellPairArgs a = make_args(c, cl, c->el_R2);
        cell_force_kernel<1, 0><<<blocks, 128, 0, c->stream>>>(a, c->el_k, c->mass, 1, c->charge, (int)lo, (int)hi,
                                                             acc_out, ld_out, acc_flag, c->dyn);
    } else {
        const CellPairArgs a = make_args(c, cl, c->el_R2);
        cell_force_kernel<1, 1><<<blocks, 128, 0, c->stream>>>(a, c->el_k, c->mass, 1, c->charge, (int)lo, (int)hi,
                                                             acc_out, ld_out, acc_flag, c->dyn);
    }
    timer_end(c, NBX_T_PAIR_CELLS);
    NBX_CUDA(c, cudaGetLastError());
    return NBX_OK;
}

// ------------------------------------------------------------------------------------------------
// one cutoff potential over a cell list: plan -> (re)build -> forces.  *used = false when the box cannot be
// cell-listed for this cutoff (the caller falls back to the all-pairs kernel).
// ------------------------------------------------------------------------------------------------
int cells_pairs(nbx_ctx *c, CellList *cl, double R, int pot, const double *px, const double *w, const int *gid, int64_t n,
                int64_t nplan, int64_t ld, int key_div, int64_t lo, int64_t hi, int mstride, double *acc_out,
                int64_t ld_out, bool accumulate, bool *used)
{
    *used = false;
    const double R2 = R * R;
    // Verlet lists need stable slots: the whole system in one context, or a slab that renumbers only at collective
    // rebuilds (SlabState::verlet: the host layer decides when, nbx_slab.cu)
    const bool slabv = c->slab.on && c->slab.verlet;
    bool verlet = c->opt_verlet_permille > 0 && c->opt_prefilter && ((!c->slab.on && c->dyn == nullptr) || slabv);
    const double skin = verlet ? R * 1e-3 * (double)c->opt_verlet_permille : 0.0;
    if (verlet) {
        NBX_TRY(cells_plan(c, R + skin, nplan, &cl->grid));
        if (!cl->grid.valid) verlet = false;
    }
    if (!verlet) {
        NBX_TRY(cells_plan(c, R, nplan, &cl->grid));
        if (!cl->grid.valid) return NBX_OK;
        *used = true;
        cl->v_valid = false;
        NBX_TRY(cells_build(c, cl, px, w, gid, n, ld, key_div, nullptr));
        return launch_cells_force(c, cl, pot, lo, hi, mstride, acc_out, ld_out, accumulate, nullptr);
    }
    *used = true;
    const int ni = (int)n;
    NBX_TRY(ensure_cells(c, cl, n, cl->grid.ncell));
    // list capacity from the mean density: 1.5 x the expected partners within R + skin (+ slack)
    const double L = c->bc[0];
    const double expect = (double)nplan / (L * L * L) * 4.18879020478639 * (R + skin) * (R + skin) * (R + skin);
    int cap = (int)(1.5 * expect) + 24;
    if (cap > ni - 1) cap = ni > 1 ? ni - 1 : 1;
    cap = (cap + 3) & ~3; // whole blocks of four (banked layout)
    // (the position of a record is its LOCAL slot number mod 4: a slab would order its lists differently from the single
    // context and lose the bit-for-bit equality of the two, for a few per cent of a kernel that is a third of its step)
    const int banked = (c->opt_verlet_banked && !c->slab.on && cap < 4 * 65535) ? 1 : 0; // (16-bit counters in the build)
    const bool same = cl->v_valid && cl->v_n == n && (slabv || cl->v_px == px) && cl->v_R == R && cl->v_skin == skin && cl->v_L == L &&
                      cl->v_key_div == key_div && cl->v_nc == cl->grid.nc[0] && cl->v_cap == cap && cl->v_banked == banked;
    if (!same) {
        const int64_t rows = banked ? 2 * (int64_t)cap : (int64_t)cap; // (banked: the build places entries beyond the final length first)
        if (cl->v_cap_alloc < rows * cl->cap_n) {
            NBX_TRY(dev_alloc(c, &cl->v_list, (size_t)rows * (size_t)cl->cap_n));
            cl->v_cap_alloc = rows * cl->cap_n;
        }
        if (cl->v_ref_n < cl->cap_n) {
            NBX_TRY(dev_alloc(c, &cl->v_ref, (size_t)3 * (size_t)cl->cap_n));
            NBX_TRY(dev_alloc(c, &cl->v_nlist, (size_t)cl->cap_n));
            cl->v_ref_n = cl->cap_n;
        }
        if (!cl->v_flags) NBX_TRY(dev_alloc(c, &cl->v_flags, (size_t)4));
        NBX_CUDA(c, cudaMemsetAsync(cl->v_flags, 0, sizeof(int) * 4, c->stream));
        NBX_CUDA(c, cudaMemsetAsync(cl->v_flags, 1, sizeof(int), c->stream)); // [0] != 0: build now
        NBX_CUDA(c, cudaMemsetAsync(cl->v_ref, 0, sizeof(double) * 3 * (size_t)cl->cap_n, c->stream));
        cl->v_valid = true; cl->v_n = n; cl->v_px = px; cl->v_R = R; cl->v_skin = skin; cl->v_L = L;
        cl->v_key_div = key_div; cl->v_nc = cl->grid.nc[0]; cl->v_cap = cap; cl->v_banked = banked;
        if (slabv) c->slab.rebuild_now = true;
    }
    const int blocks256 = (ni + 255) / 256;
    if (slabv) {
        // the host layer decided (collectively, SlabStepper) whether this evaluation follows a migration round: then the
        // local numbering is new and cells + lists are rebuilt here, unconditionally; otherwise nothing is even launched
        CellPairArgs a = make_args(c, cl, (R + skin) * (R + skin));
        VerletArgs v{};
        v.list = cl->v_list; v.nlist = cl->v_nlist; v.flags = cl->v_flags; v.cap = cap; v.stride = cl->cap_n; v.dyn = c->dyn;
        v.banked = banked; v.branchfree = c->opt_verlet_branchfree;
        // phase 1 / 2 (slab_enqueue): the rebuild chain and the evaluation are enqueued separately, and whether the chain
        // RUNS is decided on the device (slab.cond == v_flags: every kernel of it returns at once unless flags[0] is set)
        if (c->slab.phase != 2 && c->slab.rebuild_now) {
            if (!c->slab.cond) NBX_CUDA(c, cudaMemsetAsync(cl->v_flags, 1, sizeof(int), c->stream));
            NBX_TRY(cells_build(c, cl, px, w, gid, n, ld, key_div, cl->v_flags));
            timer_begin(c, NBX_T_CELL_BUILD);
            verlet_build_kernel<<<map_grid(c, ni, 128), 128, 0, c->stream>>>(a, v);
            verlet_ref_kernel<<<map_grid(c, ni, 256), 256, 0, c->stream>>>(px, ld, cl->v_ref, cl->cap_n, ni, cl->v_flags, c->dyn);
            timer_end(c, NBX_T_CELL_BUILD);
            c->slab.rebuild_now = false;
        }
        if (c->slab.phase == 1) { NBX_CUDA(c, cudaGetLastError()); return NBX_OK; }
        if (!c->slab.records_fresh) { // (slab_enqueue: the position update and the halo receive already wrote the records)
            timer_begin(c, NBX_T_CELL_BUILD);
            verlet_refresh_kernel<<<blocks256, 256, 0, c->stream>>>(px, ld, w, cl->sorted_idx, ni, cl->sp4, cl->v_flags, c->dyn);
            timer_end(c, NBX_T_CELL_BUILD);
        }
        c->slab.records_fresh = false;
        a = make_args(c, cl, R2);
        const int acc_flag = accumulate ? 1 : 0;
        timer_begin(c, NBX_T_PAIR_CELLS);
        if (pot == 0) launch_verlet_force<0>(c, a, v, 24.0 * c->lj_eps, mstride, (int)lo, (int)hi, acc_out, ld_out, acc_flag);
        else launch_verlet_force<1>(c, a, v, c->el_k, 1, (int)lo, (int)hi, acc_out, ld_out, acc_flag);
        timer_end(c, NBX_T_PAIR_CELLS);
        NBX_CUDA(c, cudaGetLastError());
        return NBX_OK; // a list overflow is reported by nbx_slab_verlet_check (no scan fallback on stale cells)
    }
    const double lim = 0.5 * skin * (1.0 - 1e-9);
    // the position update already checked the displacements and refreshed the records (vv_pos_lists_kernel)
    const bool pre = cl->prechecked && same;
    cl->prechecked = false;
    if (!pre) {
        timer_begin(c, NBX_T_CELL_BUILD);
        verlet_check_kernel<<<blocks256, 256, 0, c->stream>>>(px, ld, cl->v_ref, cl->cap_n, ni, lim * lim, cl->v_flags);
        timer_end(c, NBX_T_CELL_BUILD);
    }
    CondScope scope;
    NBX_TRY(cond_scope_begin(c, cl->v_flags, &scope, nullptr)); // (graph capture: the chain below becomes the body of an IF node)
    {
        const int rc = cells_build(c, cl, px, w, gid, n, ld, key_div, cl->v_flags);
        if (rc != NBX_OK) { cond_scope_end(c, &scope); return rc; }
    }
    CellPairArgs a = make_args(c, cl, (R + skin) * (R + skin)); // scan threshold of the list build
    VerletArgs v{};
    v.list = cl->v_list; v.nlist = cl->v_nlist; v.flags = cl->v_flags; v.cap = cap; v.stride = cl->cap_n; v.dyn = nullptr;
    v.banked = banked; v.branchfree = c->opt_verlet_branchfree;
    v.consume = pre ? 1 : 0;
    timer_begin(c, NBX_T_CELL_BUILD);
    verlet_build_kernel<<<map_grid(c, ni, 128), 128, 0, c->stream>>>(a, v);
    verlet_ref_kernel<<<map_grid(c, ni, 256), 256, 0, c->stream>>>(px, ld, cl->v_ref, cl->cap_n, ni, cl->v_flags, nullptr);
    NBX_TRY(cond_scope_end(c, &scope));
    if (!v.consume) verlet_refresh_kernel<<<blocks256, 256, 0, c->stream>>>(px, ld, w, cl->sorted_idx, ni, cl->sp4, cl->v_flags, nullptr);
    timer_end(c, NBX_T_CELL_BUILD);
    a = make_args(c, cl, R2);
    const int acc_flag = accumulate ? 1 : 0;
    timer_begin(c, NBX_T_PAIR_CELLS);
    if (pot == 0) launch_verlet_force<0>(c, a, v, 24.0 * c->lj_eps, mstride, (int)lo, (int)hi, acc_out, ld_out, acc_flag);
    else launch_verlet_force<1>(c, a, v, c->el_k, 1, (int)lo, (int)hi, acc_out, ld_out, acc_flag);
    timer_end(c, NBX_T_PAIR_CELLS);
    NBX_CUDA(c, cudaGetLastError());
    // fallback when a list overflowed (dense clusters): the cells were rebuilt above (flags[1] forces it)
    return launch_cells_force(c, cl, pot, lo, hi, mstride, acc_out, ld_out, accumulate, cl->v_flags + 1);
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
        cl->v_valid = false; // the cells are rebuilt on this call's own grid
        int rc = cells_build(c, cl, px, nullptr, nullptr, n, ld, 1, nullptr);
        if (rc != NBX_OK) { cudaFree(d_counts); return rc; }
        a = make_args(c, cl, R2);
        if (c->opt_prefilter)
            cell_pairs2_kernel<0, 1><<<blocks, 128, 0, c->stream>>>(a, 0.0, nullptr, 1, nullptr, 0, ni, nullptr, 0, 0, d_counts,
                                                                  nullptr, nullptr, nullptr, nullptr);
        else
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
            if (use_cells && c->opt_prefilter)
                cell_pairs2_kernel<0, 2><<<blocks, 128, 0, c->stream>>>(a, 0.0, nullptr, 1, nullptr, 0, ni, nullptr, 0, 0,
                                                                      d_counts, d_off, d_list, nullptr, nullptr);
            else if (use_cells)
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
    cudaFree(cl->cell_of); cudaFree(cl->count); cudaFree(cl->start); cudaFree(cl->sums);
    cudaFree(cl->arrival); cudaFree(cl->tmp_idx); cudaFree(cl->tmp_key);
    cudaFree(cl->v_list); cudaFree(cl->v_nlist); cudaFree(cl->v_ref); cudaFree(cl->v_flags);
    cudaFree(cl->sorted_idx); cudaFree(cl->scell); cudaFree(cl->sp4); cudaFree(cl->sl4); cudaFree(cl->slot_of);
    *cl = CellList{};
}

// CUDA loads kernels lazily, at their first launch, and loading may synchronise the context: a kernel that is first
// launched while another member's kernel spins on a flag would deadlock the pair (CUDA programming guide, "Lazy
// Loading": concurrent execution).  Every kernel of the distributed loops is therefore loaded when a context is created.
void preload_cells()
{
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, cell_id_kernel);
    cudaFuncGetAttributes(&a, scan_block_kernel);
    cudaFuncGetAttributes(&a, scan_sums_kernel);
    cudaFuncGetAttributes(&a, scan_add_kernel);
    cudaFuncGetAttributes(&a, scatter_kernel);
    cudaFuncGetAttributes(&a, rank_gather_kernel);
    cudaFuncGetAttributes(&a, cell_pairs2_kernel<0, 0>);
    cudaFuncGetAttributes(&a, cell_pairs2_kernel<1, 0>);
    cudaFuncGetAttributes(&a, verlet_check_kernel);
    cudaFuncGetAttributes(&a, verlet_build_kernel);
    cudaFuncGetAttributes(&a, verlet_ref_kernel);
    cudaFuncGetAttributes(&a, verlet_refresh_kernel);
    cudaFuncGetAttributes(&a, verlet_force_kernel<0, 1>);
    cudaFuncGetAttributes(&a, verlet_force_kernel<0, 2>);
    cudaFuncGetAttributes(&a, verlet_force_kernel<0, 4>);
    cudaFuncGetAttributes(&a, verlet_force_kernel<0, 8>);
    cudaFuncGetAttributes(&a, verlet_force_kernel<1, 1>);
    cudaFuncGetAttributes(&a, verlet_force_kernel<1, 2>);
    cudaFuncGetAttributes(&a, verlet_force_kernel<1, 4>);
    cudaFuncGetAttributes(&a, verlet_force_kernel<1, 8>);
    cudaFuncGetAttributes(&a, vv_pos_lists_kernel);
    cudaFuncGetAttributes(&a, cond_set_kernel);
    cudaGetLastError();
}

} // namespace nbx
