// nbx_fused.cu -- one kernel per velocity-Verlet step for a single cutoff potential (sm_100a).
//
// nbx_step_vv on a box with one cutoff pair potential (Lennard-Jones, src/basic_potentials.jl:240-272, or
// Coulomb with a finite cutoff, :274-304), cubic periodic boundary, no thermostat or Berendsen
// (src/thermostats.jl:76-83): BASELINE config 3.  The unfused path spends a third of the step in O(N)
// kernels around the force kernel (position update, velocity update, temperature sum, displacement check,
// refresh of the cell-order records) and in launches of the rebuild chain that return at once.  Here
//
//  * the state lives in CELL ORDER between list builds ("slots"; a cell's slot count is padded to a
//    multiple of the cluster size C): positions + charge (double4, double buffered), velocity + mass
//    (double4), acceleration and build-time positions (SoA rows).  Every access of the step kernel except
//    the neighbour gathers is coalesced; particle order is restored when the run ends or a list is rebuilt.
//  * CLUSTER lists: the C slots of a cluster (same cell) share ONE neighbour list -- the union of the
//    slots within R + skin of any member -- dealt round-robin to the C lanes.  A lane gathers each of its
//    entries once (one 256-bit load) and applies the reference's exact predicate and force to all C
//    targets; the C lanes' partial sums meet in a butterfly.  Per target that is 27 gathers instead of 51
//    at C = 4 for twice the FP64 work: the unfused list kernel is bound by the L1 gather rate with the
//    FP64 pipe at 34 %, so the trade is the right way round.
//  * fz_step_kernel, step j: a_j from x_j (listed pairs, exact predicate: `ri - rj` on the unwrapped
//    coordinates, the wrap loops, un-fused r2, strict <), + Berendsen term from v_{j-1}; v_j = v_{j-1} +
//    dt/2 (a_{j-1} + a_j); x_{j+1} = x_j + dt v_j + dt^2/2 a_j into the other position buffer; displacement
//    of x_{j+1} from the build-time position > skin/2 raises the rebuild request of step j+1; block
//    partials of sum m v_j^2, summed in block order by the last block to finish (the next temperature).
//  * the rebuild chain (restore particle order, bin, padded scan, place, cluster lists) is enqueued every
//    step and returns at once unless the request flag is up; two steps form a CUDA graph.
//
// The in-cutoff pair set is the reference's by construction (the list is a superset while no particle has
// moved more than skin/2; the predicate decides).  A list overflow freezes the state at the last complete
// step and hands the run back to the unfused path.
#include "nbx_internal.cuh"

#include <cmath>
#include <cstring>

namespace nbx {

enum { FZ_REQ0 = 0, FZ_REQ1 = 1, FZ_OVER = 2, FZ_REBUILDS = 3, FZ_STEPS = 4, FZ_DONE = 5, FZ_NSLOTS = 6, FZ_NFLAGS = 16 };

struct FzArgs {
    const double4 *xr;  // positions this step reads (all slots)
    double4 *xw;        // positions of the next step
    double4 *xa, *xb;   // both buffers (final restore picks by the device step counter)
    double4 *vm;
    double *sa, *ref;
    float4 *sl;
    int *pid, *scell, *cell_of, *arrival, *tmp_idx, *count, *start, *list, *nlist, *flags;
    double *partial, *scal; // partial: [cap_slots / 128] block sums, then the group sums
    int *gcount;            // arrivals per group of blocks
    double *pos, *vel, *acc;          // particle order, row stride ld
    const double *mass, *charge;
    int64_t ld, cap_slots;
    int n, nc, ncell, cap_e;
    double L, radius, R2, sigma2, scale;
    float R2f;      // prefilter threshold of the list build, (R + skin)^2 in cell units with the fp32 margin
    int hi_radius;  // high word of L/2: a displacement component with a smaller high word needs no wrapping
    double dt, hdt, hdt2, lim2;
    double kB_ndf, T0, gamma;
    int req;        // flag index of the rebuild request this step consumes; the step kernel raises the other one
    int nsteps;
};

constexpr unsigned FULL = 0xffffffffu;
constexpr int FZ_GROUP = 64; // blocks per group of the temperature sum

// rebuild chain guard: nothing to do without a request, and nothing more to do once a list has overflowed
__device__ __forceinline__ bool fz_skip(const FzArgs &a) { return !a.flags[a.req] || a.flags[FZ_OVER]; }

// ------------------------------------------------------------------------------------------------
// rebuild chain (every kernel returns at once unless the request flag is up)
// ------------------------------------------------------------------------------------------------
// slots -> particle order (the state the binning reads), and zero the cell histogram
__global__ void fz_unsort_kernel(const FzArgs a, int conditional, int final_pick)
{
    if (conditional && fz_skip(a)) return;
    const int nslots = a.flags[FZ_NSLOTS];
    const double4 *x = a.xr;
    if (final_pick) { // end of a run: x_j of the last complete step j sits in buffer j & 1, x_{j+1} of a frozen run in (j+1) & 1
        const int sd = a.flags[FZ_STEPS];
        const int b = sd == a.nsteps ? (a.nsteps & 1) : ((sd + 1) & 1);
        x = b ? a.xb : a.xa;
    }
    const int stride = gridDim.x * blockDim.x;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nslots; k += stride) {
        const int i = a.pid[k];
        if (i < 0) continue;
        const double4 p = x[k], v = a.vm[k];
        a.pos[i] = p.x; a.pos[a.ld + i] = p.y; a.pos[2 * a.ld + i] = p.z;
        a.vel[i] = v.x; a.vel[a.ld + i] = v.y; a.vel[2 * a.ld + i] = v.z;
        a.acc[i] = a.sa[k]; a.acc[a.ld + i] = a.sa[a.cap_slots + k]; a.acc[2 * a.ld + i] = a.sa[2 * a.cap_slots + k];
    }
    if (!final_pick)
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c <= a.ncell; c += stride) a.count[c] = 0;
}

__global__ void fz_count_kernel(const FzArgs a)
{
    if (fz_skip(a)) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
        const int cx = cell_coord(a.pos[i], a.L, a.nc), cy = cell_coord(a.pos[a.ld + i], a.L, a.nc),
                  cz = cell_coord(a.pos[2 * a.ld + i], a.L, a.nc);
        const int cid = (cz * a.nc + cy) * a.nc + cx; // x fastest: the three x-neighbours of a cell are contiguous
        a.cell_of[i] = cid;
        a.arrival[i] = atomicAdd(&a.count[cid], 1);
    }
}

__global__ void fz_scatter_kernel(const FzArgs a)
{
    if (fz_skip(a)) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x)
        a.tmp_idx[a.start[a.cell_of[i]] + a.arrival[i]] = i;
}

// Final slot = padded cell start + rank of the particle index among the cell's members (deterministic order);
// the same thread lays the particle's state down in slot order.  The cell's last member also fills the padding
// slots: never a neighbour (prefilter record far away), as a target a copy of its own position (results unused).
template <int C>
__global__ void fz_place_kernel(const FzArgs a)
{
    if (fz_skip(a)) return;
    const double s = (double)a.nc / a.L;
    double4 *x = const_cast<double4 *>(a.xr);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.flags[FZ_NSLOTS] = a.start[a.ncell];
        a.flags[FZ_REBUILDS] += 1;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
        const int cid = a.cell_of[i];
        const int b = a.start[cid], cnt = a.count[cid];
        int rank = 0;
        for (int m = b; m < b + cnt; ++m) rank += a.tmp_idx[m] < i ? 1 : 0;
        const int dst = b + rank;
        const double px = a.pos[i], py = a.pos[a.ld + i], pz = a.pos[2 * a.ld + i];
        const double4 rec = make_double4(px, py, pz, a.charge ? a.charge[i] : 0.0);
        x[dst] = rec;
        a.vm[dst] = make_double4(a.vel[i], a.vel[a.ld + i], a.vel[2 * a.ld + i], a.mass[i]);
        a.sa[dst] = a.acc[i]; a.sa[a.cap_slots + dst] = a.acc[a.ld + i]; a.sa[2 * a.cap_slots + dst] = a.acc[2 * a.ld + i];
        a.ref[dst] = px; a.ref[a.cap_slots + dst] = py; a.ref[2 * a.cap_slots + dst] = pz;
        a.sl[dst] = make_float4((float)(wrapped_coord(px, a.L) * s), (float)(wrapped_coord(py, a.L) * s),
                                (float)(wrapped_coord(pz, a.L) * s), __int_as_float(i));
        a.pid[dst] = i;
        a.scell[dst] = cid;
        if (rank == cnt - 1) {
            const int e = b + (cnt + C - 1) / C * C;
            for (int d = b + cnt; d < e; ++d) {
                x[d] = rec;
                a.vm[d] = make_double4(0.0, 0.0, 0.0, 1.0);
                a.sa[d] = 0.0; a.sa[a.cap_slots + d] = 0.0; a.sa[2 * a.cap_slots + d] = 0.0;
                a.ref[d] = px; a.ref[a.cap_slots + d] = py; a.ref[2 * a.cap_slots + d] = pz;
                a.sl[d] = make_float4(1.0e9f, 1.0e9f, 1.0e9f, __int_as_float(-2 - d));
                a.pid[d] = -1;
                a.scell[d] = cid;
            }
        }
    }
}

// Cluster lists.  The C lanes of a cluster walk the 27 surrounding cells together (9 x-rows: one contiguous
// slot range + at most one periodic wrap-around cell); lane c tests the candidates b + c, b + c + C, ... against
// ALL C members in fp32 (wrapped cell-unit coordinates, threshold with the margin of make_args in nbx_cells.cu:
// it can only over-accept) and keeps those within reach of any member.  A candidate is thus tested once per
// cluster and the lanes' lists come out balanced.
template <int C>
__global__ void __launch_bounds__(128) fz_build_kernel(const FzArgs a)
{
    if (!a.flags[a.req]) return;
    const int nslots = a.start[a.ncell];
    const int lane = threadIdx.x & 31, gl = lane & ~(C - 1), c0 = lane & (C - 1);
    const int nc = a.nc;
    const float fnc = (float)nc;
    bool over = false;
    for (int base = blockIdx.x * 128; base < nslots; base += gridDim.x * 128) {
        const int k = base + threadIdx.x;
        const bool active = k < nslots; // a cluster is active or inactive as a whole (nslots is a multiple of C)
        const int kk = active ? k : nslots - 1;
        const float4 me = a.sl[kk];
        const int cid = a.scell[kk];
        float mx[C], my[C], mz[C];
        int mk[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            mx[c] = __shfl_sync(FULL, me.x, gl + c);
            my[c] = __shfl_sync(FULL, me.y, gl + c);
            mz[c] = __shfl_sync(FULL, me.z, gl + c);
            mk[c] = __shfl_sync(FULL, __float_as_int(me.w), gl + c);
        }
        const int cx = cid % nc, cy = (cid / nc) % nc, cz = cid / (nc * nc);
        int cnt = 0;
        auto scan = [&](int b, int e, float sx, float sy, float sz) {
            float tx[C], ty[C], tz[C];
#pragma unroll
            for (int c = 0; c < C; ++c) { tx[c] = mx[c] + sx; ty[c] = my[c] + sy; tz[c] = mz[c] + sz; }
            for (int m = b + c0; m < e; m += C) {
                const float4 cj = __ldg(&a.sl[m]);
                const int kj = __float_as_int(cj.w);
                if (kj < 0) continue; // padding slot (all of them sit at the same far point: test the key, not the distance)
                bool pass = false;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float dx = tx[c] - cj.x, dy = ty[c] - cj.y, dz = tz[c] - cj.z;
                    const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    pass = pass || (r2 < a.R2f && kj != mk[c] && mk[c] >= 0);
                }
                if (pass) {
                    if (cnt < a.cap_e) a.list[(size_t)cnt * a.cap_slots + k] = m;
                    else over = true;
                    ++cnt;
                }
            }
        };
        if (active) {
#pragma unroll 1
            for (int r = 0; r < 9; ++r) {
                const int dz = r / 3 - 1, dy = r - (r / 3) * 3 - 1;
                int z = cz + dz, y = cy + dy;
                float sz = 0.f, sy = 0.f;
                // a neighbour cell reached through a periodic face holds the image w_j -+ nc of its particles
                if (z < 0) { z += nc; sz = fnc; } else if (z >= nc) { z -= nc; sz = -fnc; }
                if (y < 0) { y += nc; sy = fnc; } else if (y >= nc) { y -= nc; sy = -fnc; }
                const int row = (z * nc + y) * nc;
                const int xa = cx > 0 ? cx - 1 : 0, xb = cx < nc - 1 ? cx + 1 : nc - 1;
                scan(a.start[row + xa], a.start[row + xb + 1], 0.f, sy, sz);
                if (cx == 0) scan(a.start[row + nc - 1], a.start[row + nc], fnc, sy, sz);
                else if (cx == nc - 1) scan(a.start[row], a.start[row + 1], -fnc, sy, sz);
            }
            a.nlist[k] = cnt < a.cap_e ? cnt : a.cap_e;
        }
    }
    if (over) a.flags[FZ_OVER] = 1;
}

// ------------------------------------------------------------------------------------------------
// the step
// ------------------------------------------------------------------------------------------------
template <int POT, int C, bool BEREND, bool ADVANCE>
__global__ void __launch_bounds__(128) fz_step_kernel(const FzArgs a)
{
    __shared__ double wsum[4];
    __shared__ int s_cnt;
    if (a.flags[FZ_OVER]) return; // a list overflowed: the state stays frozen at the last complete step
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads(); // the only block barrier: before any work
    const int nslots = a.flags[FZ_NSLOTS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double sc = 0.0;
    if (BEREND) { // src/thermostats.jl:76-83 with the temperature of v_{j-1} (only the last block overwrites scal[0], at its end)
        const double T = *(volatile const double *)a.scal / a.kB_ndf;
        sc = (1.0 / T == INFINITY) ? a.gamma : a.gamma * (a.T0 / T - 1.0);
    }
    double mv2 = 0.0;
    if (blockIdx.x * 128 < nslots) {
        const int k = blockIdx.x * 128 + tid;
        const bool active = k < nslots;
        const int kk = active ? k : nslots - 1;
        const uint64_t keep = l2_policy_keep(), stream = l2_policy_stream();
        const double4 xi = load_rec_nc_hint(a.xr + kk, keep);
        const int gl = lane & ~(C - 1), c0 = lane & (C - 1), gbase = kk & ~(C - 1);
        double tx[C], ty[C], tz[C], f[C][3];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            tx[c] = __shfl_sync(FULL, xi.x, gl + c);
            ty[c] = __shfl_sync(FULL, xi.y, gl + c);
            tz[c] = __shfl_sync(FULL, xi.z, gl + c);
            f[c][0] = f[c][1] = f[c][2] = 0.0;
        }
        const int cnt = active ? __ldcs(a.nlist + kk) : 0;
        // One batch = FLY gathered records x C targets, evaluated WITHOUT a branch per pair: the displacement of
        // every pair first, one (rare) branch for the whole batch if any component needs the wrap loops, then the
        // exact predicate as a select (out-of-cutoff, self and past-the-end pairs contribute g = 0).  The
        // independent chains interleave, so the FP64 latency is covered with few warps per SM.
        constexpr int FLY = C >= 8 ? 1 : (C == 4 ? 2 : (C == 2 ? 3 : 4));
        auto batch = [&](const double4 (&p)[FLY], const int (&m)[FLY], int e) {
            double rx[FLY][C], ry[FLY][C], rz[FLY][C];
            int hmax = 0;
#pragma unroll
            for (int u = 0; u < FLY; ++u) {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    rx[u][c] = __dsub_rn(tx[c], p[u].x);
                    ry[u][c] = __dsub_rn(ty[c], p[u].y);
                    rz[u][c] = __dsub_rn(tz[c], p[u].z);
                    const int hx = __double2hiint(rx[u][c]) & 0x7fffffff, hy = __double2hiint(ry[u][c]) & 0x7fffffff,
                              hz = __double2hiint(rz[u][c]) & 0x7fffffff;
                    hmax = max(hmax, max(hx, max(hy, hz)));
                }
            }
            if (hmax >= a.hi_radius) { // rare: some pair of the batch straddles a periodic face
#pragma unroll
                for (int u = 0; u < FLY; ++u) {
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        rx[u][c] = wrap_cubic(rx[u][c], a.radius, a.L);
                        ry[u][c] = wrap_cubic(ry[u][c], a.radius, a.L);
                        rz[u][c] = wrap_cubic(rz[u][c], a.radius, a.L);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < FLY; ++u) {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const double r2 = r2_unfused(rx[u][c], ry[u][c], rz[u][c]);
                    // r2 >= 0 and R2 > 0: IEEE order == order of the bit patterns (keeps the test off the FP64 pipe)
                    const bool in = __double_as_longlong(r2) < __double_as_longlong(a.R2) && m[u] != gbase + c && e + u < cnt;
                    const double r2s = in ? r2 : 1.0;
                    double g;
                    if (POT == 0) {
                        const double inv = rcp_fast(r2s);
                        const double qq = a.sigma2 * inv;
                        const double s6 = qq * qq * qq;
                        g = (s6 * inv) * fma(2.0, s6, -1.0); // (2 s12 - s6) / r2
                    } else {
                        g = w_rinv3(r2s, p[u].w);
                    }
                    g = in ? g : 0.0;
                    f[c][0] = fma(g, rx[u][c], f[c][0]);
                    f[c][1] = fma(g, ry[u][c], f[c][1]);
                    f[c][2] = fma(g, rz[u][c], f[c][2]);
                }
            }
        };
        // Software pipeline over the lane's entries, FLY at a time: the indices run two batches ahead of the
        // arithmetic and the 256-bit gathers one batch ahead, so neither latency sits on the critical path.
        // Entries past the end read the lane's own slot (a valid address) and are not evaluated.
        const int *lp = a.list + kk;
        int m0[FLY], m1[FLY];
        double4 p0[FLY];
#pragma unroll
        for (int u = 0; u < FLY; ++u) m0[u] = u < cnt ? __ldcs(lp + (size_t)u * a.cap_slots) : kk;
#pragma unroll
        for (int u = 0; u < FLY; ++u) m1[u] = FLY + u < cnt ? __ldcs(lp + (size_t)(FLY + u) * a.cap_slots) : kk;
#pragma unroll
        for (int u = 0; u < FLY; ++u) p0[u] = load_rec_nc_hint(a.xr + m0[u], keep);
        for (int e = 0; e < cnt; e += FLY) {
            double4 p1[FLY];
            int m2[FLY];
#pragma unroll
            for (int u = 0; u < FLY; ++u) p1[u] = load_rec_nc_hint(a.xr + m1[u], keep);
#pragma unroll
            for (int u = 0; u < FLY; ++u)
                m2[u] = e + 2 * FLY + u < cnt ? __ldcs(lp + (size_t)(e + 2 * FLY + u) * a.cap_slots) : kk;
            batch(p0, m0, e);
#pragma unroll
            for (int u = 0; u < FLY; ++u) { p0[u] = p1[u]; m0[u] = m1[u]; m1[u] = m2[u]; }
        }
        // the C lanes' partial sums of every target (butterfly: the same value in every lane of the cluster)
        double F0 = 0.0, F1 = 0.0, F2 = 0.0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                double v = f[c][d];
#pragma unroll
                for (int o = 1; o < C; o <<= 1) v += __shfl_xor_sync(FULL, v, o);
                f[c][d] = v;
            }
            if (c == c0) { F0 = f[c][0]; F1 = f[c][1]; F2 = f[c][2]; }
        }
        const int i = active ? a.pid[kk] : -1;
        if (i >= 0) {
            const double4 v0 = load_rec_hint(a.vm + kk, stream);
            double coeff = a.scale / v0.w;
            if (POT == 1) coeff *= xi.w;
            double an0 = coeff * F0, an1 = coeff * F1, an2 = coeff * F2;
            if (BEREND) { an0 += sc * v0.x; an1 += sc * v0.y; an2 += sc * v0.z; }
            const double vx = fma(a.hdt, __ldcs(a.sa + kk) + an0, v0.x);
            const double vy = fma(a.hdt, __ldcs(a.sa + a.cap_slots + kk) + an1, v0.y);
            const double vz = fma(a.hdt, __ldcs(a.sa + 2 * a.cap_slots + kk) + an2, v0.z);
            __stcs(a.sa + kk, an0); __stcs(a.sa + a.cap_slots + kk, an1); __stcs(a.sa + 2 * a.cap_slots + kk, an2);
            store_rec_hint(a.vm + kk, make_double4(vx, vy, vz, v0.w), stream);
            mv2 = v0.w * fma(vz, vz, fma(vy, vy, vx * vx));
            if (ADVANCE) {
                const double nx = fma(a.hdt2, an0, fma(a.dt, vx, xi.x));
                const double ny = fma(a.hdt2, an1, fma(a.dt, vy, xi.y));
                const double nz = fma(a.hdt2, an2, fma(a.dt, vz, xi.z));
                store_rec_hint(a.xw + kk, make_double4(nx, ny, nz, xi.w), keep); // the next step's gather target
                const double dx = nx - __ldcs(a.ref + kk), dy = ny - __ldcs(a.ref + a.cap_slots + kk),
                             dz = nz - __ldcs(a.ref + 2 * a.cap_slots + kk);
                const double d2 = dx * dx + dy * dy + dz * dz;
                if (!(d2 <= a.lim2)) a.flags[a.req ^ 1] = 1; // also catches NaN
            }
        } else if (active && ADVANCE) {
            store_rec_hint(a.xw + kk, xi, keep); // padding slot
        }
    }
    // sum m v^2 without a block barrier (a warp that is done must not hold its slot for the slowest one): warp sums
    // -> the block's last warp adds the four in warp order -> the last block of a group of 64 adds the group's
    // block sums in block order -> the last group adds the group sums in group order.  Who comes last varies;
    // what is added, and in which order, does not.
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mv2 += __shfl_down_sync(FULL, mv2, o);
    int last = 0;
    if (lane == 0) {
        wsum[warp] = mv2;
        __threadfence_block();
        last = atomicAdd(&s_cnt, 1) == 3;
    }
    last = __shfl_sync(FULL, last, 0);
    if (!last) return;
    const int nb = (int)gridDim.x, ng = (nb + FZ_GROUP - 1) / FZ_GROUP, grp = (int)blockIdx.x / FZ_GROUP;
    if (lane == 0) {
        __threadfence_block();
        const volatile double *ws = wsum;
        a.partial[blockIdx.x] = ((ws[0] + ws[1]) + ws[2]) + ws[3];
        __threadfence();
        const int members = min(FZ_GROUP, nb - grp * FZ_GROUP);
        last = atomicAdd(&a.gcount[grp], 1) == members - 1;
    }
    last = __shfl_sync(FULL, last, 0);
    if (!last) return;
    __threadfence();
    {
        const int b0 = grp * FZ_GROUP;
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < FZ_GROUP / 32; ++q) {
            const int b = b0 + q * 32 + lane;
            v += b < nb ? __ldcg(a.partial + b) : 0.0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
        if (lane == 0) {
            a.partial[a.cap_slots / 128 + grp] = v;
            a.gcount[grp] = 0;
            __threadfence();
            last = atomicAdd(&a.flags[FZ_DONE], 1) == ng - 1;
        }
        last = __shfl_sync(FULL, last, 0);
    }
    if (!last) return;
    __threadfence();
    double v = 0.0;
    for (int g = lane; g < ng; g += 32) v += __ldcg(a.partial + a.cap_slots / 128 + g);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
    if (lane == 0) {
        a.scal[0] = v;
        a.flags[FZ_DONE] = 0;
        a.flags[FZ_STEPS] += 1;
        a.flags[a.req] = 0; // the request this step's chain consumed
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int fz_pot(const nbx_ctx *c) { return c->has_lj ? 0 : 1; }

static bool fz_grid(nbx_ctx *c, CellGrid *g, double *R, double *skin)
{
    *R = fz_pot(c) == 0 ? c->lj_R : c->el_R;
    *skin = *R * 1e-3 * (double)c->opt_verlet_permille;
    if (!(*R > 0.0) || !std::isfinite(*R)) return false;
    if (cells_plan(c, *R + *skin, c->n, g) != NBX_OK) return false;
    return g->valid && g->ncell < (int64_t)1 << 30;
}

bool fused_eligible(nbx_ctx *c, int64_t nsteps)
{
    if (!c->opt_fused || c->fz.disabled || nsteps < c->fused_min_steps || nsteps > 2000000000) return false;
    if (c->water || c->slab.on || c->dyn || c->gid || c->pair_nranks > 1 || c->tgt_lo != 0 || c->tgt_hi != c->n) return false;
    if (c->ncols != c->n || c->n < 2 || c->n > (int64_t)1 << 30) return false;
    if (c->has_grav || c->has_dip || c->has_spcfw || (c->has_lj == c->has_coul)) return false;
    if (c->has_coul && !c->has_q) return false;
    if (c->thermo != NBX_THERMO_NONE && c->thermo != NBX_THERMO_BERENDSEN) return false;
    if (c->bc_kind != NBX_BC_CUBIC || !c->opt_prefilter || c->opt_verlet_permille <= 0) return false;
    const int C = c->opt_fused_cluster;
    if (C != 1 && C != 2 && C != 4 && C != 8) return false;
    CellGrid g;
    double R, skin;
    return fz_grid(c, &g, &R, &skin);
}

void fused_free(nbx_ctx *c)
{
    FusedState &z = c->fz;
    cudaFree(z.x[0]); cudaFree(z.x[1]); cudaFree(z.vm); cudaFree(z.sa); cudaFree(z.ref); cudaFree(z.sl);
    cudaFree(z.pid); cudaFree(z.scell); cudaFree(z.cell_of); cudaFree(z.arrival); cudaFree(z.tmp_idx);
    cudaFree(z.count); cudaFree(z.start); cudaFree(z.sums); cudaFree(z.list); cudaFree(z.nlist); cudaFree(z.flags);
    cudaFree(z.partial); cudaFree(z.gcount);
    const bool disabled = z.disabled;
    const int64_t st = z.steps_total, rb = z.rebuilds_total;
    z = FusedState();
    z.disabled = disabled; z.steps_total = st; z.rebuilds_total = rb;
}

static int fz_ensure(nbx_ctx *c, int C, int64_t ncell, int cap_e)
{
    FusedState &z = c->fz;
    if (z.C == C && z.n == c->n && z.ncell == ncell && z.cap_e >= cap_e && z.flags) return NBX_OK;
    fused_free(c);
    const int64_t pad = (int64_t)(C - 1) * (c->n < ncell ? c->n : ncell);
    const int64_t cap = ((c->n + pad + 127) / 128) * 128;
    NBX_TRY(dev_alloc(c, &z.x[0], (size_t)cap));
    NBX_TRY(dev_alloc(c, &z.x[1], (size_t)cap));
    NBX_TRY(dev_alloc(c, &z.vm, (size_t)cap));
    NBX_TRY(dev_alloc(c, &z.sa, (size_t)3 * cap));
    NBX_TRY(dev_alloc(c, &z.ref, (size_t)3 * cap));
    NBX_TRY(dev_alloc(c, &z.sl, (size_t)cap));
    NBX_TRY(dev_alloc(c, &z.pid, (size_t)cap));
    NBX_TRY(dev_alloc(c, &z.scell, (size_t)cap));
    NBX_TRY(dev_alloc(c, &z.cell_of, (size_t)c->n));
    NBX_TRY(dev_alloc(c, &z.arrival, (size_t)c->n));
    NBX_TRY(dev_alloc(c, &z.tmp_idx, (size_t)cap));
    NBX_TRY(dev_alloc(c, &z.count, (size_t)ncell + 1));
    NBX_TRY(dev_alloc(c, &z.start, (size_t)ncell + 1));
    NBX_TRY(dev_alloc(c, &z.sums, (size_t)(ncell / 1024 + 3)));
    NBX_TRY(dev_alloc(c, &z.list, (size_t)cap_e * (size_t)cap));
    NBX_TRY(dev_alloc(c, &z.nlist, (size_t)cap));
    NBX_TRY(dev_alloc(c, &z.flags, (size_t)FZ_NFLAGS));
    NBX_TRY(dev_alloc(c, &z.partial, (size_t)(cap / 128) + (size_t)(cap / 128 / FZ_GROUP + 2)));
    NBX_TRY(dev_alloc(c, &z.gcount, (size_t)(cap / 128 / FZ_GROUP + 2)));
    NBX_CUDA(c, cudaMemsetAsync(z.gcount, 0, sizeof(int) * (size_t)(cap / 128 / FZ_GROUP + 2), c->stream));
    z.C = C; z.n = c->n; z.ncell = ncell; z.cap_slots = cap; z.cap_e = cap_e;
    return NBX_OK;
}

template <int C>
static int fz_chain(nbx_ctx *c, const FzArgs &a)
{
    const int sm = c->sm_count;
    auto grid = [&](int64_t work, int threads) {
        const int64_t full = (work + threads - 1) / threads, cap = (int64_t)sm * (2048 / threads);
        return (unsigned)(full < cap ? (full > 0 ? full : 1) : cap);
    };
    timer_begin(c, NBX_T_CELL_BUILD);
    fz_unsort_kernel<<<grid(a.cap_slots, 256), 256, 0, c->stream>>>(a, 1, 0);
    fz_count_kernel<<<grid(a.n, 256), 256, 0, c->stream>>>(a);
    NBX_TRY(cells_scan(c, a.count, a.start, a.ncell, c->fz.sums, C, a.flags + a.req));
    fz_scatter_kernel<<<grid(a.n, 256), 256, 0, c->stream>>>(a);
    fz_place_kernel<C><<<grid(a.n, 128), 128, 0, c->stream>>>(a);
    fz_build_kernel<C><<<grid(a.cap_slots, 128), 128, 0, c->stream>>>(a);
    timer_end(c, NBX_T_CELL_BUILD);
    NBX_CUDA(c, cudaGetLastError());
    return NBX_OK;
}

template <int C>
static int fz_step(nbx_ctx *c, const FzArgs &a, bool advance)
{
    const unsigned blocks = (unsigned)(a.cap_slots / 128);
    const bool ber = c->thermo == NBX_THERMO_BERENDSEN;
    const int pot = fz_pot(c);
    timer_begin(c, NBX_T_PAIR_CELLS);
#define FZ_LAUNCH(P, B, A) fz_step_kernel<P, C, B, A><<<blocks, 128, 0, c->stream>>>(a)
    if (pot == 0) {
        if (ber) { if (advance) FZ_LAUNCH(0, true, true); else FZ_LAUNCH(0, true, false); }
        else { if (advance) FZ_LAUNCH(0, false, true); else FZ_LAUNCH(0, false, false); }
    } else {
        if (ber) { if (advance) FZ_LAUNCH(1, true, true); else FZ_LAUNCH(1, true, false); }
        else { if (advance) FZ_LAUNCH(1, false, true); else FZ_LAUNCH(1, false, false); }
    }
#undef FZ_LAUNCH
    timer_end(c, NBX_T_PAIR_CELLS);
    NBX_CUDA(c, cudaGetLastError());
    return NBX_OK;
}

template <int C>
static int fz_run(nbx_ctx *c, double dt, int64_t nsteps, int64_t *steps_done)
{
    FusedState &z = c->fz;
    CellGrid g;
    double R, skin;
    if (!fz_grid(c, &g, &R, &skin)) return fail(c, NBX_ERR_INVALID, "fused step: the box cannot be cell-listed");
    const double L = c->bc[0];
    // list capacity from the mean density: 1.5 x the expected entries per lane (+ slack).  The union of C spheres of
    // radius R + skin around points of one cell holds about uf(C) x one sphere (Monte Carlo, uniform fluid).
    const double uf = C == 1 ? 1.0 : (C == 2 ? 1.55 : (C == 4 ? 2.2 : 3.0));
    const double expect = (double)c->n / (L * L * L) * 4.18879020478639 * (R + skin) * (R + skin) * (R + skin) * uf / C;
    int cap_e = (int)(1.5 * expect) + 16;
    if (cap_e > c->n) cap_e = (int)c->n;
    NBX_TRY(fz_ensure(c, C, g.ncell, cap_e));

    FzArgs a{};
    a.xa = z.x[0]; a.xb = z.x[1];
    a.vm = z.vm; a.sa = z.sa; a.ref = z.ref; a.sl = z.sl;
    a.pid = z.pid; a.scell = z.scell; a.cell_of = z.cell_of; a.arrival = z.arrival; a.tmp_idx = z.tmp_idx;
    a.count = z.count; a.start = z.start; a.list = z.list; a.nlist = z.nlist; a.flags = z.flags;
    a.partial = z.partial; a.scal = c->d_scal; a.gcount = z.gcount;
    a.pos = c->pos; a.vel = c->vel; a.acc = c->acc; a.mass = c->mass; a.charge = c->has_q ? c->charge : nullptr;
    a.ld = c->npad; a.cap_slots = z.cap_slots;
    a.n = (int)c->n; a.nc = g.nc[0]; a.ncell = (int)g.ncell; a.cap_e = z.cap_e;
    a.L = L; a.radius = 0.5 * L; a.R2 = R * R; a.sigma2 = c->lj_sigma2;
    a.scale = fz_pot(c) == 0 ? 24.0 * c->lj_eps : c->el_k;
    {   // same threshold as make_args (nbx_cells.cu) for the cutoff R + skin
        const double cell = L / (double)a.nc;
        const double margin = 32.0 * (double)a.nc * 5.9604644775390625e-8 + 1e-5;
        a.R2f = nextafterf((float)((R + skin) * (R + skin) / (cell * cell) * (1.0 + margin)), INFINITY);
        int64_t bits;
        memcpy(&bits, &a.radius, sizeof bits);
        a.hi_radius = (int)(bits >> 32);
    }
    a.dt = dt; a.hdt = 0.5 * dt; a.hdt2 = 0.5 * dt * dt;
    const double lim = 0.5 * skin * (1.0 - 1e-9);
    a.lim2 = lim * lim;
    a.kB_ndf = c->kB * (double)(3 * c->thN - c->thNc);
    a.T0 = c->T0; a.gamma = c->thermo == NBX_THERMO_BERENDSEN ? 0.5 / c->tparam : 0.0;
    a.nsteps = (int)nsteps;

    // step j reads buffer j & 1, consumes request flag j & 1, writes buffer / raises flag (j + 1) & 1
    auto args_of = [&](int64_t j) {
        FzArgs s = a;
        s.req = (int)(j & 1);
        s.xr = (j & 1) ? z.x[1] : z.x[0];
        s.xw = (j & 1) ? z.x[0] : z.x[1];
        return s;
    };
    auto enqueue = [&](int64_t j) -> int {
        const FzArgs s = args_of(j);
        NBX_TRY(fz_chain<C>(c, s));
        return fz_step<C>(c, s, j < nsteps);
    };

    NBX_TRY(launch_vv_pos(c, dt)); // x_1 in particle order; the first chain lays everything down in slot order
    NBX_CUDA(c, cudaMemsetAsync(z.flags, 0, sizeof(int) * FZ_NFLAGS, c->stream));
    NBX_CUDA(c, cudaMemsetAsync(z.flags + FZ_REQ1, 1, sizeof(int), c->stream)); // != 0: build before step 1
    int64_t j = 1;
    if (c->opt_fused_debug & 4) { // test hook ("fused_debug" bit 2): lists only, the unfused path runs the steps
        NBX_TRY(fz_chain<C>(c, args_of(1)));
        NBX_CUDA(c, cudaStreamSynchronize(c->stream));
        *steps_done = 0;
        return NBX_OK;
    }
    NBX_TRY(enqueue(j++));
    const bool graphable = c->opt_graph && !c->timing && nsteps >= 32 && c->stream != nullptr &&
                           c->stream != cudaStreamLegacy && c->stream != cudaStreamPerThread;
    if (graphable) {
        // steps 2 .. nsteps-1 advance the positions: (even, odd) pairs replay one captured graph
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        cudaError_t e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
            int rc = enqueue(2);
            if (rc == NBX_OK) rc = enqueue(3);
            e = cudaStreamEndCapture(c->stream, &graph);
            if (rc != NBX_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
        }
        if (e == cudaSuccess && exec) {
            for (; j + 2 <= nsteps; j += 2) { // steps j, j+1 < nsteps
                e = cudaGraphLaunch(exec, c->stream);
                if (e != cudaSuccess) break;
            }
        }
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        if (e != cudaSuccess) return cuda_fail(c, e, "CUDA graph of the fused step");
    }
    for (; j <= nsteps; ++j) NBX_TRY(enqueue(j));

    // back to particle order (buffer picked on the device from the step counter), then the verdict
    {
        FzArgs s = a;
        s.xr = z.x[0];
        const int64_t full = (z.cap_slots + 255) / 256, cap = (int64_t)c->sm_count * 8;
        fz_unsort_kernel<<<(unsigned)(full < cap ? full : cap), 256, 0, c->stream>>>(s, 0, 1);
        NBX_CUDA(c, cudaGetLastError());
    }
    int h[FZ_NFLAGS];
    NBX_CUDA(c, cudaMemcpyAsync(h, z.flags, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    NBX_CUDA(c, cudaStreamSynchronize(c->stream));
    *steps_done = h[FZ_STEPS];
    z.steps_total += h[FZ_STEPS];
    z.rebuilds_total += h[FZ_REBUILDS];
    if (h[FZ_OVER]) z.disabled = true;
    else if (h[FZ_STEPS] != nsteps) return fail(c, NBX_ERR_CUDA, "fused step: %d of %lld steps ran", h[FZ_STEPS], (long long)nsteps);
    return NBX_OK;
}

// nsteps velocity-Verlet steps from the resident state.  On return *steps_done steps are complete; if that is
// fewer than nsteps (a cluster list overflowed) pos already holds x of step *steps_done + 1 and the caller
// continues on the unfused path without its first position update.
int fused_run(nbx_ctx *c, double dt, int64_t nsteps, int64_t *steps_done)
{
    *steps_done = 0;
    switch (c->opt_fused_cluster) {
    case 1: return fz_run<1>(c, dt, nsteps, steps_done);
    case 2: return fz_run<2>(c, dt, nsteps, steps_done);
    case 8: return fz_run<8>(c, dt, nsteps, steps_done);
    default: return fz_run<4>(c, dt, nsteps, steps_done);
    }
}

} // namespace nbx
