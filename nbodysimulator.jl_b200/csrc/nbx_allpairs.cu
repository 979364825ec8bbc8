// nbx_allpairs.cu -- target-tiled all-pairs kernels (sm_100a).
//
// Replaces the O(N) inner loops of
//   gravitational_acceleration!             src/basic_potentials.jl:306-331
//   pairwise_electrostatic_acceleration!    src/basic_potentials.jl:274-304
//   magnetostatic_dipdip_acceleration!      src/basic_potentials.jl:333-365
//   pairwise_lennard_jones_acceleration!    src/basic_potentials.jl:240-272 (boxes too small for a cell list)
// each of which the reference runs once per target particle from soode_system!
// (src/nbody_to_ode.jl:474-488).
//
// Decomposition: work item = (target tile of THREADS*T bodies) x (source chunk).  Each thread
// keeps T targets in registers; the CTA streams the chunk's sources through a NST-stage shared
// memory ring filled by TMA 1-D bulk copies (cp.async.bulk + mbarrier complete_tx) of the SoA
// rows.  All lanes read the same source at the same time (shared-memory broadcast, LDS.128 of
// two consecutive sources).  Per-item partial sums go to part[chunk][comp][target]; a second
// kernel adds the chunks in ascending order, applies the per-target prefactor and writes the SoA
// acceleration rows -- so the result is bit-reproducible whatever the grid or the scheduling.
#include "nbx_internal.cuh"

namespace nbx {

struct APParams {
    const double *src[6];   // SoA source rows (padded to a multiple of S doubles)
    int nsrc_pad;           // padded source count
    int n;                  // real source count
    int tgt_lo;             // first target column
    int ntgt;               // targets in this launch
    int ntiles, nchunk;
    int chunk_len;          // sources per chunk (multiple of S)
    double *part;           // [nchunk][NACC][part_ld]
    int part_ld;
};

// ------------------------------------------------------------------------------------------------
// Pair policies
// ------------------------------------------------------------------------------------------------

// Unbounded 1/r^2 central force: gravity (w = m) and Coulomb with R = Inf in an InfiniteBox (w = q).
// Accumulates sum_j w_j (r_j - r_i) / |r_j - r_i|^3 ; the reduce kernel applies G or -k q_i / m_i.
struct GravPolicy {
    static constexpr int NA = 4;   // x y z w
    static constexpr int NACC = 3;
    struct Args { const double *tx, *ty, *tz; };
    struct Tgt { double x, y, z; int i; };
    __device__ static __forceinline__ Tgt load(const Args &a, int i)
    {
        return Tgt{a.tx[i], a.ty[i], a.tz[i], i};
    }
    template <bool MASK>
    __device__ static __forceinline__ void pair(const Args &, const Tgt &t, double (&acc)[NACC],
                                                const double (&s)[NA], int jg, int)
    {
        const double dx = s[0] - t.x, dy = s[1] - t.y, dz = s[2] - t.z;
        double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
        if (MASK) r2 = (jg == t.i) ? 1.0 : r2; // self pair: dx = dy = dz = 0 -> contributes exactly 0
        const double f = w_rinv3(r2, s[3]);
        acc[0] = fma(f, dx, acc[0]);
        acc[1] = fma(f, dy, acc[1]);
        acc[2] = fma(f, dz, acc[2]);
    }
};

// Dipole-dipole force, src/basic_potentials.jl:344-357.  With rh = rij/|rij|, mir = mi.rh, mjr = mj.rh:
//   contrib = (mi mjr + mj mir + rh (mi.mj) - 5 rh mir mjr) / |rij|^4
// acc[0..2] collect the mj- and rh-directed parts, acc[3] the scalar sum_j mjr/|rij|^4 that multiplies
// the (constant) target moment mi; the reduce kernel recombines them.
struct DipolePolicy {
    static constexpr int NA = 6;   // x y z mx my mz
    static constexpr int NACC = 4;
    struct Args { const double *tx, *ty, *tz, *tmx, *tmy, *tmz; };
    struct Tgt { double x, y, z, mx, my, mz; int i; };
    __device__ static __forceinline__ Tgt load(const Args &a, int i)
    {
        return Tgt{a.tx[i], a.ty[i], a.tz[i], a.tmx[i], a.tmy[i], a.tmz[i], i};
    }
    template <bool MASK>
    __device__ static __forceinline__ void pair(const Args &, const Tgt &t, double (&acc)[NACC],
                                                const double (&s)[NA], int jg, int n)
    {
        const double dx = t.x - s[0], dy = t.y - s[1], dz = t.z - s[2]; // rij = ri - rj
        double d = fma(dz, dz, fma(dy, dy, dx * dx));
        bool live = true;
        if (MASK) { live = (jg != t.i) && (jg < n); d = live ? d : 1.0; }
        // full-precision 1/sqrt(d): y = y0 (1 + e (1/2 + 3/8 e)), e = 1 - d y0^2
        const double y0 = rsqrt_seed(d);
        const double a = y0 * y0;
        const double e = fma(-d, a, 1.0);
        const double y = fma(fma(0.375, e, 0.5), e * y0, y0);
        const double i2 = y * y;
        double i4 = i2 * i2;
        if (MASK) i4 = live ? i4 : 0.0;
        const double rx = dx * y, ry = dy * y, rz = dz * y;
        const double mir = fma(t.mz, rz, fma(t.my, ry, t.mx * rx));
        const double mjr = fma(s[5], rz, fma(s[4], ry, s[3] * rx));
        const double mimj = fma(t.mz, s[5], fma(t.my, s[4], t.mx * s[3]));
        const double g = i4 * fma(-5.0 * mir, mjr, mimj);
        const double h = i4 * mir;
        acc[0] = fma(s[3], h, fma(rx, g, acc[0]));
        acc[1] = fma(s[4], h, fma(ry, g, acc[1]));
        acc[2] = fma(s[5], h, fma(rz, g, acc[2]));
        acc[3] = fma(i4, mjr, acc[3]);
    }
};

// Exact-predicate pair loop for cutoff potentials under any boundary kind: rij = ri - rj wrapped by
// the reference's loops, r2 un-fused, `r2 < R2` strict (src/boundary_conditions.jl:111-172,
// src/basic_potentials.jl:258, :292).  POT 0: Lennard-Jones, POT 1: Coulomb.
// EXCL 0: j != i (src/nbody_to_ode.jl:316-329), EXCL 1: own molecule (j/3 == i/3, :331-351).
template <int POT, int EXCL>
struct CutoffPolicy {
    static constexpr int NA = POT == 1 ? 4 : 3;
    static constexpr int NACC = 3;
    struct Args {
        const double *tx, *ty, *tz;
        int bc_kind;
        double b0, b1, b2, b3, b4, b5; // cubic: b0 = L, b1 = 0.5 L ; periodic: lo/hi per dimension
        double R2, sigma2;
    };
    struct Tgt { double x, y, z; int i; };
    __device__ static __forceinline__ Tgt load(const Args &a, int i)
    {
        return Tgt{a.tx[i], a.ty[i], a.tz[i], i};
    }
    template <bool MASK>
    __device__ static __forceinline__ void pair(const Args &a, const Tgt &t, double (&acc)[NACC],
                                                const double (&s)[NA], int jg, int n)
    {
        if (MASK) {
            const bool excl = EXCL == 0 ? (jg == t.i) : ((jg / 3) == (t.i / 3));
            if (excl || jg >= n) return;
        }
        double x = __dsub_rn(t.x, s[0]), y = __dsub_rn(t.y, s[1]), z = __dsub_rn(t.z, s[2]);
        if (a.bc_kind == NBX_BC_CUBIC) {
            x = wrap_cubic(x, a.b1, a.b0);
            y = wrap_cubic(y, a.b1, a.b0);
            z = wrap_cubic(z, a.b1, a.b0);
        } else if (a.bc_kind == NBX_BC_PERIODIC) {
            x = wrap_range(x, a.b0, a.b1);
            y = wrap_range(y, a.b2, a.b3);
            z = wrap_range(z, a.b4, a.b5);
        }
        const double r2 = r2_unfused(x, y, z);
        if (r2 < a.R2) {
            double f;
            if (POT == 0) {
                const double inv = 1.0 / r2;
                const double q = a.sigma2 * inv;
                const double s6 = q * q * q;
                const double s12 = s6 * s6;
                f = (2.0 * s12 - s6) * inv;
            } else {
                f = w_rinv3(r2, s[3]);
            }
            acc[0] = fma(f, x, acc[0]);
            acc[1] = fma(f, y, acc[1]);
            acc[2] = fma(f, z, acc[2]);
        }
    }
};

// ------------------------------------------------------------------------------------------------
// The tiled kernel
// ------------------------------------------------------------------------------------------------
template <class P, int T, int THREADS, int S, int NST, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) allpairs_kernel(const APParams p, const typename P::Args args)
{
    constexpr int NA = P::NA;
    constexpr int NACC = P::NACC;
    constexpr int TILE = THREADS * T;
    constexpr uint32_t STAGE_BYTES = NA * S * sizeof(double);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *ring = reinterpret_cast<double *>(smem_raw); // [NST][NA][S]
    __shared__ __align__(8) uint64_t full[NST];

    const int tid = threadIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t parity = 0; // bit s = phase the next wait on stage s must see completed

    const int nitems = p.ntiles * p.nchunk;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int tile = item / p.nchunk;
        const int chunk = item - tile * p.nchunk;
        const int s_begin = chunk * p.chunk_len;
        const int s_end = min(s_begin + p.chunk_len, p.nsrc_pad);
        const int nsub = (s_end - s_begin) / S;
        const int tbase = p.tgt_lo + tile * TILE;
        const int tlast = p.tgt_lo + p.ntgt - 1;

        typename P::Tgt tg[T];
        double acc[T][NACC];
#pragma unroll
        for (int k = 0; k < T; ++k) {
            tg[k] = P::load(args, min(tbase + k * THREADS + tid, tlast));
#pragma unroll
            for (int c = 0; c < NACC; ++c) acc[k][c] = 0.0;
        }

        auto issue = [&](int sub, int st) {
            double *dst = ring + (size_t)st * NA * S;
            const int j0 = s_begin + sub * S;
            mbar_expect_tx(&full[st], STAGE_BYTES);
#pragma unroll
            for (int a = 0; a < NA; ++a) bulk_g2s(dst + a * S, p.src[a] + j0, S * sizeof(double), &full[st]);
        };
        if (tid == 0) {
            const int npro = min(NST, nsub);
            for (int s = 0; s < npro; ++s) issue(s, s);
        }

        for (int sub = 0; sub < nsub; ++sub) {
            const int st = sub % NST;
            mbar_wait(&full[st], (parity >> st) & 1u);
            parity ^= 1u << st;
            const double *b = ring + (size_t)st * NA * S;
            const int j0 = s_begin + sub * S;
            // sub-tiles that can hold a self / excluded / padding source take the masked body
            const bool masked = (j0 < tbase + TILE + 3 && j0 + S + 3 > tbase) || (j0 + S > p.n);
            if (masked) {
#pragma unroll 1
                for (int j = 0; j < S; j += 2) {
                    double2 v[NA];
#pragma unroll
                    for (int a = 0; a < NA; ++a) v[a] = *reinterpret_cast<const double2 *>(b + a * S + j);
                    double s0[NA], s1[NA];
#pragma unroll
                    for (int a = 0; a < NA; ++a) { s0[a] = v[a].x; s1[a] = v[a].y; }
#pragma unroll
                    for (int k = 0; k < T; ++k) P::template pair<true>(args, tg[k], acc[k], s0, j0 + j, p.n);
#pragma unroll
                    for (int k = 0; k < T; ++k) P::template pair<true>(args, tg[k], acc[k], s1, j0 + j + 1, p.n);
                }
            } else {
#pragma unroll 2
                for (int j = 0; j < S; j += 2) {
                    double2 v[NA];
#pragma unroll
                    for (int a = 0; a < NA; ++a) v[a] = *reinterpret_cast<const double2 *>(b + a * S + j);
                    double s0[NA], s1[NA];
#pragma unroll
                    for (int a = 0; a < NA; ++a) { s0[a] = v[a].x; s1[a] = v[a].y; }
#pragma unroll
                    for (int k = 0; k < T; ++k) P::template pair<false>(args, tg[k], acc[k], s0, j0 + j, p.n);
#pragma unroll
                    for (int k = 0; k < T; ++k) P::template pair<false>(args, tg[k], acc[k], s1, j0 + j + 1, p.n);
                }
            }
            __syncthreads(); // every lane is done with stage st before TMA refills it
            if (tid == 0 && sub + NST < nsub) issue(sub + NST, st);
        }

#pragma unroll
        for (int k = 0; k < T; ++k) {
            const int i = tbase + k * THREADS + tid;
            if (i <= tlast) {
#pragma unroll
                for (int c = 0; c < NACC; ++c)
                    p.part[((size_t)chunk * NACC + c) * p.part_ld + (i - p.tgt_lo)] = acc[k][c];
            }
        }
    }
}

// Sum the chunk partials in ascending chunk order and apply the per-target prefactor.
//   kind 0: scale                      (gravity: G)
//   kind 1: scale * q_i / m_i          (Coulomb: k q_i / m_i, sign folded into scale)
//   kind 2: scale / m_i                (LJ: 24 eps / m_i)
//   kind 3: dipole recombination, scale / m_i with scale = 3 mu/4pi
__global__ void allpairs_reduce_kernel(const double *__restrict__ part, int nchunk, int nacc, int part_ld,
                                       int tgt_lo, int ntgt, int kind, double scale,
                                       const double *__restrict__ mass, int mstride,
                                       const double *__restrict__ charge, const double *__restrict__ mmx, const double *__restrict__ mmy,
                                       const double *__restrict__ mmz, double *__restrict__ ax,
                                       double *__restrict__ ay, double *__restrict__ az, int accumulate)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntgt) return;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int c = 0; c < nchunk; ++c)
        for (int k = 0; k < nacc; ++k) s[k] += part[((size_t)c * nacc + k) * part_ld + t];
    const int i = tgt_lo + t;
    double f = scale;
    if (kind == 1) f = scale * charge[i] / mass[i];
    if (kind == 2 || kind == 3) f = scale / mass[(size_t)i * mstride];
    if (kind == 3) {
        s[0] = fma(mmx[i], s[3], s[0]);
        s[1] = fma(mmy[i], s[3], s[1]);
        s[2] = fma(mmz[i], s[3], s[2]);
    }
    if (accumulate) {
        ax[i] += f * s[0]; ay[i] += f * s[1]; az[i] += f * s[2];
    } else {
        ax[i] = f * s[0]; ay[i] = f * s[1]; az[i] = f * s[2];
    }
}

// Small systems (the reference's own examples: 216 argon atoms, 216 water molecules): the tiled kernel would put all targets
// into ONE CTA that walks the sources serially (221 us for 216 atoms with R = L/2: one warp per scheduler, every FP64
// latency exposed).  Here a warp owns a target, its lanes stride over the sources straight from L2 and meet in a butterfly
// of fixed order; the partial sums land in chunk 0 of the layout the reduce kernel reads.
template <class P>
__global__ void __launch_bounds__(128) allpairs_small_kernel(const APParams p, const typename P::Args args)
{
    constexpr int NA = P::NA, NACC = P::NACC;
    const int lane = threadIdx.x & 31, w = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (w >= p.ntgt) return; // (warp-uniform)
    const typename P::Tgt tg = P::load(args, p.tgt_lo + w);
    double acc[NACC];
#pragma unroll
    for (int c = 0; c < NACC; ++c) acc[c] = 0.0;
    for (int j = lane; j < p.n; j += 32) {
        double s[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) s[a] = p.src[a][j];
        P::template pair<true>(args, tg, acc, s, j, p.n);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int c = 0; c < NACC; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    }
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < NACC; ++c) p.part[(size_t)c * p.part_ld + w] = acc[c];
    }
}
constexpr int kSmallSystem = 1024; // sources and targets up to which the warp-per-target kernel is used

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
template <int T, int THREADS, int S>
static void plan_items(nbx_ctx *c, int ctas_per_sm, APParams *p, int *grid)
{
    const int tile = T * THREADS;
    p->ntiles = (p->ntgt + tile - 1) / tile;
    const int nsub_total = p->nsrc_pad / S;
    const int grid_full = c->sm_count * ctas_per_sm;
    // enough items that the static round-robin over a full grid is balanced to ~3 %,
    // but never chunks shorter than 4 sub-tiles (pipeline prologue amortisation)
    const int want = (32 * grid_full + p->ntiles - 1) / p->ntiles;
    const int max_chunks = nsub_total / 4 > 0 ? nsub_total / 4 : 1;
    const int nchunk = want < 1 ? 1 : (want > max_chunks ? max_chunks : want);
    const int sub_per_chunk = (nsub_total + nchunk - 1) / nchunk;
    p->chunk_len = sub_per_chunk * S;
    p->nchunk = (nsub_total + sub_per_chunk - 1) / sub_per_chunk;
    const long items = (long)p->ntiles * p->nchunk;
    *grid = (int)(items < grid_full ? items : grid_full);
    p->part_ld = ((p->ntgt + 31) / 32) * 32;
}

static int ensure_part(nbx_ctx *c, size_t bytes)
{
    if (bytes <= c->part_bytes) return NBX_OK;
    if (c->part) { cudaFree(c->part); c->part = nullptr; c->part_bytes = 0; }
    cudaError_t e = cudaMalloc((void **)&c->part, bytes);
    if (e != cudaSuccess) return cuda_fail(c, e, "cudaMalloc(partials)");
    c->part_bytes = bytes;
    return NBX_OK;
}

// Launch the tiled kernel for policy P; *p comes back with the item plan filled in.
template <class P, int T, int THREADS, int S, int NST, int MINB>
static int run_allpairs(nbx_ctx *c, APParams *p, const typename P::Args &args)
{
    int grid = 1;
    if (p->n <= kSmallSystem && p->ntgt <= kSmallSystem && p->ntgt > 0) {
        p->ntiles = 1; p->nchunk = 1; p->chunk_len = p->nsrc_pad;
        p->part_ld = ((p->ntgt + 31) / 32) * 32;
        NBX_TRY(ensure_part(c, (size_t)P::NACC * p->part_ld * sizeof(double)));
        p->part = c->part;
        timer_begin(c, NBX_T_PAIR_ALLPAIRS);
        allpairs_small_kernel<P><<<(p->ntgt + 3) / 4, 128, 0, c->stream>>>(*p, args);
        timer_end(c, NBX_T_PAIR_ALLPAIRS);
        NBX_CUDA(c, cudaGetLastError());
        c->last_grid = (p->ntgt + 3) / 4;
        c->last_nchunk = 1;
        return NBX_OK;
    }
    plan_items<T, THREADS, S>(c, MINB, p, &grid);
    NBX_TRY(ensure_part(c, (size_t)p->nchunk * P::NACC * p->part_ld * sizeof(double)));
    p->part = c->part;
    auto kern = allpairs_kernel<P, T, THREADS, S, NST, MINB>;
    const size_t smem = (size_t)NST * P::NA * S * sizeof(double);
    const void *key = reinterpret_cast<const void *>(kern);
    bool done = false;
    for (const void *k : c->attr_done) done = done || (k == key);
    if (!done) {
        NBX_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        c->attr_done.push_back(key);
    }
    timer_begin(c, NBX_T_PAIR_ALLPAIRS);
    kern<<<grid, THREADS, smem, c->stream>>>(*p, args);
    timer_end(c, NBX_T_PAIR_ALLPAIRS);
    NBX_CUDA(c, cudaGetLastError());
    c->last_grid = grid;
    c->last_nchunk = p->nchunk;
    return NBX_OK;
}

static int run_reduce(nbx_ctx *c, const APParams &p, int nacc, int kind, double scale, const double *mass,
                      int mstride, double *acc_out, int64_t ld_out, bool accumulate)
{
    const int threads = 256;
    const int blocks = (p.ntgt + threads - 1) / threads;
    allpairs_reduce_kernel<<<blocks, threads, 0, c->stream>>>(
        c->part, p.nchunk, nacc, p.part_ld, p.tgt_lo, p.ntgt, kind, scale, mass, mstride, c->charge, c->mm,
        c->mm ? c->mm + c->npad : nullptr, c->mm ? c->mm + 2 * c->npad : nullptr, acc_out, acc_out + ld_out,
        acc_out + 2 * ld_out, accumulate ? 1 : 0);
    NBX_CUDA(c, cudaGetLastError());
    return NBX_OK;
}

// Tunables: targets per thread, CTA size, sources per stage, ring depth, CTAs per SM.
constexpr int G_T = 4, G_THREADS = 128, G_S = 256, G_NST = 3, G_MINB = 4;
constexpr int D_T = 2, D_THREADS = 128, D_S = 256, D_NST = 3, D_MINB = 4;
constexpr int C_T = 2, C_THREADS = 128, C_S = 256, C_NST = 3, C_MINB = 4;

int launch_allpairs_grav(nbx_ctx *c, const double *w, int scale_kind, double scale, double *acc_out, bool accumulate)
{
    APParams p{};
    const double *x = c->pos, *y = c->pos + c->npad, *z = c->pos + 2 * c->npad;
    p.src[0] = x; p.src[1] = y; p.src[2] = z; p.src[3] = w;
    p.nsrc_pad = (int)c->npad;
    p.n = (int)c->n;
    p.tgt_lo = (int)c->tgt_lo;
    p.ntgt = (int)(c->tgt_hi - c->tgt_lo);
    if (p.ntgt <= 0) return NBX_OK;
    const bool whole = c->tgt_lo == 0 && c->tgt_hi == c->n;
    if (c->pair_nranks > 1 || (c->opt_sym && whole && c->n >= c->sym_min_n)) {
        const bool is_mass = (w == c->mass);
        return launch_sympairs(c, w, is_mass ? c->mass_uniform : c->charge_uniform, is_mass ? c->h_m1 : c->h_q1,
                               scale_kind, scale, acc_out, accumulate);
    }
    GravPolicy::Args a{x, y, z};
    NBX_TRY((run_allpairs<GravPolicy, G_T, G_THREADS, G_S, G_NST, G_MINB>(c, &p, a)));
    return run_reduce(c, p, GravPolicy::NACC, scale_kind, scale, c->mass, 1, acc_out, c->npad, accumulate);
}

int launch_allpairs_dipole(nbx_ctx *c, double *acc_out, bool accumulate)
{
    APParams p{};
    const double *x = c->pos, *y = c->pos + c->npad, *z = c->pos + 2 * c->npad;
    const double *mx = c->mm, *my = c->mm + c->npad, *mz = c->mm + 2 * c->npad;
    p.src[0] = x; p.src[1] = y; p.src[2] = z; p.src[3] = mx; p.src[4] = my; p.src[5] = mz;
    p.nsrc_pad = (int)c->npad;
    p.n = (int)c->n;
    p.tgt_lo = (int)c->tgt_lo;
    p.ntgt = (int)(c->tgt_hi - c->tgt_lo);
    if (p.ntgt <= 0) return NBX_OK;
    DipolePolicy::Args a{x, y, z, mx, my, mz};
    NBX_TRY((run_allpairs<DipolePolicy, D_T, D_THREADS, D_S, D_NST, D_MINB>(c, &p, a)));
    // coeff = 3 mu/4pi / m_i  (src/basic_potentials.jl:360)
    return run_reduce(c, p, DipolePolicy::NACC, 3, 3.0 * c->mu_4pi, c->mass, 1, acc_out, c->npad, accumulate);
}

template <int POT, int EXCL>
static int run_cutoff(nbx_ctx *c, const double *px, int64_t ld, double R2, APParams p, const double *mass,
                      int mstride, double *acc_out, int64_t ld_out, bool accumulate)
{
    using P = CutoffPolicy<POT, EXCL>;
    typename P::Args a{};
    a.tx = px; a.ty = px + ld; a.tz = px + 2 * ld;
    a.bc_kind = c->bc_kind;
    if (c->bc_kind == NBX_BC_CUBIC) { a.b0 = c->bc[0]; a.b1 = 0.5 * c->bc[0]; }
    else { a.b0 = c->bc[0]; a.b1 = c->bc[1]; a.b2 = c->bc[2]; a.b3 = c->bc[3]; a.b4 = c->bc[4]; a.b5 = c->bc[5]; }
    a.R2 = R2;
    a.sigma2 = c->lj_sigma2;
    NBX_TRY((run_allpairs<P, C_T, C_THREADS, C_S, C_NST, C_MINB>(c, &p, a)));
    // LJ prefactor 24 eps / m_i (src/basic_potentials.jl:267); Coulomb k q_i / m_i (:299)
    const int kind = POT == 0 ? 2 : 1;
    const double scale = POT == 0 ? 24.0 * c->lj_eps : c->el_k;
    return run_reduce(c, p, 3, kind, scale, mass, mstride, acc_out, ld_out, accumulate);
}

// pot 0: LJ over the n columns of px (self exclusion; mass of column i = mass[i * mstride], so the
// compact oxygen sub-system of water passes mstride = 3); pot 1: Coulomb over all atoms (self
// exclusion); pot 2: Coulomb with own-molecule exclusion (water).  px = SoA rows of stride ld,
// padded to a multiple of kPad.
int launch_allpairs_pbc(nbx_ctx *c, int pot, const double *px, int64_t n, int64_t ld, int64_t lo, int64_t hi,
                        int mstride, double *acc_out, int64_t ld_out, bool accumulate)
{
    APParams p{};
    p.src[0] = px; p.src[1] = px + ld; p.src[2] = px + 2 * ld; p.src[3] = c->charge;
    p.nsrc_pad = (int)(((n + kPad - 1) / kPad) * kPad);
    p.n = (int)n;
    p.tgt_lo = (int)lo;
    p.ntgt = (int)(hi - lo);
    if (p.ntgt <= 0) return NBX_OK;
    switch (pot) {
    case 0: return run_cutoff<0, 0>(c, px, ld, c->lj_R2, p, c->mass, mstride, acc_out, ld_out, accumulate);
    case 1: return run_cutoff<1, 0>(c, px, ld, c->el_R2, p, c->mass, 1, acc_out, ld_out, accumulate);
    case 2: return run_cutoff<1, 1>(c, px, ld, c->el_R2, p, c->mass, 1, acc_out, ld_out, accumulate);
    default: return fail(c, NBX_ERR_INVALID, "launch_allpairs_pbc: bad potential id %d", pot);
    }
}

} // namespace nbx
