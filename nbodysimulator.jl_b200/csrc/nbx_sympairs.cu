// nbx_sympairs.cu -- all-pairs 1/r^2 central forces with Newton's third law (sm_100a).
//
// Same result as the ordered kernel of nbx_allpairs.cu for gravitational_acceleration!
// (src/basic_potentials.jl:306-331) and the unbounded Coulomb case (:274-304), but every UNORDERED
// pair {i,j} is evaluated once and applied to both bodies: 20 FP64-pipe instructions per unordered
// pair (18 when all weights are equal) instead of 2 x 16.
//
// Decomposition (half ring over tiles): the padded index range is cut into NT tiles of TS = 128*T
// bodies.  Tile A interacts with tiles B = A+k (mod NT), k = 1..NT/2 (for even NT, k = NT/2 only for
// A < NT/2, so each unordered tile pair is visited once) and with itself (k = 0, ordered, self-masked).
// Multi-GPU: rank r of R owns the ring offsets k = r, r+R, r+2R, ... ("pair sharding"); every rank then
// holds a partial acceleration for ALL bodies and the host sums them with one reduce-scatter.
// A work item is (A, segment of the rank's offset list); a persistent grid walks the items round-robin.
//
// Inside a CTA (4 warps): every lane keeps T bodies of A stationary in registers (positions, weights,
// accumulators).  The B tile arrives in shared memory by TMA bulk copies (2-stage ring).  It is cut
// into sets of 32*U bodies; a warp loads one set (U bodies per lane) and passes it around the warp
// with shuffles: 32 ring steps visit all (32T) x (32U) pairs, the B-side accumulators travel with the
// bodies.  The four warps' B-side sums of a set are added in a fixed order through shared memory and
// written to the partial slot of that ring offset; the A-side sums go to the slot of the item's
// segment.  sym_reduce_kernel adds all slots of a body in ascending slot order -> bit-reproducible.
#include "nbx_internal.cuh"

#include <cstring>

namespace nbx {

struct SymParams {
    const double *x, *y, *z, *w;
    int n, npad;
    int NT, K;             // tiles; largest ring offset
    int kr0, kstride, M;   // this rank's offsets k(m) = kr0 + m * kstride for m < M0; m = M0 (if M > M0): the half offset K
    int M0;                // of an even ring, which EVERY rank takes a share of (tiles A = rank mod kstride): balanced shares
    int S, seg_len;        // segments of the local offset list
    double *part;          // [(S + M)][3][npad]: A-side slots per segment, then B-side slots per local offset
    // periodic cutoff variant (PBC = 1): cubic box edge, half edge, squared cutoff, high word of the half edge
    double L, radius, R2;
    int hi_radius;
    const int *gate;       // the kernel runs iff gate == nullptr or gate[0] == gate_want (fast / general periodic variant)
    int gate_want;
    int *ticket;           // work items are handed out by this counter (zeroed on the stream before the launch)
};

// PBC = 1: Coulomb with a cutoff under CubicPeriodicBoundaryConditions (src/basic_potentials.jl:288-297 with the distance of
// src/boundary_conditions.jl:138-165): rij = ri - rj in the reference's orientation (i = the stationary body), the compare-
// and-subtract wrap, un-fused r2, strict r2 < R2.  For R < L/2 no accepted pair has a component on the +-L/2 tie, so the
// partner's own evaluation, wrap(rj - ri), is exactly -wrap(ri - rj): one evaluation serves both bodies with the reference's
// pair set.  EXCL3: pairs inside one molecule (i/3 == j/3) are skipped (src/nbody_to_ode.jl:331-351).
template <int T, int U, bool UNIFORM, bool DIAG, int UNR, int PBC = 0, bool EXCL3 = false>
__device__ __forceinline__ void ring_pass(const double (&ax)[T], const double (&ay)[T], const double (&az)[T],
                                          const double (&aw)[T], double (&fx)[T], double (&fy)[T], double (&fz)[T],
                                          double (&bx)[U], double (&by)[U], double (&bz)[U], double (&bw)[U],
                                          double (&gx)[U], double (&gy)[U], double (&gz)[U], int lane, int ibase,
                                          int jbase, const SymParams *pp = nullptr)
{
    const int src = (lane + 1) & 31;
#pragma unroll(UNR)
    for (int s = 0; s < 32; ++s) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int t = 0; t < T; ++t) {
                double dx, dy, dz, r2;
                bool ok = true;
                if (PBC == 2) {
                    // every coordinate lies within [-L/4, 5L/4) (checked on the device before the launch): |ri - rj| < 3L/2,
                    // so ONE round of the reference's wrap loop settles every component; a component exactly on the tie
                    // (-L/2 stays, +L/2 wraps) follows the loop's own conditions.  No branch, no integer work.
                    double x = __dsub_rn(ax[t], bx[u]), y = __dsub_rn(ay[t], by[u]), z = __dsub_rn(az[t], bz[u]);
                    // (one compare of |c| per component: a component exactly on the tie -L/2, which the loop leaves alone,
                    // may be folded to +L/2 here -- either way r2 >= L^2/4 > R2 and the pair is outside)
                    const int Lh = __double2hiint(pp->L), Ll = __double2loint(pp->L);
                    if (fabs(x) >= pp->radius) x = __dsub_rn(x, __hiloint2double(Lh | (__double2hiint(x) & 0x80000000), Ll));
                    if (fabs(y) >= pp->radius) y = __dsub_rn(y, __hiloint2double(Lh | (__double2hiint(y) & 0x80000000), Ll));
                    if (fabs(z) >= pp->radius) z = __dsub_rn(z, __hiloint2double(Lh | (__double2hiint(z) & 0x80000000), Ll));
                    r2 = r2_unfused(x, y, z);
                    ok = __double_as_longlong(r2) < __double_as_longlong(pp->R2); // (r2 >= 0: IEEE order == order of the bit patterns)
                    if (EXCL3) {
                        const int j = jbase + u * 32 + ((lane + s) & 31);
                        const int i = ibase + t * 32 + lane;
                        ok = ok && (i / 3 != j / 3);
                    }
                    dx = -x; dy = -y; dz = -z;
                    r2 = ok ? r2 : 1.0;
                } else if (PBC) {
                    const double x0 = __dsub_rn(ax[t], bx[u]), y0 = __dsub_rn(ay[t], by[u]), z0 = __dsub_rn(az[t], bz[u]);
                    // In an all-pairs sweep most pairs need the wrap (any component beyond L/2), so its first round is
                    // branch-free: |c| > L/2 is decided on the high words, c -+ L is the loop's own rounded subtraction.
                    // High words equal to the half edge's, or a result still beyond it (coordinates that drifted by more
                    // than a box), take the reference's loops (rare).
                    const int wx = __double2hiint(x0), wy = __double2hiint(y0), wz = __double2hiint(z0);
                    const int hx = wx & 0x7fffffff, hy = wy & 0x7fffffff, hz = wz & 0x7fffffff;
                    const int hm = max(hx, max(hy, hz));
                    ok = hm < 0x5d000000; // a padding body (parked at 1e150): weight 0, never in the cutoff
                    const int Lh = __double2hiint(pp->L), Ll = __double2loint(pp->L);
                    const double xw = __dsub_rn(x0, __hiloint2double(Lh | (wx & 0x80000000), Ll));
                    const double yw = __dsub_rn(y0, __hiloint2double(Lh | (wy & 0x80000000), Ll));
                    const double zw = __dsub_rn(z0, __hiloint2double(Lh | (wz & 0x80000000), Ll));
                    double x = hx > pp->hi_radius ? xw : x0, y = hy > pp->hi_radius ? yw : y0, z = hz > pp->hi_radius ? zw : z0;
                    const int nx = __double2hiint(x) & 0x7fffffff, ny = __double2hiint(y) & 0x7fffffff, nz = __double2hiint(z) & 0x7fffffff;
                    if (ok && max(nx, max(ny, nz)) >= pp->hi_radius) {
                        x = wrap_cubic(x0, pp->radius, pp->L); y = wrap_cubic(y0, pp->radius, pp->L); z = wrap_cubic(z0, pp->radius, pp->L);
                    }
                    r2 = r2_unfused(x, y, z);
                    ok = ok && (__double_as_longlong(r2) < __double_as_longlong(pp->R2));
                    if (EXCL3) {
                        const int j = jbase + u * 32 + ((lane + s) & 31);
                        const int i = ibase + t * 32 + lane;
                        ok = ok && (i / 3 != j / 3);
                    }
                    dx = -x; dy = -y; dz = -z;
                    r2 = ok ? r2 : 1.0;
                } else {
                    dx = bx[u] - ax[t]; dy = by[u] - ay[t]; dz = bz[u] - az[t];
                    r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                }
                if (DIAG) {
                    // body index of source u at this step vs target t (same tile)
                    const int j = jbase + u * 32 + ((lane + s) & 31);
                    const int i = ibase + t * 32 + lane;
                    r2 = (i == j) ? 1.0 : r2; // dx = dy = dz = 0 -> contributes exactly 0
                }
                const double y0 = rsqrt_seed(r2);
                const double a = y0 * y0;
                const double e = fma(-r2, a, 1.0);
                const double y3 = a * y0;
                const double p = fma(1.875, e, 1.5);
                const double q = p * e;
                double g = fma(y3, q, y3); // r^-3 to ~3 ulp
                if (PBC) g = ok ? g : 0.0;
                if (UNIFORM) {
                    fx[t] = fma(g, dx, fx[t]); fy[t] = fma(g, dy, fy[t]); fz[t] = fma(g, dz, fz[t]);
                    if (!DIAG) { gx[u] = fma(-g, dx, gx[u]); gy[u] = fma(-g, dy, gy[u]); gz[u] = fma(-g, dz, gz[u]); }
                } else {
                    const double sj = g * bw[u];
                    fx[t] = fma(sj, dx, fx[t]); fy[t] = fma(sj, dy, fy[t]); fz[t] = fma(sj, dz, fz[t]);
                    if (!DIAG) {
                        const double si = -(g * aw[t]);
                        gx[u] = fma(si, dx, gx[u]); gy[u] = fma(si, dy, gy[u]); gz[u] = fma(si, dz, gz[u]);
                    }
                }
            }
        }
        // pass the B bodies (and their accumulators) to the neighbouring lane
#pragma unroll
        for (int u = 0; u < U; ++u) {
            bx[u] = __shfl_sync(0xffffffffu, bx[u], src);
            by[u] = __shfl_sync(0xffffffffu, by[u], src);
            bz[u] = __shfl_sync(0xffffffffu, bz[u], src);
            if (!UNIFORM) bw[u] = __shfl_sync(0xffffffffu, bw[u], src);
            if (!DIAG) {
                gx[u] = __shfl_sync(0xffffffffu, gx[u], src);
                gy[u] = __shfl_sync(0xffffffffu, gy[u], src);
                gz[u] = __shfl_sync(0xffffffffu, gz[u], src);
            }
        }
    }
    // after 32 passes every body (and its accumulator) is back in the lane that loaded it
}

// offset k of tile A is skipped when it would visit a tile pair twice (even NT, k = NT/2, upper half); the lower half of
// that offset is dealt out over the ranks by tile (a whole offset more on one rank would cost it 1 / (2 offsets per rank))
__host__ __device__ __forceinline__ bool sym_live(int k, int A, int NT, int K, int rank, int nranks)
{
    if ((NT & 1) == 0 && k == K && K > 0) return A < K && (A % nranks) == rank;
    return true;
}

template <int T, int U, bool UNIFORM, int MINB, int UNR, int PBC = 0, bool EXCL3 = false>
__global__ void __launch_bounds__(128, MINB) sym_kernel(const SymParams p)
{
    constexpr int TS = 128 * T;       // bodies per tile
    constexpr int SET = 32 * U;       // bodies per ring set
    constexpr int NSET = TS / SET;
    constexpr uint32_t STAGE_BYTES = 4 * TS * sizeof(double);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *ring = reinterpret_cast<double *>(smem_raw);      // [2][4][TS]
    double *red = ring + 2 * 4 * TS;                            // [4 warps][3][SET]
    __shared__ __align__(8) uint64_t full[2];

    __shared__ int s_item[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (PBC && p.gate && p.gate[0] != p.gate_want) return;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
        // Work items (tile A, segment of ring offsets) are taken from a counter, two at a time: the one in hand and the next,
        // whose first B tile is fetched during the last tile of the current one.  Items differ in cost (the half offset of an
        // even ring, the diagonal block, dead items of a rank's share), and a rank of 8 has only ~14 tiles per CTA: the
        // static round-robin of r01 cost it 16 tile times.  Every item writes its own slots: the sums do not depend on
        // which CTA took it.
        s_item[0] = atomicAdd(p.ticket, 1);
        s_item[1] = atomicAdd(p.ticket, 1);
    }
    __syncthreads();
    uint32_t parity = 0;

    const int nitems = p.NT * p.S;
    auto kof = [&](int m) { return m < p.M0 ? p.kr0 + m * p.kstride : p.K; };
    // B tile of (tile A, local offset index m) -> stage st
    auto issue = [&](int A, int m, int st) {
        const int B0 = ((A + kof(m)) % p.NT) * TS;
        double *dst = ring + (size_t)st * 4 * TS;
        mbar_expect_tx(&full[st], STAGE_BYTES);
        bulk_g2s(dst, p.x + B0, TS * sizeof(double), &full[st]);
        bulk_g2s(dst + TS, p.y + B0, TS * sizeof(double), &full[st]);
        bulk_g2s(dst + 2 * TS, p.z + B0, TS * sizeof(double), &full[st]);
        bulk_g2s(dst + 3 * TS, p.w + B0, TS * sizeof(double), &full[st]);
    };
    auto next_live = [&](int A, int m, int mend) { while (m < mend && !sym_live(kof(m), A, p.NT, p.K, p.kr0, p.kstride)) ++m; return m; };
    int st = 0;              // ring stage of the next tile: runs on across items
    bool have_first = false; // the previous item already issued this item's first tile (a short segment must not
                             // expose the copy latency once per item: multi-GPU shares have two offsets per segment)
    int item = s_item[0], item_next = s_item[1];
    __syncthreads();
    while (item < nitems) {
        const int A = item / p.S, seg = item - A * p.S;
        const int m0 = seg * p.seg_len;
        const int m1 = min(m0 + p.seg_len, p.M);
        const int A0 = A * TS;

        double ax[T], ay[T], az[T], aw[T], fx[T], fy[T], fz[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int i = A0 + warp * (32 * T) + t * 32 + lane;
            ax[t] = p.x[i]; ay[t] = p.y[i]; az[t] = p.z[i];
            aw[t] = UNIFORM ? 1.0 : p.w[i];
            fx[t] = fy[t] = fz[t] = 0.0;
        }

        int m = next_live(A, m0, m1);
        if (!have_first && tid == 0 && m < m1) issue(A, m, st);
        have_first = false;
        while (m < m1) {
            const int mn = next_live(A, m + 1, m1);
            if (mn < m1) {
                if (tid == 0) issue(A, mn, st ^ 1); // stage st^1 was released by the barrier that ended the previous block
            } else {
                const int item2 = item_next; // last tile of this item: fetch the first tile of the next one meanwhile
                if (item2 < nitems) {
                    const int A2 = item2 / p.S, seg2 = item2 - A2 * p.S;
                    const int e2 = min((seg2 + 1) * p.seg_len, p.M);
                    const int m2 = next_live(A2, seg2 * p.seg_len, e2);
                    if (m2 < e2) {
                        if (tid == 0) issue(A2, m2, st ^ 1);
                        have_first = true;
                    }
                }
            }
            mbar_wait(&full[st], (parity >> st) & 1u);
            parity ^= 1u << st;
            const double *bsx = ring + (size_t)st * 4 * TS, *bsy = bsx + TS, *bsz = bsy + TS, *bsw = bsz + TS;
            const int k = kof(m);
            const int B0 = ((A + k) % p.NT) * TS;
            const int ibase = A0 + warp * (32 * T);
            for (int set = 0; set < NSET; ++set) {
                double bx[U], by[U], bz[U], bw[U], gx[U], gy[U], gz[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int o = set * SET + u * 32 + lane;
                    bx[u] = bsx[o]; by[u] = bsy[o]; bz[u] = bsz[o];
                    bw[u] = UNIFORM ? 1.0 : bsw[o];
                    gx[u] = gy[u] = gz[u] = 0.0;
                }
                if (k == 0) {
                    ring_pass<T, U, UNIFORM, true, 1, PBC, EXCL3>(ax, ay, az, aw, fx, fy, fz, bx, by, bz, bw, gx, gy, gz, lane, ibase,
                                                                  B0 + set * SET, &p);
                } else {
                    // (molecules only straddle ADJACENT tiles: the exclusion test is compiled into the k = 1 pass alone)
                    if (EXCL3 && k == 1)
                        ring_pass<T, U, UNIFORM, false, 1, PBC, true>(ax, ay, az, aw, fx, fy, fz, bx, by, bz, bw, gx, gy, gz, lane,
                                                                      ibase, B0 + set * SET, &p);
                    else
                        ring_pass<T, U, UNIFORM, false, UNR, PBC, false>(ax, ay, az, aw, fx, fy, fz, bx, by, bz, bw, gx, gy, gz, lane,
                                                                         ibase, B0 + set * SET, &p);
                    // B-side sums of the four warps, added in warp order, go to the slot of this ring offset
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        red[(warp * 3 + 0) * SET + u * 32 + lane] = gx[u];
                        red[(warp * 3 + 1) * SET + u * 32 + lane] = gy[u];
                        red[(warp * 3 + 2) * SET + u * 32 + lane] = gz[u];
                    }
                    __syncthreads();
                    double *slot = p.part + (size_t)(p.S + m) * 3 * p.npad;
                    for (int e = tid; e < 3 * SET; e += 128) {
                        const int c = e / SET, o = e - c * SET;
                        const double s = ((red[(0 * 3 + c) * SET + o] + red[(1 * 3 + c) * SET + o]) +
                                          red[(2 * 3 + c) * SET + o]) + red[(3 * 3 + c) * SET + o];
                        slot[(size_t)c * p.npad + B0 + set * SET + o] = s;
                    }
                    __syncthreads();
                }
            }
            __syncthreads(); // all warps are done with stage st
            st ^= 1;
            m = mn;
        }

        double *slot = p.part + (size_t)seg * 3 * p.npad;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int i = A0 + warp * (32 * T) + t * 32 + lane;
            slot[i] = fx[t];
            slot[(size_t)p.npad + i] = fy[t];
            slot[2 * (size_t)p.npad + i] = fz[t];
        }
        if (tid == 0) s_item[0] = atomicAdd(p.ticket, 1);
        __syncthreads();
        item = item_next;
        item_next = s_item[0];
        __syncthreads();
    }
}

// acc_i (+)= f_i * (A-side slots in segment order, then the B-side slots of the offsets that reached body i)
__global__ void sym_reduce_kernel(const SymParams p, int TS, int kind, double scale, const double *__restrict__ mass,
                                  const double *__restrict__ charge, double *__restrict__ ax, double *__restrict__ ay,
                                  double *__restrict__ az, int accumulate)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const int tile = i / TS;
    const size_t np = (size_t)p.npad;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int sl = 0; sl < p.S; ++sl) {
        const double *b = p.part + (size_t)sl * 3 * np;
        s0 += b[i]; s1 += b[np + i]; s2 += b[2 * np + i];
    }
    for (int m = 0; m < p.M; ++m) {
        const int k = m < p.M0 ? p.kr0 + m * p.kstride : p.K;
        if (k == 0) continue;                                       // the diagonal block has no B side
        const int A = (tile - k % p.NT + p.NT) % p.NT;               // the tile that visited us at this offset
        if (!sym_live(k, A, p.NT, p.K, p.kr0, p.kstride)) continue;
        const double *b = p.part + (size_t)(p.S + m) * 3 * np;
        s0 += b[i]; s1 += b[np + i]; s2 += b[2 * np + i];
    }
    double f = scale;
    if (kind == 1) f = scale * charge[i] / mass[i];
    if (accumulate) { ax[i] += f * s0; ay[i] += f * s1; az[i] += f * s2; }
    else { ax[i] = f * s0; ay[i] = f * s1; az[i] = f * s2; }
}

template <int T, int U, bool UNIFORM, int MINB, int UNR, int PBC = 0, bool EXCL3 = false>
static int run_sym(nbx_ctx *c, const double *w, double wval, int scale_kind, double scale, double *acc_out,
                   bool accumulate, double R2 = 0.0, const int *gate = nullptr, int gate_want = 0, bool reduce = true)
{
    constexpr int TS = 128 * T, SET = 32 * U;
    SymParams p{};
    p.x = c->pos; p.y = c->pos + c->npad; p.z = c->pos + 2 * c->npad; p.w = w;
    p.n = (int)c->n; p.npad = (int)c->npad;
    p.NT = p.npad / TS;
    p.K = p.NT / 2;
    p.kr0 = c->pair_rank;
    p.kstride = c->pair_nranks;
    if (PBC) {
        p.L = c->bc[0]; p.radius = 0.5 * c->bc[0]; p.R2 = R2;
        int64_t bits;
        memcpy(&bits, &p.radius, sizeof bits);
        p.hi_radius = (int)(bits >> 32);
        p.gate = gate; p.gate_want = gate_want;
    }
    const bool half = (p.NT % 2 == 0) && p.K > 0;      // even ring: offset K is half an offset, shared by all ranks
    const int kfull = half ? p.K - 1 : p.K;            // largest whole offset
    p.M0 = p.kr0 <= kfull ? (kfull - p.kr0) / p.kstride + 1 : 0;
    p.M = p.M0 + (half ? 1 : 0);
    const int grid_full = c->sm_count * MINB;
    // Segment length: the CTAs take items from a counter, so a launch lasts about (all tiles) / grid + one item + a fixed
    // cost per item (A-side loads and slot writes; measured small: 262,144 bodies, whole ring of 129 offsets, ms per launch
    // for 1 / 2 / 3 / 5 offsets per item: 40.13 / 40.31 / 40.37 / 40.50 -- but every segment is another 6 MB slot for the
    // reduction to read; one rank of eight, 17 offsets: 5.19 / 5.35 / 5.66; the static round-robin of r01: 42.04 and 5.63)
    int best_len = 1;
    double best_cost = 1e300;
    for (int len = 1; len <= 16 && len <= (p.M > 0 ? p.M : 1); ++len) {
        const double segs = (double)((p.M + len - 1) / len);
        const double cost = (double)p.NT * p.M / grid_full + len + 0.05 * (double)p.NT * segs / grid_full;
        if (cost < best_cost - 1e-9) { best_cost = cost; best_len = len; }
    }
    if (c->opt_sym_seg_len > 0) best_len = c->opt_sym_seg_len;
    p.seg_len = best_len;
    p.S = p.M > 0 ? (p.M + p.seg_len - 1) / p.seg_len : 1;
    if (!c->sym_ticket) NBX_TRY(dev_alloc(c, &c->sym_ticket, (size_t)2));
    p.ticket = c->sym_ticket;
    NBX_CUDA(c, cudaMemsetAsync(c->sym_ticket, 0, sizeof(int), c->stream));
    const size_t bytes = (size_t)(p.S + p.M) * 3 * p.npad * sizeof(double);
    if (bytes > c->part_bytes) {
        if (c->part) { cudaFree(c->part); c->part = nullptr; c->part_bytes = 0; }
        cudaError_t e = cudaMalloc((void **)&c->part, bytes);
        if (e != cudaSuccess) return cuda_fail(c, e, "cudaMalloc(symmetric partials)");
        c->part_bytes = bytes;
    }
    p.part = c->part;
    auto kern = sym_kernel<T, U, UNIFORM, MINB, UNR, PBC, EXCL3>;
    const size_t smem = (size_t)(2 * 4 * TS + 4 * 3 * SET) * sizeof(double);
    const void *key = reinterpret_cast<const void *>(kern);
    bool done = false;
    for (const void *k : c->attr_done) done = done || (k == key);
    if (!done) {
        NBX_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        c->attr_done.push_back(key);
    }
    const long items = (long)p.NT * p.S;
    const int grid = (int)(items < grid_full ? items : grid_full);
    timer_begin(c, NBX_T_PAIR_ALLPAIRS);
    kern<<<grid, 128, smem, c->stream>>>(p);
    timer_end(c, NBX_T_PAIR_ALLPAIRS);
    NBX_CUDA(c, cudaGetLastError());
    c->last_grid = grid;
    c->last_nchunk = p.S;
    double f = scale;
    if (UNIFORM) f *= wval;
    if (reduce) sym_reduce_kernel<<<(p.n + 255) / 256, 256, 0, c->stream>>>(p, TS, scale_kind, f, c->mass, c->charge, acc_out,
                                                               acc_out + c->npad, acc_out + 2 * c->npad,
                                                               accumulate ? 1 : 0);
    NBX_CUDA(c, cudaGetLastError());
    return NBX_OK;
}

// Evaluates this rank's share of the unordered pairs (everything when pair_nranks == 1) and writes /
// accumulates the (partial) accelerations of ALL n bodies.  uniform = all weights equal to wval.
int launch_sympairs(nbx_ctx *c, const double *w, bool uniform, double wval, int scale_kind, double scale,
                    double *acc_out, bool accumulate)
{
    // variants (option "sym_variant"): T stationary and U travelling bodies per lane, CTAs per SM, ring unroll.
    // Measured on the 262,144-body sphere, equal masses (r01c, ms per launch): <8,2,2,4> 42.7 (default), <8,1,2,2> 43.5,
    // <8,2,2,2> 44.2, <8,2,2,1> 44.2, <8,4,2,1> 44.9, <8,1,3,4> 46.8, <4,4,3,2> 47.2; ordered kernel 70.8.  Deeper unrolls
    // lose: <8,2,2,8> 48.8, <8,2,2,16> 59.1 (instruction cache); <8,1,2,8> 42.7 ties with the default.
    switch (c->opt_sym_variant) {
    case 1:
        if (uniform) return run_sym<4, 4, true, 3, 2>(c, w, wval, scale_kind, scale, acc_out, accumulate);
        return run_sym<4, 4, false, 3, 2>(c, w, wval, scale_kind, scale, acc_out, accumulate);
    case 2:
        if (uniform) return run_sym<8, 2, true, 2, 2>(c, w, wval, scale_kind, scale, acc_out, accumulate);
        return run_sym<8, 2, false, 2, 2>(c, w, wval, scale_kind, scale, acc_out, accumulate);
    case 3:
        if (uniform) return run_sym<8, 1, true, 3, 4>(c, w, wval, scale_kind, scale, acc_out, accumulate);
        return run_sym<8, 1, false, 3, 4>(c, w, wval, scale_kind, scale, acc_out, accumulate);
    case 4:
        if (uniform) return run_sym<8, 2, true, 2, 1>(c, w, wval, scale_kind, scale, acc_out, accumulate);
        return run_sym<8, 2, false, 2, 1>(c, w, wval, scale_kind, scale, acc_out, accumulate);
    case 5:
        if (uniform) return run_sym<8, 1, true, 2, 2>(c, w, wval, scale_kind, scale, acc_out, accumulate);
        return run_sym<8, 1, false, 2, 2>(c, w, wval, scale_kind, scale, acc_out, accumulate);
    case 6:
        if (uniform) return run_sym<8, 4, true, 2, 1>(c, w, wval, scale_kind, scale, acc_out, accumulate);
        return run_sym<8, 4, false, 2, 1>(c, w, wval, scale_kind, scale, acc_out, accumulate);
    case 7:
        if (uniform) return run_sym<8, 1, true, 2, 8>(c, w, wval, scale_kind, scale, acc_out, accumulate);
        return run_sym<8, 1, false, 2, 8>(c, w, wval, scale_kind, scale, acc_out, accumulate);
    default:
        if (uniform) return run_sym<8, 2, true, 2, 4>(c, w, wval, scale_kind, scale, acc_out, accumulate);
        return run_sym<8, 2, false, 2, 4>(c, w, wval, scale_kind, scale, acc_out, accumulate);
    }
}

// flag[0] = 1 when some coordinate lies outside [-L/4, 5L/4) (then |ri - rj| can reach 3L/2 and one wrap round is not enough)
__global__ void pbc_window_kernel(const double *__restrict__ pos, int64_t ld, int n, double lo, double hi, int *__restrict__ flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = pos[i], y = pos[ld + i], z = pos[2 * ld + i];
    if (!(x >= lo && x < hi && y >= lo && y < hi && z >= lo && z < hi)) flag[0] = 1;
}

// Coulomb with a cutoff in a cubic periodic box as all UNORDERED pairs (the boxes a cell list cannot serve: R >= L/3):
// half the evaluations of the ordered kernel with the exact periodic predicate.  excl3: own-molecule exclusion (water).
// The caller has checked R < L/2 and that the context evaluates all targets.  Two launches, one of which returns at
// once: the branch-free variant while every coordinate is within a quarter box of the cell, else the general one.
int launch_sympairs_coulomb_pbc(nbx_ctx *c, bool excl3, double *acc_out, bool accumulate)
{
    NBX_TRY(ensure_red(c));
    int *flag = reinterpret_cast<int *>(c->d_red + c->red_cap - 4); // [0] drifted, [1] scratch (the tail of the reduction scratch)
    const double L = c->bc[0];
    NBX_CUDA(c, cudaMemsetAsync(flag, 0, 2 * sizeof(int), c->stream));
    pbc_window_kernel<<<(unsigned)((c->n + 255) / 256), 256, 0, c->stream>>>(c->pos, c->npad, (int)c->n, -0.25 * L, 1.25 * L, flag);
    NBX_CUDA(c, cudaGetLastError());
    if (excl3) {
        // (other shapes of the branch-free variant, r02, ms per water step at 98,304 atoms against 10.2 for this one:
        // <4,4,3,2> 11.2, <8,2,2,4> 13.1, <8,1,3,4> 13.7, <8,2,2,1> 10.3, <8,1,2,2> 10.4, <4,2,4,2> 11.7, <4,2,3,4> 10.2)
        NBX_TRY((run_sym<8, 2, false, 2, 2, 2, true>(c, c->charge, 1.0, 1, -c->el_k, acc_out, accumulate, c->el_R2, flag, 0, false)));
        return run_sym<8, 2, false, 2, 2, 1, true>(c, c->charge, 1.0, 1, -c->el_k, acc_out, accumulate, c->el_R2, flag, 1, true);
    }
    NBX_TRY((run_sym<8, 2, false, 2, 2, 2, false>(c, c->charge, 1.0, 1, -c->el_k, acc_out, accumulate, c->el_R2, flag, 0, false)));
    return run_sym<8, 2, false, 2, 2, 1, false>(c, c->charge, 1.0, 1, -c->el_k, acc_out, accumulate, c->el_R2, flag, 1, true);
}

} // namespace nbx
