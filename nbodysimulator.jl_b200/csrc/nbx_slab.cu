// nbx_slab.cu -- slab decomposition of cutoff systems over one process per GPU (sm_100a).
//
// The reference is serial; this is the multi-GPU form of the target loop of soode_system!
// (src/nbody_to_ode.jl:474-488) for short-range potentials in a CubicPeriodicBoundaryConditions box
// (src/boundary_conditions.jl:101-103).  The box is cut along x into slabs of whole cell layers (the
// cells of the binning grid, edge >= cutoff).  A rank OWNS the particles whose wrapped x lies in its
// layers [c0, c1) and keeps, as GHOSTS, copies of the positions of the two adjacent layers.  Per step:
//
//   nbx_vv_begin         x += dt v + dt^2/2 a on the own particles
//   nbx_slab_pack        classify the own particles; stayers are compacted (stable) into the alternate
//                        state arrays, leavers go with their full state into the send buffer of the
//                        neighbour they moved to, the own boundary layers go as halo records
//   (host)               send-to-left/right <-> recv-from-right/left, one message per neighbour
//   nbx_slab_unpack      arrivals are appended to the own particles, then the ghost set is laid down:
//                        the leavers just sent (they sit in the neighbour's boundary layer now) and the
//                        two received halos
//   nbx_vv_forces/finish cell list over own + ghosts in the GLOBAL grid, pair kernel on the own targets
//
// Every local particle carries its global id; the cell order is ranked by that id (nbx_cells.cu), so the
// force sums are bit-identical to the single-GPU result whatever the local numbering.
// The particle counts never visit the host inside a step: they live in d_n ([0] own, [1] ghosts), every kernel of
// the step is launched for the capacity bound cap_loc and clamps to d_n (nbx_ctx::dyn), and errors (a particle that
// jumped past the neighbouring slab, a full message or state buffer) accumulate in d_n[2..3] until
// nbx_slab_check / nbx_slab_unpack(counts != NULL) reads them.  A step is therefore a pure stream of launches
// and copies, i.e. capturable in a CUDA graph together with the NCCL send/recv of the host layer.
// Exchange, direct mode (nbx_slab_connect): the receive area of a rank (flags + 2 x 2 message buffers, double
// buffered by message parity) is mapped into its neighbours (CUDA IPC over NVLink, or plain pointers inside one
// process).  The pack kernel then stores its records straight into the neighbours' memory; the last block to
// finish fences and raises the neighbours' flags, and the unpack kernel of the neighbour spins on its flags.
// Compute and transfer are one kernel, nothing is staged, no collective library and no host in the loop.
// Without a connection the host moves the send buffers (NCCL / gloo send-recv, or a copy) -- same layout.
// A slab needs >= 2 layers so that one exchange per step suffices (an arrival from the left lands in the
// leftmost layer and is a ghost of the left neighbour only, which kept it).
#include "nbx_internal.cuh"

#include <algorithm>
#include <cstring>

namespace nbx {

constexpr int kHdr = 8;    // message header doubles: [0] migrants, [1] halo records
constexpr int kMigW = 12;  // gid, x y z, vx vy vz, ax ay az, m, q
constexpr int kHaloW = 5;  // gid, x y z, q
constexpr int kSlabBlock = 256;

enum { CAT_STAY = 0, CAT_MIGL = 1, CAT_MIGR = 2, CAT_HALOL = 3, CAT_HALOR = 4, CAT_N = 5 };
// device counters
enum { CNT_ARRL = 5, CNT_ARRR = 6, CNT_N = 8 };
// persistent counters d_n
enum { DN_OWN = 0, DN_GHOST = 1, DN_LOST = 2, DN_CAP = 3, DN_MSG = 4, DN_TICKET = 5, DN_TIMEOUT = 6, DN_HALOL = 8, DN_HALOR = 9,
       DN_RHALOL = 10, DN_RHALOR = 11, // halo records RECEIVED from the left / right at the last rebuild
       DN_N = 12 };
constexpr int kLLWords = 6; // x, y, z as LL words
constexpr int kRxHdr = 16; // doubles in front of the receive area: [0] 'message m from the left is complete', [1] same from the right


struct SlabGeom {
    double L;
    int nc, c0, width, wL, wR; // own layers [c0, c0 + width), neighbour widths
    int init;                  // 1: particles outside the own layers are dropped, not migrated
};

__device__ __forceinline__ int slab_layer(double x, double L, int nc)
{
    double w = x - L * floor(x / L);
    if (w < 0.0) w += L;
    if (w >= L) w -= L;
    int cx = (int)(w * ((double)nc / L));
    return cx < 0 ? 0 : (cx >= nc ? nc - 1 : cx); // same expression as the binning of nbx_cells.cu
}

// bit c set: the particle counts in category c;  bit 8: lost (moved past the neighbouring slab)
__device__ __forceinline__ unsigned slab_classify(double x, const SlabGeom &g)
{
    const int layer = slab_layer(x, g.L, g.nc);
    int rel = layer - g.c0;
    if (rel < 0) rel += g.nc;
    if (rel < g.width) {
        unsigned f = 1u << CAT_STAY;
        if (rel == 0) f |= 1u << CAT_HALOL;
        if (rel == g.width - 1) f |= 1u << CAT_HALOR;
        return f;
    }
    if (g.init) return 0u;
    const int dl = g.nc - rel, dr = rel - g.width + 1; // layers beyond the left / right face
    if (dl <= dr) return (1u << CAT_MIGL) | (dl > g.wL ? 256u : 0u);
    return (1u << CAT_MIGR) | (dr > g.wR ? 256u : 0u);
}

// cond (all kernels of the migration round): run iff cond == nullptr or cond[0] != 0 -- the collective rebuild decision taken
// on the device (slab_enqueue); skip (the halo refresh kernels): return iff skip && skip[0] != 0
__global__ void __launch_bounds__(kSlabBlock) slab_count_kernel(const double *__restrict__ px, int n, SlabGeom g,
                                                                int *__restrict__ blockcnt, int *__restrict__ dn,
                                                                const int *__restrict__ dyn, const int *__restrict__ cond)
{
    __shared__ int cnt[CAT_N];
    if (cond && !cond[0]) return;
    n = dyn_own(dyn, n);
    if (threadIdx.x < CAT_N) cnt[threadIdx.x] = 0;
    __syncthreads();
    const int i = blockIdx.x * kSlabBlock + threadIdx.x;
    const unsigned f = i < n ? slab_classify(px[i], g) : 0u;
#pragma unroll
    for (int c = 0; c < CAT_N; ++c) {
        const unsigned m = __ballot_sync(0xffffffffu, (f >> c) & 1u);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(&cnt[c], __popc(m));
    }
    if (f & 256u) atomicAdd(&dn[DN_LOST], 1);
    __syncthreads();
    if (threadIdx.x < CAT_N) blockcnt[blockIdx.x * CAT_N + threadIdx.x] = cnt[threadIdx.x];
}

// exclusive scan over the blocks, one warp per category; totals -> counts[0..4]
__global__ void slab_scan_kernel(const int *__restrict__ blockcnt, int *__restrict__ blockoff, int nb,
                                 int *__restrict__ counts, const int *__restrict__ cond)
{
    if (cond && !cond[0]) return;
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (c >= CAT_N) return;
    int carry = 0;
    for (int b0 = 0; b0 < nb; b0 += 32) {
        const int b = b0 + lane;
        const int v = b < nb ? blockcnt[b * CAT_N + c] : 0;
        int s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (b < nb) blockoff[b * CAT_N + c] = carry + s - v;
        carry += __shfl_sync(0xffffffffu, s, 31);
    }
    if (lane == 0) counts[c] = carry;
}

struct SlabArrays {
    double *pos, *vel, *acc, *mass, *charge; // SoA rows of stride ld (charge may be null)
    int *gid;                                // may be null: identity
};

__device__ __forceinline__ double *rx_msg(double *rx, int which, int parity, int64_t msg_doubles)
{
    return rx + kRxHdr + (size_t)(which * 2 + parity) * (size_t)msg_doubles; // which: 0 from-left, 1 from-right
}

// peerL / peerR: receive areas of the left / right neighbour (direct mode), else null.  A message to the left
// lands in the left neighbour's "from-right" buffer and vice versa.
__global__ void __launch_bounds__(kSlabBlock) slab_pack_kernel(SlabArrays src, SlabArrays dst, int64_t ld, int n,
                                                               SlabGeom g, const int *__restrict__ blockoff,
                                                               int *__restrict__ counts, double *__restrict__ sendL,
                                                               double *__restrict__ sendR, int capM, int capH,
                                                               int *__restrict__ dn, const int *__restrict__ dyn,
                                                               double *peerL, double *peerR, int64_t msg_doubles,
                                                               int *__restrict__ halo_idxL, int *__restrict__ halo_idxR,
                                                               const int *__restrict__ cond)
{
    __shared__ int wcnt[kSlabBlock / 32][CAT_N];
    __shared__ int last_block;
    if (cond && !cond[0]) return;
    n = dyn_own(dyn, n);
    const int msg = dn[DN_MSG] + 1; // number of the message this kernel produces (the last block publishes it)
    double *remL = peerL ? rx_msg(peerL, 1, msg & 1, msg_doubles) : nullptr;
    double *remR = peerR ? rx_msg(peerR, 0, msg & 1, msg_doubles) : nullptr;
    const int i = blockIdx.x * kSlabBlock + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned f = i < n ? slab_classify(src.pos[i], g) : 0u;
    int rank[CAT_N];
#pragma unroll
    for (int c = 0; c < CAT_N; ++c) {
        const unsigned m = __ballot_sync(0xffffffffu, (f >> c) & 1u);
        rank[c] = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) wcnt[warp][c] = __popc(m);
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CAT_N; ++c) {
        int base = blockoff[blockIdx.x * CAT_N + c];
        for (int w = 0; w < warp; ++w) base += wcnt[w][c];
        rank[c] += base;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const double hl0 = (double)min(counts[CAT_MIGL], capM), hl1 = (double)min(counts[CAT_HALOL], capH);
        const double hr0 = (double)min(counts[CAT_MIGR], capM), hr1 = (double)min(counts[CAT_HALOR], capH);
        sendL[0] = hl0; sendL[1] = hl1; sendR[0] = hr0; sendR[1] = hr1;
        if (remL) { remL[0] = hl0; remL[1] = hl1; }
        if (remR) { remR[0] = hr0; remR[1] = hr1; }
        if (counts[CAT_MIGL] > capM || counts[CAT_MIGR] > capM || counts[CAT_HALOL] > capH || counts[CAT_HALOR] > capH)
            dn[DN_CAP] = 1;
        if (halo_idxL) { dn[DN_HALOL] = (int)hl1; dn[DN_HALOR] = (int)hr1; }
    }
    if (i < n && f != 0u) {
        const double x = src.pos[i], y = src.pos[ld + i], z = src.pos[2 * ld + i];
        const int id = src.gid ? src.gid[i] : i;
        const double q = src.charge ? src.charge[i] : 0.0;
        if (f & (1u << CAT_STAY)) {
            const int d = rank[CAT_STAY];
            dst.pos[d] = x; dst.pos[ld + d] = y; dst.pos[2 * ld + d] = z;
            dst.vel[d] = src.vel[i]; dst.vel[ld + d] = src.vel[ld + i]; dst.vel[2 * ld + d] = src.vel[2 * ld + i];
            dst.acc[d] = src.acc[i]; dst.acc[ld + d] = src.acc[ld + i]; dst.acc[2 * ld + d] = src.acc[2 * ld + i];
            dst.mass[d] = src.mass[i];
            if (dst.charge) dst.charge[d] = q;
            dst.gid[d] = id;
            for (int side = 0; side < 2; ++side) {
                const int cat = side == 0 ? CAT_HALOL : CAT_HALOR;
                if (!(f & (1u << cat)) || rank[cat] >= capH) continue;
                // halo records go to the neighbour directly when connected, else into the send buffer
                double *out = side == 0 ? (remL ? remL : sendL) : (remR ? remR : sendR);
                double *rec = out + kHdr + (size_t)capM * kMigW + (size_t)rank[cat] * kHaloW;
                rec[0] = (double)id; rec[1] = x; rec[2] = y; rec[3] = z; rec[4] = q;
                if (halo_idxL) (side == 0 ? halo_idxL : halo_idxR)[rank[cat]] = d; // slot of this record from now on
            }
        } else {
            const int cat = (f & (1u << CAT_MIGL)) ? CAT_MIGL : CAT_MIGR;
            if (rank[cat] < capM) {
                const double v0 = src.vel[i], v1 = src.vel[ld + i], v2 = src.vel[2 * ld + i];
                const double a0 = src.acc[i], a1 = src.acc[ld + i], a2 = src.acc[2 * ld + i];
                const double m = src.mass[i];
                double *rem = cat == CAT_MIGL ? remL : remR;
                // the local copy stays: the migrant is a ghost here from now on (unpack reads it back)
                for (int copy = 0; copy < 2; ++copy) {
                    double *out = copy == 0 ? (cat == CAT_MIGL ? sendL : sendR) : rem;
                    if (!out) continue;
                    double *rec = out + kHdr + (size_t)rank[cat] * kMigW;
                    rec[0] = (double)id; rec[1] = x; rec[2] = y; rec[3] = z;
                    rec[4] = v0; rec[5] = v1; rec[6] = v2; rec[7] = a0; rec[8] = a1; rec[9] = a2;
                    rec[10] = m; rec[11] = q;
                }
            }
        }
    }
    // completion: every block fences its (remote) stores and takes a ticket; the last one publishes the message
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last_block = atomicAdd(&dn[DN_TICKET], 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (last_block && threadIdx.x == 0) {
        __threadfence_system();
        if (peerL) *reinterpret_cast<volatile long long *>(peerL + 1) = (long long)msg; // left's "from the right"
        if (peerR) *reinterpret_cast<volatile long long *>(peerR + 0) = (long long)msg; // right's "from the left"
        dn[DN_TICKET] = 0;
        dn[DN_MSG] = msg;
    }
}

// direct mode: wait until both neighbours have published message `msg` in this rank's receive area (thread 0 of every
// block spins, bounded by 10 s of %globaltimer: a dead neighbour must not hang the GPU).  Returns false on a timeout.
__device__ __forceinline__ bool slab_wait_messages(double *rx, int msg, unsigned long long timeout_ns)
{
    __shared__ int timed_out;
    if (threadIdx.x == 0) {
        timed_out = 0;
        const volatile long long *flag = reinterpret_cast<const volatile long long *>(rx);
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (flag[0] < msg || flag[1] < msg) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) { timed_out = 1; break; }
        }
        __threadfence_system();
    }
    __syncthreads();
    return !timed_out;
}

// ------------------------------------------------------------------------------------------------
// between rebuilds (Verlet lists inside the slab): nothing migrates and nothing is renumbered; the positions of the
// boundary-layer particles recorded at the rebuild travel in the same message layout, by the same flag protocol
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSlabBlock) slab_halo_send_kernel(SlabArrays src, int64_t ld, const int *__restrict__ idxL,
                                                                    const int *__restrict__ idxR, int *__restrict__ dn,
                                                                    double *__restrict__ sendL, double *__restrict__ sendR,
                                                                    int capM, int capH, double *peerL, double *peerR,
                                                                    int64_t msg_doubles, const int *__restrict__ skip)
{
    __shared__ int last_block;
    if (skip && skip[0]) return;
    const int msg = dn[DN_MSG] + 1;
    double *outL = peerL ? rx_msg(peerL, 1, msg & 1, msg_doubles) : sendL;
    double *outR = peerR ? rx_msg(peerR, 0, msg & 1, msg_doubles) : sendR;
    const int nL = dn[DN_HALOL], nR = dn[DN_HALOR];
    const int t = blockIdx.x * kSlabBlock + threadIdx.x;
    if (t == 0) {
        outL[0] = 0.0; outL[1] = (double)nL; outR[0] = 0.0; outR[1] = (double)nR;
        sendL[0] = 0.0; sendL[1] = (double)nL; sendR[0] = 0.0; sendR[1] = (double)nR; // no migrants kept as ghosts
    }
    const int side = t / capH, k = t - side * capH;
    if (side < 2 && k < (side == 0 ? nL : nR)) {
        const int i = (side == 0 ? idxL : idxR)[k];
        double *rec = (side == 0 ? outL : outR) + kHdr + (size_t)capM * kMigW + (size_t)k * kHaloW;
        rec[0] = (double)(src.gid ? src.gid[i] : i);
        rec[1] = src.pos[i]; rec[2] = src.pos[ld + i]; rec[3] = src.pos[2 * ld + i];
        rec[4] = src.charge ? src.charge[i] : 0.0;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last_block = atomicAdd(&dn[DN_TICKET], 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (last_block && threadIdx.x == 0) {
        __threadfence_system();
        if (peerL) *reinterpret_cast<volatile long long *>(peerL + 1) = (long long)msg;
        if (peerR) *reinterpret_cast<volatile long long *>(peerR + 0) = (long long)msg;
        dn[DN_TICKET] = 0;
        dn[DN_MSG] = msg;
    }
}

// the ghost slots were laid down by the rebuild's unpack: halos from the left, then from the right, after the own
// (slot_of != null: the ghost's cell-order record is refreshed here too, so no separate refresh pass is needed)
__global__ void slab_halo_recv_kernel(SlabArrays dst, int64_t ld, int *__restrict__ dn, double *rx, int64_t msg_doubles,
                                      int direct, int capM, int capH, const int *__restrict__ skip, unsigned long long timeout_ns,
                                      const int *__restrict__ slot_of, double4 *__restrict__ sp4)
{
    if (skip && skip[0]) return;
    const int msg = dn[DN_MSG];
    if (direct && !slab_wait_messages(rx, msg, timeout_ns)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) dn[DN_TIMEOUT] = 1;
        return;
    }
    const double *recvL = rx_msg(rx, 0, direct ? (msg & 1) : 0, msg_doubles);
    const double *recvR = rx_msg(rx, 1, direct ? (msg & 1) : 0, msg_doubles);
    const int haloL = min((int)recvL[1], capH), haloR = min((int)recvR[1], capH);
    const int n_own = dn[DN_OWN];
    if (haloL + haloR != dn[DN_GHOST]) { // a neighbour changed its halo without a collective rebuild
        if (blockIdx.x == 0 && threadIdx.x == 0) dn[DN_CAP] = 1;
        return;
    }
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int side = t / capH, k = t - side * capH;
    if (side >= 2 || k >= (side == 0 ? haloL : haloR)) return;
    const double *rec = (side == 0 ? recvL : recvR) + kHdr + (size_t)capM * kMigW + (size_t)k * kHaloW;
    const int d = n_own + (side == 0 ? 0 : haloL) + k;
    dst.pos[d] = rec[1]; dst.pos[ld + d] = rec[2]; dst.pos[2 * ld + d] = rec[3];
    if (slot_of) sp4[slot_of[d]] = make_double4(rec[1], rec[2], rec[3], dst.charge ? dst.charge[d] : 0.0);
}

// Halo refresh of slab_enqueue between rebuilds, LL protocol: the current positions of the recorded boundary-layer particles
// go straight into the neighbours' LL areas as self-validating 64-bit words tagged with the step's exchange number (the
// all-reduce that precedes it in every rank's stream is the barrier that makes one buffer enough).  No fence, no flag, no
// completion count; the receiver's threads each wait for their own six words and write position + cell-order record.
__global__ void __launch_bounds__(256) slab_ll_send_kernel(const double *__restrict__ pos, int64_t ld, const int *__restrict__ idxL,
                                                           const int *__restrict__ idxR, const int *__restrict__ dn,
                                                           unsigned long long *outL, unsigned long long *outR, int capH,
                                                           const int *__restrict__ seq, const int *__restrict__ skip)
{
    if (skip && skip[0]) return;
    const unsigned tag = (unsigned)seq[0];
    const int nL = dn[DN_HALOL], nR = dn[DN_HALOR];
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < 2 * capH; t += gridDim.x * blockDim.x) {
        const int side = t / capH, k = t - side * capH;
        if (k >= (side == 0 ? nL : nR)) continue;
        const int i = (side == 0 ? idxL : idxR)[k];
        // word j of halo particle k at [j][capH] + k: a warp's store is 256 contiguous bytes = eight full sectors on the
        // link (particle-major words were 32 partial sectors per store)
        unsigned long long *w = (side == 0 ? outL : outR) + k;
        ll_store_strided(w, (size_t)capH, pos[i], tag);
        ll_store_strided(w + 2 * (size_t)capH, (size_t)capH, pos[ld + i], tag);
        ll_store_strided(w + 4 * (size_t)capH, (size_t)capH, pos[2 * ld + i], tag);
    }
}

__global__ void __launch_bounds__(256) slab_ll_recv_kernel(double *__restrict__ pos, const double *__restrict__ charge, int64_t ld,
                                                           int *__restrict__ dn, const unsigned long long *__restrict__ ll, int capH,
                                                           const int *__restrict__ seq, const int *__restrict__ skip,
                                                           unsigned long long timeout_ns, const int *__restrict__ slot_of,
                                                           double4 *__restrict__ sp4)
{
    if (skip && skip[0]) return;
    const unsigned tag = (unsigned)seq[0];
    const int haloL = dn[DN_RHALOL], haloR = dn[DN_RHALOR], n_own = dn[DN_OWN];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int side = t / capH, k = t - side * capH;
    if (side >= 2 || k >= (side == 0 ? haloL : haloR)) return;
    const unsigned long long *w = ll + (size_t)side * capH * kLLWords + k;
    double x, y, z;
    if (!ll_load_strided(w, (size_t)capH, tag, timeout_ns, &x) || !ll_load_strided(w + 2 * (size_t)capH, (size_t)capH, tag, timeout_ns, &y) ||
        !ll_load_strided(w + 4 * (size_t)capH, (size_t)capH, tag, timeout_ns, &z)) {
        dn[DN_TIMEOUT] = 1;
        return;
    }
    const int d = n_own + (side == 0 ? 0 : haloL) + k;
    pos[d] = x; pos[ld + d] = y; pos[2 * ld + d] = z;
    sp4[slot_of[d]] = make_double4(x, y, z, charge ? charge[d] : 0.0);
}

// Position update of the own particles fused with what a regular step needs next: the displacement check against the
// build-time positions (-> flags[1] = 1.0 when a particle moved more than skin/2, or a list overflowed) and the refresh of
// the particle's cell-order record.  Same arithmetic as vv_pos_kernel / slab_verlet_check_kernel / verlet_refresh_kernel.
__global__ void slab_pos_kernel(double *__restrict__ pos, const double *__restrict__ vel, const double *__restrict__ acc,
                                const double *__restrict__ w, int64_t ld, const int *__restrict__ dn, double dt, double hdt2,
                                const double *__restrict__ ref, int64_t rld, double lim2, const int *__restrict__ slot_of,
                                double4 *__restrict__ sp4, const int *__restrict__ vflags, double *__restrict__ flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool moved = i == 0 && vflags[1];
    if (i < dn[DN_OWN]) {
        const double x = fma(hdt2, acc[i], fma(dt, vel[i], pos[i]));
        const double y = fma(hdt2, acc[ld + i], fma(dt, vel[ld + i], pos[ld + i]));
        const double z = fma(hdt2, acc[2 * ld + i], fma(dt, vel[2 * ld + i], pos[2 * ld + i]));
        pos[i] = x; pos[ld + i] = y; pos[2 * ld + i] = z;
        const double dx = x - ref[i], dy = y - ref[rld + i], dz = z - ref[2 * rld + i];
        const double d2 = dx * dx + dy * dy + dz * dz;
        moved = moved || !(d2 <= lim2); // also catches NaN
        sp4[slot_of[i]] = make_double4(x, y, z, w ? w[i] : 0.0);
    }
    if (moved) flags[1] = 1.0;
}

// displacement of the own particles from the positions the lists were built from: out[0] |= beyond the soft limit
// (ask for a collective rebuild), out[1] |= beyond skin/2 or a list overflowed (the lists are no longer a superset)
__global__ void slab_verlet_check_kernel(const double *__restrict__ px, int64_t ld, const double *__restrict__ ref, int64_t rld,
                                         const int *__restrict__ dn, double soft2, double hard2,
                                         const int *__restrict__ vflags, int *__restrict__ out, double *__restrict__ outd)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool soft = false, hard = i == 0 && vflags[1];
    if (i < dn[DN_OWN]) {
        const double dx = px[i] - ref[i], dy = px[ld + i] - ref[rld + i], dz = px[2 * ld + i] - ref[2 * rld + i];
        const double d2 = dx * dx + dy * dy + dz * dz;
        soft = !(d2 <= soft2);
        hard = hard || !(d2 <= hard2);
    }
    if (soft) { if (out) out[0] = 1; else outd[0] = 1.0; }
    if (hard) { if (out) out[1] = 1; else outd[1] = 1.0; }
}

// segments, in the order they are laid down: arrivals (left, right) extend the own particles; the ghosts are
// the migrants just sent (left, right) and the received halos (left, right)
__global__ void slab_unpack_kernel(SlabArrays dst, int64_t ld, int64_t cap_cols, int *__restrict__ counts,
                                   int *__restrict__ dn, const double *__restrict__ sendL,
                                   const double *__restrict__ sendR, double *rx, int64_t msg_doubles, int direct,
                                   int capM, int capH, const int *__restrict__ cond, unsigned long long timeout_ns)
{
    if (cond && !cond[0]) return;
    // direct mode: the neighbours store into this rank's receive area and raise its flags (message number)
    const int msg = dn[DN_MSG];
    if (direct && !slab_wait_messages(rx, msg, timeout_ns)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) { dn[DN_TIMEOUT] = 1; dn[DN_OWN] = 0; dn[DN_GHOST] = 0; }
        return;
    }
    const double *recvL = rx_msg(rx, 0, direct ? (msg & 1) : 0, msg_doubles);
    const double *recvR = rx_msg(rx, 1, direct ? (msg & 1) : 0, msg_doubles);
    const int nstay = counts[CAT_STAY];
    const int arrL = min((int)recvL[0], capM), arrR = min((int)recvR[0], capM);
    const int keptL = min((int)sendL[0], capM), keptR = min((int)sendR[0], capM);
    const int haloL = min((int)recvL[1], capH), haloR = min((int)recvR[1], capH);
    const int n_own = nstay + arrL + arrR;
    const int n_ghost = keptL + keptR + haloL + haloR;
    const bool fits = (int64_t)n_own + n_ghost <= cap_cols;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        counts[CNT_ARRL] = arrL; counts[CNT_ARRR] = arrR;
        dn[DN_RHALOL] = haloL; dn[DN_RHALOR] = haloR;
        dn[DN_OWN] = fits ? n_own : 0; dn[DN_GHOST] = fits ? n_ghost : 0;
        if (!fits) dn[DN_CAP] = 1;
    }
    if (!fits) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    int seg, idx;
    if (t < 2 * capM) { seg = t / capM; idx = t - seg * capM; }                 // 0,1: arrivals
    else if (t < 4 * capM) { seg = 2 + (t - 2 * capM) / capM; idx = (t - 2 * capM) % capM; } // 2,3: kept
    else { const int u = t - 4 * capM; if (u >= 2 * capH) return; seg = 4 + u / capH; idx = u % capH; } // 4,5: halos
    const int cnt = seg == 0 ? arrL : seg == 1 ? arrR : seg == 2 ? keptL : seg == 3 ? keptR : seg == 4 ? haloL : haloR;
    if (idx >= cnt) return;
    const double *buf = seg == 0 ? recvL : seg == 1 ? recvR : seg == 2 ? sendL : seg == 3 ? sendR : seg == 4 ? recvL : recvR;
    if (seg < 4) {
        const double *rec = buf + kHdr + (size_t)idx * kMigW;
        int d;
        if (seg == 0) d = nstay + idx;
        else if (seg == 1) d = nstay + arrL + idx;
        else if (seg == 2) d = n_own + idx;
        else d = n_own + keptL + idx;
        dst.gid[d] = (int)rec[0];
        dst.pos[d] = rec[1]; dst.pos[ld + d] = rec[2]; dst.pos[2 * ld + d] = rec[3];
        if (seg < 2) { // full state
            dst.vel[d] = rec[4]; dst.vel[ld + d] = rec[5]; dst.vel[2 * ld + d] = rec[6];
            dst.acc[d] = rec[7]; dst.acc[ld + d] = rec[8]; dst.acc[2 * ld + d] = rec[9];
            dst.mass[d] = rec[10];
        } else {
            dst.mass[d] = 1.0;
        }
        if (dst.charge) dst.charge[d] = rec[11];
    } else {
        const double *rec = buf + kHdr + (size_t)capM * kMigW + (size_t)idx * kHaloW;
        const int d = n_own + keptL + keptR + (seg == 4 ? 0 : haloL) + idx;
        dst.gid[d] = (int)rec[0];
        dst.pos[d] = rec[1]; dst.pos[ld + d] = rec[2]; dst.pos[2 * ld + d] = rec[3];
        dst.mass[d] = 1.0;
        if (dst.charge) dst.charge[d] = rec[4];
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
void slab_free(nbx_ctx *c)
{
    SlabState &s = c->slab;
    for (void *p : s.ipc_opened) cudaIpcCloseMemHandle(p);
    cudaFree(s.msg[0]); cudaFree(s.msg[1]); cudaFree(s.rx);
    cudaFree(s.pos2); cudaFree(s.vel2); cudaFree(s.acc2); cudaFree(s.mass2); cudaFree(s.charge2);
    cudaFree(s.gid_a); cudaFree(s.gid_b); cudaFree(s.blockcnt); cudaFree(s.blockoff); cudaFree(s.d_counts);
    cudaFree(s.d_n); cudaFree(s.halo_idx[0]); cudaFree(s.halo_idx[1]);
    if (s.h_counts) cudaFreeHost(s.h_counts);
    if (s.ev_counts) cudaEventDestroy(s.ev_counts);
    s = SlabState{};
    c->gid = nullptr;
    c->dyn = nullptr;
}

static SlabGeom geom(const nbx_ctx *c, int init)
{
    const SlabState &s = c->slab;
    SlabGeom g{};
    g.L = c->bc[0]; g.nc = s.nc; g.c0 = s.c0; g.width = s.c1 - s.c0; g.wL = s.wL; g.wR = s.wR; g.init = init;
    return g;
}

static SlabArrays arrays(double *pos, double *vel, double *acc, double *mass, double *charge, int *gid)
{
    SlabArrays a{};
    a.pos = pos; a.vel = vel; a.acc = acc; a.mass = mass; a.charge = charge; a.gid = gid;
    return a;
}

static int run_pack(nbx_ctx *c, int init)
{
    SlabState &s = c->slab;
    const int n = (int)(init ? s.n_total : s.cap_loc); // launch bound; the kernels clamp to d_n[0] after the init
    const int nb = (n + kSlabBlock - 1) / kSlabBlock;
    const SlabGeom g = geom(c, init);
    const int *dyn = init ? nullptr : s.d_n;
    NBX_CUDA(c, cudaMemsetAsync(s.d_counts, 0, sizeof(int) * CNT_N, c->stream));
    int *gid_dst = c->gid == s.gid_a ? s.gid_b : s.gid_a;
    const SlabArrays src = arrays(c->pos, c->vel, c->acc, c->mass, c->charge, c->gid);
    const SlabArrays dst = arrays(s.pos2, s.vel2, s.acc2, s.mass2, c->charge ? s.charge2 : nullptr, gid_dst);
    slab_count_kernel<<<nb, kSlabBlock, 0, c->stream>>>(c->pos, n, g, s.blockcnt, s.d_n, dyn, s.cond);
    slab_scan_kernel<<<1, 32 * CAT_N, 0, c->stream>>>(s.blockcnt, s.blockoff, nb, s.d_counts, s.cond);
    slab_pack_kernel<<<nb, kSlabBlock, 0, c->stream>>>(src, dst, c->npad, n, g, s.blockoff, s.d_counts, s.msg[0], s.msg[1],
                                                      (int)s.capM, (int)s.capH, s.d_n, dyn, s.direct ? s.peer[0] : nullptr,
                                                      s.direct ? s.peer[1] : nullptr, s.msg_doubles,
                                                      s.record_halo ? s.halo_idx[0] : nullptr, s.record_halo ? s.halo_idx[1] : nullptr,
                                                      s.cond);
    NBX_CUDA(c, cudaGetLastError());
    s.record_halo = false;
    // the compacted state is the state from here on (stream-ordered: later kernels see the new pointers)
    std::swap(c->pos, s.pos2); std::swap(c->vel, s.vel2); std::swap(c->acc, s.acc2); std::swap(c->mass, s.mass2);
    if (c->charge) std::swap(c->charge, s.charge2);
    c->gid = gid_dst;
    s.packed = true;
    return NBX_OK;
}

int slab_init(nbx_ctx *c, int rank, int nranks)
{
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(c, NBX_ERR_INVALID, "nbx_slab_init: rank %d of %d", rank, nranks);
    if (c->slab.on) return fail(c, NBX_ERR_INVALID, "nbx_slab_init: already decomposed (nbx_system starts over)");
    if (c->bc_kind != NBX_BC_CUBIC) return fail(c, NBX_ERR_UNSUPPORTED, "nbx_slab_init: slabs need CubicPeriodicBoundaryConditions");
    if (c->water || c->has_grav || c->has_dip || c->has_spcfw || c->thermo == NBX_THERMO_NOSEHOOVER)
        return fail(c, NBX_ERR_UNSUPPORTED, "nbx_slab_init: slabs cover atomic systems with cutoff Lennard-Jones / Coulomb terms");
    const bool coul_cut = c->has_coul && isfinite(c->el_R);
    if (c->has_coul && !coul_cut) return fail(c, NBX_ERR_UNSUPPORTED, "nbx_slab_init: unbounded Coulomb is an all-pairs problem (use nbx_shard)");
    if (!c->has_lj && !coul_cut) return fail(c, NBX_ERR_INVALID, "nbx_slab_init: no cutoff potential configured");
    if (c->tgt_lo != 0 || c->tgt_hi != c->n || c->pair_nranks > 1) return fail(c, NBX_ERR_INVALID, "nbx_slab_init: context is already sharded");
    const double R = fmax(c->has_lj ? c->lj_R : 0.0, coul_cut ? c->el_R : 0.0);
    CellGrid grid;
    // Verlet lists inside the slab: one cutoff potential and a grid of edge >= R + skin that still gives every rank two
    // layers (the ghost layer must hold everything within R + skin of the face)
    bool verlet = c->opt_verlet_permille > 0 && c->opt_prefilter && (c->has_lj != coul_cut);
    if (verlet) {
        const double skin = R * 1e-3 * (double)c->opt_verlet_permille; // as cells_pairs computes it
        NBX_TRY(cells_plan(c, R + skin, c->n, &grid));
        if (!grid.valid || (nranks > 1 && grid.nc[0] < 2 * nranks)) verlet = false;
    }
    if (!verlet) NBX_TRY(cells_plan(c, R, c->n, &grid));
    if (!grid.valid) return fail(c, NBX_ERR_UNSUPPORTED, "nbx_slab_init: the cutoff does not allow a cell grid (R >= L/3)");
    const int nc = grid.nc[0];
    if (nranks > 1 && nc < 2 * nranks)
        return fail(c, NBX_ERR_UNSUPPORTED, "nbx_slab_init: %d cell layers cannot give %d slabs of >= 2 layers", nc, nranks);
    SlabState &s = c->slab;
    s.verlet = verlet;
    s.rebuild_now = true;
    s.rank = rank; s.nranks = nranks; s.nc = nc;
    auto lo_of = [&](int r) { return (int)((int64_t)r * nc / nranks); };
    s.c0 = lo_of(rank); s.c1 = lo_of(rank + 1);
    const int left = (rank + nranks - 1) % nranks, right = (rank + 1) % nranks;
    s.wL = lo_of(left + 1) - lo_of(left); s.wR = lo_of(right + 1) - lo_of(right);
    s.n_total = c->n;
    const int64_t layer = (c->n + nc - 1) / nc;
    // a boundary layer's population with 50 % head room; a step moves a small fraction of a layer across a face
    s.capH = std::min<int64_t>(c->n, layer + layer / 2 + 1024);
    s.capM = std::min<int64_t>(c->n, layer / 8 + 1024);
    s.msg_doubles = kHdr + s.capM * kMigW + s.capH * kHaloW;
    // own + ghosts: twice the slab's share of a uniform box (+ the ghost layers); all of it for one slab
    s.cap_loc = nranks == 1 ? c->n : std::min<int64_t>(c->n, 2 * layer * (s.c1 - s.c0) + 2 * s.capH + 4096);
    const size_t np = (size_t)c->npad;
    for (int k = 0; k < 2; ++k) {
        NBX_TRY(dev_alloc(c, &s.msg[k], (size_t)s.msg_doubles));
        NBX_CUDA(c, cudaMemsetAsync(s.msg[k], 0, sizeof(double) * (size_t)s.msg_doubles, c->stream));
    }
    // receive area: flags + {from-left, from-right} x {parity 0, 1}; host-driven exchanges use parity 0
    s.ll_off = kRxHdr + 4 * s.msg_doubles;                      // LL area: [from-left, from-right][capH][6 words]
    s.rx_doubles = s.ll_off + 2 * s.capH * kLLWords;
    NBX_TRY(dev_alloc(c, &s.rx, (size_t)s.rx_doubles));
    NBX_CUDA(c, cudaMemsetAsync(s.rx, 0, sizeof(double) * (size_t)s.rx_doubles, c->stream));
    s.msg[2] = s.rx + kRxHdr;
    s.msg[3] = s.rx + kRxHdr + 2 * s.msg_doubles;
    NBX_TRY(dev_alloc(c, &s.pos2, 3 * np)); NBX_TRY(dev_alloc(c, &s.vel2, 3 * np)); NBX_TRY(dev_alloc(c, &s.acc2, 3 * np));
    NBX_TRY(dev_alloc(c, &s.mass2, np));
    if (c->charge) NBX_TRY(dev_alloc(c, &s.charge2, np));
    NBX_TRY(dev_alloc(c, &s.gid_a, np)); NBX_TRY(dev_alloc(c, &s.gid_b, np));
    const size_t nbmax = (size_t)((c->n + kSlabBlock - 1) / kSlabBlock) + 1;
    NBX_TRY(dev_alloc(c, &s.blockcnt, nbmax * CAT_N)); NBX_TRY(dev_alloc(c, &s.blockoff, nbmax * CAT_N));
    NBX_TRY(dev_alloc(c, &s.d_counts, (size_t)CNT_N));
    NBX_TRY(dev_alloc(c, &s.d_n, (size_t)DN_N));
    NBX_TRY(dev_alloc(c, &s.halo_idx[0], (size_t)s.capH));
    NBX_TRY(dev_alloc(c, &s.halo_idx[1], (size_t)s.capH));
    NBX_CUDA(c, cudaMemsetAsync(s.d_n, 0, sizeof(int) * DN_N, c->stream));
    NBX_CUDA(c, cudaMallocHost((void **)&s.h_counts, sizeof(int) * (CNT_N + DN_N)));
    NBX_CUDA(c, cudaEventCreateWithFlags(&s.ev_counts, cudaEventDisableTiming));
    // padding of the alternate rows as in nbx_system
    NBX_TRY(launch_fill(c, s.pos2, kFarAway, 3 * c->npad));
    NBX_CUDA(c, cudaMemsetAsync(s.vel2, 0, sizeof(double) * 3 * np, c->stream));
    NBX_CUDA(c, cudaMemsetAsync(s.acc2, 0, sizeof(double) * 3 * np, c->stream));
    NBX_CUDA(c, cudaMemsetAsync(s.mass2, 0, sizeof(double) * np, c->stream));
    NBX_CUDA(c, cudaStreamSynchronize(c->stream)); // the receive area is zeroed before a neighbour may write to it
    s.on = true;
    s.first = true;
    c->gid = nullptr; // the first pack reads the full uploaded state (ids = column numbers)
    return NBX_OK;
}

// Direct exchange: the neighbours' receive areas, either as device pointers valid in this process or as CUDA IPC
// handles of another process (64 bytes each, from nbx_slab_ipc_handle).
int slab_connect(nbx_ctx *c, const void *left_handle, const void *right_handle, void *left_ptr, void *right_ptr)
{
    SlabState &s = c->slab;
    if (!s.on || !s.first) return fail(c, NBX_ERR_INVALID, "nbx_slab_connect: call it between nbx_slab_init and the first nbx_slab_pack");
    if (!left_handle && !right_handle && !left_ptr && !right_ptr) { // back to the host-driven exchange
        s.direct = false;
        return NBX_OK;
    }
    if (s.nranks == 1) return NBX_OK;
    void *ptr[2] = {left_ptr, right_ptr};
    const void *hdl[2] = {left_handle, right_handle};
    for (int k = 0; k < 2; ++k) {
        if (ptr[k]) continue;
        if (!hdl[k]) return fail(c, NBX_ERR_INVALID, "nbx_slab_connect: neither a pointer nor a handle for side %d", k);
        if (k == 1 && hdl[0] && !left_ptr && !memcmp(hdl[0], hdl[1], sizeof(cudaIpcMemHandle_t))) { ptr[1] = ptr[0]; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, hdl[k], sizeof h);
        cudaError_t e = cudaIpcOpenMemHandle(&ptr[k], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return cuda_fail(c, e, "cudaIpcOpenMemHandle (neighbour receive area)");
        s.ipc_opened.push_back(ptr[k]);
    }
    s.peer[0] = static_cast<double *>(ptr[0]);
    s.peer[1] = static_cast<double *>(ptr[1]);
    s.direct = true;
    return NBX_OK;
}

int slab_pack(nbx_ctx *c)
{
    SlabState &s = c->slab;
    if (!s.on) return fail(c, NBX_ERR_INVALID, "nbx_slab_pack: call nbx_slab_init first");
    if (s.packed) return fail(c, NBX_ERR_INVALID, "nbx_slab_pack: the previous pack was not completed by nbx_slab_unpack");
    if (s.first) {
        NBX_TRY(run_pack(c, 1));
        s.first = false;
        // from here on the counts are device-side and the host sizes are bounds
        c->dyn = s.d_n;
        c->n = s.cap_loc;
        c->tgt_lo = 0;
        c->tgt_hi = s.cap_loc;
        return NBX_OK;
    }
    return run_pack(c, 0);
}

// halo refresh between rebuilds (see slab_halo_send_kernel)
int slab_refresh_send(nbx_ctx *c)
{
    SlabState &s = c->slab;
    if (!s.on || !s.verlet || s.first) return fail(c, NBX_ERR_INVALID, "nbx_slab_refresh_send: no Verlet-list slab state");
    if (s.packed) return fail(c, NBX_ERR_INVALID, "nbx_slab_refresh_send: the previous message was not received");
    const int64_t threads = 2 * s.capH;
    const SlabArrays src = arrays(c->pos, c->vel, c->acc, c->mass, c->charge, c->gid);
    slab_halo_send_kernel<<<(unsigned)((threads + kSlabBlock - 1) / kSlabBlock), kSlabBlock, 0, c->stream>>>(
        src, c->npad, s.halo_idx[0], s.halo_idx[1], s.d_n, s.msg[0], s.msg[1], (int)s.capM, (int)s.capH,
        s.direct ? s.peer[0] : nullptr, s.direct ? s.peer[1] : nullptr, s.msg_doubles, s.cond);
    NBX_CUDA(c, cudaGetLastError());
    s.packed = true;
    return NBX_OK;
}

int slab_refresh_recv(nbx_ctx *c)
{
    SlabState &s = c->slab;
    if (!s.on || !s.packed) return fail(c, NBX_ERR_INVALID, "nbx_slab_refresh_recv: nothing was sent");
    const int64_t threads = 2 * s.capH;
    const SlabArrays dst = arrays(c->pos, c->vel, c->acc, c->mass, c->charge, c->gid);
    slab_halo_recv_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, c->stream>>>(dst, c->npad, s.d_n, s.rx, s.msg_doubles,
                                                                                  s.direct ? 1 : 0, (int)s.capM, (int)s.capH, s.cond,
                                                                                  (unsigned long long)c->spin_timeout_ms * 1000000ull,
                                                                                  s.refresh_cl ? s.refresh_cl->slot_of : nullptr,
                                                                                  s.refresh_cl ? s.refresh_cl->sp4 : nullptr);
    NBX_CUDA(c, cudaGetLastError());
    s.packed = false;
    return NBX_OK;
}

// out2_dev[0] |= some own particle moved more than soft_fraction x skin/2 since the lists were built (or there are no
// lists yet); out2_dev[1] |= more than skin/2, or a list overflowed.  Enqueued on the context's stream.
// (out2_dbl instead of out2_dev: the same two flags as doubles, 0.0 / 1.0, for a driver that sums them with sum m v^2)
int slab_verlet_check(nbx_ctx *c, double soft_fraction, int *out2_dev, double *out2_dbl)
{
    SlabState &s = c->slab;
    if (!s.on || !s.verlet) return fail(c, NBX_ERR_INVALID, "nbx_slab_verlet_check: the slab does not keep Verlet lists");
    if (!out2_dev && !out2_dbl) return fail(c, NBX_ERR_INVALID, "nbx_slab_verlet_check: out is NULL");
    CellList *cl = c->has_lj ? &c->cl_lj : &c->cl_el;
    if (!cl->v_valid || !cl->v_ref || s.rebuild_now) { // nothing to compare with: ask for a rebuild
        if (out2_dev) NBX_CUDA(c, cudaMemsetAsync(out2_dev, 1, sizeof(int), c->stream));
        else NBX_TRY(launch_fill(c, out2_dbl, 1.0, 1));
        return NBX_OK;
    }
    const double lim = 0.5 * cl->v_skin * (1.0 - 1e-9), soft = lim * soft_fraction;
    const int n = (int)s.cap_loc;
    slab_verlet_check_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->pos, c->npad, cl->v_ref, cl->cap_n, s.d_n, soft * soft,
                                                                    lim * lim, cl->v_flags, out2_dev, out2_dbl);
    NBX_CUDA(c, cudaGetLastError());
    return NBX_OK;
}

// waits for the last asynchronous read-back and reports accumulated errors
int slab_check(nbx_ctx *c, int64_t *out)
{
    SlabState &s = c->slab;
    if (!s.on) return fail(c, NBX_ERR_INVALID, "nbx_slab_check: call nbx_slab_init first");
    if (s.packed) return fail(c, NBX_ERR_INVALID, "nbx_slab_check: a pack is waiting for nbx_slab_unpack");
    NBX_CUDA(c, cudaMemcpyAsync(s.h_counts, s.d_counts, sizeof(int) * CNT_N, cudaMemcpyDeviceToHost, c->stream));
    NBX_CUDA(c, cudaMemcpyAsync(s.h_counts + CNT_N, s.d_n, sizeof(int) * DN_N, cudaMemcpyDeviceToHost, c->stream));
    NBX_CUDA(c, cudaStreamSynchronize(c->stream));
    const int *h = s.h_counts, *hn = s.h_counts + CNT_N;
    if (hn[DN_LOST] > 0)
        return fail(c, NBX_ERR_INVALID, "slab exchange: %d particle(s) moved past the neighbouring slab in one step", hn[DN_LOST]);
    if (hn[DN_TIMEOUT] != 0)
        return fail(c, NBX_ERR_CUDA, "slab exchange: timed out waiting for a neighbour's message (direct mode)");
    if (hn[DN_CAP] != 0)
        return fail(c, NBX_ERR_CAPACITY, "slab exchange: message or state capacity exceeded (last step: migrants %d/%d of %lld, "
                    "halo %d/%d of %lld; own + ghosts bound %lld)", h[CAT_MIGL], h[CAT_MIGR], (long long)s.capM, h[CAT_HALOL],
                    h[CAT_HALOR], (long long)s.capH, (long long)s.cap_loc);
    s.n_own = hn[DN_OWN];
    s.n_ghost = hn[DN_GHOST];
    if (out) {
        out[0] = s.n_own; out[1] = s.n_ghost; out[2] = h[CAT_MIGL]; out[3] = h[CAT_MIGR]; out[4] = h[CNT_ARRL]; out[5] = h[CNT_ARRR];
    }
    return NBX_OK;
}

// counts == NULL: asynchronous (errors surface at the next nbx_slab_check); else also synchronise and report
int slab_unpack(nbx_ctx *c, int64_t *out)
{
    SlabState &s = c->slab;
    if (!s.on || !s.packed) return fail(c, NBX_ERR_INVALID, "nbx_slab_unpack: nothing was packed");
    const int64_t threads = 4 * s.capM + 2 * s.capH;
    const SlabArrays dst = arrays(c->pos, c->vel, c->acc, c->mass, c->charge, c->gid);
    slab_unpack_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, c->stream>>>(dst, c->npad, s.cap_loc, s.d_counts, s.d_n,
                                                                              s.msg[0], s.msg[1], s.rx, s.msg_doubles,
                                                                              s.direct ? 1 : 0, (int)s.capM, (int)s.capH, s.cond,
                                                                              (unsigned long long)c->spin_timeout_ms * 1000000ull);
    NBX_CUDA(c, cudaGetLastError());
    s.packed = false;
    return out ? slab_check(c, out) : NBX_OK;
}

// ------------------------------------------------------------------------------------------------
// the slab step loop behind nbx_step_vv: no host decision, no collective library
// ------------------------------------------------------------------------------------------------
int compute_pairs(nbx_ctx *c); // nbx_api.cu

// Initial distribution: the first pack selects the own particles out of the full upload, the neighbours' messages lay the
// ghosts down; with Verlet lists a second round records the halo index lists and the lists are built from the
// distributed positions right away.  Enqueues only (the kernels wait for the neighbours on the device): synchronise
// with nbx_slab_check once every rank has made the call.
int slab_start(nbx_ctx *c)
{
    SlabState &s = c->slab;
    if (!s.on || !s.first) return fail(c, NBX_ERR_INVALID, "nbx_group_start: the slab was started already");
    if (s.nranks > 1 && !s.direct) return fail(c, NBX_ERR_INVALID, "nbx_group_start: connect the neighbours first (nbx_group_connect)");
    NBX_TRY(slab_pack(c));
    NBX_TRY(slab_unpack(c, nullptr));
    if (s.verlet) {
        s.record_halo = true;
        NBX_TRY(slab_pack(c));
        NBX_TRY(slab_unpack(c, nullptr));
        s.rebuild_now = true;
        std::swap(c->acc, c->acc_old); // cells and lists get built, a(t) stays what it is
        const int rc = compute_pairs(c);
        std::swap(c->acc, c->acc_old);
        if (rc != NBX_OK) return rc;
    }
    if (c->T_slot != 12) { // the sum of the upload is the global one
        NBX_CUDA(c, cudaMemcpyAsync(c->d_scal + 12, c->d_scal, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        c->T_slot = 12;
    }
    s.scal0_global = true;
    s.started = true;
    return NBX_OK;
}

// One velocity-Verlet step of the slab as a fixed launch sequence.  With Verlet lists: position update, displacement
// check against the build-time positions, ONE peer-memory all-reduce of (sum m v^2, moved-too-far flags) that every rank
// evaluates identically; the migration + halo-recording rounds and the cell / list rebuild run iff the reduced flag is
// set (each kernel of the chain returns at once otherwise; while a graph is captured the chain is the body of an IF
// node), the plain halo refresh iff it is not.  Same criterion, same step as a single context: with the bit-identical
// cell order (ranked by global id) the trajectory IS the single-context trajectory.
static int slab_one_step(nbx_ctx *c, double dt)
{
    SlabState &s = c->slab;
    CellList *cl = c->has_lj ? &c->cl_lj : &c->cl_el;
    const bool needT = c->thermo == NBX_THERMO_BERENDSEN;
    if (s.verlet) {
        // regular step = 6 launches: position update + check + own records | all-reduce + IF decision | halo send |
        // halo receive + ghost records | listed pairs | velocity update + sum m v^2
        const bool lists = cl->v_valid && cl->v_ref && cl->slot_of && !s.rebuild_now;
        if (lists) {
            const double lim = 0.5 * cl->v_skin * (1.0 - 1e-9);
            const int n = (int)s.cap_loc;
            timer_begin(c, NBX_T_INTEGRATE);
            slab_pos_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->pos, c->vel, c->acc, c->has_lj ? nullptr : c->charge, c->npad, s.d_n, dt,
                                                                   0.5 * dt * dt, cl->v_ref, cl->cap_n, lim * lim, cl->slot_of, cl->sp4,
                                                                   cl->v_flags, c->d_scal + 13);
            timer_end(c, NBX_T_INTEGRATE);
            NBX_CUDA(c, cudaGetLastError());
        } else {
            NBX_TRY(launch_vv_pos(c, dt));
            NBX_TRY(slab_verlet_check(c, 1.0, nullptr, c->d_scal + 13));
        }
        cudaGraphConditionalHandle h{};
        bool have_h = false;
        NBX_TRY(cond_handle_create(c, &h, &have_h));
        NBX_TRY(comm_allreduce3(c, c->d_scal, c->d_scal + 12, cl->v_flags, have_h ? &h : nullptr));
        s.cond = cl->v_flags;
        CondScope scope;
        NBX_TRY(cond_scope_begin(c, cl->v_flags, &scope, have_h ? &h : nullptr));
        int rc = slab_pack(c);                         // migration round
        if (rc == NBX_OK) rc = slab_unpack(c, nullptr);
        s.record_halo = true;                          // halo round incl. the arrivals, remembered
        if (rc == NBX_OK) rc = slab_pack(c);
        if (rc == NBX_OK) rc = slab_unpack(c, nullptr);
        s.rebuild_now = true;
        s.phase = 1;                                   // cells + lists, no evaluation
        if (rc == NBX_OK) { std::swap(c->acc, c->acc_old); rc = compute_pairs(c); std::swap(c->acc, c->acc_old); }
        s.phase = 0;
        const int rc2 = cond_scope_end(c, &scope);
        if (rc != NBX_OK || rc2 != NBX_OK) { s.cond = nullptr; return rc != NBX_OK ? rc : rc2; }
        // (Measured and dropped, r02: the targets that cannot have a ghost partner -- all own layers but the first and the
        // last -- evaluated on a second, low-priority stream while the halo refresh travels, the boundary layers after the
        // receive.  Bit-identical, but slower on 2 GPUs: 0.214 vs 0.186 ms per step at 1,048,576 atoms, 0.067 vs 0.058 at
        // 131,072 -- boundary targets are two cells of EVERY x-row, so both halves launch the whole grid and nine tenths
        // of the second launch's threads exit at once; a compact list of boundary slots would be needed.)
        if (lists) { // LL halo refresh (skipped on the device when the rebuild ran); one slab has no neighbours at all
            if (s.nranks > 1) {
                const int *seq = c->comm.d_seq + SEQ_SCAL;
                unsigned long long *toL = reinterpret_cast<unsigned long long *>(s.peer[0] + s.ll_off) + (size_t)s.capH * kLLWords; // left's from-right
                unsigned long long *toR = reinterpret_cast<unsigned long long *>(s.peer[1] + s.ll_off);                           // right's from-left
                const int64_t threads = 2 * s.capH;
                const unsigned sgrid = (unsigned)std::min<int64_t>((threads + 255) / 256, (int64_t)c->sm_count * 2);
                timer_begin(c, NBX_T_INTEGRATE);
                slab_ll_send_kernel<<<sgrid, 256, 0, c->stream>>>(c->pos, c->npad, s.halo_idx[0], s.halo_idx[1], s.d_n, toL, toR, (int)s.capH,
                                                                 seq, s.cond);
                slab_ll_recv_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, c->stream>>>(
                    c->pos, c->charge, c->npad, s.d_n, reinterpret_cast<const unsigned long long *>(s.rx + s.ll_off), (int)s.capH, seq, s.cond,
                    (unsigned long long)c->spin_timeout_ms * 1000000ull, cl->slot_of, cl->sp4);
                timer_end(c, NBX_T_INTEGRATE);
                NBX_CUDA(c, cudaGetLastError());
            }
        } else {
            rc = slab_refresh_send(c);
            if (rc == NBX_OK) rc = slab_refresh_recv(c);
        }
        s.cond = nullptr;
        NBX_TRY(rc);
        s.records_fresh = lists;                       // (a rebuild writes the records from the same positions)
        s.phase = 2;
    } else {
        NBX_TRY(launch_vv_pos(c, dt));
        if (needT) NBX_TRY(comm_allreduce3(c, c->d_scal, c->d_scal + 12, nullptr));
        NBX_TRY(slab_pack(c));
        NBX_TRY(slab_unpack(c, nullptr));
    }
    std::swap(c->acc, c->acc_old);
    const int rc = compute_pairs(c);
    s.phase = 0;
    s.records_fresh = false;
    NBX_TRY(rc);
    return launch_vv_vel(c, dt, true);
}

} // namespace nbx

#include "nbx_graph.inl"

namespace nbx {

int slab_enqueue(nbx_ctx *c, double dt, int64_t nsteps)
{
    SlabState &s = c->slab;
    if (!s.on || !s.started) return fail(c, NBX_ERR_INVALID, "nbx_step_vv: start the slab decomposition first (nbx_group_start)");
    if (s.packed) return fail(c, NBX_ERR_INVALID, "nbx_step_vv: a pack is waiting for nbx_slab_unpack");
    if (s.nranks > 1 && !s.direct) return fail(c, NBX_ERR_INVALID, "nbx_step_vv: the slabs are not connected (nbx_group_connect); drive the exchange from the host");
    if (c->thermo == NBX_THERMO_ANDERSEN || c->thermo == NBX_THERMO_LANGEVIN || c->thermo == NBX_THERMO_NOSEHOOVER)
        return fail(c, NBX_ERR_UNSUPPORTED, "nbx_step_vv: slabs run NVE or with the Berendsen thermostat");
    NBX_TRY(comm_arm_global0(c));
    NBX_CUDA(c, cudaMemsetAsync(c->d_scal + 13, 0, 2 * sizeof(double), c->stream)); // the all-reduce re-zeroes the flags it has read
    NBX_TRY(steps_graphed(c, 13, dt, nsteps, true, s.verlet ? 2 : 6, [&]() { return slab_one_step(c, dt); }));
    // leave the scalar block as a single context does: [0] = the sum over all ranks
    NBX_CUDA(c, cudaMemsetAsync(c->d_scal + 13, 0, 2 * sizeof(double), c->stream));
    NBX_TRY(comm_allreduce3(c, c->d_scal, c->d_scal + 12, nullptr));
    NBX_CUDA(c, cudaMemcpyAsync(c->d_scal, c->d_scal + 12, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    s.scal0_global = true;
    return NBX_OK;
}

int slab_finish(nbx_ctx *c)
{
    NBX_TRY(slab_check(c, nullptr));
    int h[SEQ_N] = {0};
    if (c->comm.d_seq) NBX_CUDA(c, cudaMemcpy(h, c->comm.d_seq, sizeof h, cudaMemcpyDeviceToHost));
    if (h[SEQ_TIMEOUT]) return fail(c, NBX_ERR_CUDA, "slab step: timed out waiting for a peer's flags (a rank did not make the same call?)");
    return NBX_OK;
}

// CUDA loads kernels lazily, at their first launch, and loading may synchronise the context: a kernel that is first
// launched while another member's kernel spins on a flag would deadlock the pair (CUDA programming guide, "Lazy
// Loading": concurrent execution).  Every kernel of the distributed loops is therefore loaded when a context is created.
void preload_slab()
{
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, slab_pos_kernel);
    cudaFuncGetAttributes(&a, slab_ll_send_kernel);
    cudaFuncGetAttributes(&a, slab_ll_recv_kernel);
    cudaFuncGetAttributes(&a, slab_count_kernel);
    cudaFuncGetAttributes(&a, slab_scan_kernel);
    cudaFuncGetAttributes(&a, slab_pack_kernel);
    cudaFuncGetAttributes(&a, slab_halo_send_kernel);
    cudaFuncGetAttributes(&a, slab_halo_recv_kernel);
    cudaFuncGetAttributes(&a, slab_verlet_check_kernel);
    cudaFuncGetAttributes(&a, slab_unpack_kernel);
    cudaGetLastError();
}

} // namespace nbx
