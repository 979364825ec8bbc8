// nbx_multi.cu -- multi-GPU behind the C ABI (sm_100a): a peer-memory communicator and the distributed step loops.
//
// The reference is serial; the data-parallel axis is the target index i of soode_system! (src/nbody_to_ode.jl:474-488,
// :502-532).  A GROUP is nranks contexts, one per GPU, in one process (nbx_create_multi) or one per process
// (nbx_group_init / nbx_group_export / nbx_group_connect with CUDA IPC handles exchanged by the host).  Three
// decompositions, chosen from the potentials:
//   pairs   unbounded gravity / Coulomb: every rank evaluates its ring offsets of the Newton's-third-law kernel, PUSHES its
//           partial acceleration rows into the owners' staging areas, the owner adds the nranks partials in rank order
//   targets any potential: rank r evaluates the target columns [lo_r, hi_r) against all sources
//   slabs   cutoff Lennard-Jones / Coulomb in a cubic periodic box: x-slabs with halo exchange (nbx_slab.cu)
// Exchanges are done by the kernels themselves over NVLink peer memory: the position update stores the new positions of
// its block into EVERY rank's position rows (the all-gather), the last block fences and raises a flag in every peer's
// window; consumers spin on their own window.  Scalars (sum m v^2, rebuild flags) travel the same way.  No collective
// library, no host in the loop: a velocity-Verlet step is a fixed launch sequence, replayed as a CUDA graph of two steps.
#include "nbx_internal.cuh"

#include <cmath>
#include <cstring>

namespace nbx {

int compute_pairs(nbx_ctx *c); // nbx_api.cu

// ------------------------------------------------------------------------------------------------
// device side
// ------------------------------------------------------------------------------------------------
// spin until *flag >= want; bounded by 10 s of %globaltimer (a dead peer must not hang the GPU)
__device__ __forceinline__ bool spin_ge(const volatile long long *flag, long long want, unsigned long long timeout_ns)
{
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (*flag < want) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) return false;
    }
    return true;
}
__device__ __forceinline__ bool spin_ge_d(const volatile double *flag, double want, unsigned long long timeout_ns)
{
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (*flag < want) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) return false;
    }
    return true;
}

// Sum over the ranks of three doubles, identical on every rank (added in rank order): every rank stores its values and a
// sequence number into slot [parity][rank] of EVERY window, then waits for the nranks slots of its own window.  Double
// buffered by the parity of the sequence number: a peer can be at most one exchange ahead.
//   in0 (nullable; taken from rank 0 only while seq[SEQ_GLOBAL0] is set -- the value is then already the global one: the sum
//   m v^2 of an upload or of the end of a run; the flag is consumed here, so every step is the same launch sequence),
//   in12 (nullable: two values); out0, out12 (nullable); flag_out[0] = (sum of in12[0..1] != 0): the collective rebuild decision
// The two flag inputs are zeroed once read (the next step's check raises them again: no memset per step), and with a
// conditional handle the kernel also decides the graph's IF node (cudaGraphSetConditional): one launch less per step.
__global__ void comm_allreduce3_kernel(CommDev cd, const double *in0, double *in12, double *out0, double *out12,
                                       int *flag_out, cudaGraphConditionalHandle cond, int has_cond)
{
    const int in0_rank0_only = cd.seq[SEQ_GLOBAL0];
    __shared__ double v[kMaxRanks][3];
    __shared__ int bad;
    const int t = threadIdx.x;
    const int s = cd.seq[SEQ_SCAL] + 1;
    if (t == 0) bad = 0;
    __syncthreads();
    if (t < cd.nranks) {
        double a = in0 ? in0[0] : 0.0;
        if (in0_rank0_only && cd.rank != 0) a = 0.0;
        const double b = in12 ? in12[0] : 0.0, e = in12 ? in12[1] : 0.0;
        // LL words: no fence, no flag -- every 64-bit word carries its own tag (the sequence number)
        unsigned long long *slot = reinterpret_cast<unsigned long long *>(cd.win[t] + kWinScal) + ((s & 1) * kMaxRanks + cd.rank) * 8;
        ll_store(slot, a, (unsigned)s); ll_store(slot + 2, b, (unsigned)s); ll_store(slot + 4, e, (unsigned)s);
        const unsigned long long *mine = reinterpret_cast<const unsigned long long *>(cd.win[cd.rank] + kWinScal) + ((s & 1) * kMaxRanks + t) * 8;
        if (!ll_load(mine, (unsigned)s, cd.timeout_ns, &v[t][0]) || !ll_load(mine + 2, (unsigned)s, cd.timeout_ns, &v[t][1]) ||
            !ll_load(mine + 4, (unsigned)s, cd.timeout_ns, &v[t][2])) {
            bad = 1; v[t][0] = v[t][1] = v[t][2] = 0.0;
        }
    }
    __syncthreads();
    if (t == 0) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int r = 0; r < cd.nranks; ++r) { s0 += v[r][0]; s1 += v[r][1]; s2 += v[r][2]; }
        if (bad) { cd.seq[SEQ_TIMEOUT] = 1; s1 = 1.0; }
        if (out0) out0[0] = s0;
        if (out12) { out12[0] = s1; out12[1] = s2; }
        const int any = (s1 != 0.0 || s2 != 0.0) ? 1 : 0;
        if (flag_out) flag_out[0] = any;
        if (has_cond) cudaGraphSetConditional(cond, (unsigned)any);
        if (in12) { in12[0] = 0.0; in12[1] = 0.0; }
        cd.seq[SEQ_SCAL] = s;
        cd.seq[SEQ_GLOBAL0] = 0;
    }
}

struct PushArgs {
    double *dst[kMaxRanks]; // position rows (or staging areas) of every rank
    double *win[kMaxRanks];
    int rank, nranks;
    int *seq;
};

// completion of a pushing kernel: every block fences its remote stores and takes a ticket; the last one raises this
// rank's flag (sequence number s) in every window
__device__ __forceinline__ void push_complete(const PushArgs &pa, int s, int win_off, int seq_slot, int ticket_slot)
{
    __shared__ int last_block;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last_block = atomicAdd(&pa.seq[ticket_slot], 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!last_block) return;
    if (threadIdx.x < pa.nranks) {
        __threadfence_system();
        reinterpret_cast<volatile long long *>(pa.win[threadIdx.x] + win_off)[pa.rank] = (long long)s;
    }
    if (threadIdx.x == 0) { pa.seq[ticket_slot] = 0; pa.seq[seq_slot] = s; }
}

// x+ = x + dt v + dt^2/2 a on the own block [lo, hi) (update != 0; else the positions as they are), stored into the
// position rows of EVERY rank: the per-step position all-gather of SURVEY 8(e), done by the update kernel itself
__global__ void __launch_bounds__(256) vv_pos_push_kernel(PushArgs pa, const double *__restrict__ vel, const double *__restrict__ acc,
                                                          int64_t ld, int64_t lo, int64_t hi, double dt, double hdt2, int update)
{
    const int s = pa.seq[SEQ_POS] + 1;
    const double *own = pa.dst[pa.rank];
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int64_t k = d * ld + i;
            const double x = update ? fma(hdt2, acc[k], fma(dt, vel[k], own[k])) : own[k];
            for (int p = 0; p < pa.nranks; ++p)
                if (update || p != pa.rank) pa.dst[p][k] = x;
        }
    }
    push_complete(pa, s, kWinPos, SEQ_POS, SEQ_TICKET);
}

// Euler-Maruyama already moved the own block (em_kernel); this only distributes it
// (same kernel, update = 0)

// one block: wait until every rank's flag of kind `win_off` has reached this rank's own sequence number
__global__ void comm_wait_kernel(CommDev cd, int win_off, int seq_slot)
{
    const int t = threadIdx.x;
    if (t < cd.nranks) {
        const long long want = cd.seq[seq_slot];
        if (!spin_ge(reinterpret_cast<const volatile long long *>(cd.win[cd.rank] + win_off) + t, want, cd.timeout_ns)) cd.seq[SEQ_TIMEOUT] = 1;
        __threadfence_system();
    }
}

// pair sharding: this rank's partial accelerations of ALL bodies go to the owners: stage[owner][(rank*3 + d)*per + off]
__global__ void __launch_bounds__(256) acc_push_kernel(PushArgs pa, const double *__restrict__ acc, int64_t ld, int64_t n, int64_t per)
{
    const int s = pa.seq[SEQ_ACC] + 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int owner = (int)(i / per);
        const int64_t off = i - (int64_t)owner * per;
        double *dst = pa.dst[owner] + (size_t)pa.rank * 3 * per + off;
#pragma unroll
        for (int d = 0; d < 3; ++d) dst[(size_t)d * per] = acc[d * ld + i];
    }
    push_complete(pa, s, kWinAcc, SEQ_ACC, SEQ_TICKET2);
}

// the owner adds the staged partials of its block in rank order (deterministic, the same on every run)
__global__ void __launch_bounds__(256) acc_sum_kernel(CommDev cd, const double *__restrict__ stage, int64_t per,
                                                      double *__restrict__ acc, int64_t ld, int64_t lo, int64_t hi)
{
    __shared__ int bad;
    if (threadIdx.x == 0) {
        bad = 0;
        const long long want = cd.seq[SEQ_ACC];
        const volatile long long *f = reinterpret_cast<const volatile long long *>(cd.win[cd.rank] + kWinAcc);
        for (int r = 0; r < cd.nranks; ++r)
            if (!spin_ge(f + r, want, cd.timeout_ns)) bad = 1;
        __threadfence_system();
        if (bad) cd.seq[SEQ_TIMEOUT] = 1;
    }
    __syncthreads();
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t off = i - lo;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            double s = 0.0;
            for (int r = 0; r < cd.nranks; ++r) s += __ldcg(stage + ((size_t)r * 3 + d) * per + off);
            acc[d * ld + i] = s;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side: communicator
// ------------------------------------------------------------------------------------------------
CommDev comm_dev(const nbx_ctx *c)
{
    CommDev cd{};
    const Comm &m = c->comm;
    for (int r = 0; r < kMaxRanks; ++r) cd.win[r] = m.peer_win[r];
    cd.rank = m.rank; cd.nranks = m.nranks; cd.seq = m.d_seq;
    cd.timeout_ns = (unsigned long long)(c->spin_timeout_ms > 0 ? c->spin_timeout_ms : 1) * 1000000ull;
    return cd;
}

static PushArgs push_args(const nbx_ctx *c, double *const *dst)
{
    PushArgs pa{};
    const Comm &m = c->comm;
    for (int r = 0; r < kMaxRanks; ++r) { pa.dst[r] = dst[r]; pa.win[r] = m.peer_win[r]; }
    pa.rank = m.rank; pa.nranks = m.nranks; pa.seq = m.d_seq;
    return pa;
}

// window + counters of a one-rank communicator; nbx_group_connect fills in the peers
int comm_alloc(nbx_ctx *c)
{
    Comm &m = c->comm;
    if (m.win) return NBX_OK;
    NBX_TRY(dev_alloc(c, &m.win, (size_t)kWinDoubles));
    NBX_TRY(dev_alloc(c, &m.d_seq, (size_t)SEQ_N));
    NBX_CUDA(c, cudaMemsetAsync(m.win, 0, sizeof(double) * kWinDoubles, c->stream));
    NBX_CUDA(c, cudaMemsetAsync(m.d_seq, 0, sizeof(int) * SEQ_N, c->stream));
    NBX_CUDA(c, cudaStreamSynchronize(c->stream)); // zeroed before a peer may write to it
    m.rank = 0; m.nranks = 1;
    for (int r = 0; r < kMaxRanks; ++r) { m.peer_win[r] = nullptr; m.peer_pos[r] = nullptr; m.peer_stage[r] = nullptr; }
    m.peer_win[0] = m.win;
    return NBX_OK;
}

void comm_free(nbx_ctx *c)
{
    Comm &m = c->comm;
    for (void *p : m.ipc_opened) cudaIpcCloseMemHandle(p);
    cudaFree(m.win); cudaFree(m.stage); cudaFree(m.d_seq);
    m = Comm{};
}

void graph_drop(nbx_ctx *c)
{
    if (c->mg_exec) { cudaGraphExecDestroy(c->mg_exec); c->mg_exec = nullptr; }
    c->mg_kind = 0;
}

// sum over the ranks of (in0[0], d_scal[13], d_scal[14]) -> (out0[0], d_scal[13], d_scal[14]); flag_out[0] = any of the two flags.
// While the host knows that in0 already holds the global value (slab.scal0_global) it raises the device flag first.
int comm_arm_global0(nbx_ctx *c)
{
    NBX_TRY(comm_alloc(c));
    if (c->slab.scal0_global) {
        NBX_CUDA(c, cudaMemsetAsync(c->comm.d_seq + SEQ_GLOBAL0, 1, sizeof(int), c->stream));
        c->slab.scal0_global = false;
    }
    return NBX_OK;
}

int comm_allreduce3(nbx_ctx *c, const double *in0, double *out0, int *flag_out, const cudaGraphConditionalHandle *cond)
{
    NBX_TRY(comm_arm_global0(c)); // (a replayed graph never comes through here: the step loops arm the flag before they enqueue)
    comm_allreduce3_kernel<<<1, 32, 0, c->stream>>>(comm_dev(c), in0, c->d_scal + 13, out0, nullptr, flag_out,
                                                    cond ? *cond : cudaGraphConditionalHandle{}, cond ? 1 : 0);
    NBX_CUDA(c, cudaGetLastError());
    return NBX_OK;
}

static bool slab_capable(const nbx_ctx *c)
{
    const bool coul_cut = c->has_coul && std::isfinite(c->el_R);
    return c->bc_kind == NBX_BC_CUBIC && !c->water && !c->has_grav && !c->has_dip && !c->has_spcfw && (c->has_lj || coul_cut) &&
           (!c->has_coul || coul_cut) && c->thermo != NBX_THERMO_NOSEHOOVER && c->thermo != NBX_THERMO_ANDERSEN &&
           c->thermo != NBX_THERMO_LANGEVIN;
}

// ------------------------------------------------------------------------------------------------
// group set-up (the same four calls whether the peers live in this process or in others)
// ------------------------------------------------------------------------------------------------
int group_init(nbx_ctx *c, int rank, int nranks, int mode)
{
    if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks)
        return fail(c, NBX_ERR_INVALID, "nbx_group_init: rank %d of %d (at most %d ranks)", rank, nranks, kMaxRanks);
    if (c->comm.on) return fail(c, NBX_ERR_INVALID, "nbx_group_init: the context already belongs to a group (nbx_system starts over)");
    if (c->slab.on || c->tgt_lo != 0 || c->tgt_hi != c->n || c->pair_nranks > 1)
        return fail(c, NBX_ERR_INVALID, "nbx_group_init: the context is already sharded");
    if (c->thermo == NBX_THERMO_NOSEHOOVER) return fail(c, NBX_ERR_UNSUPPORTED, "nbx_group_init: the Nose-Hoover thermostat is not distributed");
    if (mode == 0) {
        if (pair_capable(c)) mode = 1;
        else if (slab_capable(c)) mode = 3;
        else mode = 2;
    }
    if (mode == 1 && !pair_capable(c))
        return fail(c, NBX_ERR_UNSUPPORTED, "nbx_group_init: pair sharding covers unbounded gravity / Coulomb, and Coulomb with L/3 <= R < L/2 in a cubic box");
    graph_drop(c);
    NBX_TRY(comm_alloc(c));
    Comm &m = c->comm;
    if (mode == 3) {
        const int rc = slab_init(c, rank, nranks);
        if (rc != NBX_OK) return rc;
    } else {
        const int64_t mult = c->water ? 3 : 1;
        int64_t per = (c->n + nranks - 1) / nranks;
        per = (per + mult - 1) / mult * mult;
        m.per = per;
        const int64_t lo = std::min<int64_t>(c->n, (int64_t)rank * per), hi = std::min<int64_t>(c->n, lo + per);
        c->tgt_lo = lo; c->tgt_hi = hi;
        if (mode == 1) {
            c->pair_rank = rank; c->pair_nranks = nranks;
            m.stage_doubles = (int64_t)nranks * 3 * per;
            NBX_TRY(dev_alloc(c, &m.stage, (size_t)m.stage_doubles));
            NBX_CUDA(c, cudaMemsetAsync(m.stage, 0, sizeof(double) * (size_t)m.stage_doubles, c->stream));
        }
    }
    m.rank = rank; m.nranks = nranks; m.mode = mode;
    for (int r = 0; r < kMaxRanks; ++r) { m.peer_win[r] = nullptr; m.peer_pos[r] = nullptr; m.peer_stage[r] = nullptr; }
    m.peer_win[rank] = m.win; m.peer_pos[rank] = c->pos; m.peer_stage[rank] = m.stage;
    m.warm = false;
    m.on = true;
    NBX_TRY(ensure_red(c));
    if (mode != 3 && c->resident) {
        // one sharded evaluation now, into the spare rows: every buffer the step loop needs exists before the first
        // kernel waits for a peer (an allocation between two members' launches would serialise their streams)
        std::swap(c->acc, c->acc_old);
        const int rc = compute_pairs(c);
        std::swap(c->acc, c->acc_old);
        NBX_TRY(rc);
        m.warm = true;
    }
    if (c->thermo == NBX_THERMO_BERENDSEN && mode != 3 && c->T_slot != 12) { // the summed sum m v^2 lives in slot 12
        NBX_CUDA(c, cudaMemcpyAsync(c->d_scal + 12, c->d_scal, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        c->T_slot = 12;
    }
    NBX_CUDA(c, cudaStreamSynchronize(c->stream));
    return NBX_OK;
}

static void *export_ptr(nbx_ctx *c, int kind)
{
    switch (kind) {
    case 0: return c->comm.win;
    case 1: return c->slab.on ? c->slab.rx : nullptr;
    case 2: return c->comm.mode == 3 ? nullptr : c->pos;
    case 3: return c->comm.stage;
    default: return nullptr;
    }
}

int group_export(nbx_ctx *c, int kind, void **ptr, void *handle64)
{
    if (!c->comm.on) return fail(c, NBX_ERR_INVALID, "nbx_group_export: call nbx_group_init first");
    if (kind < 0 || kind > 3) return fail(c, NBX_ERR_INVALID, "nbx_group_export: kind %d", kind);
    void *p = export_ptr(c, kind);
    if (ptr) *ptr = p;
    if (handle64) {
        memset(handle64, 0, 64);
        if (p) {
            cudaIpcMemHandle_t h;
            NBX_CUDA(c, cudaIpcGetMemHandle(&h, p));
            memcpy(handle64, &h, sizeof h);
        }
    }
    return NBX_OK;
}

// handles: [nranks][4][64] bytes (other processes), ptrs: [nranks][4] (this process); the entry of a rank may be given
// either way, a pointer wins.  Entries of kinds this decomposition does not use are ignored.
int group_connect(nbx_ctx *c, const void *handles, void *const *ptrs)
{
    Comm &m = c->comm;
    if (!m.on) return fail(c, NBX_ERR_INVALID, "nbx_group_connect: call nbx_group_init first");
    if (m.nranks == 1) return NBX_OK;
    const unsigned char *hb = static_cast<const unsigned char *>(handles);
    auto resolve = [&](int r, int kind, void **out) -> int {
        *out = nullptr;
        if (ptrs && ptrs[r * 4 + kind]) { *out = ptrs[r * 4 + kind]; return NBX_OK; }
        if (!hb) return NBX_OK;
        const unsigned char *h = hb + ((size_t)r * 4 + kind) * 64;
        bool zero = true;
        for (int k = 0; k < 64 && zero; ++k) zero = h[k] == 0;
        if (zero) return NBX_OK;
        cudaIpcMemHandle_t ih;
        memcpy(&ih, h, sizeof ih);
        void *p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return cuda_fail(c, e, "cudaIpcOpenMemHandle (peer memory of the group)");
        m.ipc_opened.push_back(p);
        *out = p;
        return NBX_OK;
    };
    for (int r = 0; r < m.nranks; ++r) {
        if (r == m.rank) continue;
        void *p = nullptr;
        NBX_TRY(resolve(r, 0, &p));
        if (!p) return fail(c, NBX_ERR_INVALID, "nbx_group_connect: no window for rank %d", r);
        m.peer_win[r] = static_cast<double *>(p);
        if (m.mode != 3) {
            NBX_TRY(resolve(r, 2, &p));
            if (!p) return fail(c, NBX_ERR_INVALID, "nbx_group_connect: no position rows for rank %d", r);
            m.peer_pos[r] = static_cast<double *>(p);
        }
        if (m.mode == 1) {
            NBX_TRY(resolve(r, 3, &p));
            if (!p) return fail(c, NBX_ERR_INVALID, "nbx_group_connect: no staging area for rank %d", r);
            m.peer_stage[r] = static_cast<double *>(p);
        }
    }
    if (m.mode == 3) {
        const int left = (m.rank + m.nranks - 1) % m.nranks, right = (m.rank + 1) % m.nranks;
        void *pl = nullptr, *pr = nullptr;
        NBX_TRY(resolve(left, 1, &pl));
        if (right == left) pr = pl; else NBX_TRY(resolve(right, 1, &pr));
        if (!pl || !pr) return fail(c, NBX_ERR_INVALID, "nbx_group_connect: no receive area for a neighbouring slab");
        NBX_TRY(slab_connect(c, nullptr, nullptr, pl, pr));
    }
    return NBX_OK;
}

// enqueue the initial distribution (slabs); the caller synchronises every rank afterwards (nbx_slab_check / nbx_synchronize)
int group_start(nbx_ctx *c)
{
    if (!c->comm.on) return fail(c, NBX_ERR_INVALID, "nbx_group_start: call nbx_group_init first");
    if (c->comm.mode == 3) return slab_start(c);
    return NBX_OK;
}

// ------------------------------------------------------------------------------------------------
// step loops of the all-pairs decompositions
// ------------------------------------------------------------------------------------------------
static int red_grid(const nbx_ctx *c, int64_t n)
{
    int64_t b = (n + 255) / 256;
    const int64_t cap = (int64_t)c->sm_count * 4;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

static int push_positions(nbx_ctx *c, double dt, int update)
{
    Comm &m = c->comm;
    const PushArgs pa = push_args(c, m.peer_pos);
    timer_begin(c, NBX_T_INTEGRATE);
    vv_pos_push_kernel<<<red_grid(c, c->tgt_hi - c->tgt_lo), 256, 0, c->stream>>>(pa, c->vel, c->acc, c->npad, c->tgt_lo, c->tgt_hi, dt,
                                                                                0.5 * dt * dt, update);
    comm_wait_kernel<<<1, 32, 0, c->stream>>>(comm_dev(c), kWinPos, SEQ_POS);
    timer_end(c, NBX_T_INTEGRATE);
    NBX_CUDA(c, cudaGetLastError());
    return NBX_OK;
}

// pos (all bodies) -> acc rows of the own block: pair shares pushed to the owners and added there, or the own targets
static int group_forces(nbx_ctx *c)
{
    Comm &m = c->comm;
    NBX_TRY(compute_pairs(c));
    if (m.mode == 1) {
        const PushArgs pa = push_args(c, m.peer_stage);
        timer_begin(c, NBX_T_INTEGRATE);
        acc_push_kernel<<<red_grid(c, c->n), 256, 0, c->stream>>>(pa, c->acc, c->npad, c->n, m.per);
        acc_sum_kernel<<<red_grid(c, c->tgt_hi - c->tgt_lo), 256, 0, c->stream>>>(comm_dev(c), m.stage, m.per, c->acc, c->npad, c->tgt_lo,
                                                                                c->tgt_hi);
        timer_end(c, NBX_T_INTEGRATE);
        NBX_CUDA(c, cudaGetLastError());
    }
    return NBX_OK;
}

static bool needs_T(const nbx_ctx *c) { return c->thermo == NBX_THERMO_BERENDSEN; }

// the local sum m v^2 of d_scal[0] -> the sum over the ranks in d_scal[12] (and [0] when `both`)
static int sum_T(nbx_ctx *c, bool to_slot0)
{
    NBX_CUDA(c, cudaMemsetAsync(c->d_scal + 13, 0, 2 * sizeof(double), c->stream));
    NBX_TRY(comm_allreduce3(c, c->d_scal, c->d_scal + 12, nullptr));
    if (to_slot0) NBX_CUDA(c, cudaMemcpyAsync(c->d_scal, c->d_scal + 12, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    c->slab.scal0_global = to_slot0;
    return NBX_OK;
}

static int group_vv_step(nbx_ctx *c, double dt)
{
    NBX_TRY(push_positions(c, dt, 1));
    if (needs_T(c)) NBX_TRY(sum_T(c, false)); // T of v(t), summed over the ranks, before the velocity update reads it
    double *t = c->acc_old; c->acc_old = c->acc; c->acc = t;
    NBX_TRY(group_forces(c));
    NBX_TRY(launch_vv_vel(c, dt, true));
    if (c->thermo == NBX_THERMO_ANDERSEN) NBX_TRY(launch_andersen(c, dt));
    return NBX_OK;
}

} // namespace nbx

#include "nbx_graph.inl"

namespace nbx {

// nbx_step_vv of a group member (pairs / targets); enqueues only
int multi_enqueue_vv(nbx_ctx *c, double dt, int64_t nsteps)
{
    Comm &m = c->comm;
    if (!m.on) return fail(c, NBX_ERR_INVALID, "not a group member");
    if (m.mode == 3) return slab_enqueue(c, dt, nsteps);
    if (c->thermo == NBX_THERMO_LANGEVIN)
        return fail(c, NBX_ERR_UNSUPPORTED, "nbx_step_vv: the Langevin thermostat is an SDE (use nbx_step_em), as in run_simulation");
    if (m.nranks > 1 && !m.peer_pos[(m.rank + 1) % m.nranks]) return fail(c, NBX_ERR_INVALID, "nbx_step_vv: call nbx_group_connect first");
    if (needs_T(c)) NBX_TRY(comm_arm_global0(c));
    // Andersen draws from a host-side step counter: eager
    NBX_TRY(steps_graphed(c, 10 + m.mode, dt, nsteps, c->thermo != NBX_THERMO_ANDERSEN, 2, [&]() { return group_vv_step(c, dt); }));
    if (needs_T(c)) NBX_TRY(sum_T(c, true)); // leave the scalar block as a single context does: [0] = the global sum
    return NBX_OK;
}

// Euler-Maruyama on the Langevin SDE over the group (src/nbody_to_ode.jl:575-595): the noise is keyed by the global column
// index, so the trajectory is the single-context one whatever the number of ranks
int multi_enqueue_em(nbx_ctx *c, double dt, int64_t nsteps)
{
    Comm &m = c->comm;
    if (m.mode == 3) return fail(c, NBX_ERR_UNSUPPORTED, "nbx_step_em: not available on a slab decomposition (use group mode 2)");
    const bool pairs = m.mode == 1;
    auto eval = [&]() -> int {
        if (pairs) return group_forces(c);
        NBX_TRY(compute_pairs(c));
        return NBX_OK;
    };
    for (int64_t s = 0; s < nsteps; ++s) {
        if (s > 0) NBX_TRY(eval()); // a(x_s); the first one is resident already
        NBX_TRY(launch_em_step(c, dt));   // own block: x += dt v; v += dt (a - gamma v) + sigma sqrt(dt) xi
        NBX_TRY(push_positions(c, dt, 0));
    }
    NBX_TRY(eval()); // leave a(x_end) resident
    return NBX_OK;
}

int multi_finish(nbx_ctx *c)
{
    NBX_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->comm.mode == 3) return slab_finish(c);
    int h[SEQ_N];
    NBX_CUDA(c, cudaMemcpy(h, c->comm.d_seq, sizeof h, cudaMemcpyDeviceToHost));
    if (h[SEQ_TIMEOUT]) return fail(c, NBX_ERR_CUDA, "group exchange: timed out waiting for a peer (a rank did not make the same call?)");
    return NBX_OK;
}

// ------------------------------------------------------------------------------------------------
// RHS drop-in over the group: soode_system!(dv, v, u, p, t) with HOST u, v; every rank uploads only its own block,
// the blocks are all-gathered over NVLink by the push kernel, and only the own columns of dv go back to the host
// ------------------------------------------------------------------------------------------------
void maybe_pin(nbx_ctx *c, const void *p, size_t bytes)
{
    if (!c->opt_pin_host || !p) return;
    if (c->leader && c != c->leader->members[0]) return; // one registration (portable) serves every member of a leader
    for (auto &e : c->pinned)
        if (e.first == p && e.second >= bytes) return;
    if (cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterPortable) == cudaSuccess) c->pinned.emplace_back(p, bytes);
    else cudaGetLastError();
}

// Warm: the own block of u goes up, the push kernel all-gathers it.  Cold (the first call of a group that has not evaluated
// anything yet): every rank uploads ALL of u and evaluates before any kernel waits for a peer, then synchronises -- the
// evaluation allocates its buffers, and an allocation between two members' launches would serialise their streams.
int multi_accel_enqueue(nbx_ctx *c, const double *u, const double *v)
{
    Comm &m = c->comm;
    if (m.mode == 3) return fail(c, NBX_ERR_UNSUPPORTED, "nbx_accel: the context is slab-decomposed (group mode 2 serves the RHS drop-in of cutoff systems)");
    const bool cold = !m.warm;
    const int64_t lo = cold ? 0 : c->tgt_lo, hi = cold ? c->n : c->tgt_hi, cnt = hi - lo;
    const bool have_v = c->thermo == NBX_THERMO_BERENDSEN;
    if (have_v && !v) return fail(c, NBX_ERR_INVALID, "nbx_accel: this thermostat needs v");
    const size_t bytes = sizeof(double) * 3 * (size_t)c->ncols;
    maybe_pin(c, u, bytes);
    if (have_v) maybe_pin(c, v, bytes);
    if (cnt > 0) {
        NBX_CUDA(c, cudaMemcpyAsync(c->aos_u + 3 * lo, u + 3 * lo, sizeof(double) * 3 * (size_t)cnt, cudaMemcpyHostToDevice, c->stream));
        NBX_TRY(launch_aos_to_soa(c, c->aos_u + 3 * lo, c->pos + lo, cnt));
        NBX_TRY(check_finite(c, c->pos + lo, cnt));
    }
    if (!cold) NBX_TRY(push_positions(c, 0.0, 0));
    if (have_v && c->tgt_hi > c->tgt_lo) {
        const int64_t vlo = c->tgt_lo, vcnt = c->tgt_hi - vlo;
        NBX_CUDA(c, cudaMemcpyAsync(c->aos_v + 3 * vlo, v + 3 * vlo, sizeof(double) * 3 * (size_t)vcnt, cudaMemcpyHostToDevice, c->stream));
        NBX_TRY(launch_aos_to_soa(c, c->aos_v + 3 * vlo, c->vel + vlo, vcnt));
    }
    NBX_TRY(compute_pairs(c));
    if (have_v) NBX_TRY(launch_sum_mv2(c, c->vel, c->tgt_lo, c->tgt_hi));
    if (cold) {
        NBX_CUDA(c, cudaStreamSynchronize(c->stream));
        m.warm = true;
    }
    return NBX_OK;
}

// second half: everything that waits for the peers (a leader enqueues the first half of EVERY member before this one)
int multi_accel_exchange(nbx_ctx *c)
{
    Comm &m = c->comm;
    const bool have_v = c->thermo == NBX_THERMO_BERENDSEN;
    if (m.mode == 1) {
        const PushArgs pa = push_args(c, m.peer_stage);
        acc_push_kernel<<<red_grid(c, c->n), 256, 0, c->stream>>>(pa, c->acc, c->npad, c->n, m.per);
        acc_sum_kernel<<<red_grid(c, c->tgt_hi - c->tgt_lo), 256, 0, c->stream>>>(comm_dev(c), m.stage, m.per, c->acc, c->npad, c->tgt_lo,
                                                                                c->tgt_hi);
        NBX_CUDA(c, cudaGetLastError());
    }
    if (have_v) {
        c->slab.scal0_global = false;
        NBX_TRY(sum_T(c, true));   // d_scal[0] = the sum over the ranks, which the RHS term reads
        NBX_TRY(launch_thermostat_rhs(c, c->acc, c->vel));
    }
    const int64_t olo = c->tgt_lo, ocnt = c->tgt_hi - olo;
    if (ocnt > 0) NBX_TRY(launch_soa_to_aos(c, c->acc + olo, c->aos_dv + 3 * olo, ocnt, ocnt, 0, ocnt));
    c->resident = false;
    return NBX_OK;
}

int multi_accel_finish(nbx_ctx *c, double *dv)
{
    const int64_t lo = c->tgt_lo, cnt = c->tgt_hi - lo;
    if (cnt > 0)
        NBX_CUDA(c, cudaMemcpyAsync(dv + 3 * lo, c->aos_dv + 3 * lo, sizeof(double) * 3 * (size_t)cnt, cudaMemcpyDeviceToHost, c->stream));
    int flag[2] = {0, 0};
    NBX_CUDA(c, cudaMemcpyAsync(flag, c->d_scal + 15, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    NBX_CUDA(c, cudaStreamSynchronize(c->stream));
    if (flag[0]) return fail(c, NBX_ERR_NONFINITE, "non-finite coordinate in u (the reference's wrap loop would not terminate)");
    return multi_finish(c);
}

// CUDA loads kernels lazily, at their first launch, and loading may synchronise the context: a kernel that is first
// launched while another member's kernel spins on a flag would deadlock the pair (CUDA programming guide, "Lazy
// Loading": concurrent execution).  Every kernel of the distributed loops is therefore loaded when a context is created.
void preload_multi()
{
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, comm_allreduce3_kernel);
    cudaFuncGetAttributes(&a, vv_pos_push_kernel);
    cudaFuncGetAttributes(&a, comm_wait_kernel);
    cudaFuncGetAttributes(&a, acc_push_kernel);
    cudaFuncGetAttributes(&a, acc_sum_kernel);
    cudaGetLastError();
}

} // namespace nbx
