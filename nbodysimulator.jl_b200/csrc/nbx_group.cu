// nbx_group.cu -- nbx_create_multi: ONE context handle for several GPUs of one process.
//
// What a Julia task needs (north star: "Julia host code calls a thin C-ABI shared library through ccall"): the same
// nbx_system / nbx_add_* / nbx_accel / nbx_upload / nbx_step_vv / nbx_download calls as on one GPU, fanned out over the
// members by the library.  The leader holds no device state; every member is an ordinary context on its GPU, joined into
// a group by the calls another process would make (nbx_group_init / _export / _connect / _start, nbx_multi.cu) -- here
// with plain device pointers and peer access instead of CUDA IPC handles.  All exchanges are device-side (peer-memory
// stores + flags), so the single host thread only has to ENQUEUE every member's work before it waits for any of them.
#include "nbx_internal.cuh"

#include <algorithm>
#include <cmath>
#include <new>

namespace nbx {

#define NBX_MEMBER(c, m, expr)                                   \
    do {                                                         \
        const int rc__ = (expr);                                 \
        if (rc__ != NBX_OK) { (c)->err = (m)->err; return rc__; } \
    } while (0)

int leader_create(nbx_ctx **out, int ndev, const int *devs)
{
    if (!out) return fail(nullptr, NBX_ERR_INVALID, "nbx_create_multi: out is NULL");
    *out = nullptr;
    if (ndev < 1 || ndev > kMaxRanks || !devs) return fail(nullptr, NBX_ERR_INVALID, "nbx_create_multi: 1 .. %d devices", kMaxRanks);
    nbx_ctx *lead = new (std::nothrow) nbx_ctx();
    if (!lead) return fail(nullptr, NBX_ERR_INVALID, "nbx_create_multi: out of host memory");
    lead->is_group = true;
    lead->device = devs[0];
    for (int k = 0; k < ndev; ++k) {
        nbx_ctx *m = nullptr;
        const int rc = nbx_create(&m, devs[k]);
        if (rc != NBX_OK) { // nbx_last_error(NULL) holds the message
            for (nbx_ctx *x : lead->members) nbx_destroy(x);
            delete lead;
            return rc;
        }
        m->leader = lead;
        // a slab never orders its lists by record position (local slot numbers); the members evaluate a(0) before they
        // become slabs, and that evaluation must add in the same order as the steps that follow
        m->opt_verlet_banked = 0;
        lead->members.push_back(m);
    }
    // peer access between every pair of distinct devices (the kernels store into each other's memory)
    for (int a = 0; a < ndev; ++a)
        for (int b = 0; b < ndev; ++b) {
            if (devs[a] == devs[b]) continue;
            int can = 0;
            cudaSetDevice(devs[a]);
            cudaDeviceCanAccessPeer(&can, devs[a], devs[b]);
            if (!can) {
                for (nbx_ctx *x : lead->members) nbx_destroy(x);
                delete lead;
                return fail(nullptr, NBX_ERR_CUDA, "nbx_create_multi: device %d cannot access device %d (no peer path)", devs[a], devs[b]);
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(devs[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                for (nbx_ctx *x : lead->members) nbx_destroy(x);
                delete lead;
                return cuda_fail(nullptr, e, "cudaDeviceEnablePeerAccess");
            }
            cudaGetLastError();
        }
    *out = lead;
    return NBX_OK;
}

int leader_destroy(nbx_ctx *c)
{
    for (nbx_ctx *m : c->members) nbx_destroy(m);
    delete c;
    return NBX_OK;
}

int leader_system(nbx_ctx *c, int64_t n, const double *m, const double *q, const double *mm, int water)
{
    if (n <= 0 || !m) return fail(c, NBX_ERR_INVALID, "nbx_system: n > 0 and masses are required");
    c->g_m.assign(m, m + n);
    if (q) c->g_q.assign(q, q + n); else c->g_q.clear();
    if (mm) c->g_mm.assign(mm, mm + 3 * n); else c->g_mm.clear();
    c->g_water = water;
    c->n = n;
    c->g_ready = false;
    for (nbx_ctx *x : c->members) NBX_MEMBER(c, x, nbx_system(x, n, m, q, mm, water));
    c->ncols = c->members[0]->ncols;
    return NBX_OK;
}

// members -> one group: init, exchange the device pointers, connect, start (slabs: the initial distribution)
static int leader_join(nbx_ctx *c, int mode)
{
    const int R = (int)c->members.size();
    for (int k = 0; k < R; ++k) NBX_MEMBER(c, c->members[k], (cudaSetDevice(c->members[k]->device), group_init(c->members[k], k, R, mode)));
    std::vector<void *> ptrs((size_t)R * 4, nullptr);
    for (int k = 0; k < R; ++k)
        for (int kind = 0; kind < 4; ++kind) NBX_MEMBER(c, c->members[k], group_export(c->members[k], kind, &ptrs[(size_t)k * 4 + kind], nullptr));
    for (int k = 0; k < R; ++k) NBX_MEMBER(c, c->members[k], (cudaSetDevice(c->members[k]->device), group_connect(c->members[k], nullptr, ptrs.data())));
    for (int k = 0; k < R; ++k) NBX_MEMBER(c, c->members[k], (cudaSetDevice(c->members[k]->device), group_start(c->members[k])));
    for (int k = 0; k < R; ++k) {
        nbx_ctx *x = c->members[k];
        cudaSetDevice(x->device);
        if (x->comm.mode == 3) NBX_MEMBER(c, x, slab_check(x, nullptr));
        else NBX_MEMBER(c, x, nbx_synchronize(x));
    }
    c->g_ready = true;
    return NBX_OK;
}

int leader_upload(nbx_ctx *c, const double *u, const double *v)
{
    if (c->g_m.empty()) return fail(c, NBX_ERR_INVALID, "nbx_upload: call nbx_system first");
    const bool regroup = c->g_ready || c->members[0]->comm.on;
    for (nbx_ctx *x : c->members) {
        if (regroup) // the members were sharded / compacted: start over from the system description
            NBX_MEMBER(c, x, nbx_system(x, c->n, c->g_m.data(), c->g_q.empty() ? nullptr : c->g_q.data(),
                                        c->g_mm.empty() ? nullptr : c->g_mm.data(), c->g_water));
        NBX_MEMBER(c, x, nbx_upload(x, u, v));
    }
    c->g_ready = false;
    return leader_join(c, c->opt_group_mode);
}

int leader_accel(nbx_ctx *c, const double *u, double *v, double *dv)
{
    if (c->g_m.empty()) return fail(c, NBX_ERR_INVALID, "nbx_accel: call nbx_system first");
    if (!u || !dv) return fail(c, NBX_ERR_INVALID, "nbx_accel: u and dv are required");
    if (c->g_ready && c->members[0]->comm.mode == 3)
        return fail(c, NBX_ERR_UNSUPPORTED, "nbx_accel: the group is slab-decomposed (nbx_upload chose slabs); set option group_mode = 2 for the RHS drop-in");
    if (!c->g_ready) {
        int mode = c->opt_group_mode;
        if (mode == 0 || mode == 3) mode = -1; // pairs where possible, else targets
        if (mode == -1) {
            mode = pair_capable(c->members[0]) ? 1 : 2;
        }
        NBX_TRY(leader_join(c, mode));
    }
    // page-locking (option pin_host) is an allocation-class call: before anything is enqueued, never between two members' launches
    const size_t bytes = sizeof(double) * 3 * (size_t)c->ncols;
    maybe_pin(c->members[0], u, bytes);
    maybe_pin(c->members[0], dv, bytes);
    if (v) maybe_pin(c->members[0], v, bytes);
    for (nbx_ctx *x : c->members) NBX_MEMBER(c, x, (cudaSetDevice(x->device), multi_accel_enqueue(x, u, v)));
    for (nbx_ctx *x : c->members) NBX_MEMBER(c, x, (cudaSetDevice(x->device), multi_accel_exchange(x)));
    for (nbx_ctx *x : c->members) NBX_MEMBER(c, x, (cudaSetDevice(x->device), multi_accel_finish(x, dv)));
    return NBX_OK;
}

int leader_step_vv(nbx_ctx *c, double dt, int64_t nsteps)
{
    if (!c->g_ready || !c->members[0]->resident) return fail(c, NBX_ERR_INVALID, "nbx_step_vv: no resident state (call nbx_upload)");
    // every member's kernels wait for the other members' kernels: enqueue in bounded chunks, member after member, so
    // that no launch queue fills up while its device waits for work that has not been enqueued yet
    // (graph replay: a chunk is a few dozen launches per member; eager -- option graph = 0, timers on, Andersen --: a step is
    // some 25 launches, so the chunks are short)
    const nbx_ctx *m0 = c->members[0];
    const bool eager = !m0->opt_graph || m0->timing || m0->thermo == NBX_THERMO_ANDERSEN;
    const int64_t chunk = eager ? 2 : 64;
    for (int64_t done = 0; done < nsteps; done += chunk) {
        const int64_t k = std::min(chunk, nsteps - done);
        for (nbx_ctx *x : c->members) NBX_MEMBER(c, x, (cudaSetDevice(x->device), multi_enqueue_vv(x, dt, k)));
    }
    for (nbx_ctx *x : c->members) NBX_MEMBER(c, x, (cudaSetDevice(x->device), multi_finish(x)));
    return NBX_OK;
}

int leader_step_em(nbx_ctx *c, double dt, int64_t nsteps, uint64_t seed)
{
    if (!c->g_ready || !c->members[0]->resident) return fail(c, NBX_ERR_INVALID, "nbx_step_em: no resident state (call nbx_upload)");
    for (nbx_ctx *x : c->members) {
        if (x->thermo != NBX_THERMO_LANGEVIN) return fail(c, NBX_ERR_INVALID, "nbx_step_em: needs the Langevin thermostat");
        if (x->water) return fail(c, NBX_ERR_UNSUPPORTED, "nbx_step_em: the water SDE variant (src/nbody_to_ode.jl:600-680) is not built");
        if (seed) x->seed = seed;
    }
    const int64_t chunk = 4;
    for (int64_t done = 0; done < nsteps; done += chunk) {
        const int64_t k = std::min(chunk, nsteps - done);
        for (nbx_ctx *x : c->members) NBX_MEMBER(c, x, (cudaSetDevice(x->device), multi_enqueue_em(x, dt, k)));
    }
    for (nbx_ctx *x : c->members) NBX_MEMBER(c, x, (cudaSetDevice(x->device), multi_finish(x)));
    return NBX_OK;
}

int leader_download(nbx_ctx *c, double *u, double *v, double *dv)
{
    if (!c->g_ready || !c->members[0]->resident) return fail(c, NBX_ERR_INVALID, "nbx_download: no resident state (call nbx_upload)");
    const int64_t n = c->n;
    if (c->members[0]->comm.mode == 3) {
        std::vector<int32_t> gid((size_t)n);
        std::vector<double> bu(u ? 3 * (size_t)n : 0), bv(v ? 3 * (size_t)n : 0), ba(dv ? 3 * (size_t)n : 0);
        std::vector<char> seen((size_t)n, 0);
        for (nbx_ctx *x : c->members) {
            int64_t own = 0;
            NBX_MEMBER(c, x, nbx_slab_download(x, &own, gid.data(), u ? bu.data() : nullptr, v ? bv.data() : nullptr, dv ? ba.data() : nullptr));
            for (int64_t k = 0; k < own; ++k) {
                const int64_t g = gid[(size_t)k];
                if (g < 0 || g >= n || seen[(size_t)g]) return fail(c, NBX_ERR_INVALID, "nbx_download: slab ownership is not a partition (particle %lld)", (long long)g);
                seen[(size_t)g] = 1;
                for (int d = 0; d < 3; ++d) {
                    if (u) u[3 * g + d] = bu[3 * (size_t)k + d];
                    if (v) v[3 * g + d] = bv[3 * (size_t)k + d];
                    if (dv) dv[3 * g + d] = ba[3 * (size_t)k + d];
                }
            }
        }
        for (int64_t g = 0; g < n; ++g)
            if (!seen[(size_t)g]) return fail(c, NBX_ERR_INVALID, "nbx_download: particle %lld is owned by no slab", (long long)g);
        return NBX_OK;
    }
    for (nbx_ctx *x : c->members) NBX_MEMBER(c, x, nbx_download(x, x == c->members[0] ? u : nullptr, v, dv));
    return NBX_OK;
}

int leader_energy(nbx_ctx *c, double *ekin, double *epot, double *temperature)
{
    if (!c->g_ready) return fail(c, NBX_ERR_INVALID, "nbx_energy: no resident state (call nbx_upload)");
    if (c->members[0]->comm.mode == 3) { // every slab holds the global sum m v^2 after a run
        if (epot) return fail(c, NBX_ERR_UNSUPPORTED, "nbx_energy: the potential energy of a slab-decomposed system is not available");
        NBX_MEMBER(c, c->members[0], nbx_energy(c->members[0], ekin, nullptr, temperature));
        return NBX_OK;
    }
    double ek = 0.0, T = 0.0;
    for (nbx_ctx *x : c->members) {
        double e = 0.0, t = 0.0;
        NBX_MEMBER(c, x, nbx_energy(x, &e, nullptr, &t));
        ek += e; T += t;
    }
    if (ekin) *ekin = ek;
    if (temperature) *temperature = T;
    if (epot) NBX_MEMBER(c, c->members[0], nbx_energy(c->members[0], nullptr, epot, nullptr));
    return NBX_OK;
}

} // namespace nbx
