// nbx_api.cu -- the C ABI of include/nbody_b200.h: context lifetime, system description, the RHS
// drop-in (soode_system!, src/nbody_to_ode.jl:474-488 and :502-532) and device-resident stepping.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <new>

#include "nbx_internal.cuh"

#include <algorithm>
#include <cmath>
#include <utility>

namespace nbx {

static thread_local std::string g_create_error;

int fail(nbx_ctx *c, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

int cuda_fail(nbx_ctx *c, cudaError_t e, const char *what)
{
    cudaGetLastError(); // clear the sticky-free error state
    return fail(c, NBX_ERR_CUDA, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
}

// ---- timers ----------------------------------------------------------------------------------
static void timer_flush(nbx_ctx *c, Timer &t)
{
    if (t.used == 0) return;
    cudaStreamSynchronize(c->stream);
    for (size_t k = 0; k + 1 < t.used; k += 2) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, t.ev[k], t.ev[k + 1]) == cudaSuccess) { t.total_ms += ms; t.count++; }
    }
    t.used = 0;
}

void timer_begin(nbx_ctx *c, int phase)
{
    if (!c->timing) return;
    Timer &t = c->timers[phase];
    if (t.used + 2 > t.ev.size()) {
        if (t.ev.size() >= (size_t)2 * kMaxTimers) timer_flush(c, t);
        else {
            const size_t grow = t.ev.size() + 64;
            while (t.ev.size() < grow) { cudaEvent_t e; cudaEventCreate(&e); t.ev.push_back(e); }
        }
    }
    cudaEventRecord(t.ev[t.used], c->stream);
}

void timer_end(nbx_ctx *c, int phase)
{
    if (!c->timing) return;
    Timer &t = c->timers[phase];
    cudaEventRecord(t.ev[t.used + 1], c->stream);
    t.used += 2;
}

// ---- small kernels that belong to the RHS assembly -----------------------------------------------
// oxygen columns of water (every third, src/nbody_to_ode.jl:302-314) <-> compact SoA
__global__ void gather_oxygen_kernel(const double *__restrict__ pos, int64_t ld, int nmol, double *__restrict__ opos,
                                     int64_t old, double far)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= old) return;
#pragma unroll
    for (int d = 0; d < 3; ++d) opos[d * old + m] = m < nmol ? pos[d * ld + 3 * (int64_t)m] : far;
}

__global__ void scatter_oxygen_kernel(const double *__restrict__ oacc, int64_t old, int mlo, int mhi,
                                      double *__restrict__ acc, int64_t ld)
{
    const int m = mlo + blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= mhi) return;
#pragma unroll
    for (int d = 0; d < 3; ++d) acc[d * ld + 3 * (int64_t)m] += oacc[d * old + m];
}

static int zero_rows(nbx_ctx *c, double *rows, int64_t lo, int64_t hi)
{
    if (hi <= lo) return NBX_OK;
    for (int d = 0; d < 3; ++d)
        NBX_CUDA(c, cudaMemsetAsync(rows + d * c->npad + lo, 0, sizeof(double) * (size_t)(hi - lo), c->stream));
    return NBX_OK;
}

// Pair potentials: pos -> acc.  Target sharding: the columns [tgt_lo, tgt_hi) get their full
// accelerations.  Pair sharding (pair_nranks > 1): ALL columns get this rank's partial sums, which the
// host adds across ranks (reduce-scatter) before nbx_vv_finish.  Everything stays on the stream.
int compute_pairs(nbx_ctx *c)
{
    const bool pair_mode = c->pair_nranks > 1;
    if (pair_mode) {
        if (!pair_capable(c))
            return fail(c, NBX_ERR_UNSUPPORTED, "pair sharding covers unbounded gravity / Coulomb, and Coulomb with a cutoff of "
                        "L/3 <= R < L/2 in a cubic box; use nbx_shard");
        NBX_TRY(zero_rows(c, c->acc, 0, c->n));
    }
    const int64_t lo = c->tgt_lo, hi = c->tgt_hi;
    // the first potential overwrites its targets' rows; only terms that can but add need zeroed rows
    bool written = pair_mode;
    auto acc_flag = [&]() { const bool a = written; written = true; return a; };
    auto ensure_zero = [&]() -> int {
        if (!written) NBX_TRY(zero_rows(c, c->acc, lo, hi));
        written = true;
        return NBX_OK;
    };

    if (c->has_lj) {
        if (c->water) {
            const int nmol = (int)(c->n / 3);
            const int mlo = (int)((lo + 2) / 3), mhi = (int)((hi + 2) / 3);
            NBX_TRY(ensure_zero());
            gather_oxygen_kernel<<<(unsigned)((c->opad + 255) / 256), 256, 0, c->stream>>>(c->pos, c->npad, nmol, c->opos,
                                                                                         c->opad, kFarAway);
            NBX_CUDA(c, cudaGetLastError());
            bool used = false;
            NBX_TRY(cells_pairs(c, &c->cl_lj, c->lj_R, 0, c->opos, nullptr, nullptr, nmol, nmol, c->opad, 1, mlo, mhi, 3, c->oacc,
                                c->opad, false, &used));
            if (!used) NBX_TRY(launch_allpairs_pbc(c, 0, c->opos, nmol, c->opad, mlo, mhi, 3, c->oacc, c->opad, false));
            if (mhi > mlo) {
                scatter_oxygen_kernel<<<(unsigned)((mhi - mlo + 255) / 256), 256, 0, c->stream>>>(c->oacc, c->opad, mlo, mhi,
                                                                                                c->acc, c->npad);
                NBX_CUDA(c, cudaGetLastError());
            }
        } else {
            bool used = false;
            const bool accum = acc_flag();
            NBX_TRY(cells_pairs(c, &c->cl_lj, c->lj_R, 0, c->pos, nullptr, c->gid, c->n, c->slab.on ? c->slab.n_total : c->n,
                                c->npad, 1, lo, hi, 1, c->acc, c->npad, accum, &used));
            if (!used) NBX_TRY(launch_allpairs_pbc(c, 0, c->pos, c->n, c->npad, lo, hi, 1, c->acc, c->npad, accum));
        }
    }
    if (c->has_coul) {
        const int pot = c->water ? 2 : 1;
        if (!c->water && c->bc_kind == NBX_BC_INFINITE && isinf(c->el_R2)) {
            // F += q_j (ri - rj)/r^3, dv += k q_i / m_i F  ==  -k q_i/m_i * sum q_j (rj - ri)/r^3
            NBX_TRY(launch_allpairs_grav(c, c->charge, 1, -c->el_k, c->acc, acc_flag()));
        } else {
            bool used = false;
            const bool accum = acc_flag();
            NBX_TRY(cells_pairs(c, &c->cl_el, c->el_R, pot, c->pos, c->charge, c->gid, c->n, c->slab.on ? c->slab.n_total : c->n,
                                c->npad, c->water ? 3 : 1, lo, hi, 1, c->acc, c->npad, accum, &used));
            if (!used) {
                // no cell list for this cutoff (R >= L/3): all pairs with the exact periodic predicate -- once per UNORDERED pair
                // when the context evaluates every target, the box is cubic and R < L/2 (no accepted pair on the wrap tie)
                // (pair sharding: this rank's ring offsets, for all bodies -- the other terms above covered the own block only)
                const bool whole = (pair_mode || (lo == 0 && hi == c->n)) && !c->slab.on;
                if (whole && c->opt_sym && c->n >= c->sym_min_n && c->bc_kind == NBX_BC_CUBIC && c->el_R < 0.5 * c->bc[0])
                    NBX_TRY(launch_sympairs_coulomb_pbc(c, pot == 2, c->acc, accum));
                else
                    NBX_TRY(launch_allpairs_pbc(c, pot, c->pos, c->n, c->npad, lo, hi, 1, c->acc, c->npad, accum));
            }
        }
    }
    if (c->has_dip) NBX_TRY(launch_allpairs_dipole(c, c->acc, acc_flag()));
    if (c->has_grav) NBX_TRY(launch_allpairs_grav(c, c->mass, 0, c->G, c->acc, acc_flag()));
    if (c->has_spcfw) {
        NBX_TRY(ensure_zero());
        NBX_TRY(launch_spcfw_bonded(c, c->acc));
    }
    NBX_TRY(ensure_zero()); // no potential at all
    return NBX_OK;
}

// pos, vel (+ d_scal) -> acc: pair potentials, then the RHS thermostat terms (src/nbody_to_ode.jl:484-486)
int compute_accel(nbx_ctx *c)
{
    if (c->pair_nranks > 1)
        return fail(c, NBX_ERR_INVALID, "pair-sharded context: drive nbx_vv_forces / reduce / nbx_vv_finish");
    NBX_TRY(compute_pairs(c));
    return launch_thermostat_rhs(c, c->acc, c->vel);
}

static void free_system(nbx_ctx *c)
{
    cudaFree(c->mass); cudaFree(c->charge); cudaFree(c->mm);
    cudaFree(c->pos); cudaFree(c->vel); cudaFree(c->acc); cudaFree(c->acc_old);
    cudaFree(c->aos_u); cudaFree(c->aos_v); cudaFree(c->aos_dv);
    cudaFree(c->opos); cudaFree(c->oacc);
    c->mass = c->charge = c->mm = c->pos = c->vel = c->acc = c->acc_old = nullptr;
    c->aos_u = c->aos_v = c->aos_dv = c->opos = c->oacc = nullptr;
    cells_free(&c->cl_lj);
    cells_free(&c->cl_el);
    analysis_free(c);
    c->T_slot = 0;
    slab_free(c);
    comm_free(c);
    graph_drop(c);
    c->tgt_lo = c->tgt_hi = 0;
    c->pair_rank = 0; c->pair_nranks = 1;
    for (auto &e : c->pinned) cudaHostUnregister(const_cast<void *>(e.first));
    c->pinned.clear();
    cudaGetLastError();
    c->resident = false;
}

static int upload_row(nbx_ctx *c, double *dst, const double *src, int64_t n, double padval)
{
    NBX_CUDA(c, cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    return launch_fill(c, dst + n, padval, c->npad - n);
}

static bool needs_velocity(const nbx_ctx *c)
{
    return c->thermo == NBX_THERMO_BERENDSEN || c->thermo == NBX_THERMO_NOSEHOOVER;
}

static int guard(nbx_ctx *c)
{
    if (!c) return NBX_ERR_INVALID;
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return cuda_fail(c, e, "cudaSetDevice");
    return NBX_OK;
}

static int need_system(nbx_ctx *c, const char *who)
{
    if (c->n <= 0) return fail(c, NBX_ERR_INVALID, "%s: call nbx_system first", who);
    return NBX_OK;
}

// entry points that assume the context holds the whole system
static int no_slab(nbx_ctx *c, const char *who)
{
    if (c->slab.on)
        return fail(c, NBX_ERR_INVALID, "%s: the context is slab-decomposed (drive nbx_vv_begin / nbx_slab_pack / exchange / "
                    "nbx_slab_unpack / nbx_vv_forces / nbx_vv_finish; nbx_system starts over)", who);
    return NBX_OK;
}

static int need_resident(nbx_ctx *c, const char *who)
{
    NBX_TRY(need_system(c, who));
    if (!c->resident) return fail(c, NBX_ERR_INVALID, "%s: no resident state (call nbx_upload)", who);
    return NBX_OK;
}

// shared tail of nbx_accel / nbx_accel_device: aos_u/aos_v hold the inputs on the device
static int accel_from_staging(nbx_ctx *c, bool have_v)
{
    NBX_TRY(launch_aos_to_soa(c, c->aos_u, c->pos, c->n));
    NBX_TRY(check_finite(c, c->pos, c->n));
    if (have_v) {
        NBX_TRY(launch_aos_to_soa(c, c->aos_v, c->vel, c->n));
        NBX_TRY(launch_sum_mv2(c, c->vel, 0, c->n));
        if (c->thermo == NBX_THERMO_NOSEHOOVER) // zeta = u[zeta_ind], linear index 3n (src/thermostats.jl:122)
            NBX_CUDA(c, cudaMemcpyAsync(c->d_scal + 1, c->aos_u + 3 * c->n, sizeof(double), cudaMemcpyDeviceToDevice,
                                        c->stream));
    }
    {   // the RHS drop-in always returns complete accelerations of the own targets (no pair sharding)
        const int pr = c->pair_rank, pn = c->pair_nranks;
        c->pair_rank = 0; c->pair_nranks = 1;
        const int rc = compute_accel(c);
        c->pair_rank = pr; c->pair_nranks = pn;
        if (rc != NBX_OK) return rc;
    }
    NBX_TRY(launch_soa_to_aos(c, c->acc, c->aos_dv, c->n, c->ncols, c->tgt_lo, c->tgt_hi));
    c->resident = false;
    return NBX_OK;
}

} // namespace nbx

using namespace nbx;

// leader of a single-process group (nbx_create_multi): configuration calls go to every member
#define NBX_FANOUT(c, expr)                                              \
    if ((c) && (c)->is_group) {                                          \
        for (nbx_ctx *m__ : (c)->members) {                              \
            nbx_ctx *x = m__;                                            \
            const int rc__ = (expr);                                     \
            if (rc__ != NBX_OK) { (c)->err = x->err; return rc__; }      \
        }                                                                \
        return NBX_OK;                                                   \
    }

extern "C" {

int nbx_version(void) { return 100; }

const char *nbx_last_error(const nbx_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int nbx_create(nbx_ctx **out, int device)
{
    if (!out) return fail(nullptr, NBX_ERR_INVALID, "nbx_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, NBX_ERR_CUDA, "nbx_create: no CUDA device (%s); libnbody_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count) return fail(nullptr, NBX_ERR_INVALID, "nbx_create: device %d of %d", device, count);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major != 10)
        return fail(nullptr, NBX_ERR_CUDA, "nbx_create: device %d is sm_%d%d; this library holds sm_100a code only", device,
                    prop.major, prop.minor);
    nbx_ctx *c = new (std::nothrow) nbx_ctx();
    if (!c) return fail(nullptr, NBX_ERR_INVALID, "nbx_create: out of host memory");
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->d_scal, 16 * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(c->d_scal, 0, 16 * sizeof(double));
    if (e != cudaSuccess) {
        int rc = cuda_fail(nullptr, e, "nbx_create");
        delete c;
        return rc;
    }
    c->stream = c->own_stream;
    preload_multi(); preload_slab(); preload_cells(); preload_integrate();
    *out = c;
    return NBX_OK;
}

int nbx_create_multi(nbx_ctx **out, int ndev, const int *devs) { return leader_create(out, ndev, devs); }

int nbx_destroy(nbx_ctx *c)
{
    if (!c) return NBX_OK;
    if (c->is_group) return leader_destroy(c);
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    free_system(c);
    cudaFree(c->sym_ticket); cudaFree(c->d_scal); cudaFree(c->d_red); cudaFree(c->part);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    for (auto &t : c->timers)
        for (cudaEvent_t ev : t.ev) cudaEventDestroy(ev);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    delete c;
    return NBX_OK;
}

int nbx_system(nbx_ctx *c, int64_t n, const double *m, const double *q, const double *mm, int water)
{
    if (c && c->is_group) return leader_system(c, n, m, q, mm, water);
    NBX_TRY(guard(c));
    if (n <= 0 || n > (int64_t)1 << 30) return fail(c, NBX_ERR_INVALID, "nbx_system: n = %lld out of range", (long long)n);
    if (!m) return fail(c, NBX_ERR_INVALID, "nbx_system: masses are required");
    if (water && n % 3 != 0) return fail(c, NBX_ERR_INVALID, "nbx_system: water needs n %% 3 == 0 (O,H1,H2 triples)");
    cudaStreamSynchronize(c->stream);
    free_system(c);
    c->n = n;
    c->ncols = n + (c->thermo == NBX_THERMO_NOSEHOOVER ? 1 : 0);
    c->npad = ((n + kPad - 1) / kPad) * kPad;
    c->water = water ? 1 : 0;
    c->has_q = q != nullptr;
    c->has_mm = mm != nullptr;
    c->h_m1 = m[0];
    c->mass_uniform = true;
    for (int64_t i = 1; i < n && c->mass_uniform; ++i) c->mass_uniform = (m[i] == m[0]);
    c->charge_uniform = q != nullptr;
    c->h_q1 = q ? q[0] : 0.0;
    for (int64_t i = 1; q && i < n && c->charge_uniform; ++i) c->charge_uniform = (q[i] == q[0]);
    c->tgt_lo = 0;
    c->tgt_hi = n;
    const size_t np = (size_t)c->npad;
    NBX_TRY(dev_alloc(c, &c->mass, np));
    NBX_TRY(dev_alloc(c, &c->pos, 3 * np));
    NBX_TRY(dev_alloc(c, &c->vel, 3 * np));
    NBX_TRY(dev_alloc(c, &c->acc, 3 * np));
    NBX_TRY(dev_alloc(c, &c->acc_old, 3 * np));
    NBX_TRY(dev_alloc(c, &c->aos_u, 3 * (size_t)(n + 1)));
    NBX_TRY(dev_alloc(c, &c->aos_v, 3 * (size_t)(n + 1)));
    NBX_TRY(dev_alloc(c, &c->aos_dv, 3 * (size_t)(n + 1)));
    // padding: weights (mass, charge, moments) 0, positions far away, velocities 0
    NBX_TRY(upload_row(c, c->mass, m, n, 0.0));
    NBX_TRY(launch_fill(c, c->pos, kFarAway, 3 * c->npad));
    NBX_CUDA(c, cudaMemsetAsync(c->vel, 0, sizeof(double) * 3 * np, c->stream));
    NBX_CUDA(c, cudaMemsetAsync(c->acc, 0, sizeof(double) * 3 * np, c->stream));
    NBX_CUDA(c, cudaMemsetAsync(c->acc_old, 0, sizeof(double) * 3 * np, c->stream));
    if (q) {
        NBX_TRY(dev_alloc(c, &c->charge, np));
        NBX_TRY(upload_row(c, c->charge, q, n, 0.0));
    }
    if (mm) {
        // host moments are 3 x n AoS like the coordinates
        NBX_TRY(dev_alloc(c, &c->mm, 3 * np));
        NBX_CUDA(c, cudaMemsetAsync(c->mm, 0, sizeof(double) * 3 * np, c->stream));
        NBX_CUDA(c, cudaMemcpyAsync(c->aos_u, mm, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
        NBX_TRY(launch_aos_to_soa(c, c->aos_u, c->mm, n));
    }
    if (water) {
        c->opad = ((n / 3 + kPad - 1) / kPad) * kPad;
        NBX_TRY(dev_alloc(c, &c->opos, 3 * (size_t)c->opad));
        NBX_TRY(dev_alloc(c, &c->oacc, 3 * (size_t)c->opad));
        NBX_CUDA(c, cudaMemsetAsync(c->oacc, 0, sizeof(double) * 3 * (size_t)c->opad, c->stream));
    }
    NBX_CUDA(c, cudaStreamSynchronize(c->stream));
    return NBX_OK;
}

int nbx_boundary(nbx_ctx *c, int kind, const double *b)
{
    NBX_FANOUT(c, nbx_boundary(x, kind, b));
    NBX_TRY(guard(c));
    graph_drop(c);
    if (kind == NBX_BC_INFINITE) { c->bc_kind = kind; return NBX_OK; }
    if (!b) return fail(c, NBX_ERR_INVALID, "nbx_boundary: box is NULL");
    if (kind == NBX_BC_CUBIC) {
        if (!(b[0] > 0.0) || !isfinite(b[0])) return fail(c, NBX_ERR_INVALID, "nbx_boundary: cubic box needs 0 < L < Inf");
        c->bc_kind = kind;
        c->bc[0] = b[0];
        return NBX_OK;
    }
    if (kind == NBX_BC_PERIODIC) {
        for (int d = 0; d < 3; ++d)
            if (!(b[2 * d + 1] > b[2 * d]) || !isfinite(b[2 * d]) || !isfinite(b[2 * d + 1]))
                return fail(c, NBX_ERR_INVALID, "nbx_boundary: periodic box needs lo < hi, finite");
        c->bc_kind = kind;
        for (int k = 0; k < 6; ++k) c->bc[k] = b[k];
        return NBX_OK;
    }
    return fail(c, NBX_ERR_INVALID, "nbx_boundary: unknown kind %d", kind);
}

int nbx_add_gravity(nbx_ctx *c, double G)
{
    NBX_FANOUT(c, nbx_add_gravity(x, G));
    NBX_TRY(guard(c));
    graph_drop(c);
    c->has_grav = true; c->G = G;
    return NBX_OK;
}

int nbx_add_lj(nbx_ctx *c, double eps, double sigma, double R)
{
    NBX_FANOUT(c, nbx_add_lj(x, eps, sigma, R));
    NBX_TRY(guard(c));
    if (!(R > 0.0)) return fail(c, NBX_ERR_INVALID, "nbx_add_lj: R must be > 0");
    graph_drop(c);
    c->has_lj = true; c->lj_eps = eps;
    c->lj_sigma2 = sigma * sigma; // LennardJonesParameters caches sigma^2 and R^2 (src/basic_potentials.jl:67-69)
    c->lj_R = R; c->lj_R2 = R * R;
    return NBX_OK;
}

int nbx_add_coulomb(nbx_ctx *c, double k, double R)
{
    NBX_FANOUT(c, nbx_add_coulomb(x, k, R));
    NBX_TRY(guard(c));
    NBX_TRY(need_system(c, "nbx_add_coulomb"));
    if (!c->has_q) return fail(c, NBX_ERR_INVALID, "nbx_add_coulomb: the system has no charges");
    if (!(R > 0.0)) return fail(c, NBX_ERR_INVALID, "nbx_add_coulomb: R must be > 0 (Inf allowed)");
    graph_drop(c);
    c->has_coul = true; c->el_k = k; c->el_R = R; c->el_R2 = R * R;
    return NBX_OK;
}

int nbx_add_dipole(nbx_ctx *c, double mu_4pi)
{
    NBX_FANOUT(c, nbx_add_dipole(x, mu_4pi));
    NBX_TRY(guard(c));
    NBX_TRY(need_system(c, "nbx_add_dipole"));
    if (!c->has_mm) return fail(c, NBX_ERR_INVALID, "nbx_add_dipole: the system has no magnetic moments");
    c->has_dip = true; c->mu_4pi = mu_4pi;
    return NBX_OK;
}

int nbx_add_spcfw(nbx_ctx *c, double rOH, double aHOH, double kb, double ka)
{
    NBX_FANOUT(c, nbx_add_spcfw(x, rOH, aHOH, kb, ka));
    NBX_TRY(guard(c));
    NBX_TRY(need_system(c, "nbx_add_spcfw"));
    if (!c->water) return fail(c, NBX_ERR_INVALID, "nbx_add_spcfw: the system is not water (O,H1,H2 triples)");
    c->has_spcfw = true; c->rOH = rOH; c->aHOH = aHOH; c->k_bond = kb; c->k_angle = ka;
    return NBX_OK;
}

int nbx_clear_potentials(nbx_ctx *c)
{
    NBX_FANOUT(c, nbx_clear_potentials(x));
    NBX_TRY(guard(c));
    c->has_grav = c->has_lj = c->has_coul = c->has_dip = c->has_spcfw = false;
    return NBX_OK;
}

int nbx_thermostat(nbx_ctx *c, int kind, double T0, double param, double kB, int64_t N, int64_t Nc)
{
    NBX_FANOUT(c, nbx_thermostat(x, kind, T0, param, kB, N, Nc));
    NBX_TRY(guard(c));
    if (kind < NBX_THERMO_NONE || kind > NBX_THERMO_LANGEVIN) return fail(c, NBX_ERR_INVALID, "nbx_thermostat: kind %d", kind);
    if (kind != NBX_THERMO_NONE && (3 * N - Nc <= 0 || !(kB > 0.0)))
        return fail(c, NBX_ERR_INVALID, "nbx_thermostat: needs 3N - Nc > 0 and kB > 0");
    graph_drop(c);
    c->thermo = kind; c->T0 = T0; c->tparam = param; c->kB = kB; c->thN = N; c->thNc = Nc;
    if (c->n > 0) c->ncols = c->n + (kind == NBX_THERMO_NOSEHOOVER ? 1 : 0);
    return NBX_OK;
}

int nbx_shard_pairs(nbx_ctx *c, int rank, int nranks)
{
    NBX_TRY(guard(c));
    NBX_TRY(need_system(c, "nbx_shard_pairs"));
    NBX_TRY(no_slab(c, "nbx_shard_pairs"));
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(c, NBX_ERR_INVALID, "nbx_shard_pairs: rank %d of %d", rank, nranks);
    c->pair_rank = rank; c->pair_nranks = nranks;
    return NBX_OK;
}

int nbx_shard(nbx_ctx *c, int64_t lo, int64_t hi)
{
    NBX_TRY(guard(c));
    NBX_TRY(need_system(c, "nbx_shard"));
    NBX_TRY(no_slab(c, "nbx_shard"));
    if (lo < 0 || hi > c->n || lo > hi) return fail(c, NBX_ERR_INVALID, "nbx_shard: [%lld,%lld) outside [0,%lld)", (long long)lo, (long long)hi, (long long)c->n);
    if (c->water && (lo % 3 || hi % 3)) return fail(c, NBX_ERR_INVALID, "nbx_shard: water shards must hold whole molecules");
    c->tgt_lo = lo; c->tgt_hi = hi;
    return NBX_OK;
}

int nbx_slab_init(nbx_ctx *c, int rank, int nranks)
{
    NBX_TRY(guard(c));
    NBX_TRY(need_resident(c, "nbx_slab_init"));
    return slab_init(c, rank, nranks);
}

int nbx_slab_pack(nbx_ctx *c)
{
    NBX_TRY(guard(c));
    NBX_TRY(need_resident(c, "nbx_slab_pack"));
    return slab_pack(c);
}

int nbx_slab_unpack(nbx_ctx *c, int64_t *counts)
{
    NBX_TRY(guard(c));
    NBX_TRY(need_resident(c, "nbx_slab_unpack"));
    return slab_unpack(c, counts);
}

int nbx_slab_prime(nbx_ctx *c)
{
    NBX_TRY(guard(c));
    NBX_TRY(need_resident(c, "nbx_slab_prime"));
    if (!c->slab.on || !c->slab.verlet || c->slab.packed) return fail(c, NBX_ERR_INVALID, "nbx_slab_prime: no complete Verlet-list slab state");
    // evaluate the pair terms into the spare acceleration rows: cells and lists get built, a(t) stays what it is
    c->slab.rebuild_now = true;
    std::swap(c->acc, c->acc_old);
    const int rc = compute_pairs(c);
    std::swap(c->acc, c->acc_old);
    return rc;
}

int nbx_slab_refresh_send(nbx_ctx *c)
{
    NBX_TRY(guard(c));
    NBX_TRY(need_resident(c, "nbx_slab_refresh_send"));
    return slab_refresh_send(c);
}

int nbx_slab_refresh_recv(nbx_ctx *c)
{
    NBX_TRY(guard(c));
    NBX_TRY(need_resident(c, "nbx_slab_refresh_recv"));
    return slab_refresh_recv(c);
}

int nbx_slab_verlet_check(nbx_ctx *c, double soft_fraction, void *out2_dev)
{
    NBX_TRY(guard(c));
    NBX_TRY(need_resident(c, "nbx_slab_verlet_check"));
    return slab_verlet_check(c, soft_fraction, static_cast<int *>(out2_dev), nullptr);
}

int nbx_slab_step_begin(nbx_ctx *c, double dt, double soft_fraction)
{
    NBX_TRY(guard(c));
    NBX_TRY(need_resident(c, "nbx_slab_step_begin"));
    if (!c->slab.on || !c->slab.verlet || c->slab.first) return fail(c, NBX_ERR_INVALID, "nbx_slab_step_begin: no Verlet-list slab state");
    if (c->T_slot != 12) { // the sum of the upload is the global one: it becomes the summed slot as it is
        NBX_CUDA(c, cudaMemcpyAsync(c->d_scal + 12, c->d_scal, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        c->T_slot = 12;
    }
    NBX_TRY(launch_vv_pos(c, dt));
    NBX_CUDA(c, cudaMemsetAsync(c->d_scal + 13, 0, 2 * sizeof(double), c->stream));
    return slab_verlet_check(c, soft_fraction, nullptr, c->d_scal + 13);
}

int nbx_slab_step_end(nbx_ctx *c, double dt, int refresh)
{
    NBX_TRY(guard(c));
    NBX_TRY(need_resident(c, "nbx_slab_step_end"));
    if (!c->slab.on) return fail(c, NBX_ERR_INVALID, "nbx_slab_step_end: call nbx_slab_init first");
    if (refresh) {
        if (!c->slab.direct && c->slab.nranks > 1)
            return fail(c, NBX_ERR_INVALID, "nbx_slab_step_end: without nbx_slab_connect the host has to move the messages");
        NBX_TRY(slab_refresh_send(c));
        NBX_TRY(slab_refresh_recv(c));
    }
    NBX_TRY(nbx_vv_forces(c));
    return nbx_vv_finish(c, dt);
}

int nbx_slab_rx(nbx_ctx *c, void **ptr, int64_t *ndoubles, void *ipc_handle64)
{
    NBX_TRY(guard(c));
    if (!c->slab.on) return fail(c, NBX_ERR_INVALID, "nbx_slab_rx: call nbx_slab_init first");
    if (ptr) *ptr = c->slab.rx;
    if (ndoubles) *ndoubles = c->slab.rx_doubles;
    if (ipc_handle64) {
        cudaIpcMemHandle_t h;
        NBX_CUDA(c, cudaIpcGetMemHandle(&h, c->slab.rx));
        static_assert(sizeof(h) == 64, "CUDA IPC handles are 64 bytes");
        memcpy(ipc_handle64, &h, sizeof h);
    }
    return NBX_OK;
}

int nbx_slab_connect(nbx_ctx *c, const void *left_handle64, const void *right_handle64, void *left_ptr, void *right_ptr)
{
    NBX_TRY(guard(c));
    return slab_connect(c, left_handle64, right_handle64, left_ptr, right_ptr);
}

int nbx_slab_check(nbx_ctx *c, int64_t *counts)
{
    NBX_TRY(guard(c));
    return slab_check(c, counts);
}

int nbx_slab_buffer(nbx_ctx *c, int which, void **ptr, int64_t *ndoubles)
{
    NBX_TRY(guard(c));
    if (!c->slab.on) return fail(c, NBX_ERR_INVALID, "nbx_slab_buffer: call nbx_slab_init first");
    if (which < 0 || which > 3 || !ptr) return fail(c, NBX_ERR_INVALID, "nbx_slab_buffer: which = %d", which);
    if (c->slab.direct) return fail(c, NBX_ERR_INVALID, "nbx_slab_buffer: the exchange is direct (nbx_slab_connect), nothing for the host to move");
    *ptr = c->slab.msg[which];
    if (ndoubles) *ndoubles = c->slab.msg_doubles;
    return NBX_OK;
}

int nbx_slab_download(nbx_ctx *c, int64_t *n_own, int32_t *gid, double *u, double *v, double *dv)
{
    NBX_TRY(guard(c));
    if (!c->slab.on || c->slab.packed) return fail(c, NBX_ERR_INVALID, "nbx_slab_download: no complete slab state");
    NBX_TRY(slab_check(c, nullptr));
    const int64_t m = c->slab.n_own;
    const size_t bytes = sizeof(double) * 3 * (size_t)m;
    if (n_own) *n_own = m;
    if (gid && m) NBX_CUDA(c, cudaMemcpyAsync(gid, c->gid, sizeof(int32_t) * (size_t)m, cudaMemcpyDeviceToHost, c->stream));
    double *const rows[3] = {c->pos, c->vel, c->acc};
    double *const stage[3] = {c->aos_u, c->aos_v, c->aos_dv};
    double *const host[3] = {u, v, dv};
    for (int k = 0; k < 3; ++k) {
        if (!host[k] || m == 0) continue;
        NBX_TRY(launch_soa_to_aos(c, rows[k], stage[k], m, m, 0, m));
        NBX_CUDA(c, cudaMemcpyAsync(host[k], stage[k], bytes, cudaMemcpyDeviceToHost, c->stream));
    }
    NBX_CUDA(c, cudaStreamSynchronize(c->stream));
    return NBX_OK;
}

int nbx_group_init(nbx_ctx *c, int rank, int nranks, int mode)
{
    NBX_TRY(guard(c));
    if (c->is_group) return fail(c, NBX_ERR_INVALID, "nbx_group_init: the handle is a group leader (nbx_create_multi joins its members itself)");
    NBX_TRY(need_system(c, "nbx_group_init"));
    if (mode < 0 || mode > 3) return fail(c, NBX_ERR_INVALID, "nbx_group_init: mode 0 .. 3");
    return group_init(c, rank, nranks, mode);
}

int nbx_group_export(nbx_ctx *c, int kind, void **ptr, void *ipc_handle64)
{
    NBX_TRY(guard(c));
    return group_export(c, kind, ptr, ipc_handle64);
}

int nbx_group_connect(nbx_ctx *c, const void *handles, void *const *ptrs)
{
    NBX_TRY(guard(c));
    return group_connect(c, handles, ptrs);
}

int nbx_group_start(nbx_ctx *c)
{
    NBX_TRY(guard(c));
    NBX_TRY(need_system(c, "nbx_group_start"));
    return group_start(c);
}

int nbx_set_stream(nbx_ctx *c, void *stream)
{
    NBX_FANOUT(c, nbx_set_stream(x, stream));
    NBX_TRY(guard(c));
    cudaStreamSynchronize(c->stream);
    graph_drop(c);
    c->stream = stream ? (cudaStream_t)stream : c->own_stream;
    return NBX_OK;
}

int nbx_synchronize(nbx_ctx *c)
{
    NBX_FANOUT(c, nbx_synchronize(x));
    NBX_TRY(guard(c));
    NBX_CUDA(c, cudaStreamSynchronize(c->stream));
    return NBX_OK;
}

static int finish_and_check(nbx_ctx *c)
{
    int flag[2] = {0, 0};
    NBX_CUDA(c, cudaMemcpyAsync(flag, c->d_scal + 15, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    NBX_CUDA(c, cudaStreamSynchronize(c->stream));
    if (flag[0]) return fail(c, NBX_ERR_NONFINITE, "non-finite coordinate in u (the reference's wrap loop would not terminate)");
    return NBX_OK;
}

int nbx_accel(nbx_ctx *c, const double *u, double *v, double t, double *dv)
{
    (void)t;
    if (c && c->is_group) return leader_accel(c, u, v, dv);
    NBX_TRY(guard(c));
    NBX_TRY(need_system(c, "nbx_accel"));
    if (c->comm.on) { // group member: upload the own block, all-gather over NVLink, return the own columns of dv
        if (!u || !dv) return fail(c, NBX_ERR_INVALID, "nbx_accel: u and dv are required");
        maybe_pin(c, dv, sizeof(double) * 3 * (size_t)c->ncols);
        NBX_TRY(multi_accel_enqueue(c, u, v));
        NBX_TRY(multi_accel_exchange(c));
        return multi_accel_finish(c, dv);
    }
    NBX_TRY(no_slab(c, "nbx_accel"));
    if (!u || !dv) return fail(c, NBX_ERR_INVALID, "nbx_accel: u and dv are required");
    const bool have_v = needs_velocity(c);
    if (have_v && !v) return fail(c, NBX_ERR_INVALID, "nbx_accel: this thermostat needs v");
    const size_t bytes = sizeof(double) * 3 * (size_t)c->ncols;
    NBX_CUDA(c, cudaMemcpyAsync(c->aos_u, u, bytes, cudaMemcpyHostToDevice, c->stream));
    if (have_v) NBX_CUDA(c, cudaMemcpyAsync(c->aos_v, v, bytes, cudaMemcpyHostToDevice, c->stream));
    NBX_TRY(accel_from_staging(c, have_v));
    NBX_CUDA(c, cudaMemcpyAsync(dv, c->aos_dv, bytes, cudaMemcpyDeviceToHost, c->stream));
    if (c->thermo == NBX_THERMO_NOSEHOOVER) // v[zeta_ind] = ...  (src/thermostats.jl:126)
        NBX_CUDA(c, cudaMemcpyAsync(v + 3 * c->n, c->d_scal + 2, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    return finish_and_check(c);
}

int nbx_accel_begin(nbx_ctx *c, const double *u)
{
    NBX_TRY(guard(c));
    NBX_TRY(need_system(c, "nbx_accel_begin"));
    NBX_TRY(no_slab(c, "nbx_accel_begin"));
    if (!u) return fail(c, NBX_ERR_INVALID, "nbx_accel_begin: u is required");
    if (c->pair_nranks <= 1) return fail(c, NBX_ERR_INVALID, "nbx_accel_begin: the context is not pair-sharded (use nbx_accel)");
    if (needs_velocity(c)) return fail(c, NBX_ERR_UNSUPPORTED, "nbx_accel_begin: thermostats with an RHS term need nbx_accel");
    const size_t bytes = sizeof(double) * 3 * (size_t)c->ncols;
    NBX_CUDA(c, cudaMemcpyAsync(c->aos_u, u, bytes, cudaMemcpyHostToDevice, c->stream));
    NBX_TRY(launch_aos_to_soa(c, c->aos_u, c->pos, c->n));
    NBX_TRY(check_finite(c, c->pos, c->n));
    NBX_TRY(compute_pairs(c)); // partial sums of all bodies (pair sharding)
    c->resident = false;
    return NBX_OK;
}

int nbx_accel_end(nbx_ctx *c, double *dv)
{
    NBX_TRY(guard(c));
    NBX_TRY(need_system(c, "nbx_accel_end"));
    if (!dv) return fail(c, NBX_ERR_INVALID, "nbx_accel_end: dv is required");
    NBX_TRY(launch_soa_to_aos(c, c->acc, c->aos_dv, c->n, c->ncols, c->tgt_lo, c->tgt_hi));
    const size_t bytes = sizeof(double) * 3 * (size_t)c->ncols;
    NBX_CUDA(c, cudaMemcpyAsync(dv, c->aos_dv, bytes, cudaMemcpyDeviceToHost, c->stream));
    return finish_and_check(c);
}

int nbx_accel_device(nbx_ctx *c, const double *u_dev, double *v_dev, double t, double *dv_dev)
{
    (void)t;
    NBX_TRY(guard(c));
    NBX_TRY(need_system(c, "nbx_accel_device"));
    NBX_TRY(no_slab(c, "nbx_accel_device"));
    if (!u_dev || !dv_dev) return fail(c, NBX_ERR_INVALID, "nbx_accel_device: u and dv are required");
    const bool have_v = needs_velocity(c);
    if (have_v && !v_dev) return fail(c, NBX_ERR_INVALID, "nbx_accel_device: this thermostat needs v");
    const size_t bytes = sizeof(double) * 3 * (size_t)c->ncols;
    NBX_CUDA(c, cudaMemcpyAsync(c->aos_u, u_dev, bytes, cudaMemcpyDeviceToDevice, c->stream));
    if (have_v) NBX_CUDA(c, cudaMemcpyAsync(c->aos_v, v_dev, bytes, cudaMemcpyDeviceToDevice, c->stream));
    NBX_TRY(accel_from_staging(c, have_v));
    NBX_CUDA(c, cudaMemcpyAsync(dv_dev, c->aos_dv, bytes, cudaMemcpyDeviceToDevice, c->stream));
    if (c->thermo == NBX_THERMO_NOSEHOOVER)
        NBX_CUDA(c, cudaMemcpyAsync(v_dev + 3 * c->n, c->d_scal + 2, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    return NBX_OK; // asynchronous on ctx's stream; nbx_synchronize() to wait
}

int nbx_upload(nbx_ctx *c, const double *u, const double *v)
{
    if (c && c->is_group) return leader_upload(c, u, v);
    NBX_TRY(guard(c));
    NBX_TRY(need_system(c, "nbx_upload"));
    NBX_TRY(no_slab(c, "nbx_upload"));
    if (c->comm.on) return fail(c, NBX_ERR_INVALID, "nbx_upload: the context belongs to a group (nbx_system starts over)");
    graph_drop(c);
    if (!u || !v) return fail(c, NBX_ERR_INVALID, "nbx_upload: u and v are required");
    const size_t bytes = sizeof(double) * 3 * (size_t)c->ncols;
    NBX_CUDA(c, cudaMemcpyAsync(c->aos_u, u, bytes, cudaMemcpyHostToDevice, c->stream));
    NBX_CUDA(c, cudaMemcpyAsync(c->aos_v, v, bytes, cudaMemcpyHostToDevice, c->stream));
    NBX_TRY(launch_aos_to_soa(c, c->aos_u, c->pos, c->n));
    NBX_TRY(launch_aos_to_soa(c, c->aos_v, c->vel, c->n));
    NBX_TRY(check_finite(c, c->pos, c->n));
    NBX_TRY(launch_sum_mv2(c, c->vel, 0, c->n)); // every rank holds all velocities at upload time
    if (c->thermo == NBX_THERMO_NOSEHOOVER) {
        NBX_CUDA(c, cudaMemcpyAsync(c->d_scal + 1, c->aos_u + 3 * c->n, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        NBX_CUDA(c, cudaMemcpyAsync(c->d_scal + 2, c->aos_v + 3 * c->n, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    }
    c->rng_step = 0;
    {   // a(0): every rank holds the whole state at upload time, so evaluate it unsharded
        const int pr = c->pair_rank, pn = c->pair_nranks;
        c->pair_rank = 0; c->pair_nranks = 1;
        const int rc = compute_accel(c);
        c->pair_rank = pr; c->pair_nranks = pn;
        if (rc != NBX_OK) return rc;
    }
    c->forces_done = false;
    c->resident = true;
    return finish_and_check(c);
}

int nbx_eval_resident(nbx_ctx *c)
{
    NBX_TRY(guard(c));
    NBX_TRY(need_resident(c, "nbx_eval_resident"));
    NBX_TRY(no_slab(c, "nbx_eval_resident"));
    return compute_accel(c);
}

int nbx_vv_begin(nbx_ctx *c, double dt)
{
    NBX_TRY(guard(c));
    NBX_TRY(need_resident(c, "nbx_vv_begin"));
    return launch_vv_pos(c, dt);
}

int nbx_vv_forces(nbx_ctx *c)
{
    NBX_TRY(guard(c));
    NBX_TRY(need_resident(c, "nbx_vv_forces"));
    double *t = c->acc_old; c->acc_old = c->acc; c->acc = t;
    NBX_TRY(compute_pairs(c)); // a(t+dt) from x(t+dt): full for the own targets, or partial for all (pair sharding)
    c->forces_done = true;
    return NBX_OK;
}

int nbx_vv_finish(nbx_ctx *c, double dt)
{
    NBX_TRY(guard(c));
    NBX_TRY(need_resident(c, "nbx_vv_finish"));
    if (!c->forces_done) {
        if (c->pair_nranks > 1)
            return fail(c, NBX_ERR_INVALID, "nbx_vv_finish: pair-sharded contexts need nbx_vv_forces + a cross-rank sum first");
        double *t = c->acc_old; c->acc_old = c->acc; c->acc = t;
        NBX_TRY(compute_pairs(c));
    }
    c->forces_done = false;
    NBX_TRY(launch_vv_vel(c, dt, true)); // RHS thermostats use v(t) and the temperature of v(t)
    if (c->thermo == NBX_THERMO_ANDERSEN) NBX_TRY(launch_andersen(c, dt));
    return NBX_OK;
}

// A captured two-step graph that outlives one call: nbx_run_vv keeps it across its chunks (nothing but nbx_download
// happens in between, so every buffer and every host-side decision baked into the capture still holds).
struct StepGraphKeep {
    cudaGraphExec_t exec = nullptr;
    double dt = 0.0;
    double *acc0 = nullptr; // the acc / acc_old roles the capture started from (they swap every step)
};

static int step_vv_impl(nbx_ctx *c, double dt, int64_t nsteps, StepGraphKeep *keep)
{
    if (c && c->is_group) return leader_step_vv(c, dt, nsteps);
    NBX_TRY(guard(c));
    NBX_TRY(need_resident(c, "nbx_step_vv"));
    if (c->comm.on) { // group member (pairs / targets / slabs): the distributed loop, exchanges over peer memory
        NBX_TRY(multi_enqueue_vv(c, dt, nsteps));
        return multi_finish(c);
    }
    NBX_TRY(no_slab(c, "nbx_step_vv"));
    if (c->tgt_lo != 0 || c->tgt_hi != c->n)
        return fail(c, NBX_ERR_INVALID, "nbx_step_vv: sharded context; drive nbx_vv_begin / all-gather / nbx_vv_finish");
    if (c->pair_nranks > 1)
        return fail(c, NBX_ERR_INVALID, "nbx_step_vv: pair-sharded context; drive nbx_vv_forces / reduce / nbx_vv_finish");
    if (c->thermo == NBX_THERMO_LANGEVIN)
        return fail(c, NBX_ERR_UNSUPPORTED, "nbx_step_vv: the Langevin thermostat is an SDE (use nbx_step_em), as in run_simulation");
    int64_t s = 0;
    // one cutoff potential over Verlet lists: the position update also checks the displacements and refreshes the
    // cell-order records (vv_pos_lists_kernel); decided per step, the first evaluation after a (re)configuration is plain
    const bool one_cutoff = !c->water && !c->has_grav && !c->has_dip && !c->has_spcfw && c->tgt_lo == 0 && c->tgt_hi == c->n &&
                            (c->has_lj != (c->has_coul && std::isfinite(c->el_R)));
    CellList *ucl = c->has_lj ? &c->cl_lj : &c->cl_el;
    const double *uw = c->has_lj ? nullptr : c->charge;
    auto one_step = [&]() -> int {
        if (one_cutoff && lists_can_fuse_update(c, ucl, c->pos)) NBX_TRY(launch_vv_pos_lists(c, ucl, uw, dt));
        else NBX_TRY(launch_vv_pos(c, dt));
        double *t = c->acc_old; c->acc_old = c->acc; c->acc = t;
        NBX_TRY(compute_pairs(c));
        NBX_TRY(launch_vv_vel(c, dt, true));
        if (c->thermo == NBX_THERMO_ANDERSEN) NBX_TRY(launch_andersen(c, dt));
        return NBX_OK;
    };
    // Long runs replay a CUDA graph of TWO steps (the acc / acc_old swap has period two): every decision inside a
    // step (Verlet rebuild, overflow fallback) is taken on the device, so the launch sequence is the same for all
    // steps.  Andersen draws from a host-side step counter and the phase timers record events: both stay eager.
    const bool graph_ok = c->opt_graph && !c->timing && c->thermo != NBX_THERMO_ANDERSEN && c->stream != nullptr &&
                          c->stream != cudaStreamLegacy && c->stream != cudaStreamPerThread;
    // (a graph that will be kept pays for its capture over all chunks of the run: short chunks are worth it too)
    const bool graphable = graph_ok && nsteps - s >= (keep ? 6 : 32);
    if (graph_ok && keep && keep->exec && keep->dt == dt) {
        // a kept graph: one eager step first if the acc / acc_old roles are the other way round (odd chunk lengths)
        if (c->acc != keep->acc0) { NBX_TRY(one_step()); ++s; }
        cudaError_t e = cudaSuccess;
        if (c->acc == keep->acc0)
            for (; s + 2 <= nsteps; s += 2) {
                e = cudaGraphLaunch(keep->exec, c->stream);
                if (e != cudaSuccess) break;
            }
        if (e != cudaSuccess) return cuda_fail(c, e, "CUDA graph of the velocity-Verlet step (kept)");
    } else if (graphable) {
        for (int w = 0; w < 2; ++w, ++s) NBX_TRY(one_step()); // warm-up: allocations and attribute calls happen outside the capture
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        cudaError_t e = cudaSuccess;
        double *const acc_at_capture = c->acc;
        if (c->opt_cond_nodes && !c->cond_fail && !c->aux_stream &&
            cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking) != cudaSuccess) { c->aux_stream = nullptr; cudaGetLastError(); }
        for (int attempt = 0; attempt < 2; ++attempt) { // second attempt: plain capture, should the IF nodes be refused
            cudaStream_t main_stream = c->stream;
            c->cond_capture = c->opt_cond_nodes && !c->cond_fail && c->aux_stream != nullptr;
            const bool with_nodes = c->cond_capture;
            double *acc0 = c->acc, *acc_old0 = c->acc_old;
            e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
            if (e != cudaSuccess) { c->cond_capture = false; break; }
            int rc = one_step();
            if (rc == NBX_OK) rc = one_step();
            c->cond_capture = false;
            c->stream = main_stream;
            e = cudaStreamEndCapture(c->stream, &graph);
            if (rc == NBX_OK && e == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
            if (rc == NBX_OK && e == cudaSuccess) break;
            if (graph) { cudaGraphDestroy(graph); graph = nullptr; }
            exec = nullptr;
            cudaGetLastError();
            c->acc = acc0; c->acc_old = acc_old0; c->forces_done = false; // nothing of the failed capture ran
            if (!with_nodes) { if (rc != NBX_OK) return rc; break; }
            c->cond_fail = true; // try again without
            e = cudaSuccess;
        }
        if (e == cudaSuccess && exec) {
            for (; s + 2 <= nsteps; s += 2) {
                e = cudaGraphLaunch(exec, c->stream);
                if (e != cudaSuccess) break;
            }
        }
        if (exec && keep && e == cudaSuccess) {
            if (keep->exec) cudaGraphExecDestroy(keep->exec);
            keep->exec = exec; keep->dt = dt; keep->acc0 = acc_at_capture;
            exec = nullptr;
        }
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        if (e != cudaSuccess) return cuda_fail(c, e, "CUDA graph of the velocity-Verlet step");
    }
    for (; s < nsteps; ++s) NBX_TRY(one_step());
    NBX_CUDA(c, cudaStreamSynchronize(c->stream));
    return NBX_OK;
}

int nbx_step_vv(nbx_ctx *c, double dt, int64_t nsteps) { return step_vv_impl(c, dt, nsteps, nullptr); }

// run_simulation with saveat (src/nbody_simulation_result.jl:468-487): nsteps velocity-Verlet steps on the device, the state
// copied out every save_every steps (and after the last one) as consecutive 3 x ncols frames -- what the Julia shim wraps
// into the SciMLBase solution behind SimulationResult (:5-8, :49-104).
int nbx_run_vv(nbx_ctx *c, double dt, int64_t nsteps, int64_t save_every, double *u_frames, double *v_frames, int64_t max_frames,
               int64_t *nframes)
{
    if (!c) return NBX_ERR_INVALID;
    if (nframes) *nframes = 0;
    if (nsteps < 0 || save_every < 1) return fail(c, NBX_ERR_INVALID, "nbx_run_vv: nsteps >= 0 and save_every >= 1");
    const int64_t ncols = c->is_group ? c->ncols : c->ncols;
    const int64_t need = (nsteps + save_every - 1) / save_every;
    if ((u_frames || v_frames) && max_frames < need)
        return fail(c, NBX_ERR_CAPACITY, "nbx_run_vv: %lld frames needed, capacity %lld", (long long)need, (long long)max_frames);
    int64_t done = 0, k = 0;
    StepGraphKeep keep; // (single contexts; groups cache their graph themselves)
    int rc = NBX_OK;
    while (done < nsteps && rc == NBX_OK) {
        const int64_t chunk = std::min<int64_t>(save_every, nsteps - done);
        rc = step_vv_impl(c, dt, chunk, &keep);
        done += chunk;
        if (rc == NBX_OK && (u_frames || v_frames)) {
            const size_t off = (size_t)k * 3 * (size_t)ncols;
            rc = nbx_download(c, u_frames ? u_frames + off : nullptr, v_frames ? v_frames + off : nullptr, nullptr);
        }
        ++k;
    }
    if (keep.exec) cudaGraphExecDestroy(keep.exec);
    if (rc != NBX_OK) return rc;
    if (nframes) *nframes = k;
    return NBX_OK;
}

int nbx_set_seed(nbx_ctx *c, uint64_t seed)
{
    NBX_FANOUT(c, nbx_set_seed(x, seed));
    NBX_TRY(guard(c));
    c->seed = seed; c->rng_step = 0;
    return NBX_OK;
}

int nbx_step_em(nbx_ctx *c, double dt, int64_t nsteps, uint64_t seed)
{
    if (c && c->is_group) return leader_step_em(c, dt, nsteps, seed);
    NBX_TRY(guard(c));
    NBX_TRY(need_resident(c, "nbx_step_em"));
    NBX_TRY(no_slab(c, "nbx_step_em"));
    if (c->thermo != NBX_THERMO_LANGEVIN) return fail(c, NBX_ERR_INVALID, "nbx_step_em: needs the Langevin thermostat");
    if (seed) c->seed = seed;
    if (c->comm.on) {
        NBX_TRY(multi_enqueue_em(c, dt, nsteps));
        return multi_finish(c);
    }
    for (int64_t s = 0; s < nsteps; ++s) {
        if (s > 0) NBX_TRY(compute_accel(c)); // a(x_s); the first one is resident already
        NBX_TRY(launch_em_step(c, dt));
    }
    NBX_TRY(compute_accel(c)); // leave a(x_end) resident
    NBX_CUDA(c, cudaStreamSynchronize(c->stream));
    return NBX_OK;
}

int nbx_download(nbx_ctx *c, double *u, double *v, double *dv)
{
    if (c && c->is_group) return leader_download(c, u, v, dv);
    NBX_TRY(guard(c));
    NBX_TRY(need_resident(c, "nbx_download"));
    NBX_TRY(no_slab(c, "nbx_download"));
    if (c->comm.on) { // group member: all positions are here, velocities and accelerations of the own block only
        const int64_t lo = c->tgt_lo, cnt = c->tgt_hi - lo;
        if (u) {
            NBX_TRY(launch_soa_to_aos(c, c->pos, c->aos_u, c->n, c->n, 0, c->n));
            NBX_CUDA(c, cudaMemcpyAsync(u, c->aos_u, sizeof(double) * 3 * (size_t)c->n, cudaMemcpyDeviceToHost, c->stream));
        }
        double *const rows[2] = {c->vel, c->acc}, *const stage[2] = {c->aos_v, c->aos_dv}, *const host[2] = {v, dv};
        for (int k = 0; k < 2; ++k) {
            if (!host[k] || cnt <= 0) continue;
            NBX_TRY(launch_soa_to_aos(c, rows[k] + lo, stage[k] + 3 * lo, cnt, cnt, 0, cnt));
            NBX_CUDA(c, cudaMemcpyAsync(host[k] + 3 * lo, stage[k] + 3 * lo, sizeof(double) * 3 * (size_t)cnt, cudaMemcpyDeviceToHost, c->stream));
        }
        NBX_CUDA(c, cudaStreamSynchronize(c->stream));
        return NBX_OK;
    }
    const size_t bytes = sizeof(double) * 3 * (size_t)c->ncols;
    const bool nose = c->thermo == NBX_THERMO_NOSEHOOVER;
    if (u) {
        NBX_TRY(launch_soa_to_aos(c, c->pos, c->aos_u, c->n, c->ncols, 0, c->n));
        if (nose) NBX_CUDA(c, cudaMemcpyAsync(c->aos_u + 3 * c->n, c->d_scal + 1, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        NBX_CUDA(c, cudaMemcpyAsync(u, c->aos_u, bytes, cudaMemcpyDeviceToHost, c->stream));
    }
    if (v) {
        NBX_TRY(launch_soa_to_aos(c, c->vel, c->aos_v, c->n, c->ncols, 0, c->n));
        if (nose) NBX_CUDA(c, cudaMemcpyAsync(c->aos_v + 3 * c->n, c->d_scal + 2, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        NBX_CUDA(c, cudaMemcpyAsync(v, c->aos_v, bytes, cudaMemcpyDeviceToHost, c->stream));
    }
    if (dv) {
        NBX_TRY(launch_soa_to_aos(c, c->acc, c->aos_dv, c->n, c->ncols, c->tgt_lo, c->tgt_hi));
        NBX_CUDA(c, cudaMemcpyAsync(dv, c->aos_dv, bytes, cudaMemcpyDeviceToHost, c->stream));
    }
    NBX_CUDA(c, cudaStreamSynchronize(c->stream));
    return NBX_OK;
}

int nbx_energy(nbx_ctx *c, double *ekin, double *epot, double *temperature)
{
    if (c && c->is_group) return leader_energy(c, ekin, epot, temperature);
    NBX_TRY(guard(c));
    NBX_TRY(need_resident(c, "nbx_energy"));
    if (c->slab.on && c->slab.started && !epot) {
        // a started slab keeps the sum m v^2 over ALL ranks in the scalar block after nbx_group_start / nbx_step_vv
        double mv2 = 0.0;
        NBX_CUDA(c, cudaMemcpyAsync(&mv2, c->d_scal, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        NBX_CUDA(c, cudaStreamSynchronize(c->stream));
        if (ekin) *ekin = 0.5 * mv2;
        if (temperature) {
            const int64_t N = c->thN > 0 ? c->thN : c->slab.n_total;
            *temperature = mv2 / ((c->kB != 0.0 ? c->kB : 1.0) * (double)(3 * N - c->thNc));
        }
        return NBX_OK;
    }
    NBX_TRY(no_slab(c, "nbx_energy"));
    if (ekin || temperature) NBX_TRY(reduce_kinetic(c, ekin, temperature));
    if (epot) NBX_TRY(reduce_potential(c, epot));
    return NBX_OK;
}

int nbx_neighbors(nbx_ctx *c, int64_t *offsets, int32_t *list, int64_t cap)
{
    if (c && c->is_group) {
        if (!c->g_ready || c->members[0]->comm.mode == 3)
            return fail(c, NBX_ERR_UNSUPPORTED, "nbx_neighbors: needs a group whose members hold all positions (modes 1, 2) with a resident state");
        const int rc = nbx_neighbors(c->members[0], offsets, list, cap);
        if (rc != NBX_OK) c->err = c->members[0]->err;
        return rc;
    }
    NBX_TRY(guard(c));
    NBX_TRY(need_resident(c, "nbx_neighbors"));
    NBX_TRY(no_slab(c, "nbx_neighbors"));
    if (!c->has_lj) return fail(c, NBX_ERR_INVALID, "nbx_neighbors: no Lennard-Jones potential (the cutoff predicate) configured");
    if (!offsets || (!list && cap > 0)) return fail(c, NBX_ERR_INVALID, "nbx_neighbors: NULL output");
    if (c->water) {
        const int nmol = (int)(c->n / 3);
        gather_oxygen_kernel<<<(unsigned)((c->opad + 255) / 256), 256, 0, c->stream>>>(c->pos, c->npad, nmol, c->opos, c->opad, kFarAway);
        NBX_TRY(cells_plan(c, c->lj_R, nmol, &c->cl_lj.grid));
        return cells_neighbors(c, &c->cl_lj, c->opos, nmol, c->opad, c->lj_R2, offsets, list, cap);
    }
    NBX_TRY(cells_plan(c, c->lj_R, c->n, &c->cl_lj.grid));
    return cells_neighbors(c, &c->cl_lj, c->pos, c->n, c->npad, c->lj_R2, offsets, list, cap);
}

int nbx_device_ptr(nbx_ctx *c, int which, void **ptr, int64_t *ld)
{
    NBX_TRY(guard(c));
    NBX_TRY(need_system(c, "nbx_device_ptr"));
    if (!ptr) return fail(c, NBX_ERR_INVALID, "nbx_device_ptr: ptr is NULL");
    switch (which) {
    case 0: *ptr = c->pos; break;
    case 1: *ptr = c->vel; break;
    case 2: *ptr = c->acc; break; // NOTE: acc and acc_old swap every step; query after nbx_vv_forces
    case 3: *ptr = c->d_scal; break; // [0] = sum m v^2 of the shard after a step (all-reduce it across ranks)
    default: return fail(c, NBX_ERR_INVALID, "nbx_device_ptr: which = %d", which);
    }
    if (ld) *ld = which == 3 ? 16 : c->npad;
    return NBX_OK;
}

int nbx_timing_enable(nbx_ctx *c, int enable)
{
    NBX_FANOUT(c, nbx_timing_enable(x, enable));
    NBX_TRY(guard(c));
    c->timing = enable != 0;
    return NBX_OK;
}

int nbx_timing_get(nbx_ctx *c, int phase, double *total_ms, int64_t *count)
{
    if (c && c->is_group) { // the slowest member
        double best = 0.0; int64_t cnt = 0;
        for (nbx_ctx *x : c->members) {
            double t = 0.0; int64_t k = 0;
            const int rc = nbx_timing_get(x, phase, &t, &k);
            if (rc != NBX_OK) { c->err = x->err; return rc; }
            if (t >= best) { best = t; cnt = k; }
        }
        if (total_ms) *total_ms = best;
        if (count) *count = cnt;
        return NBX_OK;
    }
    NBX_TRY(guard(c));
    if (phase < 0 || phase >= NBX_T_COUNT) return fail(c, NBX_ERR_INVALID, "nbx_timing_get: phase %d", phase);
    timer_flush(c, c->timers[phase]);
    if (total_ms) *total_ms = c->timers[phase].total_ms;
    if (count) *count = c->timers[phase].count;
    return NBX_OK;
}

int nbx_timing_reset(nbx_ctx *c)
{
    NBX_FANOUT(c, nbx_timing_reset(x));
    NBX_TRY(guard(c));
    for (auto &t : c->timers) { timer_flush(c, t); t.total_ms = 0.0; t.count = 0; }
    return NBX_OK;
}

int nbx_set_option(nbx_ctx *c, const char *key, int64_t value)
{
    if (c && c->is_group) {
        if (key && !strcmp(key, "group_mode")) {
            if (value < 0 || value > 3) return fail(c, NBX_ERR_INVALID, "group_mode: 0 (chosen from the potentials), 1 pairs, 2 targets, 3 slabs");
            c->opt_group_mode = (int)value;
            return NBX_OK;
        }
        NBX_FANOUT(c, nbx_set_option(x, key, value));
    }
    NBX_TRY(guard(c));
    if (!key) return fail(c, NBX_ERR_INVALID, "nbx_set_option: key is NULL");
    graph_drop(c);
    if (!strcmp(key, "cell_list")) c->opt_cell_list = (int)value;
    else if (!strcmp(key, "prefilter")) c->opt_prefilter = (int)value;
    else if (!strcmp(key, "verlet_skin_permille")) {
        if (value < 0 || value > 1000) return fail(c, NBX_ERR_INVALID, "verlet_skin_permille: 0 .. 1000 (thousandths of the cutoff)");
        c->opt_verlet_permille = (int)value;
    }
    else if (!strcmp(key, "graph")) c->opt_graph = (int)value;
    else if (!strcmp(key, "pin_host")) c->opt_pin_host = (int)value;
    else if (!strcmp(key, "spin_timeout_ms")) {
        if (value < 1 || value > 600000) return fail(c, NBX_ERR_INVALID, "spin_timeout_ms: 1 .. 600000");
        c->spin_timeout_ms = value;
    }
    else if (!strcmp(key, "fuse_update")) c->opt_fuse_update = (int)value;
    else if (!strcmp(key, "graph_if_nodes")) { c->opt_cond_nodes = (int)value; if (value) c->cond_fail = false; }
    else if (!strcmp(key, "temperature_slot")) {
        if (value != 0 && value != 12) return fail(c, NBX_ERR_INVALID, "temperature_slot: 0 or 12");
        c->T_slot = (int)value;
    }
    else if (!strcmp(key, "slab_rebuild")) c->slab.rebuild_now = value != 0;
    else if (!strcmp(key, "slab_record_halo")) c->slab.record_halo = value != 0;
    else if (!strcmp(key, "verlet_lanes")) {
        if (value != 0 && value != 1 && value != 2 && value != 4 && value != 8) return fail(c, NBX_ERR_INVALID, "verlet_lanes: 0, 1, 2, 4 or 8");
        c->opt_verlet_lanes = (int)value;
    }
    else if (!strcmp(key, "verlet_branchfree")) c->opt_verlet_branchfree = (int)value;
    else if (!strcmp(key, "verlet_banked")) c->opt_verlet_banked = (int)value;
    else if (!strcmp(key, "symmetric_pairs")) c->opt_sym = (int)value;
    else if (!strcmp(key, "symmetric_min_n")) c->sym_min_n = value;
    else if (!strcmp(key, "sym_variant")) c->opt_sym_variant = (int)value;
    else if (!strcmp(key, "sym_seg_len")) c->opt_sym_seg_len = (int)value;
    else if (!strcmp(key, "uniform_weights")) { if (!value) c->mass_uniform = c->charge_uniform = false; }
    else return fail(c, NBX_ERR_INVALID, "nbx_set_option: unknown key '%s'", key);
    return NBX_OK;
}

int nbx_get_info(nbx_ctx *c, const char *key, int64_t *value)
{
    if (c && c->is_group) { // the members are alike: answer for member `group_member` (default 0); "group_size" for the leader
        if (!key || !value) return fail(c, NBX_ERR_INVALID, "nbx_get_info: NULL argument");
        if (!strcmp(key, "group_size")) { *value = (int64_t)c->members.size(); return NBX_OK; }
        if (!strcmp(key, "n")) { *value = c->n; return NBX_OK; }
        if (!strcmp(key, "ncols")) { *value = c->ncols; return NBX_OK; }
        if (!strcmp(key, "verlet_rebuilds") || !strcmp(key, "slab_own") || !strcmp(key, "slab_ghost")) { // max / sum over the members
            int64_t acc = 0;
            for (nbx_ctx *x : c->members) {
                int64_t t = 0;
                const int rc = nbx_get_info(x, key, &t);
                if (rc != NBX_OK) { c->err = x->err; return rc; }
                acc = !strcmp(key, "verlet_rebuilds") ? std::max(acc, t) : acc + t;
            }
            *value = acc;
            return NBX_OK;
        }
        nbx_ctx *x = c->members[0];
        const int rc = nbx_get_info(x, key, value);
        if (rc != NBX_OK) c->err = x->err;
        return rc;
    }
    NBX_TRY(guard(c));
    if (!key || !value) return fail(c, NBX_ERR_INVALID, "nbx_get_info: NULL argument");
    if (!strcmp(key, "n")) *value = c->n;
    else if (!strcmp(key, "npad")) *value = c->npad;
    else if (!strcmp(key, "ncols")) *value = c->ncols;
    else if (!strcmp(key, "water")) *value = c->water;
    else if (!strcmp(key, "sm_count")) *value = c->sm_count;
    else if (!strcmp(key, "slab_own")) *value = c->slab.on ? c->slab.n_own : c->n;
    else if (!strcmp(key, "slab_ghost")) *value = c->slab.on ? c->slab.n_ghost : 0;
    else if (!strcmp(key, "slab_layer_lo")) *value = c->slab.c0;
    else if (!strcmp(key, "slab_layer_hi")) *value = c->slab.c1;
    else if (!strcmp(key, "slab_layers")) *value = c->slab.nc;
    else if (!strcmp(key, "slab_verlet")) *value = (c->slab.on && c->slab.verlet) ? 1 : 0;
    else if (!strcmp(key, "thermostat")) *value = c->thermo;
    else if (!strcmp(key, "group_mode")) *value = c->comm.on ? c->comm.mode : 0;
    else if (!strcmp(key, "group_rank")) *value = c->comm.on ? c->comm.rank : 0;
    else if (!strcmp(key, "group_size")) *value = c->comm.on ? c->comm.nranks : 1;
    else if (!strcmp(key, "shard_lo")) *value = c->tgt_lo;
    else if (!strcmp(key, "shard_hi")) *value = c->tgt_hi;
    else if (!strcmp(key, "graph_cached")) *value = c->mg_exec ? 1 : 0;
    else if (!strcmp(key, "verlet_overflow") || !strcmp(key, "verlet_rebuilds")) {
        // device flags of the LJ list (synchronises): [1] sticky overflow, [2] rebuilds so far
        int h[4] = {0, 0, 0, 0};
        if (c->cl_lj.v_valid && c->cl_lj.v_flags) {
            NBX_CUDA(c, cudaMemcpyAsync(h, c->cl_lj.v_flags, sizeof h, cudaMemcpyDeviceToHost, c->stream));
            NBX_CUDA(c, cudaStreamSynchronize(c->stream));
        }
        *value = !strcmp(key, "verlet_overflow") ? h[1] : h[2];
    }
    else if (!strcmp(key, "graph_if_nodes")) *value = (c->opt_cond_nodes && !c->cond_fail) ? 1 : 0;
    else if (!strcmp(key, "verlet_lj")) *value = c->cl_lj.v_valid ? c->cl_lj.v_cap : 0;
    else if (!strcmp(key, "verlet_el")) *value = c->cl_el.v_valid ? c->cl_el.v_cap : 0;
    else if (!strcmp(key, "cells_lj")) *value = c->cl_lj.grid.valid ? c->cl_lj.grid.ncell : 0;
    else if (!strcmp(key, "cells_el")) *value = c->cl_el.grid.valid ? c->cl_el.grid.ncell : 0;
    else if (!strcmp(key, "allpairs_grid")) *value = c->last_grid;
    else if (!strcmp(key, "allpairs_chunks")) *value = c->last_nchunk;
    else return fail(c, NBX_ERR_INVALID, "nbx_get_info: unknown key '%s'", key);
    return NBX_OK;
}

int nbx_measure_fp64_peak(nbx_ctx *c, double *tflops, double *sm_mhz_effective)
{
    NBX_TRY(guard(c));
    return measure_fp64_peak(c, tflops, sm_mhz_effective);
}

int nbx_measure_hbm_peak(nbx_ctx *c, double *gbs)
{
    NBX_TRY(guard(c));
    return measure_hbm_peak(c, gbs);
}

int nbx_rdf_reset(nbx_ctx *c, int maxbin)
{
    NBX_TRY(guard(c));
    NBX_TRY(need_system(c, "nbx_rdf_reset"));
    return analysis_rdf_reset(c, maxbin);
}

int nbx_rdf_add(nbx_ctx *c, const double *u)
{
    NBX_TRY(guard(c));
    NBX_TRY(need_system(c, "nbx_rdf_add"));
    return analysis_rdf_add(c, u);
}

int nbx_rdf_get(nbx_ctx *c, int64_t *hist, int64_t cap, int64_t *frames)
{
    NBX_TRY(guard(c));
    return analysis_rdf_get(c, hist, cap, frames);
}

int nbx_msd(nbx_ctx *c, const double *u0, const double *u, double *out)
{
    NBX_TRY(guard(c));
    NBX_TRY(need_system(c, "nbx_msd"));
    return analysis_msd(c, u0, u, out);
}

} // extern "C"
