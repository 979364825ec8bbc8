// nbx_internal.cuh -- context, error plumbing and sm_100a PTX helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <utility>
#include <cmath>
#include <vector>

#include "../../include/nbody_b200.h"

namespace nbx {

constexpr int kPad = 1024;           // SoA rows are padded to a multiple of this many doubles
// Padding sources sit here: r^2 = 3e300 is finite, but r^-3 underflows to exactly 0, so a padding source
// contributes exactly nothing to the unbounded kernels whatever its weight.
constexpr double kFarAway = 1.0e150;
constexpr int kMaxTimers = 8192;     // event pairs kept per phase between resets

struct Timer {
    std::vector<cudaEvent_t> ev; // start/stop interleaved
    size_t used = 0;             // events used (2 per launch)
    double total_ms = 0.0;
    int64_t count = 0;
};

struct CellGrid {
    bool valid = false;      // cell list usable for the current box / cutoff
    int nc[3] = {0, 0, 0};   // cells per dimension
    int64_t ncell = 0;
    double lo[3] = {0, 0, 0};
    double len[3] = {0, 0, 0}; // box edge per dimension
};

// One cell list (grid + the particle data in cell order).  A context keeps two: the LJ list (all
// atoms, or the oxygen sub-system of water) and the Coulomb-cutoff list (all atoms).
struct CellList {
    CellGrid grid;
    int64_t n = 0;              // particles binned by the last build
    int *cell_of = nullptr;     // [n] cell id per particle
    int *count = nullptr;       // [ncell+1] histogram
    int *start = nullptr;       // [ncell+1] exclusive scan of count
    int *arrival = nullptr;     // [n] arrival rank of a particle inside its cell (histogram atomics)
    int *tmp_idx = nullptr;     // [n] particle index per slot before the in-cell ordering
    int *tmp_key = nullptr;     // [n] ordering key (global id) per slot, only when the context has ids
    int *sums = nullptr;        // scan block totals
    int *sorted_idx = nullptr;  // [n] particle index of the k-th slot in cell order
    int *slot_of = nullptr;     // [n] inverse: slot of a particle
    bool prechecked = false;    // the position update already did this evaluation's displacement check + record refresh
    int *scell = nullptr;       // [n] cell id of the k-th slot
    double4 *sp4 = nullptr;     // [n] cell order: x, y, z unwrapped (as the reference's predicate uses them), w = charge
    float4 *sl4 = nullptr;      // [n] cell order: wrapped coordinates in cell units (fp32 prefilter), w = exclusion key bits
    int64_t cap_n = 0, cap_cells = 0;
    // Verlet list kept across evaluations (nbx_cells.cu): slots listed per slot, build-time positions, flags
    int *v_list = nullptr, *v_nlist = nullptr, *v_flags = nullptr;
    double *v_ref = nullptr;
    int64_t v_cap_alloc = 0, v_ref_n = 0;
    bool v_valid = false; // the settings below describe the resident list
    int64_t v_n = 0;
    const double *v_px = nullptr;
    double v_R = 0, v_skin = 0, v_L = 0;
    int v_key_div = 0, v_nc = 0, v_cap = 0, v_banked = 0;
};

// slab decomposition state (nbx_slab.cu)
struct SlabState {
    bool on = false, packed = false, first = false; // first: the next pack selects the slab from the full upload
    bool direct = false;                            // neighbours' receive areas are mapped: the pack kernel stores into them
    double *rx = nullptr;                           // receive area: flags + 4 message buffers (2 sides x 2 parities)
    int64_t rx_doubles = 0, ll_off = 0;             // ll_off: start of the LL area (halo refresh of slab_enqueue) inside rx
    double *peer[2] = {nullptr, nullptr};           // receive areas of the left / right neighbour (direct mode)
    std::vector<void *> ipc_opened;
    int rank = 0, nranks = 1;
    int nc = 0, c0 = 0, c1 = 0, wL = 0, wR = 0; // layers of the slab grid; own layers [c0, c1); neighbour widths
    int64_t n_total = 0, n_own = 0, n_ghost = 0; // n_own / n_ghost: as of the last nbx_slab_check
    int64_t cap_loc = 0;                         // bound on own + ghosts (launch sizes); exceeding it is an error
    int *d_n = nullptr;                          // device: [0] own, [1] ghosts, [2] lost particles, [3] capacity errors
    cudaEvent_t ev_counts = nullptr;             // completion of the last asynchronous count read-back
    int64_t capM = 0, capH = 0, msg_doubles = 0; // message capacities (migrants, halo records) and size
    double *msg[4] = {nullptr, nullptr, nullptr, nullptr}; // send-left, send-right, recv-from-left, recv-from-right
    double *pos2 = nullptr, *vel2 = nullptr, *acc2 = nullptr, *mass2 = nullptr, *charge2 = nullptr; // compaction targets
    int *gid_a = nullptr, *gid_b = nullptr;
    int *blockcnt = nullptr, *blockoff = nullptr, *d_counts = nullptr, *h_counts = nullptr;
    // Verlet lists inside the slab (one cutoff potential, skin > 0): local slots stay put between COLLECTIVE
    // rebuilds; in between only the positions of a fixed halo index list travel (slab_refresh_send / _recv)
    bool verlet = false;
    bool rebuild_now = true;   // the host's decision for the next force evaluation (set after a migration round)
    bool record_halo = false;  // the next pack records the halo index lists (second round of a rebuild)
    int *halo_idx[2] = {nullptr, nullptr}; // [capH] local indices of the own boundary-layer particles, message order
    // nbx_step_vv on a slab (slab_run): the rebuild decision is taken on the device (peer-memory all-reduce of the
    // displacement flags), so a step is one fixed launch sequence -- capturable, no host in the loop
    int phase = 0;             // cells_pairs: 0 rebuild (if rebuild_now) + refresh + forces, 1 rebuild chain only, 2 refresh + forces only
    const int *cond = nullptr; // device flag while slab_run enqueues: the rebuild chain runs iff cond[0] != 0, the halo refresh iff == 0
    bool started = false;      // nbx_slab_start distributed the particles and built the first lists
    const CellList *refresh_cl = nullptr; // while slab_one_step enqueues the halo receive: write the ghosts' records too
    bool records_fresh = false; // the cell-order records hold the current positions (written by slab_pos_kernel / the halo receive)
    bool scal0_global = true;  // d_scal[0] holds the sum over ALL ranks (after an upload / at the end of a run), not the local one
};

// Peer-memory communicator (nbx_multi.cu): every rank owns a small WINDOW in device memory that all ranks map (CUDA IPC
// between processes, plain pointers + peer access inside one process).  Kernels post into the peers' windows with
// system-scope stores and spin on their own window: scalar all-reduces, "positions of step k have landed" and "partial
// accelerations of step k are complete" flags.  No collective library and no host in the loop.
constexpr int kMaxRanks = 16;
constexpr int kWinScal = 64;                        // [2 parities][kMaxRanks][8] LL words: three doubles as 6 x {32-bit half, tag}
constexpr int kWinPos = kWinScal + 2 * kMaxRanks * 8;  // [kMaxRanks] int64: positions of sequence number s have landed
constexpr int kWinAcc = kWinPos + kMaxRanks;           // [kMaxRanks] int64: partial accelerations s are staged
constexpr int kWinDoubles = kWinAcc + kMaxRanks + 32;
enum { SEQ_SCAL = 0, SEQ_POS = 1, SEQ_ACC = 2, SEQ_TICKET = 3, SEQ_TIMEOUT = 4, SEQ_TICKET2 = 5, SEQ_GLOBAL0 = 6, SEQ_N = 8 };

struct CommDev { // what the kernels need (passed by value)
    double *win[kMaxRanks];
    int rank, nranks;
    int *seq; // device counters [SEQ_N]; [SEQ_GLOBAL0] != 0: the next all-reduce takes its first value from rank 0 only
    unsigned long long timeout_ns;
};

struct Comm {
    bool on = false;
    int rank = 0, nranks = 1;
    double *win = nullptr;                 // own window [kWinDoubles]
    double *peer_win[kMaxRanks] = {};      // every rank's window (own included)
    double *peer_pos[kMaxRanks] = {};      // every rank's SoA position rows (all-pairs modes: push all-gather)
    double *stage = nullptr;               // [nranks][3][per] partial accelerations pushed by the peers (pair sharding)
    double *peer_stage[kMaxRanks] = {};
    int64_t per = 0, stage_doubles = 0;    // block size of the partition (equal blocks, the last may be short)
    int *d_seq = nullptr;
    std::vector<void *> ipc_opened;
    int mode = 0;                          // 0 none, 1 pair sharding, 2 target blocks, 3 slabs
    bool warm = false;                     // the sharded evaluation has run once: every buffer it needs is allocated
                                           // (an allocation between two members' launches would serialise their streams)
};

} // namespace nbx

struct nbx_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    std::string err;

    // ---- system -------------------------------------------------------------------------
    int64_t n = 0;      // particle columns
    double h_m1 = 0.0;  // mass of the first body (the reference's Andersen/Langevin sigma uses bodies[1].m)
    int64_t ncols = 0;  // columns of u/v/dv at the boundary (n or n+1)
    int64_t npad = 0;   // SoA row stride
    int water = 0;
    bool has_q = false, has_mm = false;
    double *mass = nullptr, *charge = nullptr, *mm = nullptr; // [npad], [npad], [3][npad]
    double *pos = nullptr, *vel = nullptr, *acc = nullptr, *acc_old = nullptr; // [3][npad]
    double *aos_u = nullptr, *aos_v = nullptr, *aos_dv = nullptr;              // [3*(n+1)] staging
    bool resident = false; // pos/vel/acc hold a simulation state

    // ---- boundary -----------------------------------------------------------------------
    int bc_kind = NBX_BC_INFINITE;
    double bc[6] = {0, 0, 0, 0, 0, 0};

    // ---- potentials ---------------------------------------------------------------------
    bool has_grav = false; double G = 0;
    bool has_lj = false; double lj_eps = 0, lj_sigma2 = 0, lj_R = 0, lj_R2 = 0;
    bool has_coul = false; double el_k = 0, el_R = 0, el_R2 = 0;
    bool has_dip = false; double mu_4pi = 0;
    bool has_spcfw = false; double rOH = 0, aHOH = 0, k_bond = 0, k_angle = 0;

    // ---- thermostat ---------------------------------------------------------------------
    int thermo = NBX_THERMO_NONE;
    double T0 = 0, tparam = 0, kB = 0;
    int64_t thN = 0, thNc = 0;
    double *d_scal = nullptr;  // device scalars: [0] sum m v^2, [1] zeta, [2] zeta_dot, [3..15] scratch
    int T_slot = 0;            // d_scal slot the velocity update reads sum m v^2 from (12 while nbx_slab_step_* drive the slab)
    double *d_red = nullptr;   // block partials for reductions
    int64_t red_cap = 0;
    uint64_t seed = 0x9E3779B97F4A7C15ull;
    uint64_t rng_step = 0;

    // ---- sharding -----------------------------------------------------------------------
    int64_t tgt_lo = 0, tgt_hi = 0;
    int *gid = nullptr;        // slab decomposition: global particle id per local column (nullptr: identity)
    nbx::SlabState slab;
    nbx::Comm comm;
    // single-process group (nbx_create_multi): the leader holds no state of its own, calls fan out over the members
    bool is_group = false;
    std::vector<nbx_ctx *> members;
    nbx_ctx *leader = nullptr;
    std::vector<double> g_m, g_q, g_mm;  // leader: the system description (a re-upload rebuilds the members' systems)
    int g_water = 0;
    bool g_ready = false;                // leader: the members are initialised, connected and started
    int opt_group_mode = 0;              // 0: chosen from the potentials, 1 pairs, 2 targets, 3 slabs
    // CUDA graph of two velocity-Verlet steps of the distributed loops (nbx_multi.cu / slab_run), kept across calls
    cudaGraphExec_t mg_exec = nullptr;
    double mg_dt = 0.0;
    double *mg_acc0 = nullptr, *mg_pos0 = nullptr;
    int mg_kind = 0;
    // pinned-host registration cache (option "pin_host"): caller buffers seen by nbx_accel are page-locked once
    int opt_pin_host = 0;
    int64_t spin_timeout_ms = 10000;     // device-side waits for a peer give up after this long (error, never a hang)
    std::vector<std::pair<const void *, size_t>> pinned;
    // slab mode: the particle counts live on the device ([0] own, [1] ghosts) so that a step needs no host
    // round trip; n / tgt_hi then are launch BOUNDS and every kernel clamps to the device counts
    const int *dyn = nullptr;

    // ---- all-pairs scratch ----------------------------------------------------------------
    double *part = nullptr; // [nchunk][3][ntgt_pad] partial sums
    size_t part_bytes = 0;
    std::vector<const void *> attr_done; // kernels whose dynamic-smem attribute is set
    int last_grid = 0, last_nchunk = 0;
    int *sym_ticket = nullptr;     // work-item counter of the symmetric all-pairs kernel (zeroed before every launch)
    int opt_sym_seg_len = 0;       // ring offsets per work item of that kernel (0: chosen from the share)

    // ---- water oxygen sub-system (compact SoA of the O columns) ----------------------------
    double *opos = nullptr, *oacc = nullptr; // [3][opad]
    int64_t opad = 0;

    // ---- cell lists ----------------------------------------------------------------------
    nbx::CellList cl_lj, cl_el;
    int opt_cell_list = 1;
    int opt_prefilter = 1;
    int opt_verlet_permille = 80;  // Verlet skin in thousandths of the cutoff (0: rescan the cells on every evaluation)
    int opt_graph = 1;
    // CUDA graph of nbx_step_vv: the conditional rebuild chain (nine launches that return at once) becomes the body of
    // an IF node decided by one single-thread kernel (cudaGraphSetConditional); falls back to plain capture if the
    // runtime refuses
    int opt_cond_nodes = 1;
    bool cond_capture = false, cond_fail = false;
    cudaStream_t aux_stream = nullptr;
    int opt_fuse_update = 1;       // nbx_step_vv: position update + displacement check + record refresh in one kernel
    int opt_verlet_banked = 1;     // Verlet lists in blocks of four by record position in a 128-byte line (conflict-free L1 gathers)
    int opt_verlet_branchfree = 1; // Verlet force kernel: batches of four list entries evaluated without branches
    int opt_verlet_lanes = 0;      // lanes per target of the Verlet force kernel (0: chosen from the system size; 1, 2, 4, 8)
    int opt_sym = 1;            // Newton's-third-law all-pairs kernel for unsharded 1/r^2 systems
    int64_t sym_min_n = 8192;
    int opt_sym_variant = 0;
    int pair_rank = 0, pair_nranks = 1; // pair sharding: this context evaluates ring offsets k = rank mod nranks
    bool forces_done = false;           // nbx_vv_forces ran; nbx_vv_finish only has to finish the step
    bool mass_uniform = false, charge_uniform = false; // all weights equal (value = first element)
    double h_q1 = 0.0;

    // ---- analysis of frames (nbx_analysis.cu) ---------------------------------------------------
    unsigned long long *an_hist = nullptr; // rdf pair-distance histogram
    int an_bins = 0;
    int64_t an_frames = 0;
    double *an_pos = nullptr, *an_pos0 = nullptr, *an_red = nullptr;

    // ---- neighbour scratch ------------------------------------------------------------------
    // ---- host staging -----------------------------------------------------------------------
    double *h_pin = nullptr; size_t h_pin_bytes = 0;

    // ---- timing -----------------------------------------------------------------------------
    bool timing = false;
    nbx::Timer timers[NBX_T_COUNT];
};

namespace nbx {

int fail(nbx_ctx *c, int code, const char *fmt, ...);
int cuda_fail(nbx_ctx *c, cudaError_t e, const char *what);

#define NBX_CUDA(ctx, expr)                                           \
    do {                                                              \
        cudaError_t e__ = (expr);                                     \
        if (e__ != cudaSuccess) return nbx::cuda_fail(ctx, e__, #expr); \
    } while (0)

#define NBX_TRY(expr)                    \
    do {                                 \
        int rc__ = (expr);               \
        if (rc__ != NBX_OK) return rc__; \
    } while (0)

// RAII-ish phase timer (records only when ctx->timing).
void timer_begin(nbx_ctx *c, int phase);
void timer_end(nbx_ctx *c, int phase);

template <typename T>
int dev_alloc(nbx_ctx *c, T **p, size_t count)
{
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (count == 0) return NBX_OK;
    cudaError_t e = cudaMalloc((void **)p, count * sizeof(T));
    if (e != cudaSuccess) return cuda_fail(c, e, "cudaMalloc");
    return NBX_OK;
}

// Pair sharding (every rank evaluates its ring offsets of the Newton's-third-law kernel for ALL bodies, the partial rows
// are summed at the owners) covers: unbounded gravity / Coulomb alone; and Coulomb with a cutoff that no cell list can
// serve (L/3 <= R < L/2 in a cubic box: the periodic variant of the kernel) together with terms that each rank
// evaluates for its own block only (Lennard-Jones over cell lists, the SPC/Fw bonds and angle) -- water with the
// reference's default electrostatic cutoff of 0.49 L.
inline bool pair_central_only(const nbx_ctx *c)
{
    return !c->has_lj && !c->has_dip && !c->has_spcfw && !c->water && (c->has_grav || c->has_coul) &&
           (!c->has_coul || (c->bc_kind == NBX_BC_INFINITE && std::isinf(c->el_R2)));
}
inline bool pair_periodic_coulomb(const nbx_ctx *c)
{
    if (!c->has_coul || c->has_grav || c->has_dip || c->bc_kind != NBX_BC_CUBIC || !c->opt_sym || c->n < c->sym_min_n) return false;
    const double L = c->bc[0], R = c->el_R;
    if (!(L > 0.0) || !std::isfinite(R) || !(R > 0.0) || !(R < 0.5 * L)) return false;
    return !c->opt_cell_list || std::floor(L / (R * (1.0 + 1e-6))) < 3.0; // (cells_plan: fewer than 3 cells per edge)
}
inline bool pair_capable(const nbx_ctx *c) { return pair_central_only(c) || pair_periodic_coulomb(c); }

// ---- kernels implemented across the .cu files (host launchers) ----------------------------
// nbx_allpairs.cu
int launch_allpairs_grav(nbx_ctx *c, const double *w, int scale_kind, double scale, double *acc_out, bool accumulate);
int launch_allpairs_dipole(nbx_ctx *c, double *acc_out, bool accumulate);
// nbx_sympairs.cu
int launch_sympairs(nbx_ctx *c, const double *w, bool uniform, double wval, int scale_kind, double scale,
                    double *acc_out, bool accumulate);
int launch_sympairs_coulomb_pbc(nbx_ctx *c, bool excl3, double *acc_out, bool accumulate);
int launch_allpairs_pbc(nbx_ctx *c, int pot, const double *px, int64_t n, int64_t ld, int64_t lo, int64_t hi,
                        int mstride, double *acc_out, int64_t ld_out, bool accumulate);
// nbx_cells.cu
int cells_plan(nbx_ctx *c, double R, int64_t n, CellGrid *g);
int cells_build(nbx_ctx *c, CellList *cl, const double *px, const double *w, const int *gid, int64_t n, int64_t ld,
                int key_div, const int *cond);
int launch_cells_force(nbx_ctx *c, CellList *cl, int pot, int64_t lo, int64_t hi, int mstride, double *acc_out,
                       int64_t ld_out, bool accumulate, const int *cond);
int cells_pairs(nbx_ctx *c, CellList *cl, double R, int pot, const double *px, const double *w, const int *gid, int64_t n,
                int64_t nplan, int64_t ld, int key_div, int64_t lo, int64_t hi, int mstride, double *acc_out,
                int64_t ld_out, bool accumulate, bool *used);
int cells_neighbors(nbx_ctx *c, CellList *cl, const double *px, int64_t n, int64_t ld, double R2, int64_t *offsets,
                    int32_t *list, int64_t cap);
void cells_free(CellList *cl);
bool lists_can_fuse_update(const nbx_ctx *c, const CellList *cl, const double *px);
int launch_vv_pos_lists(nbx_ctx *c, CellList *cl, const double *w, double dt);
int cells_scan(nbx_ctx *c, const int *in, int *out, int n, int *sums, int round_to, const int *cond);
// nbx_analysis.cu
int analysis_rdf_reset(nbx_ctx *c, int maxbin);
int analysis_rdf_add(nbx_ctx *c, const double *u_host);
int analysis_rdf_get(nbx_ctx *c, int64_t *hist, int64_t cap, int64_t *frames);
int analysis_msd(nbx_ctx *c, const double *u0_host, const double *u_host, double *out);
void analysis_free(nbx_ctx *c);
// nbx_slab.cu
int slab_init(nbx_ctx *c, int rank, int nranks);
int slab_pack(nbx_ctx *c);
int slab_unpack(nbx_ctx *c, int64_t *counts);
int slab_check(nbx_ctx *c, int64_t *counts);
int slab_connect(nbx_ctx *c, const void *left_handle, const void *right_handle, void *left_ptr, void *right_ptr);
int slab_refresh_send(nbx_ctx *c);
int slab_refresh_recv(nbx_ctx *c);
int slab_verlet_check(nbx_ctx *c, double soft_fraction, int *out2_dev, double *out2_dbl);
void slab_free(nbx_ctx *c);
int slab_start(nbx_ctx *c);
int slab_enqueue(nbx_ctx *c, double dt, int64_t nsteps); // nbx_step_vv of a slab: enqueues, does not synchronise
int slab_finish(nbx_ctx *c);
// kernel preloading (see preload_multi in nbx_multi.cu)
void preload_multi();
void preload_slab();
void preload_cells();
void preload_integrate();
// nbx_multi.cu
CommDev comm_dev(const nbx_ctx *c);
int comm_allreduce3(nbx_ctx *c, const double *in0, double *out0, int *flag_out, const cudaGraphConditionalHandle *cond = nullptr);
int ensure_red(nbx_ctx *c);
void maybe_pin(nbx_ctx *c, const void *p, size_t bytes);
int comm_alloc(nbx_ctx *c);
void comm_free(nbx_ctx *c);
int multi_enqueue_vv(nbx_ctx *c, double dt, int64_t nsteps);
int multi_enqueue_em(nbx_ctx *c, double dt, int64_t nsteps);
int multi_finish(nbx_ctx *c);
int multi_accel_enqueue(nbx_ctx *c, const double *u, const double *v);
int multi_accel_exchange(nbx_ctx *c);
int comm_arm_global0(nbx_ctx *c);
int multi_accel_finish(nbx_ctx *c, double *dv);
int group_init(nbx_ctx *c, int rank, int nranks, int mode);
int group_export(nbx_ctx *c, int kind, void **ptr, void *handle64);
int group_connect(nbx_ctx *c, const void *handles, void *const *ptrs);
int group_start(nbx_ctx *c);
// nbx_group.cu: the leader of a single-process group (nbx_create_multi)
int leader_create(nbx_ctx **out, int ndev, const int *devs);
int leader_destroy(nbx_ctx *c);
int leader_system(nbx_ctx *c, int64_t n, const double *m, const double *q, const double *mm, int water);
int leader_upload(nbx_ctx *c, const double *u, const double *v);
int leader_accel(nbx_ctx *c, const double *u, double *v, double *dv);
int leader_step_vv(nbx_ctx *c, double dt, int64_t nsteps);
int leader_step_em(nbx_ctx *c, double dt, int64_t nsteps, uint64_t seed);
int leader_download(nbx_ctx *c, double *u, double *v, double *dv);
int leader_energy(nbx_ctx *c, double *ekin, double *epot, double *temperature);
void graph_drop(nbx_ctx *c);
// nbx_cells.cu: the launches between begin and end become the body of a graph IF node on flag[0] while capturing
struct CondScope { bool active = false; cudaStream_t saved = nullptr; };
int cond_handle_create(nbx_ctx *c, cudaGraphConditionalHandle *h, bool *ok);
int cond_scope_begin(nbx_ctx *c, const int *flag, CondScope *sc, const cudaGraphConditionalHandle *pre);
int cond_scope_end(nbx_ctx *c, CondScope *sc);
// nbx_bonded.cu
int launch_spcfw_bonded(nbx_ctx *c, double *acc_out);
// nbx_integrate.cu
int launch_aos_to_soa(nbx_ctx *c, const double *aos, double *soa, int64_t ncols_used);
int launch_soa_to_aos(nbx_ctx *c, const double *soa, double *aos, int64_t n, int64_t ncols_total, int64_t lo, int64_t hi);
int launch_fill(nbx_ctx *c, double *p, double v, int64_t count);
int launch_sum_mv2(nbx_ctx *c, const double *vel, int64_t lo, int64_t hi); // -> d_scal[0]
int launch_thermostat_rhs(nbx_ctx *c, double *acc, const double *vel);
int launch_vv_pos(nbx_ctx *c, double dt);
int launch_vv_vel(nbx_ctx *c, double dt, bool with_thermostat);
int launch_em_step(nbx_ctx *c, double dt);
int launch_andersen(nbx_ctx *c, double dt);
int check_finite(nbx_ctx *c, const double *soa, int64_t n);
int reduce_kinetic(nbx_ctx *c, double *ekin, double *temp);
int reduce_potential(nbx_ctx *c, double *epot);
int measure_fp64_peak(nbx_ctx *c, double *tflops, double *mhz);
int measure_hbm_peak(nbx_ctx *c, double *gbs);
// nbx_api.cu
int compute_accel(nbx_ctx *c); // pos, vel -> acc (all potentials + RHS thermostats)

// ---- device helpers --------------------------------------------------------------------------
#ifdef __CUDACC__
// launch bound -> actual count (see nbx_ctx::dyn)
__device__ __forceinline__ int dyn_own(const int *dyn, int bound) { return dyn ? min(bound, dyn[0]) : bound; }
__device__ __forceinline__ int dyn_loc(const int *dyn, int bound) { return dyn ? min(bound, dyn[0] + dyn[1]) : bound; }
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// MUFU.RSQ64H seed: ~2^-22 relative accuracy, no special-case handling.
__device__ __forceinline__ double rsqrt_seed(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
// w * x^(-3/2) to ~3 ulp from the seed: with a = y0^2 (exact: y0 has 21 significant bits),
// e = 1 - x a,  x^(-3/2) = y0^3 (1 - e)^(-3/2) = y0^3 (1 + e (3/2 + 15/8 e) + O(e^3)), e^3 < 1e-19.
__device__ __forceinline__ double w_rinv3(double r2, double w)
{
    const double y0 = rsqrt_seed(r2);
    const double a = y0 * y0;
    const double e = fma(-r2, a, 1.0);
    const double c = a * (y0 * w);
    const double p = fma(1.875, e, 1.5);
    const double q = p * e;
    return fma(c, q, c);
}
// get_interparticle_distance wrap loops (src/boundary_conditions.jl:111-165), bit-exact:
// repeated rounded subtraction/addition of the box edge, bounded to 64 trips per side so a
// wild coordinate cannot hang the GPU (the reference would spin; NaN/Inf is rejected earlier).
__device__ __forceinline__ double wrap_cubic(double x, double radius, double size)
{
    for (int k = 0; k < 64 && x >= radius; ++k) x = __dsub_rn(x, size);
    for (int k = 0; k < 64 && x < -radius; ++k) x = __dadd_rn(x, size);
    return x;
}
__device__ __forceinline__ double wrap_range(double x, double lo, double hi)
{
    const double len = __dsub_rn(hi, lo);
    for (int k = 0; k < 64 && x < lo; ++k) x = __dadd_rn(x, len);
    for (int k = 0; k < 64 && x >= hi; ++k) x = __dsub_rn(x, len);
    return x;
}
// r2 = x^2 + y^2 + z^2, left to right, no FMA contraction (:162).
__device__ __forceinline__ double r2_unfused(double x, double y, double z)
{
    return __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
}
// binning copy of a coordinate, wrapped into [0, L) (cf. src/nbody_simulation_result.jl:571)
__device__ __forceinline__ double wrapped_coord(double x, double L)
{
    double w = x - L * floor(x / L);
    if (w < 0.0) w += L;
    if (w >= L) w -= L;
    return w;
}
__device__ __forceinline__ int cell_coord(double x, double L, int nc)
{
    int cx = (int)(wrapped_coord(x, L) * ((double)nc / L));
    return cx < 0 ? 0 : (cx >= nc ? nc - 1 : cx);
}

// 1/x to < 1 ulp-ish (error e^3, e = seed error ~2^-22) without the IEEE division's slow path
__device__ __forceinline__ double rcp_fast(double x)
{
    double y0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double e = fma(-x, y0, 1.0);
    const double p = fma(e, e, e);
    return fma(y0, p, y0);
}

// "LL" words (the low-latency protocol of collective libraries): a 64-bit store is atomic, so {32-bit payload, 32-bit tag}
// validates itself -- the receiver spins on the tag, and neither side needs a fence, a flag or a completion count.
__device__ __forceinline__ void ll_store(unsigned long long *dst, double v, unsigned tag)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    const unsigned long long t = (unsigned long long)tag << 32;
    *reinterpret_cast<volatile unsigned long long *>(dst) = t | (b & 0xffffffffull);
    *reinterpret_cast<volatile unsigned long long *>(dst + 1) = t | (b >> 32);
}
// false on a time-out
__device__ __forceinline__ bool ll_load(const unsigned long long *src, unsigned tag, unsigned long long timeout_ns, double *out)
{
    const volatile unsigned long long *p = reinterpret_cast<const volatile unsigned long long *>(src);
    unsigned long long lo = p[0], hi = p[1];
    if ((unsigned)(lo >> 32) != tag || (unsigned)(hi >> 32) != tag) {
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            lo = p[0]; hi = p[1];
            if ((unsigned)(lo >> 32) == tag && (unsigned)(hi >> 32) == tag) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) return false;
        }
    }
    *out = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
    return true;
}

// the two words of a value `stride` words apart (word-major message layouts: a warp's store is contiguous)
__device__ __forceinline__ void ll_store_strided(unsigned long long *dst, size_t stride, double v, unsigned tag)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    const unsigned long long t = (unsigned long long)tag << 32;
    *reinterpret_cast<volatile unsigned long long *>(dst) = t | (b & 0xffffffffull);
    *reinterpret_cast<volatile unsigned long long *>(dst + stride) = t | (b >> 32);
}
__device__ __forceinline__ bool ll_load_strided(const unsigned long long *src, size_t stride, unsigned tag, unsigned long long timeout_ns,
                                                double *out)
{
    const volatile unsigned long long *p = reinterpret_cast<const volatile unsigned long long *>(src);
    unsigned long long lo = p[0], hi = p[stride];
    if ((unsigned)(lo >> 32) != tag || (unsigned)(hi >> 32) != tag) {
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            lo = p[0]; hi = p[stride];
            if ((unsigned)(lo >> 32) == tag && (unsigned)(hi >> 32) == tag) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) return false;
        }
    }
    *out = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
    return true;
}

// one 32-byte cell-order record with a single 256-bit load (LDG.E.ENL2.256): half the L1 sector traffic of
// a 128 + 64 bit pair when every lane gathers a different record
__device__ __forceinline__ double4 load_rec(const double4 *p)
{
    double4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}

// a word of a stream that is read once (Verlet list entries): no L1 allocation, so that it does not displace the gathered
// records
__device__ __forceinline__ int ld_stream(const int *p)
{
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// L2 eviction policies (createpolicy): the gathered position records should stay resident, the streamed per-slot
// state and the lists should not displace them (at 1M atoms a step streams ~400 MB past a 35 MB gather target).
__device__ __forceinline__ uint64_t l2_policy_keep()
{
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_stream()
{
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ double4 load_rec_nc_hint(const double4 *p, uint64_t pol)
{
    double4 r;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
                 : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ double4 load_rec_hint(const double4 *p, uint64_t pol)
{
    double4 r;
    asm volatile("ld.global.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
                 : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p), "l"(pol) : "memory");
    return r;
}
__device__ __forceinline__ void store_rec_hint(double4 *p, const double4 v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w), "l"(pol) : "memory");
}
#endif

} // namespace nbx
