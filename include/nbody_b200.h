/*
 * nbody_b200.h -- C ABI of libnbody_b200.so: the B200-native replacement of the acceleration
 * hot path of SciML/NBodySimulator.jl (reference v1.15.0).
 *
 * The reference is pure Julia and has no FFI of its own; every entry point below states the
 * reference interface (file:line under /root/reference) whose work it takes over, i.e. what a
 * Julia `ccall` shim binds (INTEGRATION.md shows the shim).
 *
 * Conventions
 *   - Every function returns NBX_OK (0) or a negative nbx_status; the message of the last
 *     failure of a context is returned by nbx_last_error().  Nothing throws, nothing exits.
 *   - Host arrays `u`, `v`, `dv` are Julia `Matrix{Float64}` bytes: 3 x ncols, column-major
 *     (x1 y1 z1 x2 y2 z2 ...).  They are caller-owned and only read/written during the call.
 *     ncols == n, except with the Nose-Hoover thermostat where ncols == n + 1
 *     (src/nbody_to_ode.jl:6-8).  Indices are 0-based everywhere in this ABI.
 *   - One context == one simulation on one GPU (nbx_create) or on several (nbx_create_multi).  Calls on a
 *     context are blocking and must not be issued concurrently (the reference's RHS is called from one Julia task).
 *   - There is NO CPU fallback: without a CUDA device nbx_create fails with NBX_ERR_CUDA.
 */
#ifndef NBODY_B200_H
#define NBODY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NBX_API __attribute__((visibility("default")))
#else
#define NBX_API
#endif

typedef struct nbx_ctx nbx_ctx;

typedef enum {
    NBX_OK = 0,
    NBX_ERR_INVALID = -1,     /* bad argument / call order                                    */
    NBX_ERR_CUDA = -2,        /* CUDA runtime error (message holds cudaGetErrorString)        */
    NBX_ERR_NONFINITE = -3,   /* NaN/Inf coordinate (the reference's wrap loop would hang)    */
    NBX_ERR_UNSUPPORTED = -4, /* combination the reference itself cannot express              */
    NBX_ERR_CAPACITY = -5     /* caller buffer too small (nbx_neighbors)                      */
} nbx_status;

/* boundary kinds: src/boundary_conditions.jl:68-72 (InfiniteBox), :101-103 (Cubic, b[0]=L),
 * :25-29 (Periodic, b[0..5] = xlo xhi ylo yhi zlo zhi; NOT a minimum image, kept as is). */
enum { NBX_BC_INFINITE = 0, NBX_BC_CUBIC = 1, NBX_BC_PERIODIC = 2 };

/* thermostat kinds: src/thermostats.jl (Berendsen :60-83, Nose-Hoover :93-128, Andersen and
 * Langevin structs) -- Andersen/Langevin act in the stepper (nbx_step_*), not in the RHS. */
enum { NBX_THERMO_NONE = 0, NBX_THERMO_BERENDSEN = 1, NBX_THERMO_NOSEHOOVER = 2,
       NBX_THERMO_ANDERSEN = 3, NBX_THERMO_LANGEVIN = 4 };

/* timing phases for nbx_timing_get */
enum { NBX_T_PAIR_ALLPAIRS = 0,  /* tiled all-pairs kernel (gravity / Coulomb / dipole / PBC)   */
       NBX_T_CELL_BUILD = 1,     /* wrap -> cell id -> count -> scan -> scatter                 */
       NBX_T_PAIR_CELLS = 2,     /* cutoff pair kernel over the cell list                       */
       NBX_T_BONDED = 3,         /* SPC/Fw bonds + angle                                        */
       NBX_T_INTEGRATE = 4,      /* fused velocity-Verlet / thermostat update kernels           */
       NBX_T_TRANSPOSE = 5,      /* AoS <-> SoA at the boundary                                 */
       NBX_T_COUNT = 6 };

/* ---- lifetime ------------------------------------------------------------------------- */
/* One context per NBodySimulation (src/nbody_simulation.jl:42-54).  `device` is the CUDA
 * ordinal.  Fails loudly (NBX_ERR_CUDA) when no sm_100 device is present. */
NBX_API int nbx_create(nbx_ctx **out, int device);
/* ONE handle for ndev GPUs of this process (SURVEY.md 8(b) "Proposed C ABI": nbx_create(nbx_ctx**, int ndev, const int* devs)).
 * The handle takes the same calls as a single-GPU context -- nbx_system, nbx_boundary, nbx_add_*, nbx_thermostat,
 * nbx_accel, nbx_upload, nbx_step_vv, nbx_step_em, nbx_download, nbx_energy -- and fans them out: the target loop of
 * soode_system! (src/nbody_to_ode.jl:474-488, :502-532) is split over the devices (pair sharding for unbounded gravity /
 * Coulomb, x-slabs for cutoff Lennard-Jones / Coulomb in a cubic box, target blocks otherwise; option "group_mode"), all
 * exchanges run device to device over NVLink peer memory inside the step kernels, and one blocking call from one
 * host thread (a Julia task) drives all GPUs.  devs may repeat a device (several members on one GPU: testing). */
NBX_API int nbx_create_multi(nbx_ctx **out, int ndev, const int *devs);
NBX_API int nbx_destroy(nbx_ctx *ctx);
/* Message of the last failing call on ctx (ctx == NULL: of the last failing nbx_create). */
NBX_API const char *nbx_last_error(const nbx_ctx *ctx);
/* Library/ABI version (major*10000 + minor*100 + patch). */
NBX_API int nbx_version(void);

/* ---- system description (what the reference's closures capture) ----------------------- */
/* Per-particle tables built once at problem construction by obtain_data_for_*:
 * masses (src/nbody_to_ode.jl:290-300, src/nbody_simulation_result.jl:121-139), charges
 * (:316-351), magnetic moments 3 x n (src/bodies.jl:102-107).  q and mm may be NULL.
 * water != 0: the n columns are (O, H1, H2) triples (src/nbody_to_ode.jl:46): LJ acts on the
 * oxygen columns only (:302-314) and Coulomb excludes the own molecule (:331-351). */
NBX_API int nbx_system(nbx_ctx *ctx, int64_t n, const double *m, const double *q,
                       const double *mm, int water);
/* Boundary conditions used by get_interparticle_distance (src/boundary_conditions.jl:111-172). */
NBX_API int nbx_boundary(nbx_ctx *ctx, int kind, const double *b);

/* Potentials == the entries of PotentialNBodySystem.potentials (src/nbody_system.jl:72-122);
 * each nbx_add_* replaces the closure get_accelerating_function returns for that parameter
 * type (src/nbody_to_ode.jl:156-242).  Evaluation order is fixed: lennard_jones,
 * electrostatic, magnetostatic, gravitational, then SPC/Fw bonded terms. */
/* GravitationalParameters(G), gravitational_acceleration!  src/basic_potentials.jl:306-331 */
NBX_API int nbx_add_gravity(nbx_ctx *ctx, double G);
/* LennardJonesParameters(eps, sigma, R), pairwise_lennard_jones_acceleration!  :240-272 */
NBX_API int nbx_add_lj(nbx_ctx *ctx, double eps, double sigma, double R);
/* ElectrostaticParameters(k, R) (R may be +Inf), pairwise_electrostatic_acceleration! :274-304 */
NBX_API int nbx_add_coulomb(nbx_ctx *ctx, double k, double R);
/* MagnetostaticParameters(mu_4pi), magnetostatic_dipdip_acceleration!  :333-365 */
NBX_API int nbx_add_dipole(nbx_ctx *ctx, double mu_4pi);
/* SPCFwParameters(rOH, aHOH, kb, ka): harmonic_bond_potential_acceleration! :367-393 and
 * valence_angle_potential_acceleration! :395-433 (requires water != 0). */
NBX_API int nbx_add_spcfw(nbx_ctx *ctx, double rOH, double aHOH, double kb, double ka);
NBX_API int nbx_clear_potentials(nbx_ctx *ctx);

/* Thermostat (src/thermostats.jl; wiring src/nbody_to_ode.jl:377-433).
 *   BERENDSEN : param = tau  (gamma = 0.5/tau, :72-74)
 *   NOSEHOOVER: param = tau
 *   ANDERSEN  : param = nu   (src/nbody_simulation_result.jl:504-540)
 *   LANGEVIN  : param = gamma (src/nbody_to_ode.jl:567-598)
 * N, Nc: particle / constraint counts of md_temperature (:87-91): N = n, Nc = 0 for atoms;
 * N = n (=3*molecules), Nc = 2*molecules for water (src/nbody_to_ode.jl:402-408). */
NBX_API int nbx_thermostat(nbx_ctx *ctx, int kind, double T0, double param, double kB,
                           int64_t N, int64_t Nc);

/* Multi-GPU: this context evaluates only target columns [lo, hi) (all columns are sources).
 * Default [0, n).  Columns outside the range are left untouched in device state and are
 * written as zeros by nbx_accel.  (No reference equivalent: the reference is serial.) */
NBX_API int nbx_shard(nbx_ctx *ctx, int64_t lo, int64_t hi);

/* Multi-GPU, alternative for unbounded 1/r^2 systems (gravity, Coulomb with R = Inf in an InfiniteBox):
 * pair sharding.  The unordered pair set is split over nranks contexts (Newton's third law kernel,
 * ring offsets k = rank mod nranks); after nbx_vv_forces every context holds a PARTIAL acceleration
 * for ALL columns in the acc rows (nbx_device_ptr which = 2) and the host adds them across ranks
 * (reduce-scatter / all-reduce) before nbx_vv_finish.  Combine with nbx_shard(lo, hi), which then only
 * selects the columns this context integrates. */
NBX_API int nbx_shard_pairs(nbx_ctx *ctx, int rank, int nranks);

/* ---- RHS drop-in ---------------------------------------------------------------------- */
/* soode_system!(dv, v, u, p, t): src/nbody_to_ode.jl:474-488 (PotentialNBodySystem) and
 * :502-532 (WaterSPCFw).  Host pointers, 3 x ncols each.  dv is overwritten.  With the
 * Nose-Hoover thermostat v[3n] is rewritten exactly as the reference mutates it
 * (src/thermostats.jl:126); otherwise v is read-only. */
NBX_API int nbx_accel(nbx_ctx *ctx, const double *u, double *v, double t, double *dv);

/* ---- device-resident stepping (the fused drop-in for run_simulation) ------------------- */
/* run_simulation(sim, VelocityVerlet(), dt) : src/nbody_simulation_result.jl:468-487.
 * upload sets x(0), v(0) and evaluates a(0); step_vv advances nsteps with
 *   x+ = x + dt v + dt^2/2 a;  a+ = f(v, x+);  v+ = v + dt/2 (a + a+)
 * (OrdinaryDiffEqSymplecticRK VelocityVerlet, upstream); Berendsen / Nose-Hoover enter
 * through f, Andersen resamples velocities after each step.  step_em is Euler-Maruyama on
 * the Langevin SDE (src/nbody_to_ode.jl:575-595; for water the SDEProblem of WaterSPCFw, :600-680, with its
 * own drift and noise amplitudes as written there).  download copies state back (any pointer may be NULL). */
NBX_API int nbx_upload(nbx_ctx *ctx, const double *u, const double *v);
NBX_API int nbx_step_vv(nbx_ctx *ctx, double dt, int64_t nsteps);
NBX_API int nbx_step_em(nbx_ctx *ctx, double dt, int64_t nsteps, uint64_t seed);
NBX_API int nbx_download(nbx_ctx *ctx, double *u, double *v, double *dv);
/* Frames for the result accessors (run_simulation(...; saveat), src/nbody_simulation_result.jl:468-487; SimulationResult
 * :5-8, :49-104): nsteps velocity-Verlet steps on the device, the state copied out after every save_every steps and after
 * the last one: frame k = the 3 x ncols arrays at u_frames + 3 ncols k (t = t0 + min((k + 1) save_every, nsteps) dt);
 * either array may be NULL.  max_frames: capacity in frames (ceil(nsteps / save_every) are written; NBX_ERR_CAPACITY
 * otherwise); *nframes receives the count.  Works on single contexts, group members and nbx_create_multi handles. */
NBX_API int nbx_run_vv(nbx_ctx *ctx, double dt, int64_t nsteps, int64_t save_every, double *u_frames, double *v_frames,
                       int64_t max_frames, int64_t *nframes);
NBX_API int nbx_set_seed(nbx_ctx *ctx, uint64_t seed);

/* Split form of one velocity-Verlet step for multi-GPU drivers that exchange positions
 * between the halves:  begin = position update of the own shard;  the host all-gathers the
 * SoA position arrays (nbx_device_ptr);  finish = forces + velocity update of the shard. */
NBX_API int nbx_vv_begin(nbx_ctx *ctx, double dt);
/* optional middle step: evaluate the pair potentials only (needed with nbx_shard_pairs, where the host
 * sums the partial accelerations across ranks between nbx_vv_forces and nbx_vv_finish) */
NBX_API int nbx_vv_forces(nbx_ctx *ctx);
NBX_API int nbx_vv_finish(nbx_ctx *ctx, double dt);
/* Evaluate a = f(v, x) of the resident state (all potentials + RHS thermostats). */
NBX_API int nbx_eval_resident(nbx_ctx *ctx);

/* kinetic_energy src/nbody_simulation_result.jl:209-212; potential_energy :239-264
 * (lennard_jones_potential :293-319, electrostatic_potential :321-351,
 * harmonic_bonds_potential :353-372, valence_angle_harmonic_potential :374-397);
 * temperature = md_temperature src/thermostats.jl:87-91 of the resident velocities.
 * Any pointer may be NULL. */
NBX_API int nbx_energy(nbx_ctx *ctx, double *ekin, double *epot, double *temperature);

/* In-cutoff ordered pair set of the LJ predicate (get_interparticle_distance +
 * `rij_2 < p.R2`, src/basic_potentials.jl:253-258) for the resident positions, as CSR:
 * offsets[n+1], list[offsets[n]] with the partners of each column in ascending order.
 * cap = capacity of list in entries; NBX_ERR_CAPACITY if too small (offsets still valid). */
NBX_API int nbx_neighbors(nbx_ctx *ctx, int64_t *offsets, int32_t *list, int64_t cap);

/* ---- slab decomposition (multi-GPU, cutoff potentials in a cubic periodic box) ---------- */
/* No reference equivalent (the reference is serial): the target loop of soode_system!
 * (src/nbody_to_ode.jl:474-488) for cutoff Lennard-Jones (src/basic_potentials.jl:240-272) and
 * cutoff Coulomb (:274-304) under CubicPeriodicBoundaryConditions (src/boundary_conditions.jl:138-165),
 * cut along x into slabs of whole cell layers, one context per slab.
 *
 * nbx_slab_init: every rank has described and uploaded the FULL system (nbx_system, nbx_upload).
 * Rank `rank` of `nranks` keeps the particles of its layers (with their a(0)), records their global
 * ids (= column numbers of the upload); the first nbx_slab_pack then selects the own particles and
 * fills the two send buffers with the halo of the boundary layers, the host exchanges the buffers and
 * calls nbx_slab_unpack.  Needs >= 2 layers per slab.
 * One velocity-Verlet step:  nbx_vv_begin; nbx_slab_pack; exchange; nbx_slab_unpack; nbx_vv_forces;
 * nbx_vv_finish; all-reduce of the scalar block's [0] (sum m v^2) when a thermostat is set.
 * Exchange: buffer 0 (send-to-left) -> buffer 3 (recv-from-right) of the left neighbour,
 *           buffer 1 (send-to-right) -> buffer 2 (recv-from-left) of the right neighbour; periodic.
 * The particle counts stay on the device: a step is a pure stream of launches (capturable in a CUDA graph
 * with the exchange).  nbx_slab_unpack(counts = NULL) is asynchronous; errors (a particle that jumped past
 * the neighbouring slab, a full buffer) accumulate and are reported by the next nbx_slab_check, by
 * nbx_slab_unpack with counts != NULL, or by nbx_slab_download -- all three synchronise the stream.
 * counts[6] = own, ghosts, migrated out left/right, in from left/right (of the last step). */
NBX_API int nbx_slab_init(nbx_ctx *ctx, int rank, int nranks);
NBX_API int nbx_slab_pack(nbx_ctx *ctx);
NBX_API int nbx_slab_unpack(nbx_ctx *ctx, int64_t *counts);
NBX_API int nbx_slab_check(nbx_ctx *ctx, int64_t *counts);
/* Verlet lists inside a slab (one cutoff potential, option verlet_skin_permille > 0; nbx_get_info "slab_verlet" = 1):
 * the local numbering stays put between COLLECTIVE rebuilds, so the lists survive.  Every step the driver calls
 * nbx_slab_verlet_check after nbx_vv_begin (out2 = two device ints: [0] some own particle moved more than
 * soft_fraction x skin/2 since the last rebuild, [1] more than skin/2 or a list overflowed = the lists are no longer
 * a superset), combines the ints over all ranks (max) and, when it decides to rebuild -- on the same step on every
 * rank -- runs
 *     nbx_slab_pack; exchange; nbx_slab_unpack;                 (migration, as without lists)
 *     nbx_set_option("slab_record_halo", 1); nbx_slab_pack; exchange; nbx_slab_unpack;   (halo incl. the arrivals)
 *     nbx_set_option("slab_rebuild", 1); nbx_vv_forces; nbx_vv_finish
 * and otherwise only moves the positions of the recorded boundary-layer particles:
 *     nbx_slab_refresh_send; exchange; nbx_slab_refresh_recv; nbx_vv_forces; nbx_vv_finish.
 * parallel.SlabStepper reads the combined ints two steps late (no host stall) and rebuilds at a soft limit of 0.75.
 * nbx_slab_prime builds the lists from the positions as they are (after a migration + halo round) without touching
 * the resident accelerations: the driver calls it once after the initial distribution, so that the first step already
 * runs from lists and the rebuild schedule is that of a single context. */
NBX_API int nbx_slab_prime(nbx_ctx *ctx);
/* The two halves of a slab step as single calls (fewer host round trips per step; the exchange must be direct or the
 * slab alone).  The driver keeps three doubles of the scalar block (nbx_device_ptr which = 3) in flight: [12] this
 * rank's sum m v^2 of the previous step, [13] / [14] the soft / hard rebuild flags of nbx_slab_verlet_check as 0.0 / 1.0.
 *   nbx_slab_step_begin(dt, soft_fraction): nbx_vv_begin, zero [13..14], displacement check into them.
 *   (driver: ONE sum over the ranks of [12..14] -- [13..14] only on the first step --, asynchronous copy of [13..14] home)
 *   nbx_slab_step_end(dt, refresh): refresh != 0: nbx_slab_refresh_send + nbx_slab_refresh_recv; then nbx_vv_forces,
 *   nbx_vv_finish (the Berendsen term reads the summed [12]; the new local sum is written to [0] and [12]).
 * nbx_slab_step_begin switches the context to slot 12 for good (option "temperature_slot"). */
NBX_API int nbx_slab_step_begin(nbx_ctx *ctx, double dt, double soft_fraction);
NBX_API int nbx_slab_step_end(nbx_ctx *ctx, double dt, int refresh);
NBX_API int nbx_slab_refresh_send(nbx_ctx *ctx);
NBX_API int nbx_slab_refresh_recv(nbx_ctx *ctx);
NBX_API int nbx_slab_verlet_check(nbx_ctx *ctx, double soft_fraction, void *out2_dev);
/* Direct exchange over NVLink peer memory (optional, between nbx_slab_init and the first nbx_slab_pack):
 * nbx_slab_rx returns this rank's receive area (device pointer, size, and -- if ipc_handle64 != NULL -- its
 * 64-byte CUDA IPC handle for another process).  nbx_slab_connect maps the neighbours' receive areas, given
 * as IPC handles or, inside one process, as device pointers (pointer wins when both are set).  From then on
 * nbx_slab_pack stores its messages straight into the neighbours' memory and raises their flags, and
 * nbx_slab_unpack waits on its own flags on the device: no host-side exchange. */
NBX_API int nbx_slab_rx(nbx_ctx *ctx, void **ptr, int64_t *ndoubles, void *ipc_handle64);
NBX_API int nbx_slab_connect(nbx_ctx *ctx, const void *left_handle64, const void *right_handle64,
                             void *left_ptr, void *right_ptr);
/* Host-driven exchange: which = 0 send-to-left, 1 send-to-right, 2 recv-from-left, 3 recv-from-right; device
 * memory of *ndoubles doubles each, owned by the context. */
NBX_API int nbx_slab_buffer(nbx_ctx *ctx, int which, void **ptr, int64_t *ndoubles);
/* The own particles of the slab: their global ids and state as 3 x n_own column-major host arrays
 * (any pointer may be NULL; capacity: the full system's column count). */
NBX_API int nbx_slab_download(nbx_ctx *ctx, int64_t *n_own, int32_t *gid, double *u, double *v, double *dv);

/* ---- groups across processes (one context per process and GPU; nbx_create_multi does the same inside one process) ----
 * No reference equivalent (the reference is serial).  Every rank describes and uploads the FULL system, then:
 *   nbx_group_init(ctx, rank, nranks, mode)   mode 0: chosen from the potentials; 1: pair sharding (unbounded gravity /
 *       Coulomb: ring offsets of the Newton's-third-law kernel, partial accelerations pushed to their owners); 2: target
 *       blocks [lo, hi) of the columns (any potential; water keeps molecules whole); 3: x-slabs (cutoff systems, cubic box)
 *   nbx_group_export(ctx, kind, &ptr, handle64)   kind 0: the rank's window (flags + scalars), 1: slab receive area, 2: SoA
 *       position rows, 3: staging area of partial accelerations; a kind the decomposition does not use gives NULL / zeros
 *   (host: all-gather the 4 x 64-byte handles of every rank -- MPI, torch.distributed, a file: any bootstrap will do)
 *   nbx_group_connect(ctx, handles[nranks][4][64], ptrs[nranks][4])   maps the peers' memory (CUDA IPC over NVLink)
 *   (host: barrier)   nbx_group_start(ctx)   (slabs: the initial distribution; enqueues only)   nbx_slab_check / nbx_synchronize
 * From then on nbx_accel (own block of u up, all-gather over NVLink, own columns of dv back; other columns untouched),
 * nbx_step_vv, nbx_step_em, nbx_download (all positions; velocities / accelerations of the own columns), nbx_energy
 * (kinetic part: the own block's share) act on the group: every rank makes the same calls, the exchanges happen inside
 * the kernels (peer stores + flags, 10 s time-out -> NBX_ERR_CUDA, never a hang), and a step is replayed as a CUDA graph. */
NBX_API int nbx_group_init(nbx_ctx *ctx, int rank, int nranks, int mode);
NBX_API int nbx_group_export(nbx_ctx *ctx, int kind, void **ptr, void *ipc_handle64);
NBX_API int nbx_group_connect(nbx_ctx *ctx, const void *handles, void *const *ptrs);
NBX_API int nbx_group_start(nbx_ctx *ctx);

/* ---- plumbing for the host layer ------------------------------------------------------- */
/* CUDA stream (cudaStream_t as void*) all work of ctx is enqueued on; NULL = the context's own
 * non-blocking stream (the default).  To share the legacy default stream pass cudaStreamLegacy
 * ((void*)0x1), not 0. */
NBX_API int nbx_set_stream(nbx_ctx *ctx, void *stream);
NBX_API int nbx_synchronize(nbx_ctx *ctx);
/* Device pointers of the resident SoA state: which = 0 pos, 1 vel, 2 acc; each is
 * double[3][*ld] (x-row, y-row, z-row) with row stride *ld >= n. */
NBX_API int nbx_device_ptr(nbx_ctx *ctx, int which, void **ptr, int64_t *ld);
/* Split-phase RHS drop-in for a pair-sharded context (nbx_shard + nbx_shard_pairs; unbounded gravity / Coulomb).
 * nbx_accel_begin copies u (HOST, 3 x ncols as nbx_accel) to the device and enqueues this rank's share of the
 * unordered pair set: the resident acceleration rows (nbx_device_ptr which = 2) then hold PARTIAL sums for ALL
 * bodies.  The caller adds the rows across the ranks on ctx's stream (reduce-scatter; parallel.ShardedStepper does
 * it over NCCL) and nbx_accel_end returns the columns [lo, hi) of nbx_shard into dv (HOST; other columns zero) and
 * synchronises.  Replaces the same call as nbx_accel (soode_system!, src/nbody_to_ode.jl:474-488). */
NBX_API int nbx_accel_begin(nbx_ctx *ctx, const double *u);
NBX_API int nbx_accel_end(nbx_ctx *ctx, double *dv);
/* Device-pointer form of nbx_accel (u, v, dv are DEVICE pointers, AoS 3 x ncols). */
NBX_API int nbx_accel_device(nbx_ctx *ctx, const double *u_dev, double *v_dev, double t,
                             double *dv_dev);
/* Per-phase device timers (CUDA events on ctx's stream).  enable != 0 starts recording;
 * get returns the summed duration and launch count since the last reset. */
NBX_API int nbx_timing_enable(nbx_ctx *ctx, int enable);
NBX_API int nbx_timing_get(nbx_ctx *ctx, int phase, double *total_ms, int64_t *count);
NBX_API int nbx_timing_reset(nbx_ctx *ctx);
/* Tuning knobs (integers; defaults in parentheses; DESIGN.md section 3 says what each buys).  Unknown key -> NBX_ERR_INVALID.
 *   cutoff potentials : "cell_list" (1: cell lists where the box allows, 0: all-pairs kernel with the exact predicate),
 *                       "prefilter" (1: fp32 candidate scan before the exact fp64 predicate),
 *                       "verlet_skin_permille" (80: Verlet lists with skin = 0.08 R; 0: rescan the cells every evaluation),
 *                       "verlet_lanes" (0: lanes per target chosen from the system size; 1, 2, 4, 8),
 *                       "verlet_banked" (1: a target's list is stored in blocks of four ordered by the partner record's position
 *                       inside its 128-byte line, so that the gathers of four neighbouring lanes never collide in the L1 data
 *                       banks; the order depends on the LOCAL slot numbers, so slabs and the members of nbx_create_multi
 *                       always use 0, and a context that is compared bit for bit with them -- or joins a group by hand after
 *                       its upload -- sets 0 as well), "verlet_branchfree" (1: batches of four list entries evaluated without
 *                       branches; same operations, same sums),
 *   groups            : "group_mode" (leader of nbx_create_multi; 0), "spin_timeout_ms" (10000: a device-side wait for a peer gives
 *                       up after this long and raises the error flag nbx_slab_check / nbx_synchronize report), "pin_host" (0; 1: page-lock the caller's u / v / dv buffers
 *                       the first time nbx_accel sees them -- the caller must keep them alive until nbx_destroy / nbx_system),
 *   nbx_step_vv       : "graph" (1: two-step CUDA graph), "graph_if_nodes" (1: rebuild chain as the body of an IF node),
 *                       "fuse_update" (1: position update + displacement check + record refresh in one kernel),
 *   all-pairs         : "symmetric_pairs" (1: Newton's-third-law kernel), "symmetric_min_n" (8192), "sym_variant" (0),
 *                       "sym_seg_len" (0: ring offsets per work item chosen from the share; 1 .. 16),
 *                       "uniform_weights" (0 forgets that all masses / charges are equal),
 *   slab driver       : "slab_record_halo", "slab_rebuild", "temperature_slot" (see the slab section above).
 * nbx_get_info keys: "n", "npad", "ncols", "water", "sm_count", "thermostat", "cells_lj", "cells_el", "verlet_lj", "verlet_el",
 * "verlet_overflow", "verlet_rebuilds", "graph_if_nodes", "allpairs_grid", "allpairs_chunks", "slab_own", "slab_ghost", "slab_layer_lo", "slab_layer_hi",
 * "slab_layers", "slab_verlet", "group_mode", "group_rank", "group_size", "shard_lo", "shard_hi", "graph_cached". */
NBX_API int nbx_set_option(nbx_ctx *ctx, const char *key, int64_t value);
NBX_API int nbx_get_info(nbx_ctx *ctx, const char *key, int64_t *value);
/* DFMA-saturation microbenchmark: the measured FP64 roofline denominator (TFLOP/s). */
NBX_API int nbx_measure_fp64_peak(nbx_ctx *ctx, double *tflops, double *sm_mhz_effective);
/* STREAM-style copy bandwidth (GB/s, read+write) of this device, for the HBM roofline. */
NBX_API int nbx_measure_hbm_peak(nbx_ctx *ctx, double *gbs);
/* ---- analysis of saved frames (the first caller-side hotspots once the step loop is fast) ---------
 * rdf(sr), src/nbody_simulation_result.jl:664-709: O(frames x N^2) pair loop over the Lennard-Jones index set (all
 * bodies, or the oxygens of water) with get_interparticle_distance of the cubic box.  Per frame: nbx_rdf_add(u) adds the
 * frame's pairs to a device histogram (u: HOST 3 x ncols frame, or NULL for the resident positions) exactly as :676-693
 * do -- `r2 < (0.5 L)^2`, bin = ceil(r / dr), `1 < bin <= maxbin` -> += 2; integer counts, bit-exact.  nbx_rdf_get
 * returns the counts (hist[b - 1] = the reference's hist[b]) and the number of frames; the caller normalises as :695-707.
 * nbx_rdf_reset(maxbin) clears it (the reference uses maxbin = 1000, the default).
 * msd(sr), :730-783: nbx_msd(u0, u, out) = mean over the index set of |r(t) - r(0)|^2 for one frame (atoms), or of the
 * mass-weighted molecular displacement (water); u NULL = the resident positions. */
NBX_API int nbx_rdf_reset(nbx_ctx *ctx, int maxbin);
NBX_API int nbx_rdf_add(nbx_ctx *ctx, const double *u);
NBX_API int nbx_rdf_get(nbx_ctx *ctx, int64_t *hist, int64_t cap, int64_t *frames);
NBX_API int nbx_msd(nbx_ctx *ctx, const double *u0, const double *u, double *out);
#ifdef __cplusplus
}
#endif
#endif /* NBODY_B200_H */
