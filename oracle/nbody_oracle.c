/*
 * nbody_oracle.c -- CPU restatement of the NBodySimulator.jl acceleration hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under nbodysimulator.jl_b200/ may link, import or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, as the checker / the reported CPU baseline.
 *
 * Parity status: the reference (pure Julia) cannot run in this image and ships no golden
 * force vectors, so force-level parity is "UNPINNED" in the strict sense: it rests on this
 * line-by-line restatement (same operation order, no FMA contraction: build with
 * -O2 -ffp-contract=off -fno-fast-math), on closed-form checks, on an independent NumPy
 * restatement (oracle/nbody_oracle_np.py) and on the known-answer scenarios of the reference's
 * own test-suite (tests/test_oracle_kat.py lists each with file:line).
 *
 * All arrays are Julia-layout: 3 x n column-major doubles (x1 y1 z1 x2 y2 z2 ...).
 * Indices are 0-based here; the reference is 1-based.
 *
 * Citations are relative to /root/reference/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

typedef struct {
    int64_t n;          /* number of coordinate columns that carry particles            */
    int64_t ncols;      /* columns of u/v/dv (n, or n+1 with Nose-Hoover)                */
    const double *ms;   /* [n] masses                                                    */
    const double *qs;   /* [n] charges or NULL                                           */
    const double *mm;   /* [3 x n] magnetic moments or NULL                              */
    int32_t water;      /* 1: columns are (O,H1,H2) triples (src/nbody_to_ode.jl:46)     */
    int32_t bc_kind;    /* 0 InfiniteBox, 1 CubicPeriodic (bc[0]=L), 2 Periodic (6 vals) */
    double bc[6];
    int32_t has_gravity;
    double G;
    int32_t has_lj;
    double lj_eps, lj_sigma2, lj_R2;
    int32_t has_coulomb;
    double el_k, el_R2;
    int32_t has_dipole;
    double mu_4pi;
    int32_t has_spcfw;
    double rOH, aHOH, k_bond, k_angle;
    int32_t thermostat; /* 0 none, 1 Berendsen, 2 Nose-Hoover                            */
    double T0, tparam;  /* Berendsen: tparam = gamma = 0.5/tau; Nose-Hoover: tparam=tau  */
    double kB;
    int64_t N, Nc;      /* thermostat particle / constraint counts                       */
} orc_system;

/* ---------------------------------------------------------------------------------------
 * get_interparticle_distance -- src/boundary_conditions.jl:111-136 (Periodic, 6-vector,
 * wraps rij into [lo,hi): NOT a minimum image, reference quirk kept), :138-165 (Cubic,
 * wraps into [-L/2, L/2)), :167-172 (generic / InfiniteBox).
 * r2 = x^2 + y^2 + z^2 evaluated left to right, un-fused.
 * ------------------------------------------------------------------------------------- */
static inline void orc_distance_impl(const double *ri, const double *rj, int bc_kind,
                                     const double *bc, double *rij, double *r, double *r2)
{
    double x = ri[0] - rj[0], y = ri[1] - rj[1], z = ri[2] - rj[2];
    if (bc_kind == 1) {
        const double size = bc[0];
        const double radius = 0.5 * size;
        while (x >= radius) x -= size;
        while (x < -radius) x += size;
        while (y >= radius) y -= size;
        while (y < -radius) y += size;
        while (z >= radius) z -= size;
        while (z < -radius) z += size;
    } else if (bc_kind == 2) {
        while (x < bc[0]) x += bc[1] - bc[0];
        while (x >= bc[1]) x -= bc[1] - bc[0];
        while (y < bc[2]) y += bc[3] - bc[2];
        while (y >= bc[3]) y -= bc[3] - bc[2];
        while (z < bc[4]) z += bc[5] - bc[4];
        while (z >= bc[5]) z -= bc[5] - bc[4];
    }
    rij[0] = x; rij[1] = y; rij[2] = z;
    const double s = x * x + y * y + z * z;
    *r2 = s;
    *r = sqrt(s);
}

ORC_API void orc_distance(const double *ri, const double *rj, int bc_kind, const double *bc,
                          double *rij, double *r, double *r2)
{
    orc_distance_impl(ri, rj, bc_kind, bc, rij, r, r2);
}

/* StaticArrays norm(SVector{3}) = sqrt(x^2+y^2+z^2) [upstream StaticArrays, unverified] */
static inline double norm3(const double *a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
static inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(const double *a, const double *b, double *c)
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

/* gravitational_acceleration! -- src/basic_potentials.jl:306-331 */
static void gravity_i(double *dv, const double *rs, int64_t i, int64_t n, const double *ms, double G)
{
    double a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const double *ri = rs + 3 * i;
    for (int64_t j = 0; j < n; ++j) {
        if (j != i) {
            const double *rj = rs + 3 * j;
            double rij[3] = {ri[0] - rj[0], ri[1] - rj[1], ri[2] - rj[2]};
            const double nr = norm3(rij);
            const double factor = (-G * ms[j]) / (nr * nr * nr); /* :321, ^3 -> x*x*x */
            a1 += factor * rij[0];
            a2 += factor * rij[1];
            a3 += factor * rij[2];
        }
    }
    dv[0] += a1; dv[1] += a2; dv[2] += a3;
}

/* pairwise_lennard_jones_acceleration! -- src/basic_potentials.jl:240-272.
 * idx_kind 0: indxs = all columns (src/nbody_to_ode.jl:290-300); 1: every third column
 * starting at 0 = oxygen sites of water (:302-314). */
static void lj_i(double *dv, const double *rs, int64_t i, int64_t n, int idx_stride,
                 const double *ms, double eps, double sigma2, double R2, int bc_kind, const double *bc)
{
    double f1 = 0.0, f2 = 0.0, f3 = 0.0;
    const double *ri = rs + 3 * i;
    for (int64_t j = 0; j < n; j += idx_stride) {
        if (j != i) {
            double rij[3], r, r2;
            orc_distance_impl(ri, rs + 3 * j, bc_kind, bc, rij, &r, &r2);
            if (r2 < R2) {
                const double q = sigma2 / r2;
                const double s6 = q * q * q;
                const double s12 = s6 * s6;
                const double factor = (2 * s12 - s6) / r2;
                f1 += factor * rij[0];
                f2 += factor * rij[1];
                f3 += factor * rij[2];
            }
        }
    }
    const double coeff = 24 * eps / ms[i];
    dv[0] += coeff * f1; dv[1] += coeff * f2; dv[2] += coeff * f3;
}

/* pairwise_electrostatic_acceleration! -- src/basic_potentials.jl:274-304.
 * Exclusions: {i} (src/nbody_to_ode.jl:316-329) or the three atoms of the own molecule
 * (:331-351; the 3n+1 Nose-Hoover slot is never reached because j runs to n). */
static void coulomb_i(double *dv, const double *rs, int64_t i, int64_t n, const double *qs,
                      const double *ms, int water, double k, double R2, int bc_kind, const double *bc)
{
    double f1 = 0.0, f2 = 0.0, f3 = 0.0;
    const double *ri = rs + 3 * i;
    const int64_t ex_lo = water ? 3 * (i / 3) : i;
    const int64_t ex_hi = water ? ex_lo + 3 : i + 1;
    for (int64_t j = 0; j < n; ++j) {
        if (j < ex_lo || j >= ex_hi) {
            double rij[3], r, r2;
            orc_distance_impl(ri, rs + 3 * j, bc_kind, bc, rij, &r, &r2);
            if (r2 < R2) {
                const double factor = qs[j] / (r * r2);
                f1 += factor * rij[0];
                f2 += factor * rij[1];
                f3 += factor * rij[2];
            }
        }
    }
    const double coeff = k * qs[i] / ms[i];
    dv[0] += coeff * f1; dv[1] += coeff * f2; dv[2] += coeff * f3;
}

/* magnetostatic_dipdip_acceleration! -- src/basic_potentials.jl:333-365 */
static void dipole_i(double *dv, const double *rs, int64_t i, int64_t n, const double *ms,
                     const double *mm, double mu_4pi)
{
    double f1 = 0.0, f2 = 0.0, f3 = 0.0;
    const double *mi = mm + 3 * i;
    const double *ri = rs + 3 * i;
    for (int64_t j = 0; j < n; ++j) {
        if (j != i) {
            const double *mj = mm + 3 * j;
            const double *rj = rs + 3 * j;
            double rij[3] = {ri[0] - rj[0], ri[1] - rj[1], ri[2] - rj[2]};
            const double d = dot3(rij, rij);
            const double rij4 = d * d;
            const double nr = norm3(rij);
            double r[3] = {rij[0] / nr, rij[1] / nr, rij[2] / nr};
            const double mir = dot3(mi, r);
            const double mij = dot3(mj, r);
            const double mimj = dot3(mi, mj);
            double c[3];
            for (int k = 0; k < 3; ++k)
                c[k] = (((mi[k] * mij + mj[k] * mir) + r[k] * mimj) - ((5 * r[k]) * mir) * mij) / rij4;
            f1 += c[0]; f2 += c[1]; f3 += c[2];
        }
    }
    const double coeff = 3 * mu_4pi / ms[i];
    dv[0] += coeff * f1; dv[1] += coeff * f2; dv[2] += coeff * f3;
}

/* harmonic_bond_potential_acceleration! -- src/basic_potentials.jl:367-393; partner table
 * src/nbody_to_ode.jl:263-288: O -> (H1,H2), H1 -> (O), H2 -> (O); no minimum image. */
static void bond_i(double *dv, const double *rs, int64_t i, const double *ms, double rOH, double kb)
{
    double f1 = 0.0, f2 = 0.0, f3 = 0.0;
    const double *ri = rs + 3 * i;
    const int64_t o = 3 * (i / 3);
    int64_t partners[2];
    int np;
    if (i == o) { partners[0] = o + 1; partners[1] = o + 2; np = 2; }
    else { partners[0] = o; np = 1; }
    for (int p = 0; p < np; ++p) {
        const double *rj = rs + 3 * partners[p];
        double rij[3] = {ri[0] - rj[0], ri[1] - rj[1], ri[2] - rj[2]};
        const double r = norm3(rij);
        const double d = r - rOH;
        const double factor = -d * kb / r;
        f1 += factor * rij[0];
        f2 += factor * rij[1];
        f3 += factor * rij[2];
    }
    const double coeff = 1.0 / ms[i];
    dv[0] += coeff * f1; dv[1] += coeff * f2; dv[2] += coeff * f3;
}

/* valence_angle_potential_acceleration! -- src/basic_potentials.jl:395-433
 * (a=H1, b=O, c=H2; src/nbody_to_ode.jl:255-260).  normalize(v) = inv(norm(v))*v
 * [upstream StaticArrays, unverified]. Adds into three columns of the full dv. */
static void angle_abc_to(double *da, double *db, double *dc, const double *rs, int64_t a, int64_t b, int64_t c,
                         const double *ms, double ka, double aHOH0);

static void angle_abc(double *dv, const double *rs, int64_t a, int64_t b, int64_t c,
                      const double *ms, double ka, double aHOH0)
{
    angle_abc_to(dv + 3 * a, dv + 3 * b, dv + 3 * c, rs, a, b, c, ms, ka, aHOH0);
}

/* same arithmetic, the three columns it adds into given explicitly (for target subsets) */
static void angle_abc_to(double *da, double *db, double *dc, const double *rs, int64_t a, int64_t b, int64_t c,
                         const double *ms, double ka, double aHOH0)
{
    const double *ra = rs + 3 * a, *rb = rs + 3 * b, *rc = rs + 3 * c;
    double rba[3], rbc[3], rcb[3], X[3], pa[3], pc[3];
    for (int k = 0; k < 3; ++k) { rba[k] = ra[k] - rb[k]; rbc[k] = rc[k] - rb[k]; rcb[k] = rb[k] - rc[k]; }
    cross3(rba, rbc, X);
    cross3(rba, X, pa);
    cross3(rcb, X, pc);
    const double ipa = 1.0 / norm3(pa), ipc = 1.0 / norm3(pc);
    for (int k = 0; k < 3; ++k) { pa[k] = ipa * pa[k]; pc[k] = ipc * pc[k]; }
    const double nba = norm3(rba), nbc = norm3(rbc);
    double cosine = dot3(rba, rbc) / (nba * nbc);
    if (cosine > 1) cosine = 1; else if (cosine < -1) cosine = -1;
    const double theta = acos(cosine);
    const double force = -ka * (theta - aHOH0);
    for (int k = 0; k < 3; ++k) {
        const double fa = pa[k] * force / nba;
        const double fc = pc[k] * force / nbc;
        const double fb = -(fa + fc);
        da[k] += fa / ms[a];
        db[k] += fb / ms[b];
        dc[k] += fc / ms[c];
    }
}

/* md_temperature -- src/thermostats.jl:87-91 (summation order of the BLAS dot is not
 * reproducible; plain left-to-right here). */
static double md_temperature(const double *vs, const double *ms, double kb, int64_t N, int64_t Nc, int64_t ncols)
{
    double e = 0.0;
    for (int64_t i = 0; i < ncols; ++i) {
        const double *v = vs + 3 * i;
        e += ms[i] * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    }
    return e / (kb * (double)(3 * N - Nc));
}

ORC_API double orc_md_temperature(const double *vs, const double *ms, double kb, int64_t N, int64_t Nc, int64_t ncols)
{
    return md_temperature(vs, ms, kb, N, Nc, ncols);
}

/* berendsen_acceleration! -- src/thermostats.jl:76-83 */
static void berendsen(double *dv, const double *v, const double *ms, double kb, int64_t N,
                      int64_t Nc, int64_t ncols, double T0, double gamma)
{
    const double T = md_temperature(v, ms, kb, N, Nc, ncols);
    if (1.0 / T == INFINITY) {
        for (int64_t k = 0; k < 3 * ncols; ++k) dv[k] += gamma * v[k];
    } else {
        const double s = gamma * (T0 / T - 1);
        for (int64_t k = 0; k < 3 * ncols; ++k) dv[k] += s * v[k];
    }
}

/* nosehoover_acceleration! -- src/thermostats.jl:121-128.  u,v,dv have N+1 columns; zeta is
 * element (1, N+1) = linear index 3N.  MUTATES v (the reference does). */
static void nosehoover(double *dv, const double *u, double *v, const double *ms, double kb,
                       int64_t N, int64_t Nc, double T0, double tau)
{
    const int64_t zind = 3 * N;
    const double zeta = u[zind];
    for (int64_t k = 0; k < 3 * (N + 1); ++k) dv[k] -= zeta * v[k];
    dv[3 * N] = 0; dv[3 * N + 1] = 0; dv[3 * N + 2] = 0;
    const double T = md_temperature(v, ms, kb, N, Nc, N);
    const double ndf = (double)(3 * N - Nc);
    const double it = 1.0 / tau;
    v[zind] = (it * it) * (T / T0 - (ndf + 1) / ndf);
}

/* Per-particle potential terms for column i, in a fixed order.  Ordinary systems: the
 * reference iterates a Dict{Symbol,...} (src/nbody_to_ode.jl:94) whose order is a hash
 * artefact; we use lennard_jones, electrostatic, magnetostatic, gravitational.  Water:
 * O -> [electrostatic, bond, LJ], H -> [electrostatic, bond] (src/nbody_to_ode.jl:538-565). */
static void accel_column(const orc_system *s, const double *u, int64_t i, double *a)
{
    a[0] = a[1] = a[2] = 0.0;
    if (s->water) {
        const int is_o = (i % 3) == 0;
        if (s->has_coulomb) coulomb_i(a, u, i, s->n, s->qs, s->ms, 1, s->el_k, s->el_R2, s->bc_kind, s->bc);
        if (s->has_spcfw) bond_i(a, u, i, s->ms, s->rOH, s->k_bond);
        if (is_o && s->has_lj) lj_i(a, u, i, s->n, 3, s->ms, s->lj_eps, s->lj_sigma2, s->lj_R2, s->bc_kind, s->bc);
    } else {
        if (s->has_lj) lj_i(a, u, i, s->n, 1, s->ms, s->lj_eps, s->lj_sigma2, s->lj_R2, s->bc_kind, s->bc);
        if (s->has_coulomb) coulomb_i(a, u, i, s->n, s->qs, s->ms, 0, s->el_k, s->el_R2, s->bc_kind, s->bc);
        if (s->has_dipole) dipole_i(a, u, i, s->n, s->ms, s->mm, s->mu_4pi);
        if (s->has_gravity) gravity_i(a, u, i, s->n, s->ms, s->G);
    }
}

/* Accelerations of a list of target columns (per-particle potentials only: no angle term,
 * no thermostat).  out is 3 x nt.  Targets are independent, so the loop may be threaded;
 * the arithmetic per target is exactly the serial reference's. */
ORC_API void orc_accel_targets(const orc_system *s, const double *u, const int64_t *targets,
                               int64_t nt, double *out, int nthreads)
{
#ifdef _OPENMP
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
#endif
    for (int64_t t = 0; t < nt; ++t) accel_column(s, u, targets[t], out + 3 * t);
    (void)nthreads;
}

/* Water: full accelerations (pair terms, bonds AND the per-molecule angle term, src/nbody_to_ode.jl:502-532)
 * of a list of whole molecules.  out is 3 x (3 nm): columns O, H1, H2 of each listed molecule. */
ORC_API void orc_accel_molecules(const orc_system *s, const double *u, const int64_t *mols, int64_t nm,
                                 double *out, int nthreads)
{
#ifdef _OPENMP
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 2) num_threads(nthreads)
#endif
    for (int64_t t = 0; t < nm; ++t) {
        const int64_t o = 3 * mols[t];
        double *a = out + 9 * t;
        for (int k = 0; k < 3; ++k) accel_column(s, u, o + k, a + 3 * k);
        if (s->water && s->has_spcfw)
            angle_abc_to(a + 3, a, a + 6, u, o + 1, o, o + 2, s->ms, s->k_angle, s->aHOH);
    }
    (void)nthreads;
}

/* soode_system!(dv, v, u, p, t) -- src/nbody_to_ode.jl:474-488 (PotentialNBodySystem) and
 * :502-532 (WaterSPCFw).  v is non-const because Nose-Hoover writes v[zind]. */
ORC_API void orc_rhs(const orc_system *s, const double *u, double *v, double *dv, int nthreads)
{
    const int64_t n = s->n;
#ifdef _OPENMP
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
#endif
    for (int64_t i = 0; i < n; ++i) accel_column(s, u, i, dv + 3 * i);
    (void)nthreads;
    if (s->water && s->has_spcfw) {
        for (int64_t m = 0; m < n / 3; ++m)
            angle_abc(dv, u, 3 * m + 1, 3 * m, 3 * m + 2, s->ms, s->k_angle, s->aHOH);
    }
    if (s->thermostat == 1)
        berendsen(dv, v, s->ms, s->kB, s->N, s->Nc, s->ncols, s->T0, s->tparam);
    else if (s->thermostat == 2)
        nosehoover(dv, u, v, s->ms, s->kB, s->N, s->Nc, s->T0, s->tparam);
}

/* ---- extended-precision gravity (long double accumulation and arithmetic) used only to
 * judge both the restatement and the GPU result where net accelerations nearly cancel. ---- */
ORC_API void orc_gravity_targets_ld(const double *rs, const double *ms, int64_t n, double G,
                                    const int64_t *targets, int64_t nt, double *out, int nthreads)
{
#ifdef _OPENMP
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
#endif
    for (int64_t t = 0; t < nt; ++t) {
        const int64_t i = targets[t];
        long double a1 = 0, a2 = 0, a3 = 0;
        for (int64_t j = 0; j < n; ++j) {
            if (j == i) continue;
            const long double x = (long double)rs[3 * i] - rs[3 * j];
            const long double y = (long double)rs[3 * i + 1] - rs[3 * j + 1];
            const long double z = (long double)rs[3 * i + 2] - rs[3 * j + 2];
            const long double nr = sqrtl(x * x + y * y + z * z);
            const long double f = (-(long double)G * ms[j]) / (nr * nr * nr);
            a1 += f * x; a2 += f * y; a3 += f * z;
        }
        out[3 * t] = (double)a1; out[3 * t + 1] = (double)a2; out[3 * t + 2] = (double)a3;
    }
    (void)nthreads;
}

/* ---- extended-precision referee for ALL pair potentials of a system (long double arithmetic and accumulation).
 * The PAIR SET is the reference's: the cutoff predicate is evaluated exactly as above (fp64, un-fused); only the
 * force arithmetic of the accepted pairs is carried in long double, from the SAME wrapped displacement.  Used by the
 * tests to judge both the fp64 restatement and the GPU result where a body's net acceleration nearly cancels
 * (alternating charges on a lattice, r^-14 terms of both signs): neither side is allowed more than 1e-12 / twice the
 * restatement's own distance from the referee.  Not a restatement of the reference: test infrastructure only. ---- */
ORC_API void orc_accel_targets_ld(const orc_system *s, const double *rs, const int64_t *targets, int64_t nt, double *out,
                                  int nthreads)
{
#ifdef _OPENMP
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
#endif
    for (int64_t t = 0; t < nt; ++t) {
        const int64_t i = targets[t], n = s->n;
        const double *ri = rs + 3 * i;
        long double tot[3] = {0, 0, 0};
        if (s->has_lj && (!s->water || i % 3 == 0)) {
            long double f[3] = {0, 0, 0};
            const int stride = s->water ? 3 : 1;
            for (int64_t j = 0; j < n; j += stride) {
                if (j == i) continue;
                double rij[3], r, r2;
                orc_distance_impl(ri, rs + 3 * j, s->bc_kind, s->bc, rij, &r, &r2);
                if (r2 < s->lj_R2) {
                    const long double x = rij[0], y = rij[1], z = rij[2];
                    const long double d2 = x * x + y * y + z * z;
                    const long double q = (long double)s->lj_sigma2 / d2;
                    const long double s6 = q * q * q;
                    const long double fac = (2 * s6 * s6 - s6) / d2;
                    f[0] += fac * x; f[1] += fac * y; f[2] += fac * z;
                }
            }
            const long double c = 24 * (long double)s->lj_eps / s->ms[i];
            for (int k = 0; k < 3; ++k) tot[k] += c * f[k];
        }
        if (s->has_coulomb) {
            long double f[3] = {0, 0, 0};
            const int64_t ex_lo = s->water ? 3 * (i / 3) : i;
            const int64_t ex_hi = s->water ? ex_lo + 3 : i + 1;
            for (int64_t j = 0; j < n; ++j) {
                if (j >= ex_lo && j < ex_hi) continue;
                double rij[3], r, r2;
                orc_distance_impl(ri, rs + 3 * j, s->bc_kind, s->bc, rij, &r, &r2);
                if (r2 < s->el_R2) {
                    const long double x = rij[0], y = rij[1], z = rij[2];
                    const long double d2 = x * x + y * y + z * z;
                    const long double fac = (long double)s->qs[j] / (sqrtl(d2) * d2);
                    f[0] += fac * x; f[1] += fac * y; f[2] += fac * z;
                }
            }
            const long double c = (long double)s->el_k * s->qs[i] / s->ms[i];
            for (int k = 0; k < 3; ++k) tot[k] += c * f[k];
        }
        if (s->has_dipole) {
            long double f[3] = {0, 0, 0};
            const double *mi = s->mm + 3 * i;
            for (int64_t j = 0; j < n; ++j) {
                if (j == i) continue;
                const double *mj = s->mm + 3 * j, *rj = rs + 3 * j;
                const long double x[3] = {(long double)ri[0] - rj[0], (long double)ri[1] - rj[1], (long double)ri[2] - rj[2]};
                const long double d2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
                const long double nr = sqrtl(d2);
                const long double e[3] = {x[0] / nr, x[1] / nr, x[2] / nr};
                const long double mir = mi[0] * e[0] + mi[1] * e[1] + mi[2] * e[2];
                const long double mjr = mj[0] * e[0] + mj[1] * e[1] + mj[2] * e[2];
                const long double mimj = (long double)mi[0] * mj[0] + (long double)mi[1] * mj[1] + (long double)mi[2] * mj[2];
                for (int k = 0; k < 3; ++k) f[k] += (mi[k] * mjr + mj[k] * mir + e[k] * mimj - 5 * e[k] * mir * mjr) / (d2 * d2);
            }
            const long double c = 3 * (long double)s->mu_4pi / s->ms[i];
            for (int k = 0; k < 3; ++k) tot[k] += c * f[k];
        }
        if (s->has_gravity) {
            long double f[3] = {0, 0, 0};
            for (int64_t j = 0; j < n; ++j) {
                if (j == i) continue;
                const double *rj = rs + 3 * j;
                const long double x = (long double)ri[0] - rj[0], y = (long double)ri[1] - rj[1], z = (long double)ri[2] - rj[2];
                const long double nr = sqrtl(x * x + y * y + z * z);
                const long double fac = (-(long double)s->G * s->ms[j]) / (nr * nr * nr);
                f[0] += fac * x; f[1] += fac * y; f[2] += fac * z;
            }
            for (int k = 0; k < 3; ++k) tot[k] += f[k];
        }
        for (int k = 0; k < 3; ++k) out[3 * t + k] = (double)tot[k];
    }
    (void)nthreads;
}

/* In-cutoff neighbour predicate of the reference: for target i, the ordered list of j
 * (ascending) with j != i (stride as LJ index set) and r2 < R2, r2 from
 * get_interparticle_distance in un-fused fp64.  Returns the count; writes at most cap. */
ORC_API int64_t orc_neighbors_i(const double *rs, int64_t i, int64_t n, int idx_stride, double R2,
                                int bc_kind, const double *bc, int32_t *list, int64_t cap)
{
    int64_t c = 0;
    for (int64_t j = 0; j < n; j += idx_stride) {
        if (j == i) continue;
        double rij[3], r, r2;
        orc_distance_impl(rs + 3 * i, rs + 3 * j, bc_kind, bc, rij, &r, &r2);
        if (r2 < R2) { if (c < cap) list[c] = (int32_t)j; ++c; }
    }
    return c;
}

/* ------------------------------- energies ----------------------------------------------
 * kinetic_energy src/nbody_simulation_result.jl:209-212; lennard_jones_potential :293-319
 * (r2 clamped to R2 outside the cutoff); electrostatic_potential :321-351;
 * harmonic_bonds_potential :353-372 (each bond visited from both ends, /4);
 * valence_angle_harmonic_potential :374-397. */
ORC_API double orc_kinetic_energy(const double *vs, const double *ms, int64_t n)
{
    double e = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        const double *v = vs + 3 * i;
        e += (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) * (ms[i] / 2);
    }
    return e;
}

/* ---- analysis of saved frames (src/nbody_simulation_result.jl:664-783) -------------------------------
 * rdf, one frame (:676-693): pairs i < j over the Lennard-Jones index set (every idx_stride-th column: all atoms, or
 * the oxygens of water), distance through get_interparticle_distance of the cubic box, `if r2 < (0.5 L)^2:
 * bin = ceil(r / dr); if 1 < bin <= maxbin: hist[bin] += 2` with dr = L / maxbin.  hist is 0-based here:
 * hist[b - 1] is the reference's hist[b].  The normalisation (:695-707) stays with the caller. */
ORC_API void orc_rdf_hist(const double *rs, int64_t n, int idx_stride, double L, int maxbin, int64_t *hist)
{
    const double bc[6] = {L, 0, 0, 0, 0, 0};
    const double dr = L / maxbin;
    const double lim = (0.5 * L) * (0.5 * L);
    for (int64_t i = 0; i < n; i += idx_stride)
        for (int64_t j = i + idx_stride; j < n; j += idx_stride) {
            double rij[3], r, r2;
            orc_distance_impl(rs + 3 * i, rs + 3 * j, 1, bc, rij, &r, &r2);
            if (r2 < lim) {
                const double b = ceil(r / dr);
                if (b > 1 && b <= maxbin) hist[(int64_t)b - 1] += 2;
            }
        }
}

/* msd, one frame: atoms (:730-752) mean over the index set of |r(t) - r(0)|^2; water (:754-783) the same for the
 * mass-weighted centre of each molecule ((dO mO + dH1 mH + dH2 mH) / (2 mH + mO)). */
ORC_API double orc_msd(const double *rs, const double *rs0, int64_t n, int water, double mO, double mH)
{
    double s = 0.0;
    if (!water) {
        for (int64_t i = 0; i < n; ++i) {
            const double d0 = rs[3 * i] - rs0[3 * i], d1 = rs[3 * i + 1] - rs0[3 * i + 1], d2 = rs[3 * i + 2] - rs0[3 * i + 2];
            s += d0 * d0 + d1 * d1 + d2 * d2;
        }
        return s / (double)n;
    }
    const int64_t nm = n / 3;
    for (int64_t m = 0; m < nm; ++m) {
        double d[3];
        for (int k = 0; k < 3; ++k) {
            const double dO = rs[9 * m + k] - rs0[9 * m + k], dH1 = rs[9 * m + 3 + k] - rs0[9 * m + 3 + k],
                         dH2 = rs[9 * m + 6 + k] - rs0[9 * m + 6 + k];
            d[k] = (dO * mO + dH1 * mH + dH2 * mH) / (2 * mH + mO);
        }
        s += d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    }
    return s / (double)nm;
}

ORC_API double orc_lj_potential(const double *rs, int64_t n, int idx_stride, double eps, double sigma2,
                                double R2, int bc_kind, const double *bc)
{
    double e = 0.0;
    for (int64_t i = 0; i < n; i += idx_stride)
        for (int64_t j = i + idx_stride; j < n; j += idx_stride) {
            double rij[3], r, r2;
            orc_distance_impl(rs + 3 * i, rs + 3 * j, bc_kind, bc, rij, &r, &r2);
            if (!(r2 < R2)) r2 = R2;
            const double q = sigma2 / r2;
            const double s6 = q * q * q;
            const double s12 = s6 * s6;
            e += (s12 - s6);
        }
    return 4 * eps * e;
}

ORC_API double orc_coulomb_potential(const double *rs, int64_t n, const double *qs, int water, double k,
                                     double R, double R2, int bc_kind, const double *bc)
{
    double e = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        double ei = 0.0;
        const int64_t ex_hi = water ? 3 * (i / 3) + 3 : i + 1;
        for (int64_t j = i + 1; j < n; ++j) {
            if (j < ex_hi) continue;
            double rij[3], r, r2;
            orc_distance_impl(rs + 3 * i, rs + 3 * j, bc_kind, bc, rij, &r, &r2);
            if (r2 < R2) ei += qs[j] / r; else ei += qs[j] / R;
        }
        e += ei * qs[i];
    }
    return e * k;
}

ORC_API double orc_bond_potential(const double *rs, int64_t nmol, double rOH, double kb)
{
    double e = 0.0;
    for (int64_t m = 0; m < nmol; ++m) {
        const int64_t o = 3 * m;
        const int64_t pairs[4][2] = {{o, o + 1}, {o, o + 2}, {o + 1, o}, {o + 2, o}};
        for (int p = 0; p < 4; ++p) {
            const double *ri = rs + 3 * pairs[p][0], *rj = rs + 3 * pairs[p][1];
            double rij[3] = {ri[0] - rj[0], ri[1] - rj[1], ri[2] - rj[2]};
            const double d = norm3(rij) - rOH;
            e += d * d * kb;
        }
    }
    return e / 4;
}

ORC_API double orc_angle_potential(const double *rs, int64_t nmol, double aHOH0, double ka)
{
    double e = 0.0;
    for (int64_t m = 0; m < nmol; ++m) {
        const double *ra = rs + 3 * (3 * m + 1), *rb = rs + 3 * (3 * m), *rc = rs + 3 * (3 * m + 2);
        double rba[3], rbc[3];
        for (int k = 0; k < 3; ++k) { rba[k] = ra[k] - rb[k]; rbc[k] = rc[k] - rb[k]; }
        const double ang = acos(dot3(rba, rbc) / (norm3(rba) * norm3(rbc)));
        const double d = ang - aHOH0;
        e += ka * (d * d);
    }
    return e / 2;
}

ORC_API int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
