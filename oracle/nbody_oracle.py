"""ctypes binding of oracle/nbody_oracle.c plus textbook time-steppers.

TEST INFRASTRUCTURE ONLY (see the header of nbody_oracle.c).  The product package
`nbodysimulator.jl_b200` never imports this module.

Arrays follow the reference's layout: ``u``/``v``/``dv`` are float64 ``(3, n)`` arrays in
Fortran order (identical bytes to Julia's ``Matrix{Float64}``), masses etc. are ``(n,)``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnbody_oracle.so")

BC_INFINITE, BC_CUBIC, BC_PERIODIC = 0, 1, 2
THERMO_NONE, THERMO_BERENDSEN, THERMO_NOSEHOOVER = 0, 1, 2


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, -ffp-contract=off).  Returns the library path."""
    src = os.path.join(_HERE, "nbody_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class OrcSystem(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("ncols", C.c_int64),
        ("ms", C.POINTER(C.c_double)), ("qs", C.POINTER(C.c_double)), ("mm", C.POINTER(C.c_double)),
        ("water", C.c_int32), ("bc_kind", C.c_int32), ("bc", C.c_double * 6),
        ("has_gravity", C.c_int32), ("G", C.c_double),
        ("has_lj", C.c_int32), ("lj_eps", C.c_double), ("lj_sigma2", C.c_double), ("lj_R2", C.c_double),
        ("has_coulomb", C.c_int32), ("el_k", C.c_double), ("el_R2", C.c_double),
        ("has_dipole", C.c_int32), ("mu_4pi", C.c_double),
        ("has_spcfw", C.c_int32), ("rOH", C.c_double), ("aHOH", C.c_double),
        ("k_bond", C.c_double), ("k_angle", C.c_double),
        ("thermostat", C.c_int32), ("T0", C.c_double), ("tparam", C.c_double), ("kB", C.c_double),
        ("N", C.c_int64), ("Nc", C.c_int64),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip64, ip32 = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int32)
        sp = C.POINTER(OrcSystem)
        L.orc_accel_targets.argtypes = [sp, dp, ip64, C.c_int64, dp, C.c_int]
        L.orc_accel_targets.restype = None
        L.orc_accel_molecules.argtypes = [sp, dp, ip64, C.c_int64, dp, C.c_int]
        L.orc_accel_molecules.restype = None
        L.orc_rhs.argtypes = [sp, dp, dp, dp, C.c_int]
        L.orc_rhs.restype = None
        L.orc_gravity_targets_ld.argtypes = [dp, dp, C.c_int64, C.c_double, ip64, C.c_int64, dp, C.c_int]
        L.orc_gravity_targets_ld.restype = None
        L.orc_accel_targets_ld.argtypes = [C.POINTER(OrcSystem), dp, ip64, C.c_int64, dp, C.c_int]
        L.orc_accel_targets_ld.restype = None
        L.orc_neighbors_i.argtypes = [dp, C.c_int64, C.c_int64, C.c_int, C.c_double, C.c_int, dp, ip32, C.c_int64]
        L.orc_neighbors_i.restype = C.c_int64
        L.orc_distance.argtypes = [dp, dp, C.c_int, dp, dp, dp, dp]
        L.orc_distance.restype = None
        L.orc_md_temperature.argtypes = [dp, dp, C.c_double, C.c_int64, C.c_int64, C.c_int64]
        L.orc_md_temperature.restype = C.c_double
        L.orc_kinetic_energy.argtypes = [dp, dp, C.c_int64]
        L.orc_kinetic_energy.restype = C.c_double
        L.orc_lj_potential.argtypes = [dp, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, dp]
        L.orc_lj_potential.restype = C.c_double
        L.orc_coulomb_potential.argtypes = [dp, C.c_int64, dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, dp]
        L.orc_coulomb_potential.restype = C.c_double
        L.orc_bond_potential.argtypes = [dp, C.c_int64, C.c_double, C.c_double]
        L.orc_bond_potential.restype = C.c_double
        L.orc_angle_potential.argtypes = [dp, C.c_int64, C.c_double, C.c_double]
        L.orc_angle_potential.restype = C.c_double
        L.orc_rdf_hist.argtypes = [dp, C.c_int64, C.c_int, C.c_double, C.c_int, ip64]
        L.orc_rdf_hist.restype = None
        L.orc_msd.argtypes = [dp, dp, C.c_int64, C.c_int, C.c_double, C.c_double]
        L.orc_msd.restype = C.c_double
        L.orc_max_threads.argtypes = []
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f(a):
    a = np.asarray(a, dtype=np.float64)
    return a if a.flags.f_contiguous else np.asfortranarray(a)


def max_threads() -> int:
    return int(lib().orc_max_threads())


class System:
    """Plain description of one simulation's hot-path data (what the closures capture)."""

    def __init__(self, ms, *, qs=None, mm=None, water=False, bc=("infinite",), gravity=None,
                 lj=None, coulomb=None, dipole=None, spcfw=None, thermostat=None):
        self.ms = np.ascontiguousarray(ms, dtype=np.float64)
        self.n = int(self.ms.shape[0])
        self.qs = None if qs is None else np.ascontiguousarray(qs, dtype=np.float64)
        self.mm = None if mm is None else _f(mm)
        self.water = bool(water)
        self.bc = bc
        self.gravity, self.lj, self.coulomb, self.dipole, self.spcfw = gravity, lj, coulomb, dipole, spcfw
        self.thermostat = thermostat
        s = OrcSystem()
        s.n = self.n
        s.ncols = self.n
        s.ms = _dp(self.ms)
        s.qs = _dp(self.qs) if self.qs is not None else None
        s.mm = _dp(self.mm) if self.mm is not None else None
        s.water = int(self.water)
        kind = bc[0]
        if kind == "infinite":
            s.bc_kind = BC_INFINITE
        elif kind == "cubic":
            s.bc_kind = BC_CUBIC
            s.bc[0] = float(bc[1])
        elif kind == "periodic":
            s.bc_kind = BC_PERIODIC
            for k in range(6):
                s.bc[k] = float(bc[1][k])
        else:
            raise ValueError(kind)
        if gravity is not None:
            s.has_gravity, s.G = 1, float(gravity["G"])
        if lj is not None:
            # LennardJonesParameters(eps, sigma, R) caches sigma^2 and R^2 (basic_potentials.jl:67-69)
            s.has_lj, s.lj_eps = 1, float(lj["eps"])
            s.lj_sigma2 = float(lj["sigma"]) ** 2
            s.lj_R2 = float(lj["R"]) ** 2
        if coulomb is not None:
            s.has_coulomb, s.el_k = 1, float(coulomb["k"])
            R = float(coulomb.get("R", np.inf))
            s.el_R2 = R * R
        if dipole is not None:
            s.has_dipole, s.mu_4pi = 1, float(dipole["mu_4pi"])
        if spcfw is not None:
            s.has_spcfw = 1
            s.rOH, s.aHOH = float(spcfw["rOH"]), float(spcfw["aHOH"])
            s.k_bond, s.k_angle = float(spcfw["kb"]), float(spcfw["ka"])
        if thermostat is not None:
            kind = thermostat["kind"]
            s.kB = float(thermostat["kB"])
            s.T0 = float(thermostat["T"])
            s.N = int(thermostat.get("N", self.n))
            s.Nc = int(thermostat.get("Nc", 0))
            if kind == "berendsen":
                s.thermostat = THERMO_BERENDSEN
                s.tparam = 0.5 / float(thermostat["tau"])  # thermostats.jl:72-74
            elif kind == "nosehoover":
                s.thermostat = THERMO_NOSEHOOVER
                s.tparam = float(thermostat["tau"])
                s.ncols = self.n + 1
            else:
                raise ValueError(kind)
        self.c = s

    # -- force-level entry points --------------------------------------------------------
    def accel_targets(self, u, targets, nthreads=1):
        u = _f(u)
        t = np.ascontiguousarray(targets, dtype=np.int64)
        out = np.zeros((3, t.shape[0]), order="F")
        lib().orc_accel_targets(C.byref(self.c), _dp(u), t.ctypes.data_as(C.POINTER(C.c_int64)),
                                t.shape[0], _dp(out), int(nthreads))
        return out

    def accel_targets_ld(self, u, targets, nthreads=1):
        """Extended-precision referee (long double) of the pair potentials for the listed targets: same pair set as the
        reference predicate, force arithmetic and sums in long double.  Bonded terms and thermostats are not included."""
        u = _f(u)
        t = np.ascontiguousarray(targets, dtype=np.int64)
        out = np.zeros((3, t.shape[0]), order="F")
        lib().orc_accel_targets_ld(C.byref(self.c), _dp(u), t.ctypes.data_as(C.POINTER(C.c_int64)),
                                   t.shape[0], _dp(out), int(nthreads))
        return out

    def accel_molecules(self, u, mols, nthreads=1):
        """Water: full accelerations (incl. the angle term) of whole molecules; 3 x (3 len(mols))."""
        u = _f(u)
        t = np.ascontiguousarray(mols, dtype=np.int64)
        out = np.zeros((3, 3 * t.shape[0]), order="F")
        lib().orc_accel_molecules(C.byref(self.c), _dp(u), t.ctypes.data_as(C.POINTER(C.c_int64)), t.shape[0], _dp(out),
                                  int(nthreads))
        return out

    def rhs(self, u, v, nthreads=1):
        """Full soode_system!(dv, v, u, p, t).  Returns dv; ``v`` is mutated for Nose-Hoover."""
        u = _f(u)
        assert v.flags.f_contiguous and v.dtype == np.float64
        dv = np.zeros((3, self.c.ncols), order="F")
        lib().orc_rhs(C.byref(self.c), _dp(u), _dp(v), _dp(dv), int(nthreads))
        return dv

    def neighbors(self, u, i, R, idx_stride=1, cap=4096):
        u = _f(u)
        lst = np.zeros(cap, dtype=np.int32)
        bc = np.array(list(self.c.bc), dtype=np.float64)
        c = lib().orc_neighbors_i(_dp(u), int(i), self.n, int(idx_stride), float(R) ** 2, self.c.bc_kind,
                                  _dp(bc), lst.ctypes.data_as(C.POINTER(C.c_int32)), cap)
        assert c <= cap
        return lst[:c].copy()

    # -- energies (src/nbody_simulation_result.jl:209-397) -------------------------------
    def kinetic_energy(self, v):
        v = _f(v)
        return float(lib().orc_kinetic_energy(_dp(v), _dp(self.ms), self.n))

    def potential_energy(self, u):
        u = _f(u)
        bc = np.array(list(self.c.bc), dtype=np.float64)
        e = 0.0
        L = lib()
        if self.lj is not None:
            e += L.orc_lj_potential(_dp(u), self.n, 3 if self.water else 1, self.c.lj_eps, self.c.lj_sigma2,
                                    self.c.lj_R2, self.c.bc_kind, _dp(bc))
        if self.coulomb is not None:
            R = float(self.coulomb.get("R", np.inf))
            e += L.orc_coulomb_potential(_dp(u), self.n, _dp(self.qs), int(self.water), self.c.el_k, R,
                                         self.c.el_R2, self.c.bc_kind, _dp(bc))
        if self.water and self.spcfw is not None:
            e += L.orc_bond_potential(_dp(u), self.n // 3, self.c.rOH, self.c.k_bond)
            e += L.orc_angle_potential(_dp(u), self.n // 3, self.c.aHOH, self.c.k_angle)
        return float(e)

    def temperature(self, v, kB, N=None, Nc=0):
        v = _f(v)
        N = self.n if N is None else N
        return float(lib().orc_md_temperature(_dp(v), _dp(self.ms), float(kB), int(N), int(Nc), self.n))


def rdf_hist(u, L, idx_stride=1, maxbin=1000):
    """Pair-distance histogram of ONE frame exactly as rdf's inner loops build it
    (src/nbody_simulation_result.jl:676-693); hist[b - 1] is the reference's hist[b]."""
    u = _f(u)
    hist = np.zeros(maxbin, dtype=np.int64)
    lib().orc_rdf_hist(_dp(u), u.shape[1], int(idx_stride), float(L), int(maxbin), hist.ctypes.data_as(C.POINTER(C.c_int64)))
    return hist


def rdf_normalise(hist, nframes, nidx, L):
    """(rs, gr) from the accumulated histogram: src/nbody_simulation_result.jl:695-707."""
    maxbin = len(hist)
    dr = L / maxbin
    c = 4 / 3 * np.pi * nidx / L ** 3
    rlower = np.arange(maxbin) * dr
    rupper = rlower + dr
    nideal = c * (rupper ** 3 - rlower ** 3)
    return rlower + dr / 2, (hist / (nframes * nidx)) / nideal


def msd(u, u0, water=False, mO=0.0, mH=0.0):
    """Mean squared displacement of ONE frame against the first: src/nbody_simulation_result.jl:730-783."""
    u, u0 = _f(u), _f(u0)
    return float(lib().orc_msd(_dp(u), _dp(u0), u.shape[1], int(bool(water)), float(mO), float(mH)))


def gravity_targets_ld(u, ms, G, targets, nthreads=1):
    u = _f(u)
    ms = np.ascontiguousarray(ms, dtype=np.float64)
    t = np.ascontiguousarray(targets, dtype=np.int64)
    out = np.zeros((3, t.shape[0]), order="F")
    lib().orc_gravity_targets_ld(_dp(u), _dp(ms), ms.shape[0], float(G),
                                 t.ctypes.data_as(C.POINTER(C.c_int64)), t.shape[0], _dp(out), int(nthreads))
    return out


def distance(ri, rj, bc_kind, bc):
    ri = np.ascontiguousarray(ri, dtype=np.float64)
    rj = np.ascontiguousarray(rj, dtype=np.float64)
    b = np.zeros(6)
    b[: len(bc)] = bc
    rij = np.zeros(3)
    r = C.c_double()
    r2 = C.c_double()
    lib().orc_distance(_dp(ri), _dp(rj), int(bc_kind), _dp(b), _dp(rij), C.byref(r), C.byref(r2))
    return rij, r.value, r2.value


# ---------------------------------------------------------------------------------------
# Time steppers.  The reference owns no integrator (it hands closures to DiffEq); these are
# the textbook forms of the upstream schemes [upstream OrdinaryDiffEq/StochasticDiffEq,
# unverified] described in SURVEY.md Appendix A.  Step-level trajectories: parity unpinned.
# ---------------------------------------------------------------------------------------
def velocity_verlet(sys: System, u0, v0, dt, nsteps, nthreads=1, callback=None):
    """x+ = x + dt v + dt^2/2 a;  a+ = f(v, x+);  v+ = v + dt/2 (a + a+)."""
    u = np.array(u0, dtype=np.float64, order="F")
    v = np.array(v0, dtype=np.float64, order="F")
    a = sys.rhs(u, v, nthreads)
    for k in range(nsteps):
        u = u + dt * v + (0.5 * dt * dt) * a
        a_new = sys.rhs(u, v, nthreads)
        v = np.asfortranarray(v + (0.5 * dt) * (a + a_new))
        a = a_new
        if callback is not None:
            callback(k + 1, u, v)
    return u, v


def euler_maruyama(sys: System, u0, v0, dt, nsteps, gamma, sigma, rng, nthreads=1):
    """EM on the atomic SDE of src/nbody_to_ode.jl:567-598: du = v dt; dv = (a - gamma v) dt + sigma dW."""
    u = np.array(u0, dtype=np.float64, order="F")
    v = np.array(v0, dtype=np.float64, order="F")
    sq = np.sqrt(dt)
    for _ in range(nsteps):
        a = sys.rhs(u, v, nthreads)
        dW = sq * rng.standard_normal(v.shape)
        u_new = u + dt * v
        v = np.asfortranarray(v + dt * (a - gamma * v) + sigma * dW)
        u = np.asfortranarray(u_new)
    return u, v


def euler_maruyama_water(sys: System, u0, v0, dt, nsteps, gamma, kT, mO, mH, rng, nthreads=1):
    """EM on the SDE of WaterSPCFw, src/nbody_to_ode.jl:600-680, as written there: drift = a - gamma v for every column
    (:664), the oxygen columns additionally - (gamma v) / mO (:627-629; the terms subtracted from the hydrogen columns at
    :630-633 are zeroed again by the hydrogen loop :636-641); noise sqrt(2 gamma kb T) / mO and / mH (:668-676)."""
    u = np.array(u0, dtype=np.float64, order="F")
    v = np.array(v0, dtype=np.float64, order="F")
    sq = np.sqrt(dt)
    root = np.sqrt(2.0 * gamma * kT)
    sig = np.tile([root / mO, root / mH, root / mH], u.shape[1] // 3)
    for _ in range(nsteps):
        a = sys.rhs(u, v, nthreads)
        drift = a.copy()
        drift[:, 0::3] -= gamma * v[:, 0::3] / mO
        drift -= gamma * v
        dW = sq * rng.standard_normal(v.shape)
        u_new = u + dt * v
        v = np.asfortranarray(v + dt * drift + sig[None, :] * dW)
        u = np.asfortranarray(u_new)
    return u, v

