"""Host-side mirror of NBodySimulator.jl's public interface for the acceleration hot path.

Julia is not available in this image, so the layer a user touches is restated in Python with the
reference's names, argument order and error behaviour; everything below the RHS call goes to
libnbody_b200.so through the C ABI (``_lib``).  There is no CPU fallback: constructing a problem or
running a simulation without the CUDA library raises.

Mirrors (file:line under /root/reference/src):
  bodies.jl:37-113            MassBody, ChargedParticle, MagneticParticle, WaterMolecule
  bodies.jl:139-159           generate_bodies_in_cell_nodes
  basic_potentials.jl:33-238  *Parameters structs and their defaults
  boundary_conditions.jl      InfiniteBox, PeriodicBoundaryConditions, CubicPeriodicBoundaryConditions
  thermostats.jl              Null/Andersen/Berendsen/NoseHoover/Langevin thermostats
  nbody_system.jl:50-191      ChargedParticles, GravitationalSystem, PotentialNBodySystem, WaterSPCFw
  nbody_simulation.jl:42-125  NBodySimulation and its convenience constructors, kb_SI
  nbody_to_ode.jl:1-75        gather_bodies_initial_coordinates
  nbody_to_ode.jl:156-242     get_accelerating_function (per-particle closure contract)
  nbody_to_ode.jl:460-536     SecondOrderODEProblem (the RHS soode_system!)
  nbody_simulation_result.jl  run_simulation, SimulationResult accessors, energies, temperature
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import _lib

kb_SI = 1.38e-23  # nbody_simulation.jl:7


# ---------------------------------------------------------------------------------------------
# bodies
# ---------------------------------------------------------------------------------------------
def _vec3(x):
    a = np.asarray(x, dtype=np.float64).reshape(-1)
    if a.shape != (3,):
        raise ValueError("expected a 3-vector")
    return a


@dataclass
class MassBody:
    r: np.ndarray
    v: np.ndarray
    m: float

    def __post_init__(self):
        self.r, self.v, self.m = _vec3(self.r), _vec3(self.v), float(self.m)


@dataclass
class ChargedParticle:
    r: np.ndarray
    v: np.ndarray
    m: float
    q: float

    def __post_init__(self):
        self.r, self.v, self.m, self.q = _vec3(self.r), _vec3(self.v), float(self.m), float(self.q)


@dataclass
class MagneticParticle:
    r: np.ndarray
    v: np.ndarray
    m: float
    mm: np.ndarray

    def __post_init__(self):
        self.r, self.v, self.m, self.mm = _vec3(self.r), _vec3(self.v), float(self.m), _vec3(self.mm)


@dataclass
class WaterMolecule:
    O: MassBody
    H1: MassBody
    H2: MassBody


def generate_bodies_in_cell_nodes(n: int, m: float, v_dev: float, L: float, rng=None):
    """bodies.jl:139-159: simple-cubic nodes (dL/2):dL:L, z fastest, velocities v_dev * randn(3).
    The reference seeds MersenneTwister(n); those exact normals cannot be reproduced without Julia, so
    a NumPy Philox(n) stream is used instead (statistically equivalent)."""
    from .workloads import cell_node_positions

    rng = np.random.Generator(np.random.Philox(n)) if rng is None else rng
    pos = cell_node_positions(n, L)
    vel = v_dev * rng.standard_normal((3, n))
    return [MassBody(pos[:, i], vel[:, i], m) for i in range(n)]


# ---------------------------------------------------------------------------------------------
# potential parameters (defaults as in basic_potentials.jl)
# ---------------------------------------------------------------------------------------------
class PotentialParameters:
    pass


@dataclass
class LennardJonesParameters(PotentialParameters):
    ϵ: float = 1.0
    σ: float = 1.0
    R: float = 2.5

    def __post_init__(self):
        self.σ2 = self.σ ** 2
        self.R2 = self.R ** 2

    def __str__(self):
        return f"Lennard-Jones:\n\tϵ:{self.ϵ}\n\tσ:{self.σ}\n\tR:{self.R}\n"


@dataclass
class GravitationalParameters(PotentialParameters):
    G: float = 6.67408e-11

    def __str__(self):
        return f"Gravitational:\n\tG:{self.G}\n"


@dataclass
class ElectrostaticParameters(PotentialParameters):
    k: float = 9e9
    R: float = math.inf

    def __post_init__(self):
        self.R2 = self.R ** 2

    def __str__(self):
        return f"Electrostatic:\n\tk:{self.k}\n"


@dataclass
class MagnetostaticParameters(PotentialParameters):
    μ_4π: float = 1e-7

    def __str__(self):
        return f"Magnetostatic:\n\tμ/4π:{self.μ_4π}\n"


@dataclass
class SPCFwParameters(PotentialParameters):
    rOH: float
    aHOH: float
    kb: float
    ka: float


# ---------------------------------------------------------------------------------------------
# boundary conditions
# ---------------------------------------------------------------------------------------------
class BoundaryConditions:
    pass


class InfiniteBox(BoundaryConditions):
    def __repr__(self):
        return "InfiniteBox()"


class PeriodicBoundaryConditions(BoundaryConditions):
    """boundary_conditions.jl:25-29: a 6-vector (xlo, xhi, ylo, yhi, zlo, zhi); L -> (0, L) x 3."""

    def __init__(self, *b):
        if len(b) == 1 and np.isscalar(b[0]):
            L = float(b[0])
            b = (0.0, L, 0.0, L, 0.0, L)
        elif len(b) == 1:
            b = tuple(float(x) for x in b[0])
        if len(b) != 6:
            raise ValueError("PeriodicBoundaryConditions takes L or six bounds")
        self.boundary = tuple(float(x) for x in b)

    def __getitem__(self, i):
        return self.boundary[i - 1]  # Julia indexing in the reference's code is 1-based


@dataclass
class CubicPeriodicBoundaryConditions(BoundaryConditions):
    L: float


# ---------------------------------------------------------------------------------------------
# thermostats
# ---------------------------------------------------------------------------------------------
class Thermostat:
    pass


class NullThermostat(Thermostat):
    pass


@dataclass
class AndersenThermostat(Thermostat):
    T: float
    ν: float


@dataclass
class BerendsenThermostat(Thermostat):
    T: float
    τ: float

    @property
    def γ(self):
        return 0.5 / self.τ  # thermostats.jl:72-74


@dataclass
class NoseHooverThermostat(Thermostat):
    T: float
    τ: float


@dataclass
class LangevinThermostat(Thermostat):
    T: float
    γ: float


# ---------------------------------------------------------------------------------------------
# systems
# ---------------------------------------------------------------------------------------------
class NBodySystem:
    pass


@dataclass
class GravitationalSystem(NBodySystem):
    bodies: Sequence[MassBody]
    G: float


@dataclass
class ChargedParticles(NBodySystem):
    bodies: Sequence[ChargedParticle]
    k: float


_DEFAULTS = {"lennard_jones": LennardJonesParameters, "electrostatic": ElectrostaticParameters,
             "gravitational": GravitationalParameters, "magnetostatic": MagnetostaticParameters}
_ORDER = ["lennard_jones", "electrostatic", "magnetostatic", "gravitational"]


class PotentialNBodySystem(NBodySystem):
    """nbody_system.jl:72-122.  ``potentials``: dict name -> parameters, or a list of built-in names."""

    def __init__(self, bodies, potentials=None):
        if isinstance(bodies, (GravitationalSystem, ChargedParticles, PotentialNBodySystem)):
            other = bodies
            if isinstance(other, PotentialNBodySystem):
                bodies, potentials = other.bodies, dict(other.potentials)
            elif isinstance(other, GravitationalSystem):
                bodies, potentials = other.bodies, {"gravitational": GravitationalParameters(other.G)}
            else:
                bodies, potentials = other.bodies, {"electrostatic": ElectrostaticParameters(other.k)}
        if potentials is None:
            potentials = {}
        if not isinstance(potentials, dict):
            potentials = {name: _DEFAULTS[name]() for name in _ORDER if name in potentials}
        for name, p in potentials.items():
            if not isinstance(p, PotentialParameters):
                raise TypeError(f"potential {name!r} is not a PotentialParameters")
        self.bodies = list(bodies)
        self.potentials: Dict[str, PotentialParameters] = dict(potentials)

    def __str__(self):
        return "Potentials: \n" + "".join(str(self.potentials[k]) for k in _ORDER if k in self.potentials)


@dataclass
class WaterSPCFw(NBodySystem):
    bodies: Sequence
    mH: float
    mO: float
    qH: float
    qO: float
    lj_parameters: LennardJonesParameters
    e_parameters: ElectrostaticParameters
    scpfw_parameters: SPCFwParameters


# ---------------------------------------------------------------------------------------------
# simulation
# ---------------------------------------------------------------------------------------------
class NBodySimulation:
    """nbody_simulation.jl:42-125: (system, tspan[, boundary_conditions[, thermostat]][, kb])."""

    def __init__(self, system, tspan, boundary_conditions=None, thermostat=None, kb=None):
        if isinstance(thermostat, (int, float)) and kb is None:  # (system, tspan, bc, kb) form
            thermostat, kb = None, float(thermostat)
        if isinstance(system, (GravitationalSystem, ChargedParticles)):
            system = PotentialNBodySystem(system)
        if not isinstance(system, NBodySystem):
            raise TypeError("system must be an NBodySystem")
        self.system = system
        self.tspan = (float(tspan[0]), float(tspan[1]))
        self.boundary_conditions = InfiniteBox() if boundary_conditions is None else boundary_conditions
        self.thermostat = NullThermostat() if thermostat is None else thermostat
        self.kb = kb_SI if kb is None else float(kb)


# ---------------------------------------------------------------------------------------------
# lowering: gathers (nbody_to_ode.jl:1-75, :290-351) and the device context
# ---------------------------------------------------------------------------------------------
def get_masses(system):
    """nbody_simulation_result.jl:121-139."""
    if isinstance(system, WaterSPCFw):
        return np.tile([system.mO, system.mH, system.mH], len(system.bodies)).astype(np.float64)
    return np.array([b.m for b in system.bodies], dtype=np.float64)


def get_degrees_of_freedom(system):
    """nbody_simulation_result.jl:152-166 -> (n, nc, ndf)."""
    if isinstance(system, WaterSPCFw):
        n, nc = 3 * len(system.bodies), 2 * len(system.bodies)
    else:
        n, nc = len(system.bodies), 0
    return n, nc, 3 * n - nc


def gather_bodies_initial_coordinates(simulation):
    """nbody_to_ode.jl:1-75 -> (u0, v0, n): 3 x len matrices, one extra column for Nose-Hoover."""
    system = simulation.system
    bodies = system.bodies
    n = len(bodies)
    extra = 1 if isinstance(simulation.thermostat, NoseHooverThermostat) else 0
    if isinstance(system, WaterSPCFw):
        length = 3 * n + extra
        u0 = np.zeros((3, length), order="F")
        v0 = np.zeros((3, length), order="F")
        p = system.scpfw_parameters
        for i, mol in enumerate(bodies):
            o = 3 * i
            if isinstance(mol, WaterMolecule):
                for k, atom in enumerate((mol.O, mol.H1, mol.H2)):
                    u0[:, o + k], v0[:, o + k] = atom.r, atom.v
            else:
                u0[:, o] = mol.r
                u0[:, o + 1] = mol.r + np.array([p.rOH, 0.0, 0.0])
                u0[:, o + 2] = mol.r + np.array([math.cos(p.aHOH) * p.rOH, 0.0, math.sin(p.aHOH) * p.rOH])
                v0[:, o] = v0[:, o + 1] = v0[:, o + 2] = mol.v
        return u0, v0, n
    u0 = np.zeros((3, n + extra), order="F")
    v0 = np.zeros((3, n + extra), order="F")
    for i, b in enumerate(bodies):
        u0[:, i], v0[:, i] = b.r, b.v
    return u0, v0, n


def _configure_context(simulation, device=0, only: Optional[str] = None) -> "_lib.Context":
    """Builds the nbx context the closures of one simulation share.  ``only``: restrict to one
    potential name (get_accelerating_function), else every potential + RHS thermostats."""
    system = simulation.system
    ctx = _lib.Context(device)
    ms = get_masses(system)
    if isinstance(system, WaterSPCFw):
        nmol = len(system.bodies)
        qs = np.tile([system.qO, system.qH, system.qH], nmol).astype(np.float64)
        ctx.system(ms, qs=qs, water=True)
    elif isinstance(system, PotentialNBodySystem):
        pots = system.potentials
        qs = mm = None
        if "electrostatic" in pots and (only in (None, "electrostatic")):
            try:
                qs = np.array([b.q for b in system.bodies], dtype=np.float64)
            except AttributeError as e:  # the reference fails reading `.q` (nbody_to_ode.jl:324)
                raise TypeError("electrostatic potential needs ChargedParticle bodies") from e
        if "magnetostatic" in pots and (only in (None, "magnetostatic")):
            if not all(isinstance(b, MagneticParticle) for b in system.bodies):
                raise TypeError("magnetostatic potential needs MagneticParticle bodies")  # basic_potentials.jl:338
            mm = np.asfortranarray(np.stack([b.mm for b in system.bodies], axis=1))
        ctx.system(ms, qs=qs, mm=mm)
    else:
        raise TypeError(f"no problem constructor accepts {type(system).__name__}")  # e.g. CustomAccelerationSystem

    bc = simulation.boundary_conditions
    if isinstance(bc, CubicPeriodicBoundaryConditions):
        ctx.boundary(_lib.BC_CUBIC, [bc.L])
    elif isinstance(bc, PeriodicBoundaryConditions):
        ctx.boundary(_lib.BC_PERIODIC, bc.boundary)
    else:
        ctx.boundary(_lib.BC_INFINITE)

    if isinstance(system, WaterSPCFw):
        lj, el, sp = system.lj_parameters, system.e_parameters, system.scpfw_parameters
        if only in (None, "lennard_jones"):
            ctx.add_lj(lj.ϵ, lj.σ, lj.R)
        if only in (None, "electrostatic"):
            ctx.add_coulomb(el.k, el.R)
        if only in (None, "spcfw"):
            ctx.add_spcfw(sp.rOH, sp.aHOH, sp.kb, sp.ka)
    else:
        for name, p in system.potentials.items():
            if only is not None and name != only:
                continue
            if name == "lennard_jones":
                ctx.add_lj(p.ϵ, p.σ, p.R)
            elif name == "electrostatic":
                ctx.add_coulomb(p.k, p.R)
            elif name == "magnetostatic":
                ctx.add_dipole(p.μ_4π)
            elif name == "gravitational":
                ctx.add_gravity(p.G)
            else:
                raise TypeError(f"potential {name!r} has no B200 kernel (custom potentials stay on the host)")

    if only is None:
        th = simulation.thermostat
        n, nc, _ = get_degrees_of_freedom(system)
        if isinstance(th, BerendsenThermostat):
            ctx.thermostat(_lib.THERMO_BERENDSEN, th.T, th.τ, simulation.kb, n, nc)
        elif isinstance(th, NoseHooverThermostat):
            ctx.thermostat(_lib.THERMO_NOSEHOOVER, th.T, th.τ, simulation.kb, n, nc)
        elif isinstance(th, AndersenThermostat):
            ctx.thermostat(_lib.THERMO_ANDERSEN, th.T, th.ν, simulation.kb, n, nc)
        elif isinstance(th, LangevinThermostat):
            ctx.thermostat(_lib.THERMO_LANGEVIN, th.T, th.γ, simulation.kb, n, nc)
        else:
            ctx.thermostat(_lib.THERMO_NONE, 0.0, 0.0, simulation.kb, n, nc)
    return ctx


_POTENTIAL_NAME = {LennardJonesParameters: "lennard_jones", ElectrostaticParameters: "electrostatic",
                   MagnetostaticParameters: "magnetostatic", GravitationalParameters: "gravitational",
                   SPCFwParameters: "spcfw"}


def get_accelerating_function(parameters: PotentialParameters, simulation: NBodySimulation, device=0):
    """nbody_to_ode.jl:156-242: returns ``acceleration!(dv, u, v, t, i)`` that ADDS particle i's
    acceleration (0-based ``i`` here) into the 3-vector ``dv``.  Per-particle granularity is hostile to a
    GPU, so the closure evaluates the whole system on the device when it sees a new ``u`` and serves the
    cached column afterwards."""
    name = _POTENTIAL_NAME.get(type(parameters))
    if name is None:
        raise TypeError("no B200 kernel for this PotentialParameters subtype; define its closure on the host "
                        "as in test/shared/custom_potential_body.jl")
    ctx = _configure_context(simulation, device, only=name)
    cache = {"key": None, "dv": None, "last_i": -1, "sum": None}
    n = ctx.n

    def acceleration(dv, u, v, t, i):
        # a sweep visits i in ascending order (nbody_to_ode.jl:475): an index that does not increase, another array,
        # another time or other CONTENTS (a buffer mutated in place, or a freed one reused at the same address: a cheap
        # checksum of the positions guards against serving stale columns) starts a new sweep -> one device evaluation
        # for all particles.  With the Nose-Hoover thermostat the reference passes n + 1 columns (nbody_to_ode.jl:6-8):
        # the potential closures only ever look at the first n.
        key = (u.ctypes.data, float(t), u.shape[1])
        if cache["key"] != key or i <= cache["last_i"] or cache["sum"] != float(u[:, :n].sum()):
            cache["dv"] = ctx.accel(u[:, :n] if u.shape[1] != n else u)
            cache["key"] = key
            cache["sum"] = float(u[:, :n].sum())
        cache["last_i"] = i
        dv += cache["dv"][:, i]
        return dv

    def invalidate():
        cache["key"] = None

    acceleration.context = ctx
    acceleration.invalidate = invalidate
    return acceleration


class SecondOrderODEProblem:
    """nbody_to_ode.jl:460-536.  ``f(dv, v, u, p, t)`` is soode_system!, evaluated on the GPU."""

    def __init__(self, simulation: NBodySimulation, device=0):
        self.simulation = simulation
        self.u0, self.v0, self.n = gather_bodies_initial_coordinates(simulation)
        self.tspan = simulation.tspan
        self.context = _configure_context(simulation, device)

    def f(self, dv, v, u, p=None, t=0.0):
        self.context.accel(u, v if self.context_needs_v else None, t, out=dv)
        return dv

    @property
    def context_needs_v(self):
        return isinstance(self.simulation.thermostat, (BerendsenThermostat, NoseHooverThermostat))


# ---------------------------------------------------------------------------------------------
# algorithms and run_simulation (nbody_simulation_result.jl:468-502)
# ---------------------------------------------------------------------------------------------
class VelocityVerlet:
    """Fixed-step symplectic scheme, fused on the device (nbx_step_vv)."""


class EM:
    """Euler-Maruyama for the Langevin SDE, fused on the device (nbx_step_em)."""


class Tsit5:
    """Adaptive explicit RK on the HOST driving the GPU RHS (the RHS drop-in mode).  scipy's RK45
    (Dormand-Prince 5(4)) stands in for Tsit5: same order, same call pattern into soode_system!."""


class SimulationResult:
    """nbody_simulation_result.jl:5-8: saved frames + the accessors of the reference."""

    def __init__(self, simulation, ts, us, vs, naccept):
        self.simulation = simulation
        self.t = np.asarray(ts, dtype=np.float64)
        self.u = us  # list of (3, ncols) position frames
        self.v = vs
        self.naccept = naccept

    def __str__(self):
        return f"N: {len(self.simulation.system.bodies)}\n{self.simulation.tspan}\nSteps: {self.naccept}\n"

    def _frame(self, frames, rates, time):
        t = self.t
        if time <= t[0]:
            return frames[0]
        if time >= t[-1]:
            return frames[-1]
        k = int(np.searchsorted(t, time, side="right") - 1)
        if t[k] == time:
            return frames[k]
        h = t[k + 1] - t[k]
        s = (time - t[k]) / h
        if rates is None:
            return (1 - s) * frames[k] + s * frames[k + 1]
        # cubic Hermite on (x, v)
        h00, h10 = 2 * s ** 3 - 3 * s ** 2 + 1, s ** 3 - 2 * s ** 2 + s
        h01, h11 = -2 * s ** 3 + 3 * s ** 2, s ** 3 - s ** 2
        return h00 * frames[k] + h10 * h * rates[k] + h01 * frames[k + 1] + h11 * h * rates[k + 1]


def _ncoord(system):
    return 3 * len(system.bodies) if isinstance(system, WaterSPCFw) else len(system.bodies)


def get_position(sr: SimulationResult, time: float, i: int = -1):
    """nbody_simulation_result.jl:90-119 (i < 0: all particles; else 0-based particle index)."""
    n = _ncoord(sr.simulation.system)
    x = sr._frame(sr.u, sr.v, time)[:, :n]
    return x if i < 0 else x[:, i]


def get_velocity(sr: SimulationResult, time: float, i: int = -1):
    n = _ncoord(sr.simulation.system)
    v = sr._frame(sr.v, None, time)[:, :n]
    return v if i < 0 else v[:, i]


def md_temperature(vs, ms, kb, N, Nc):
    """thermostats.jl:87-91."""
    return float(np.dot(ms, (vs ** 2).sum(axis=0)) / (kb * (3 * N - Nc)))


def temperature(sr: SimulationResult, time: float):
    n, nc, _ = get_degrees_of_freedom(sr.simulation.system)
    return md_temperature(get_velocity(sr, time), get_masses(sr.simulation.system), sr.simulation.kb, n, nc)


def kinetic_energy(a, b):
    """kinetic_energy(velocities, masses) or kinetic_energy(result, time) (:209-218)."""
    if isinstance(a, SimulationResult):
        return kinetic_energy(get_velocity(a, b), get_masses(a.simulation.system))
    return float(np.dot((np.asarray(a) ** 2).sum(axis=0), np.asarray(b) / 2))


def potential_energy(a, b, device=0):
    """potential_energy(coordinates, simulation) or potential_energy(result, time) (:239-291, :399-403):
    LJ + Coulomb (+ SPC/Fw bonded terms for water), evaluated on the device."""
    if isinstance(a, SimulationResult):
        return potential_energy(get_position(a, b), a.simulation, device)
    sim = b
    ctx = _configure_context(NBodySimulation(sim.system, sim.tspan, sim.boundary_conditions, NullThermostat(), sim.kb),
                             device)
    u = np.asfortranarray(a, dtype=np.float64)
    ctx.upload(u, np.zeros_like(u))
    _, ep, _ = ctx.energy()
    ctx.close()
    return ep


def total_energy(sr: SimulationResult, time: float):
    return kinetic_energy(sr, time) + potential_energy(sr, time)


def initial_energy(simulation: NBodySimulation):
    u0, v0, _ = gather_bodies_initial_coordinates(simulation)
    n = _ncoord(simulation.system)
    return potential_energy(u0[:, :n], simulation) + kinetic_energy(v0[:, :n], get_masses(simulation.system))


def rdf(sr: SimulationResult, device: int = 0):
    """rdf(sr) -> (rs, gr): nbody_simulation_result.jl:664-709.  The O(frames x N^2) pair loop runs on the device, one
    nbx_rdf_add per saved time (the histogram is integer: identical counts); the normalisation below is :695-707."""
    sim = sr.simulation
    if not isinstance(sim.boundary_conditions, CubicPeriodicBoundaryConditions):
        raise TypeError("rdf reads pbc.L: it needs CubicPeriodicBoundaryConditions")
    L = sim.boundary_conditions.L
    n = _ncoord(sim.system)
    indlen = len(sim.system.bodies)  # LJ index set: all bodies, or one oxygen per water molecule (:302-314)
    ctx = _configure_context(NBodySimulation(sim.system, sim.tspan, sim.boundary_conditions, NullThermostat(), sim.kb), device)
    maxbin = 1000
    ctx.rdf_reset(maxbin)
    for t in sr.t:
        ctx.rdf_add(np.asfortranarray(get_position(sr, t)[:, :n]))
    hist, tlen = ctx.rdf_get()
    ctx.close()
    dr = L / maxbin
    c = 4 / 3 * np.pi * indlen / L ** 3
    gr, rs = np.zeros(maxbin), np.zeros(maxbin)
    for b in range(maxbin):
        rlower = b * dr
        rupper = rlower + dr
        nideal = c * (rupper ** 3 - rlower ** 3)
        gr[b] = (hist[b] / (tlen * indlen)) / nideal
        rs[b] = rlower + dr / 2
    return rs, gr


def msd(sr: SimulationResult, device: int = 0):
    """msd(sr) -> (ts, dr2): nbody_simulation_result.jl:730-783 (atoms: mean |r(t) - r(0)|^2; water: of the mass-weighted
    molecular displacement), one device reduction per saved time."""
    sim = sr.simulation
    n = _ncoord(sim.system)
    ctx = _configure_context(NBodySimulation(sim.system, sim.tspan, sim.boundary_conditions, NullThermostat(), sim.kb), device)
    u0 = np.asfortranarray(get_position(sr, sr.t[0])[:, :n])
    dr2 = np.array([ctx.msd(u0, np.asfortranarray(get_position(sr, t)[:, :n])) for t in sr.t])
    ctx.close()
    return sr.t.copy(), dr2


# ---------------------------------------------------------------------------------------------
# trajectory output: PDB text (nbody_simulation_result.jl:786-1001).  Host-side formatting of saved frames, as in the
# reference; nothing here touches the device.
# ---------------------------------------------------------------------------------------------
def _julia_float(x: float) -> str:
    """Julia's string interpolation of a Float64 (shortest round-trip, '5.0e-5' rather than Python's '5e-05')."""
    r = repr(float(x))
    if "e" in r:
        mant, ex = r.split("e")
        if "." not in mant:
            mant += ".0"
        return f"{mant}e{int(ex)}"
    return r


def _hetatm(serial, name, res, resseq, xyz, element):
    # "HETATM", lpad(serial, 5), "  ", rpad(name, 4), res, lpad(resseq, 6), "    ", 3 x %8.3f, "  1.00", "  0.00", 10 blanks, lpad(element, 2)
    return ("HETATM" + str(serial).rjust(5) + "  " + name.ljust(4) + res + str(resseq).rjust(6) + "    " +
            "".join(("%8.3f" % c).rjust(8) for c in xyz) + "1.00".rjust(6) + "0.00".rjust(6) + " " * 10 + element.rjust(2))


def write_pdb_data(f, sr: SimulationResult):
    """write_pdb_data(io, result): :792-865.  Coordinates x 10 (nm -> Angstrom), wrapped into [0, 10 L) with
    x - L floor(x / L); water keeps every molecule whole (hydrogens placed relative to the wrapped oxygen)."""
    sim = sr.simulation
    n = len(sim.system.bodies)
    L = 10 * sim.boundary_conditions.L
    water = isinstance(sim.system, WaterSPCFw)
    for count, t in enumerate(sr.t, start=1):
        cc0 = 10 * get_position(sr, t)
        cc = cc0 - L * np.floor(cc0 / L)
        f.write("MODEL".ljust(10) + str(count) + "\n")
        if water:
            for i in range(n):
                o = 3 * i
                cc[:, o + 1] = cc[:, o] + cc0[:, o + 1] - cc0[:, o]
                cc[:, o + 2] = cc[:, o] + cc0[:, o + 2] - cc0[:, o]
            f.write(f"REMARK 250 time={_julia_float(t)} picoseconds\n")
            for i in range(n):
                o = 3 * i
                f.write(_hetatm(o + 1, "O", "HOH", i + 1, cc[:, o], "O") + "\n")
                f.write(_hetatm(o + 2, "H1", "HOH", i + 1, cc[:, o + 1], "H") + "\n")
                f.write(_hetatm(o + 3, "H2", "HOH", i + 1, cc[:, o + 2], "H") + "\n")
        else:
            f.write(f"REMARK 250 time={_julia_float(t)} steps\n")
            for i in range(n):
                f.write(_hetatm(i + 1, "Ar", "Ar", i + 1, cc[:, i], "Ar") + "\n")
        f.write("ENDMDL\n")


def save_to_pdb(sr: SimulationResult, path):
    """save_to_pdb(result, path): :881-885."""
    with open(path, "w") as f:
        write_pdb_data(f, sr)


def extract_from_pdb(file):
    """extract_from_pdb(io): :912-1001.  Reads the first two MODELs of a water trajectory; the molecules get the
    positions of the FIRST frame (Angstrom / 10) and zero velocities (the finite-difference velocities the reference
    computes are overwritten by zeros at :992-994), masses 15.999 / 1.00794."""
    lines = iter(file.read().split("\n")) if hasattr(file, "read") else iter(file)

    def frame():
        for ln in lines:
            if ln[:5] == "MODEL":
                break
        else:
            return None, None
        parts = next(lines).split()
        t = float(parts[2][5:])
        cc = []
        for ln in lines:
            if ln[:6] == "ENDMDL":
                break
            if len(ln) > 13 and ln[13] == "O":
                rows = [ln, next(lines), next(lines)]
                for row in rows:
                    ps = row.split()
                    cc.append(np.array([float(ps[5]), float(ps[6]), float(ps[7])]) / 10)
        return t, cc

    _, cc1 = frame()
    _, cc2 = frame()
    if cc1 is None or cc2 is None:
        raise ValueError("a PDB trajectory with at least two MODEL records is required")
    zero = np.zeros(3)
    wms = []
    for m in range(len(cc2) // 3):
        wms.append(WaterMolecule(MassBody(cc1[3 * m], zero, 15.999), MassBody(cc1[3 * m + 1], zero, 1.00794),
                                 MassBody(cc1[3 * m + 2], zero, 1.00794)))
    return wms


def load_water_molecules_from_pdb(path):
    """load_water_molecules_from_pdb(path): :906-910."""
    with open(path) as f:
        return extract_from_pdb(f)


def run_simulation(s: NBodySimulation, alg=None, *, dt: Optional[float] = None, saveat: Optional[int] = None,
                   save_everystep: Optional[bool] = None, device: int = 0, seed: int = 0, rtol=1e-6, atol=1e-9):
    """nbody_simulation_result.jl:468-492.  Langevin thermostats run the SDE path (EM), everything else
    the second-order ODE path; Andersen forces save_everystep = false exactly as the reference does."""
    alg = Tsit5() if alg is None else alg
    langevin = isinstance(s.thermostat, LangevinThermostat)
    if langevin and not isinstance(alg, EM):
        raise TypeError("a LangevinThermostat simulation is an SDEProblem: use EM()")
    if isinstance(alg, EM) and not langevin:
        raise TypeError("EM() needs a LangevinThermostat")
    prob = SecondOrderODEProblem(s, device)
    ctx = prob.context
    t0, t1 = s.tspan
    if isinstance(alg, Tsit5):
        if isinstance(s.thermostat, AndersenThermostat):
            # the reference attaches the collision DiscreteCallback to ANY algorithm (nbody_simulation_result.jl:482-485);
            # the host-side adaptive path here only evaluates the RHS, so the thermostat would be dropped silently
            raise TypeError("an AndersenThermostat simulation needs a fixed-step algorithm here (VelocityVerlet(), dt=...): the "
                            "collisions are applied by the device stepper after every step")
        return _run_host_adaptive(s, prob, rtol, atol)
    if dt is None:
        raise ValueError("fixed-step algorithms need dt")
    nsteps = int(round((t1 - t0) / dt))
    if save_everystep is None:
        save_everystep = not isinstance(s.thermostat, AndersenThermostat)
    stride = 1 if saveat is None else max(1, int(saveat))
    if not save_everystep:
        stride = nsteps
    ctx.set_seed(seed or 0x9E3779B97F4A7C15)
    ctx.upload(prob.u0, prob.v0)
    ts, us, vs = [t0], [prob.u0.copy(order="F")], [prob.v0.copy(order="F")]
    done = 0
    while done < nsteps:
        k = min(stride, nsteps - done)
        if isinstance(alg, EM):
            ctx.step_em(dt, k)
        else:
            ctx.step_vv(dt, k)
        done += k
        u, v, _ = ctx.download()
        ts.append(t0 + done * dt)
        us.append(u)
        vs.append(v)
    return SimulationResult(s, ts, us, vs, nsteps)


def _run_host_adaptive(s, prob, rtol, atol):
    from scipy.integrate import solve_ivp

    ncols = prob.u0.shape[1]
    size = 3 * ncols
    dv = np.empty((3, ncols), order="F")

    def rhs(t, y):
        u = np.asfortranarray(y[size:].reshape((3, ncols), order="F"))
        v = np.asfortranarray(y[:size].reshape((3, ncols), order="F"))
        prob.f(dv, v, u, None, t)
        # ArrayPartition(v, u): d/dt = (dv, v)   (nbody_to_ode.jl:490)
        return np.concatenate([dv.ravel(order="F"), v.ravel(order="F")])

    y0 = np.concatenate([prob.v0.ravel(order="F"), prob.u0.ravel(order="F")])
    sol = solve_ivp(rhs, s.tspan, y0, method="RK45", rtol=rtol, atol=atol, dense_output=False)
    us = [np.asfortranarray(sol.y[size:, k].reshape((3, ncols), order="F")) for k in range(sol.y.shape[1])]
    vs = [np.asfortranarray(sol.y[:size, k].reshape((3, ncols), order="F")) for k in range(sol.y.shape[1])]
    return SimulationResult(s, sol.t, us, vs, int(sol.t.shape[0] - 1))
