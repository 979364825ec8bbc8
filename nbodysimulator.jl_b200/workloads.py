"""Seeded synthetic inputs for the BASELINE.json configs.

NumPy only; produces Julia-layout arrays: positions/velocities are float64 ``(3, n)`` in
Fortran order.  Every generator is deterministic in its arguments (Philox counter RNG), so
the same bytes can be regenerated on the GPU box, and dumped as raw little-endian fp64 for
a Julia run elsewhere (``dump_raw``).
"""
from __future__ import annotations

import math

import numpy as np

KB_SI = 1.38e-23  # src/nbody_simulation.jl:7


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(seed))


def cell_node_positions(n: int, L: float) -> np.ndarray:
    """Positions of generate_bodies_in_cell_nodes (src/bodies.jl:139-159): simple-cubic
    nodes (dL/2):dL:L with dL = L/ceil(cbrt(n)), z fastest, truncated to n."""
    k = int(math.ceil(n ** (1.0 / 3.0) - 1e-12))
    dL = L / k
    # Julia range (dL/2):dL:L has floor((L - dL/2)/dL) + 1 = k elements
    ax = dL / 2 + dL * np.arange(k)
    x, y, z = np.meshgrid(ax, ax, ax, indexing="ij")
    pos = np.stack([x.ravel(), y.ravel(), z.ravel()])[:, :n]
    return np.asfortranarray(pos)


def plummer(n: int, seed: int | None = None, a: float = 1.0):
    """Config 2: equal-mass Plummer sphere, G = 1, total mass 1, scale ``a``; radii beyond
    50a rejected; Aarseth-Henon-Wielen velocity sampling; centre of mass removed."""
    rng = _rng(n if seed is None else seed)
    r = np.empty(0)
    while r.shape[0] < n:
        U = rng.random(n)
        rr = a / np.sqrt(U ** (-2.0 / 3.0) - 1.0)
        r = np.concatenate([r, rr[rr <= 50.0 * a]])
    r = r[:n]

    def iso(m):
        ct = 2.0 * rng.random(m) - 1.0
        ph = 2.0 * math.pi * rng.random(m)
        st = np.sqrt(1.0 - ct * ct)
        return np.stack([st * np.cos(ph), st * np.sin(ph), ct])

    pos = iso(n) * r
    q = np.empty(0)
    while q.shape[0] < n:
        x = rng.random(2 * n)
        y = 0.1 * rng.random(2 * n)
        ok = y < x * x * (1.0 - x * x) ** 3.5
        q = np.concatenate([q, x[ok]])
    q = q[:n]
    vesc = math.sqrt(2.0) * (r * r + a * a) ** (-0.25)
    vel = iso(n) * (q * vesc)
    pos -= pos.mean(axis=1, keepdims=True)
    vel -= vel.mean(axis=1, keepdims=True)
    ms = np.full(n, 1.0 / n)
    return np.asfortranarray(pos), np.asfortranarray(vel), ms


def liquid_argon_si(n: int = 216, seed: int | None = None):
    """Config 1: examples/liquid_argon.jl:34-57 as shipped (SI units, R = 0.5 L, no thermostat)."""
    T = 120.0
    kb = KB_SI
    eps = T * kb
    sigma = 3.4e-10
    rho = 1374.0
    m = 39.95 * 1.6747 * 1e-27
    L = (m * n / rho) ** (1.0 / 3.0)
    R = 0.5 * L
    v_dev = math.sqrt(kb * T / m)
    tau = 0.5e-3 * 1e-12
    pos = cell_node_positions(n, L)
    vel = np.asfortranarray(v_dev * _rng(n if seed is None else seed).standard_normal((3, n)))
    return dict(u=pos, v=vel, ms=np.full(n, m), L=L, lj=dict(eps=eps, sigma=sigma, R=R), dt=tau, kB=kb, T=T)


def fcc_argon_reduced(cells: int = 64, seed: int | None = None, R: float = 2.25):
    """Config 3: reduced-unit argon (examples/liquid_argon_reduced.jl:11-32 scaled):
    sigma = eps = m = 1, kb = 1/120, rho* = 0.80718, FCC lattice cells^3 x 4 atoms,
    velocities N(0,1), R = 2.25 sigma, dt = 2.3136e-4."""
    n = 4 * cells ** 3
    rho = 0.80718
    L = (n / rho) ** (1.0 / 3.0)
    a = L / cells
    base = np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.0], [0.5, 0.0, 0.5], [0.0, 0.5, 0.5]]) + 0.25
    ax = np.arange(cells, dtype=np.float64)
    cx, cy, cz = np.meshgrid(ax, ax, ax, indexing="ij")
    corner = np.stack([cx.ravel(), cy.ravel(), cz.ravel()], axis=1)  # (cells^3, 3)
    pos = (corner[:, None, :] + base[None, :, :]).reshape(-1, 3) * a
    rng = _rng(n if seed is None else seed)
    vel = rng.standard_normal((3, n))
    vel -= vel.mean(axis=1, keepdims=True)
    return dict(u=np.asfortranarray(pos.T), v=np.asfortranarray(vel), ms=np.ones(n), L=L,
                lj=dict(eps=1.0, sigma=1.0, R=R), dt=2.3136e-4, kB=1.0 / 120.0, T=90.0)


def water_omm(side: int = 32, seed: int | None = None, Rel: float | None = None):
    """Config 4: SPC/Fw water, OMM units (examples/water_spc_fw_omm_units.jl:3-33), oxygen
    sites on a side^3 simple-cubic lattice, molecule geometry per src/nbody_to_ode.jl:47-49.
    ``Rel`` None -> 0.49 L (the example's rule)."""
    nmol = side ** 3
    kb = 8.3144598e-3
    T = 370.0
    mO, mH = 15.999, 1.00794
    rho = 997.0 / 1.6747  # kg/m^3 -> amu/nm^3 as in the example
    mH2O = mO + 2 * mH
    L = (mH2O * nmol / rho) ** (1.0 / 3.0)
    rOH = 0.1012
    aHOH = 113.24 * math.pi / 180.0
    opos = cell_node_positions(nmol, L)
    v_dev = math.sqrt(kb * T / mH2O)
    vmol = v_dev * _rng(nmol if seed is None else seed).standard_normal((3, nmol))
    n = 3 * nmol
    u = np.zeros((3, n), order="F")
    v = np.zeros((3, n), order="F")
    u[:, 0::3] = opos
    u[:, 1::3] = opos + np.array([[rOH], [0.0], [0.0]])
    u[:, 2::3] = opos + np.array([[math.cos(aHOH) * rOH], [0.0], [math.sin(aHOH) * rOH]])
    v[:, 0::3] = vmol
    v[:, 1::3] = vmol
    v[:, 2::3] = vmol
    ms = np.tile([mO, mH, mH], nmol).astype(np.float64)
    qs = np.tile([-0.82, 0.41, 0.41], nmol).astype(np.float64)
    return dict(u=u, v=v, ms=ms, qs=qs, L=L, nmol=nmol,
                lj=dict(eps=0.1554253 * 4.184, sigma=0.3165492, R=0.9),
                coulomb=dict(k=138.935458, R=(0.49 * L if Rel is None else Rel)),
                spcfw=dict(rOH=rOH, aHOH=aHOH, kb=1059.162 * 4.184 * 1e2, ka=75.9 * 4.184),
                dt=0.5e-4, kB=kb, T=T)


def charged_lattice(n: int = 65536, seed: int | None = None):
    """Config 5a: ChargedParticle x n on a jittered simple-cubic lattice, charges +-1e-3
    alternating (neutral), Coulomb R = inf, InfiniteBox."""
    k = int(math.ceil(n ** (1.0 / 3.0) - 1e-12))
    L = float(k)
    pos = cell_node_positions(n, L)
    rng = _rng(n if seed is None else seed)
    pos = pos + 0.05 * (2.0 * rng.random((3, n)) - 1.0)
    qs = np.where(np.arange(n) % 2 == 0, 1e-3, -1e-3)
    vel = np.zeros((3, n), order="F")
    return dict(u=np.asfortranarray(pos), v=vel, ms=np.ones(n), qs=qs, coulomb=dict(k=9e9))


def dipole_lattice(n: int = 65536, seed: int | None = None):
    """Config 5b: MagneticParticle x n, moments along z as test/magnetostaic_test.jl:7-13
    (iron: M*m/rho), jittered lattice, mu/4pi = 1e-7."""
    k = int(math.ceil(n ** (1.0 / 3.0) - 1e-12))
    d = 0.1
    L = d * k
    pos = cell_node_positions(n, L)
    rng = _rng((n if seed is None else seed) + 1)
    pos = pos + 0.05 * d * (2.0 * rng.random((3, n)) - 1.0)
    m = 5e-6
    mm_mag = 1.7e6 * m / 7800.0  # saturation magnetisation x volume
    mm = np.zeros((3, n), order="F")
    mm[2, :] = mm_mag
    # tilt a little so that all terms of the dipole kernel are exercised
    mm[0, :] = 0.1 * mm_mag * (2.0 * rng.random(n) - 1.0)
    mm[1, :] = 0.1 * mm_mag * (2.0 * rng.random(n) - 1.0)
    return dict(u=np.asfortranarray(pos), v=np.zeros((3, n), order="F"), ms=np.full(n, m), mm=mm,
                dipole=dict(mu_4pi=1e-7))


def dump_raw(path: str, **arrays) -> None:
    """Write arrays as raw little-endian fp64 (Fortran order) for consumption from Julia:
    ``reshape(reinterpret(Float64, read(path_u)), 3, n)``."""
    for name, arr in arrays.items():
        np.asfortranarray(arr, dtype="<f8").ravel(order="F").tofile(f"{path}.{name}.f64")
