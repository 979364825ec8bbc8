"""Independent pure-Python restatement of the reference kernels (small cases only).

TEST INFRASTRUCTURE ONLY.  Written separately from nbody_oracle.c, directly from the Julia
source, with Python floats (IEEE binary64, never FMA-contracted) and explicit loops, so it
must agree with the C restatement BIT FOR BIT; tests/test_oracle_kat.py checks that.

Citations are relative to /root/reference/.  Indices are 0-based.
"""
from __future__ import annotations

import math

import numpy as np


def distance(ri, rj, bc):
    """get_interparticle_distance, src/boundary_conditions.jl:111-172."""
    x, y, z = ri[0] - rj[0], ri[1] - rj[1], ri[2] - rj[2]
    kind = bc[0]
    if kind == "cubic":  # :138-165
        size = bc[1]
        radius = 0.5 * size
        while x >= radius:
            x -= size
        while x < -radius:
            x += size
        while y >= radius:
            y -= size
        while y < -radius:
            y += size
        while z >= radius:
            z -= size
        while z < -radius:
            z += size
    elif kind == "periodic":  # :111-136
        b = bc[1]
        while x < b[0]:
            x += b[1] - b[0]
        while x >= b[1]:
            x -= b[1] - b[0]
        while y < b[2]:
            y += b[3] - b[2]
        while y >= b[3]:
            y -= b[3] - b[2]
        while z < b[4]:
            z += b[5] - b[4]
        while z >= b[5]:
            z -= b[5] - b[4]
    r2 = x * x + y * y + z * z
    return (x, y, z), math.sqrt(r2), r2


def _col(a, i):
    return (float(a[0, i]), float(a[1, i]), float(a[2, i]))


def _norm(a):
    return math.sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2])


def _dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def _cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def gravity_i(rs, i, ms, G):
    """src/basic_potentials.jl:306-331"""
    n = rs.shape[1]
    a = [0.0, 0.0, 0.0]
    ri = _col(rs, i)
    for j in range(n):
        if j != i:
            rj = _col(rs, j)
            rij = (ri[0] - rj[0], ri[1] - rj[1], ri[2] - rj[2])
            nr = _norm(rij)
            factor = -G * float(ms[j]) / (nr * nr * nr)
            for k in range(3):
                a[k] += factor * rij[k]
    return a


def lj_i(rs, i, idx, ms, eps, sigma, R, bc):
    """src/basic_potentials.jl:240-272"""
    s2, R2 = sigma * sigma, R * R
    f = [0.0, 0.0, 0.0]
    ri = _col(rs, i)
    for j in idx:
        if j != i:
            rij, r, r2 = distance(ri, _col(rs, j), bc)
            if r2 < R2:
                q = s2 / r2
                s6 = q * q * q
                s12 = s6 * s6
                factor = (2 * s12 - s6) / r2
                for k in range(3):
                    f[k] += factor * rij[k]
    coeff = 24 * eps / float(ms[i])
    return [coeff * f[k] for k in range(3)]


def coulomb_i(rs, i, qs, ms, exclude, k_el, R, bc):
    """src/basic_potentials.jl:274-304"""
    n = rs.shape[1]
    R2 = R * R
    f = [0.0, 0.0, 0.0]
    ri = _col(rs, i)
    for j in range(n):
        if j not in exclude:
            rij, r, r2 = distance(ri, _col(rs, j), bc)
            if r2 < R2:
                factor = float(qs[j]) / (r * r2)
                for k in range(3):
                    f[k] += factor * rij[k]
    coeff = k_el * float(qs[i]) / float(ms[i])
    return [coeff * f[k] for k in range(3)]


def dipole_i(rs, i, ms, mm, mu_4pi):
    """src/basic_potentials.jl:333-365"""
    n = rs.shape[1]
    f = [0.0, 0.0, 0.0]
    mi = _col(mm, i)
    ri = _col(rs, i)
    for j in range(n):
        if j != i:
            mj = _col(mm, j)
            rj = _col(rs, j)
            rij = (ri[0] - rj[0], ri[1] - rj[1], ri[2] - rj[2])
            d = _dot(rij, rij)
            rij4 = d * d
            nr = _norm(rij)
            r = (rij[0] / nr, rij[1] / nr, rij[2] / nr)
            mir = _dot(mi, r)
            mij = _dot(mj, r)
            mimj = _dot(mi, mj)
            for k in range(3):
                f[k] += (((mi[k] * mij + mj[k] * mir) + r[k] * mimj) - ((5 * r[k]) * mir) * mij) / rij4
    coeff = 3 * mu_4pi / float(ms[i])
    return [coeff * f[k] for k in range(3)]


def bond_i(rs, i, ms, rOH, kb):
    """src/basic_potentials.jl:367-393 with the partner table of src/nbody_to_ode.jl:263-288"""
    o = 3 * (i // 3)
    partners = [o + 1, o + 2] if i == o else [o]
    f = [0.0, 0.0, 0.0]
    ri = _col(rs, i)
    for j in partners:
        rj = _col(rs, j)
        rij = (ri[0] - rj[0], ri[1] - rj[1], ri[2] - rj[2])
        r = _norm(rij)
        d = r - rOH
        factor = -d * kb / r
        for k in range(3):
            f[k] += factor * rij[k]
    coeff = 1.0 / float(ms[i])
    return [coeff * f[k] for k in range(3)]


def angle_abc(dv, rs, a, b, c, ms, ka, aHOH0):
    """src/basic_potentials.jl:395-433; adds into dv (3 x n numpy array) in place."""
    ra, rb, rc = _col(rs, a), _col(rs, b), _col(rs, c)
    rba = tuple(ra[k] - rb[k] for k in range(3))
    rbc = tuple(rc[k] - rb[k] for k in range(3))
    rcb = tuple(rb[k] - rc[k] for k in range(3))
    X = _cross(rba, rbc)
    pa = _cross(rba, X)
    pc = _cross(rcb, X)
    ipa, ipc = 1.0 / _norm(pa), 1.0 / _norm(pc)
    pa = tuple(ipa * pa[k] for k in range(3))
    pc = tuple(ipc * pc[k] for k in range(3))
    nba, nbc = _norm(rba), _norm(rbc)
    cosine = _dot(rba, rbc) / (nba * nbc)
    cosine = 1.0 if cosine > 1 else (-1.0 if cosine < -1 else cosine)
    theta = math.acos(cosine)
    force = -ka * (theta - aHOH0)
    for k in range(3):
        fa = pa[k] * force / nba
        fc = pc[k] * force / nbc
        fb = -(fa + fc)
        dv[k, a] += fa / float(ms[a])
        dv[k, b] += fb / float(ms[b])
        dv[k, c] += fc / float(ms[c])


def md_temperature(vs, ms, kb, N, Nc):
    """src/thermostats.jl:87-91"""
    e = 0.0
    for i in range(vs.shape[1]):
        v = _col(vs, i)
        e += float(ms[i]) * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    return e / (kb * (3 * N - Nc))


def rhs(spec, u, v):
    """soode_system! for an ordinary (src/nbody_to_ode.jl:474-488) or water (:502-532) system.

    ``spec`` is the keyword dictionary accepted by ``nbody_oracle.System`` plus ``ms``.
    Potential order for ordinary systems: lennard_jones, electrostatic, magnetostatic,
    gravitational (the reference's Dict order is a hash artefact).
    """
    ms = spec["ms"]
    n = len(ms)
    bc = spec.get("bc", ("infinite",))
    water = spec.get("water", False)
    dv = np.zeros((3, u.shape[1]), order="F")
    for i in range(n):
        acc = [0.0, 0.0, 0.0]

        def add(t):
            for k in range(3):
                acc[k] += t[k]

        if water:
            o = 3 * (i // 3)
            if spec.get("coulomb"):
                c = spec["coulomb"]
                add(coulomb_i(u, i, spec["qs"], ms, (o, o + 1, o + 2), c["k"], c.get("R", math.inf), bc))
            if spec.get("spcfw"):
                add(bond_i(u, i, ms, spec["spcfw"]["rOH"], spec["spcfw"]["kb"]))
            if i == o and spec.get("lj"):
                lj = spec["lj"]
                add(lj_i(u, i, range(0, n, 3), ms, lj["eps"], lj["sigma"], lj["R"], bc))
        else:
            if spec.get("lj"):
                lj = spec["lj"]
                add(lj_i(u, i, range(n), ms, lj["eps"], lj["sigma"], lj["R"], bc))
            if spec.get("coulomb"):
                c = spec["coulomb"]
                add(coulomb_i(u, i, spec["qs"], ms, (i,), c["k"], c.get("R", math.inf), bc))
            if spec.get("dipole"):
                add(dipole_i(u, i, ms, spec["mm"], spec["dipole"]["mu_4pi"]))
            if spec.get("gravity"):
                add(gravity_i(u, i, ms, spec["gravity"]["G"]))
        for k in range(3):
            dv[k, i] = acc[k]
    if water and spec.get("spcfw"):
        for m in range(n // 3):
            angle_abc(dv, u, 3 * m + 1, 3 * m, 3 * m + 2, ms, spec["spcfw"]["ka"], spec["spcfw"]["aHOH"])
    th = spec.get("thermostat")
    if th and th["kind"] == "berendsen":  # src/thermostats.jl:76-83
        T = md_temperature(v, ms, th["kB"], th.get("N", n), th.get("Nc", 0))
        gamma = 0.5 / th["tau"]
        s = gamma if 1.0 / T == math.inf else gamma * (th["T"] / T - 1)
        for i in range(v.shape[1]):
            for k in range(3):
                dv[k, i] += s * float(v[k, i])
    return dv


def rdf_hist(u, L, idx_stride=1, maxbin=1000):
    """One frame of rdf's histogram (src/nbody_simulation_result.jl:676-693), plain Python loops; 0-based bins."""
    n = u.shape[1]
    dr = L / maxbin
    hist = [0] * maxbin
    bc = ("cubic", L)
    for i in range(0, n, idx_stride):
        ri = [float(u[k, i]) for k in range(3)]
        for j in range(i + idx_stride, n, idx_stride):
            _, r, r2 = distance(ri, [float(u[k, j]) for k in range(3)], bc)
            if r2 < (0.5 * L) ** 2:
                b = math.ceil(r / dr)
                if 1 < b <= maxbin:
                    hist[b - 1] += 2
    return hist


def msd(u, u0):
    """src/nbody_simulation_result.jl:730-752 for one frame (atoms)."""
    n = u.shape[1]
    s = 0.0
    for i in range(n):
        d = [float(u[k, i]) - float(u0[k, i]) for k in range(3)]
        s += d[0] * d[0] + d[1] * d[1] + d[2] * d[2]
    return s / n
