"""Pins the CPU oracle (oracle/nbody_oracle.c) before anything is compared with it.

The reference ships no golden force vectors and Julia is absent, so
the oracle is pinned by
  (1) bit-for-bit agreement with an independent pure-Python restatement of the Julia source,
  (2) closed-form force values,
  (3) the known-answer scenarios of the reference's own test-suite, each cited below.
CPU only.
"""
import math

import numpy as np
import pytest

from oracle import nbody_oracle as orc
from oracle import nbody_oracle_np as onp

import nbody_b200.workloads as wl


def F(a):
    return np.asfortranarray(np.array(a, dtype=np.float64))


# ----------------------------------------------------------------------------------------
# (1) C restatement == pure-Python restatement, bit for bit
# ----------------------------------------------------------------------------------------
def _rand_sys(n, seed, L=None):
    rng = np.random.Generator(np.random.Philox(seed))
    scale = 1.0 if L is None else L
    u = F(rng.random((3, n)) * scale * 1.7 - 0.3 * scale)  # deliberately outside the box too
    v = F(rng.standard_normal((3, n)))
    ms = rng.random(n) + 0.5
    qs = rng.standard_normal(n)
    mm = F(rng.standard_normal((3, n)))
    return u, v, ms, qs, mm


@pytest.mark.parametrize("bc", [("infinite",), ("cubic", 1.3), ("periodic", (0.0, 1.3, 0.0, 1.3, 0.0, 1.3)),
                                ("periodic", (-0.2, 0.9, 0.1, 1.5, -1.0, 0.4))])
def test_c_equals_python_ordinary(bc):
    n = 23
    u, v, ms, qs, mm = _rand_sys(n, 11, 1.3)
    spec = dict(ms=ms, qs=qs, mm=mm, bc=bc, lj=dict(eps=0.7, sigma=0.31, R=0.55),
                coulomb=dict(k=2.5, R=0.6), dipole=dict(mu_4pi=1e-3), gravity=dict(G=0.3),
                thermostat=dict(kind="berendsen", T=1.5, tau=0.2, kB=0.01))
    s = orc.System(ms, qs=qs, mm=mm, bc=bc, lj=spec["lj"], coulomb=spec["coulomb"], dipole=spec["dipole"],
                   gravity=spec["gravity"], thermostat=spec["thermostat"])
    a_c = s.rhs(u, v.copy(order="F"))
    a_p = onp.rhs(spec, u, v)
    assert np.array_equal(a_c, a_p)  # bit-exact
    assert np.isfinite(a_c).all() and np.abs(a_c).max() > 0


def test_c_equals_python_water():
    w = wl.water_omm(2, seed=5)
    rng = np.random.Generator(np.random.Philox(3))
    u = F(w["u"] + 0.01 * rng.standard_normal(w["u"].shape))
    bc = ("cubic", w["L"])
    spec = dict(ms=w["ms"], qs=w["qs"], water=True, bc=bc, lj=w["lj"], coulomb=w["coulomb"], spcfw=w["spcfw"])
    s = orc.System(w["ms"], qs=w["qs"], water=True, bc=bc, lj=w["lj"], coulomb=w["coulomb"], spcfw=w["spcfw"])
    a_c = s.rhs(u, F(w["v"]))
    a_p = onp.rhs(spec, u, w["v"])
    assert np.array_equal(a_c, a_p)


def test_accel_targets_matches_rhs_and_threads():
    u, v, ms, *_ = _rand_sys(200, 2)
    s = orc.System(ms, gravity=dict(G=1.0))
    full = s.rhs(u, v)
    t = np.array([0, 7, 199, 42])
    assert np.array_equal(s.accel_targets(u, t), full[:, t])
    assert np.array_equal(s.accel_targets(u, np.arange(200), nthreads=4), full)


# ----------------------------------------------------------------------------------------
# (2) closed forms (last sentence)
# ----------------------------------------------------------------------------------------
def test_closed_form_gravity():
    u = F([[0.0, 2.0], [0, 0], [0, 0]])
    s = orc.System([3.0, 5.0], gravity=dict(G=0.5))
    a = s.rhs(u, F(np.zeros((3, 2))))
    assert a[0, 0] == pytest.approx(0.5 * 5.0 / 4.0, rel=1e-15)
    assert a[0, 1] == pytest.approx(-0.5 * 3.0 / 4.0, rel=1e-15)


def test_closed_form_lj_zero_force_at_minimum():
    r0 = 2.0 ** (1.0 / 6.0)
    u = F([[0.0, r0], [0, 0], [0, 0]])
    s = orc.System([1.0, 1.0], lj=dict(eps=1.0, sigma=1.0, R=math.inf))
    a = s.rhs(u, F(np.zeros((3, 2))))
    assert abs(a[0, 0]) < 1e-14
    # and the analytic value elsewhere: F = 24 eps (2 (s/r)^12 - (s/r)^6) / r
    u = F([[0.0, 1.3], [0, 0], [0, 0]])
    a = s.rhs(u, F(np.zeros((3, 2))))
    assert a[0, 1] == pytest.approx(24 * (2 / 1.3 ** 12 - 1 / 1.3 ** 6) / 1.3, rel=1e-14)


def test_closed_form_coulomb():
    u = F([[0.0, 3.0], [0, 0], [0, 0]])
    s = orc.System([2.0, 4.0], qs=[1e-3, -2e-3], coulomb=dict(k=9e9))
    a = s.rhs(u, F(np.zeros((3, 2))))
    assert a[0, 0] == pytest.approx(9e9 * 1e-3 * 2e-3 / (2.0 * 9.0), rel=1e-15)  # attraction -> +x
    assert a[0, 1] == pytest.approx(-9e9 * 1e-3 * 2e-3 / (4.0 * 9.0), rel=1e-15)


def test_closed_form_dipoles_side_by_side():
    # parallel moments perpendicular to the separation: F = 3 mu/4pi m1 m2 / r^4, repulsive
    d, m1, m2 = 0.01, 2e-3, 3e-3
    u = F([[-d / 2, d / 2], [0, 0], [0, 0]])
    mm = F([[0, 0], [0, 0], [m1, m2]])
    s = orc.System([5e-6, 7e-6], mm=mm, dipole=dict(mu_4pi=1e-7))
    a = s.rhs(u, F(np.zeros((3, 2))))
    f = 3 * 1e-7 * m1 * m2 / d ** 4
    assert a[0, 1] == pytest.approx(f / 7e-6, rel=1e-14)
    assert a[0, 0] == pytest.approx(-f / 5e-6, rel=1e-14)


def test_closed_form_bond_and_angle():
    w = wl.water_omm(1)
    sp = w["spcfw"]
    s = orc.System(w["ms"], qs=w["qs"], water=True, spcfw=sp)
    # equilibrium geometry (src/nbody_to_ode.jl:47-49): both terms vanish
    a = s.rhs(w["u"], F(w["v"]))
    assert np.abs(a).max() < 1e-9 * sp["kb"] * sp["rOH"]
    # stretch the O-H1 bond by d along x: force on H1 = -kb d
    u = w["u"].copy(order="F")
    dlt = 0.003
    u[0, 1] += dlt
    a = s.rhs(u, F(w["v"]))
    # the angle is unchanged by a stretch along the bond, so H1 feels the bond force only
    assert a[0, 1] * w["ms"][1] == pytest.approx(-sp["kb"] * dlt, rel=1e-9)
    # bend: total force and torque of the angle term vanish
    u = w["u"].copy(order="F")
    u[2, 2] += 0.01
    u[1, 1] += 0.004
    s_ang = orc.System(w["ms"], qs=w["qs"], water=True, spcfw=dict(sp, kb=0.0))
    a = s_ang.rhs(u, F(w["v"]))
    Fm = a * w["ms"]
    assert np.abs(Fm.sum(axis=1)).max() < 1e-9 * np.abs(Fm).max()
    tq = sum(np.cross(u[:, k] - u[:, 0], Fm[:, k]) for k in range(3))
    assert np.abs(tq).max() < 1e-9 * np.abs(Fm).max()


def test_min_image_conventions():
    # Cubic wraps into [-L/2, L/2) (boundary_conditions.jl:138-165)
    rij, r, r2 = orc.distance([0.9, 0.5, 0.0], [0.1, 0.0, 0.5], 1, [1.0])
    assert rij == pytest.approx([-0.2, -0.5, -0.5])  # +0.5 -> -0.5 (half-open upper end)
    # Periodic(L) wraps the DISPLACEMENT into [0, L): not a minimum image (reference quirk, :111-136)
    rij, r, r2 = orc.distance([0.1, 0.0, 0.0], [0.2, 0.0, 0.0], 2, [0, 1.0, 0, 1.0, 0, 1.0])
    assert rij[0] == pytest.approx(0.9)
    rij, r, r2 = orc.distance([5.3, -7.2, 0.0], [0.0, 0.0, 0.0], 1, [1.0])
    assert np.all(np.abs(rij) <= 0.5)


# ----------------------------------------------------------------------------------------
# (3) known-answer scenarios of the reference's test-suite
# ----------------------------------------------------------------------------------------
def test_figure_eight_velocity_verlet():
    """test/gravitational_test.jl:9-24, :38-40, :54-60"""
    u0 = F([[-0.995492, 0.995492, 0.0], [0, 0, 0], [0, 0, 0]])
    v0 = F([[-0.347902, -0.347902, 0.695804], [-0.53393, -0.53393, 1.06786], [0, 0, 0]])
    s = orc.System(np.ones(3), gravity=dict(G=1))
    assert s.kinetic_energy(v0) == pytest.approx(1.218, abs=1e-3)
    u, v = orc.velocity_verlet(s, u0, v0, math.pi / 130, 260)
    assert np.abs(u - u0).max() < 1e-3


def test_two_body_lj_energy():
    """test/lennard_jones_test.jl:4-33"""
    r1 = 1.3
    u0 = F([[-r1 / 2, r1 / 2], [0, 0], [0, 0]])
    s = orc.System(np.ones(2), lj=dict(eps=1.0, sigma=1.0, R=math.inf))
    u, v = orc.velocity_verlet(s, u0, F(np.zeros((3, 2))), 1e-3, 1000)
    r2 = np.linalg.norm(u[:, 1] - u[:, 0])
    v_exp = math.sqrt(4 * ((1 / r1 ** 12 - 1 / r2 ** 12) - (1 / r1 ** 6 - 1 / r2 ** 6)))
    assert np.linalg.norm(v[:, 1]) == pytest.approx(v_exp, rel=1e-3)


def test_two_charges_repelling():
    """test/electrostatics_test.jl:42-68"""
    k, q, m = 9e9, 1e-3, 1.0
    u0 = F([[-0.5, 0.5], [0, 0], [0, 0]])
    s = orc.System([m, m], qs=[q, q], coulomb=dict(k=k))
    u, v = orc.velocity_verlet(s, u0, F(np.zeros((3, 2))), 1e-3, 1000)
    r2 = np.linalg.norm(u[:, 1] - u[:, 0])
    v_exp = math.sqrt(k * q * q / m * (1 / 1.0 - 1 / r2))
    assert np.linalg.norm(v[:, 1]) == pytest.approx(v_exp, rel=1e-3)


def test_eight_charges_pbc_energy():
    """test/electrostatics_test.jl:84-119: relative total-energy drift < 1e-3 over 1000 VV steps"""
    n, L, m, q, k = 8, 1.0, 1.0, 1.0, 9e9
    dL = L / (math.ceil(n ** (1 / 3)) + 1)
    ax = np.arange(dL / 2, L, dL)
    pts = [(x, y, z) for x in ax for y in ax for z in ax][:n]
    u0 = F(np.array(pts).T)
    tau = 0.01 * dL / math.sqrt(2 * k * q * q / (dL * m))
    s = orc.System(np.full(n, m), qs=np.full(n, q), bc=("cubic", L), coulomb=dict(k=k, R=0.45 * L))
    v0 = F(np.zeros((3, n)))
    e1 = s.kinetic_energy(v0) + s.potential_energy(u0)
    u, v = orc.velocity_verlet(s, u0, v0, tau, 1000)
    e2 = s.kinetic_energy(v) + s.potential_energy(u)
    assert abs((e2 - e1) / e1) < 1e-3


def test_repelling_dipoles():
    """test/magnetostaic_test.jl:5-32"""
    d1, m1, rho, M = 0.01, 5e-6, 7800, 1.2e6
    mmv = M * m1 / rho
    u0 = F([[-d1 / 2, d1 / 2], [0, 0], [0, 0]])
    mm = F([[0, 0], [0, 0], [mmv, mmv]])
    s = orc.System([m1, m1], mm=mm, dipole=dict(mu_4pi=1e-7))
    u, v = orc.velocity_verlet(s, u0, F(np.zeros((3, 2))), 0.01, 100)
    d2 = np.linalg.norm(u[:, 1] - u[:, 0])
    v_exp = math.sqrt(1e-7 * (mmv * mmv * (1 / d1 ** 3 - 1 / d2 ** 3)) / m1)
    assert np.linalg.norm(v[:, 1]) == pytest.approx(v_exp, abs=1e-3)


def _three_argon():
    T, kb = 120.0, 8.3144598e-3
    eps, sigma, m = T * kb, 0.34, 39.95
    L = 5 * sigma
    vd = math.sqrt(3 * kb * T / m)
    u0 = F([[L / 3, L / 3, 2 * L / 3], [L / 3, 2 * L / 3, L / 3], [2 * L / 3, L / 3, L / 3]])
    v0 = F([[0, 0, -vd], [0, -vd, 0], [-vd, 0, 0]])
    return T, kb, eps, sigma, m, L, vd, u0, v0


@pytest.mark.parametrize("bc_kind", ["periodic", "cubic"])
def test_three_argon_atoms(bc_kind):
    """test/lennard_jones_test.jl:46-97"""
    T, kb, eps, sigma, m, L, vd, u0, v0 = _three_argon()
    bc = ("periodic", (0, L, 0, L, 0, L)) if bc_kind == "periodic" else ("cubic", L)
    s = orc.System(np.full(3, m), bc=bc, lj=dict(eps=eps, sigma=sigma, R=2.25 * sigma))
    assert s.temperature(v0, kb) == pytest.approx(120.0, abs=1e-6)
    assert s.kinetic_energy(v0) == m * (3 * vd * vd) / 2 or s.kinetic_energy(v0) == pytest.approx(m * 3 * vd * vd / 2, rel=1e-15)
    e1 = s.kinetic_energy(v0) + s.potential_energy(u0)
    u, v = orc.velocity_verlet(s, u0, v0, 0.5e-3, 100)
    e2 = s.kinetic_energy(v) + s.potential_energy(u)
    assert e2 == pytest.approx(e1, abs=0.1 * abs(e1))


def test_water_three_molecules():
    """test/water_test.jl:33-82 (qO = -0.84 there)"""
    T, kb = 298.16, 8.3144598e-3
    mO, mH = 15.999, 1.00794
    mH2O = mO + 2 * mH
    L = (mH2O * 216 / (997 / 1.6747)) ** (1 / 3)
    rOH, aHOH = 0.1012, 113.24 * math.pi / 180
    vd = math.sqrt(kb * T / mH2O)
    opos = np.array([[L / 3, L / 3, 2 * L / 3], [L / 3, 2 * L / 3, L / 3], [2 * L / 3, L / 3, L / 3]]).T
    ovel = np.array([[0, 0, -vd], [0, -vd, 0], [-vd, 0, 0]]).T
    u0 = np.zeros((3, 9), order="F")
    v0 = np.zeros((3, 9), order="F")
    u0[:, 0::3] = opos
    u0[:, 1::3] = opos + np.array([[rOH], [0], [0]])
    u0[:, 2::3] = opos + np.array([[math.cos(aHOH) * rOH], [0], [math.sin(aHOH) * rOH]])
    for k in range(3):
        v0[:, k::3] = ovel
    ms = np.tile([mO, mH, mH], 3)
    qs = np.tile([-0.84, 0.41, 0.41], 3)
    s = orc.System(ms, qs=qs, water=True, bc=("cubic", L),
                   lj=dict(eps=0.1554253 * 4.184, sigma=0.3165492, R=0.9),
                   coulomb=dict(k=138.935458, R=0.49 * L),
                   spcfw=dict(rOH=rOH, aHOH=aHOH, kb=1059.162 * 4.184 * 1e2, ka=75.9 * 4.184))
    e1 = s.kinetic_energy(v0) + s.potential_energy(u0)
    u, v = orc.velocity_verlet(s, u0, v0, 0.5e-3, 10)
    for i in range(3):
        o = 3 * i
        b1, b2 = u[:, o] - u[:, o + 1], u[:, o] - u[:, o + 2]
        assert abs(rOH - np.linalg.norm(b1)) / rOH < 0.01
        assert abs(rOH - np.linalg.norm(b2)) / rOH < 0.01
        ang = math.acos(b1 @ b2 / (np.linalg.norm(b1) * np.linalg.norm(b2)))
        assert abs(ang - aHOH) / aHOH < 0.01
    e2 = s.kinetic_energy(v) + s.potential_energy(u)
    assert abs((e1 - e2) / e1) < 0.01
    T_exp = 3 * vd * vd * (2 * mH + mO) / (kb * 21)
    assert s.temperature(v0, kb, N=9, Nc=6) == pytest.approx(T_exp, rel=1e-12)


def test_water_sde_stepper_follows_the_reference_formulas():
    """The oracle's Euler-Maruyama stepper for the SDEProblem of WaterSPCFw (src/nbody_to_ode.jl:600-680): without
    friction and noise it is the plain Euler step of the water RHS; with friction the oxygen columns lose (gamma v) / mO more
    than the hydrogens (:627-629, :664); the noise of one step has the amplitudes sqrt(2 gamma kb T dt) / mO and / mH
    (:668-676)."""
    T, kb = 298.16, 8.3144598e-3
    mO, mH = 15.999, 1.00794
    L = ((mO + 2 * mH) * 216 / (997 / 1.6747)) ** (1 / 3)
    rOH, aHOH = 0.1012, 113.24 * math.pi / 180
    rng = np.random.default_rng(5)
    nm = 4
    opos = rng.random((3, nm)) * L
    u0 = np.zeros((3, 3 * nm), order="F")
    u0[:, 0::3] = opos
    u0[:, 1::3] = opos + np.array([[rOH], [0], [0]])
    u0[:, 2::3] = opos + np.array([[math.cos(aHOH) * rOH], [0], [math.sin(aHOH) * rOH]])
    v0 = np.asfortranarray(rng.standard_normal((3, 3 * nm)))
    s = orc.System(np.tile([mO, mH, mH], nm), qs=np.tile([-0.82, 0.41, 0.41], nm), water=True, bc=("cubic", L),
                   lj=dict(eps=0.1554253 * 4.184, sigma=0.3165492, R=0.9), coulomb=dict(k=138.935458, R=0.49 * L),
                   spcfw=dict(rOH=rOH, aHOH=aHOH, kb=1059.162 * 4.184 * 1e2, ka=75.9 * 4.184))
    dt = 1e-4
    a0 = s.rhs(u0, v0)
    u1, v1 = orc.euler_maruyama_water(s, u0, v0, dt, 1, 0.0, 0.0, mO, mH, np.random.default_rng(1))
    assert np.array_equal(u1, u0 + dt * v0) and np.allclose(v1, v0 + dt * a0, rtol=0, atol=1e-13 * np.abs(a0).max() * dt)
    gamma = 7.0
    _, v2 = orc.euler_maruyama_water(s, u0, v0, dt, 1, gamma, 0.0, mO, mH, np.random.default_rng(1))
    lost = (v1 - v2) / (dt * gamma)                    # = v for hydrogens, v (1 + 1 / mO) for oxygens
    assert np.allclose(lost[:, 1::3], v0[:, 1::3], rtol=1e-9) and np.allclose(lost[:, 2::3], v0[:, 2::3], rtol=1e-9)
    assert np.allclose(lost[:, 0::3], v0[:, 0::3] * (1.0 + 1.0 / mO), rtol=1e-9)
    _, v3 = orc.euler_maruyama_water(s, u0, v0, dt, 1, gamma, kb * T, mO, mH, np.random.default_rng(2))
    xi = np.random.default_rng(2).standard_normal(v0.shape)
    kick = (v3 - v2) / (math.sqrt(2.0 * gamma * kb * T) * math.sqrt(dt))
    assert np.allclose(kick[:, 0::3], xi[:, 0::3] / mO, rtol=1e-9, atol=1e-12)
    assert np.allclose(kick[:, 1::3], xi[:, 1::3] / mH, rtol=1e-9, atol=1e-12)


def test_berendsen_125_atoms():
    """test/thermostat_test.jl:39-56: |T2 - T0| / T0 < 0.1 after 200 steps (tau_B = 10 dt)"""
    T, T0, kb = 120.0, 90.0, 1.38e-23
    eps, sigma, m = T * kb, 3.4e-10, 39.95 * 1.6747e-27
    N = 125
    L = (m * N / 1374) ** (1 / 3)
    tau = 0.5e-15
    u0 = wl.cell_node_positions(N, L)
    v0 = F(math.sqrt(kb * T / m) * np.random.Generator(np.random.Philox(N)).standard_normal((3, N)))
    s = orc.System(np.full(N, m), bc=("cubic", L), lj=dict(eps=eps, sigma=sigma, R=0.5 * L),
                   thermostat=dict(kind="berendsen", T=T0, tau=10 * tau, kB=kb))
    u, v = orc.velocity_verlet(s, u0, v0, tau, 200, nthreads=4)
    T2 = s.temperature(v, kb)
    assert abs(T2 - T0) / T0 < 0.1


def test_nosehoover_matches_formula():
    """src/thermostats.jl:121-128: dv -= zeta v; dv[:,end] = 0; v[zind] = (T/T0 - (ndf+1)/ndf)/tau^2"""
    n = 5
    u, v, ms, *_ = _rand_sys(n, 4)
    u1 = F(np.hstack([u, [[0.37], [0], [0]]]))
    v1 = F(np.hstack([v, np.zeros((3, 1))]))
    th = dict(kind="nosehoover", T=1.2, tau=0.3, kB=0.05)
    s0 = orc.System(ms, gravity=dict(G=1.0))
    s = orc.System(ms, gravity=dict(G=1.0), thermostat=th)
    base = s0.rhs(u, v)
    vv = v1.copy(order="F")
    a = s.rhs(u1, vv)
    assert np.allclose(a[:, :n], base - 0.37 * v, rtol=1e-14, atol=0)
    assert np.all(a[:, n] == 0)
    Tm = s0.temperature(v, 0.05)
    ndf = 3 * n
    assert vv[0, n] == pytest.approx((Tm / 1.2 - (ndf + 1) / ndf) / 0.3 ** 2, rel=1e-13)


def test_neighbor_predicate_matches_distance():
    d = wl.fcc_argon_reduced(3, seed=9)
    rng = np.random.Generator(np.random.Philox(1))
    u = F(d["u"] + 0.05 * rng.standard_normal(d["u"].shape) + 3 * d["L"] * rng.integers(-1, 2, size=d["u"].shape))
    s = orc.System(d["ms"], bc=("cubic", d["L"]), lj=d["lj"])
    nb = s.neighbors(u, 5, 2.25)
    ref = [j for j in range(u.shape[1]) if j != 5 and onp.distance(u[:, 5], u[:, j], ("cubic", d["L"]))[2] < 2.25 ** 2]
    assert nb.tolist() == ref and len(ref) > 10


# ----------------------------------------------------------------------------------------
# analysis of frames: rdf / msd (src/nbody_simulation_result.jl:664-783)
# ----------------------------------------------------------------------------------------
def test_rdf_and_msd_c_equals_python():
    rng = np.random.Generator(np.random.Philox(21))
    L = 3.0
    u = F(rng.random((3, 70)) * L * 1.6 - 0.3 * L)  # outside the box too: the wrap loops matter
    assert np.array_equal(orc.rdf_hist(u, L), np.array(onp.rdf_hist(u, L)))
    assert np.array_equal(orc.rdf_hist(u[:, :69], L, idx_stride=3), np.array(onp.rdf_hist(u[:, :69], L, idx_stride=3)))
    u1 = F(u + 0.1 * rng.standard_normal(u.shape))
    assert orc.msd(u1, u) == onp.msd(u1, u)


def test_rdf_counts_every_pair_inside_half_the_box_twice():
    """Closed form: a pair at distance d < L/2 lands in bin ceil(d / dr) with weight 2; bin 1 is never filled (`bin > 1`,
    src/nbody_simulation_result.jl:688)."""
    L = 10.0
    u = F([[1.0, 1.0 + 3.337, 1.0], [1.0, 1.0, 1.0], [1.0, 1.0, 1.0 + 0.004]])
    h = orc.rdf_hist(u, L)
    dr = L / 1000
    assert h.sum() == 4 and h[math.ceil(3.337 / dr) - 1] == 4  # pairs (0,1) and (1,2) share a bin; (0,2) at 0.004 < dr is dropped
    rs, gr = orc.rdf_normalise(h, 1, 3, L)
    assert rs[0] == dr / 2 and gr[0] == 0.0


def test_extended_precision_referee_agrees_with_the_restatement(oracle):
    """orc_accel_targets_ld (long double force arithmetic on the reference's pair set; the tests' referee where net
    accelerations cancel) and the fp64 restatement agree to rounding on a well-conditioned mixed system, potential by
    potential, under all three boundary kinds."""
    rng = np.random.default_rng(5)
    n = 200
    u = np.asfortranarray(rng.random((3, n)) * 1.3 + 4.0 * rng.integers(-2, 3, (3, n)))
    ms, qs = rng.random(n) + 0.5, rng.standard_normal(n)
    mm = np.asfortranarray(rng.standard_normal((3, n)))
    t = np.arange(n)
    cases = [dict(bc=("cubic", 1.3), lj=dict(eps=0.7, sigma=0.05, R=0.3)),
             dict(bc=("periodic", [0.0, 1.3, 0.0, 1.3, 0.0, 1.3]), lj=dict(eps=0.7, sigma=0.05, R=0.9)),
             dict(bc=("cubic", 1.3), qs=qs, coulomb=dict(k=2.5, R=0.4)),
             dict(qs=qs, coulomb=dict(k=2.5, R=np.inf)),
             dict(mm=mm, dipole=dict(mu_4pi=1e-3)),
             dict(gravity=dict(G=0.3))]
    for kw in cases:
        s = oracle.System(ms, **kw)
        a, b = s.accel_targets(u, t, 2), s.accel_targets_ld(u, t, 2)
        err = np.linalg.norm(a - b, axis=0) / np.linalg.norm(b, axis=0)
        assert err.max() < 1e-12, (kw.keys(), err.max())
    # water: Lennard-Jones acts on the oxygens only, Coulomb skips the own molecule
    s = oracle.System(np.tile([16.0, 1.0, 1.0], 20), qs=np.tile([-0.8, 0.4, 0.4], 20), water=True, bc=("cubic", 1.3),
                      lj=dict(eps=0.7, sigma=0.05, R=0.3), coulomb=dict(k=2.5, R=0.5))
    uw = np.asfortranarray(rng.random((3, 60)) * 1.3)
    a, b = s.accel_targets(uw, np.arange(60), 1), s.accel_targets_ld(uw, np.arange(60), 1)
    assert (np.linalg.norm(a - b, axis=0) <= 1e-12 * np.linalg.norm(b, axis=0)).all()
