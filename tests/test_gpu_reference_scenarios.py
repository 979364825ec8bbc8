"""The reference's own functional tests (test/*.jl), restated against the B200 path through the
reference-shaped host API (nbody_b200.api).  Each test cites the Julia test it follows and keeps its
tolerance.  GPU only."""
import math

import numpy as np
import pytest

from nbody_b200.api import (AndersenThermostat, BerendsenThermostat, ChargedParticle, ChargedParticles,
                            CubicPeriodicBoundaryConditions, EM, ElectrostaticParameters, GravitationalSystem,
                            LangevinThermostat, LennardJonesParameters, MagneticParticle, MagnetostaticParameters,
                            MassBody, NBodySimulation, NoseHooverThermostat, PeriodicBoundaryConditions,
                            PotentialNBodySystem, SPCFwParameters, SecondOrderODEProblem, Tsit5, VelocityVerlet,
                            WaterSPCFw, generate_bodies_in_cell_nodes, get_accelerating_function, get_position,
                            get_velocity, initial_energy, kinetic_energy, potential_energy, run_simulation,
                            msd, rdf, temperature, total_energy)

pytestmark = pytest.mark.gpu


def _figure_eight():
    # test/gravitational_test.jl:7-18
    m1 = MassBody([-0.995492, 0.0, 0.0], [-0.347902, -0.53393, 0.0], 1.0)
    m2 = MassBody([0.995492, 0.0, 0.0], [-0.347902, -0.53393, 0.0], 1.0)
    m3 = MassBody([0.0, 0.0, 0.0], [0.695804, 1.067860, 0.0], 1.0)
    return [m1, m2, m3]


def test_gravitational_figure_eight_velocity_verlet():
    bodies = _figure_eight()
    sim = NBodySimulation(GravitationalSystem(bodies, 1.0), (0.0, 2 * math.pi))
    sr = run_simulation(sim, VelocityVerlet(), dt=math.pi / 130)
    end = get_position(sr, 2 * math.pi)
    for i, b in enumerate(bodies):
        assert np.abs(end[:, i] - b.r).max() < 1e-3                  # :54-60
    assert kinetic_energy(sr, 0.0) == pytest.approx(1.218, abs=1e-3)  # :38-40
    assert np.array_equal(get_position(sr, 0.0, 1), bodies[1].r)      # exact echo :31-36
    assert np.array_equal(get_velocity(sr, 0.0, 2), bodies[2].v)


def test_gravitational_figure_eight_adaptive_host_integrator():
    """run_simulation(sim) with the default adaptive algorithm: the RHS drop-in mode (:19-24)."""
    bodies = _figure_eight()
    sim = NBodySimulation(GravitationalSystem(bodies, 1.0), (0.0, 2 * math.pi))
    sr = run_simulation(sim, Tsit5(), rtol=1e-9, atol=1e-11)
    end = get_position(sr, 2 * math.pi)
    for i, b in enumerate(bodies):
        assert np.abs(end[:, i] - b.r).max() < 0.1


def test_electrostatics_two_like_charges():
    # test/electrostatics_test.jl:42-68: speeds from energy conservation within 0.1 %
    k, q, m, r0 = 9e9, 1e-3, 100.0, 1.0
    p1 = ChargedParticle([-r0 / 2, 0, 0], [0, 0, 0], m, q)
    p2 = ChargedParticle([r0 / 2, 0, 0], [0, 0, 0], m, q)
    sim = NBodySimulation(ChargedParticles([p1, p2], k), (0.0, 1.0))
    sr = run_simulation(sim, VelocityVerlet(), dt=1e-3)
    x = get_position(sr, 1.0)
    r1 = np.linalg.norm(x[:, 0] - x[:, 1])
    v_expected = math.sqrt(k * q * q / m * (1 / r0 - 1 / r1))
    v = np.linalg.norm(get_velocity(sr, 1.0, 1))
    assert v == pytest.approx(v_expected, rel=1e-3)


def test_electrostatics_periodic_energy_drift():
    # test/electrostatics_test.jl:84-119: 8 charges, cubic PBC, cutoff 0.45 L, 1000 VV steps, drift < 1e-3
    L, k = 1.0, 9e9
    rng = np.random.Generator(np.random.Philox(8))
    bodies = []
    for ix in range(2):
        for iy in range(2):
            for iz in range(2):
                r = (np.array([ix, iy, iz]) + 0.25 + 0.02 * rng.standard_normal(3)) * L / 2
                bodies.append(ChargedParticle(r, [0, 0, 0], 1.0, 1e-6 * (1 if (ix + iy + iz) % 2 == 0 else -1)))
    pot = {"electrostatic": ElectrostaticParameters(k, 0.45 * L)}
    sim = NBodySimulation(PotentialNBodySystem(bodies, pot), (0.0, 1e-1), CubicPeriodicBoundaryConditions(L))
    sr = run_simulation(sim, VelocityVerlet(), dt=1e-4, saveat=1000)
    e0, e1 = total_energy(sr, 0.0), total_energy(sr, 1e-1)
    assert abs(e1 - e0) / abs(e0) < 1e-3


def test_magnetostatic_two_dipoles():
    # test/magnetostaic_test.jl:5-32: parallel dipoles repel; speed from the dipole energy, +-1e-3
    mu_4pi, d, t1 = 1e-7, 0.01, 1.0
    m = 5e-6
    mm = np.array([0.0, 0.0, 1.2e6 * m / 7800.0])
    p1 = MagneticParticle([-d / 2, 0, 0], [0, 0, 0], m, mm)
    p2 = MagneticParticle([d / 2, 0, 0], [0, 0, 0], m, mm)
    system = PotentialNBodySystem([p1, p2], {"magnetostatic": MagnetostaticParameters(mu_4pi)})
    sr = run_simulation(NBodySimulation(system, (0.0, t1)), VelocityVerlet(), dt=t1 / 100)
    x = get_position(sr, t1)
    r1 = np.linalg.norm(x[:, 1] - x[:, 0])
    # U = mu/4pi m1 m2 / r^3 for side-by-side parallel dipoles
    v_expected = math.sqrt(mu_4pi * mm[2] ** 2 / m * (1 / d ** 3 - 1 / r1 ** 3))
    assert np.linalg.norm(get_velocity(sr, t1, 1)) == pytest.approx(v_expected, abs=1e-3)


def test_lennard_jones_two_atoms():
    # test/lennard_jones_test.jl:4-33: R = Inf-like cutoff, 1000 VV steps, speed within 0.1 %
    eps, sigma, m = 1.0, 1.0, 1.0
    r0 = 1.3
    p1 = MassBody([0, 0, 0], [0, 0, 0], m)
    p2 = MassBody([r0, 0, 0], [0, 0, 0], m)
    system = PotentialNBodySystem([p1, p2], {"lennard_jones": LennardJonesParameters(eps, sigma, 1e6)})
    t1 = 0.3
    sr = run_simulation(NBodySimulation(system, (0.0, t1)), VelocityVerlet(), dt=t1 / 1000)
    x = get_position(sr, t1)
    r1 = np.linalg.norm(x[:, 0] - x[:, 1])

    def U(r):
        return 4 * eps * ((sigma / r) ** 12 - (sigma / r) ** 6)

    v_expected = math.sqrt((U(r0) - U(r1)) / m)  # two equal masses share the energy
    assert np.linalg.norm(get_velocity(sr, t1, 1)) == pytest.approx(v_expected, rel=1e-3)


def _argon(n_side=5, T=120.0):
    # test/thermostat_test.jl:5-23 (125 atoms, kJ/mol - nm - Da - ps units)
    kb = 8.3144598e-3
    eps, sigma = T * kb, 0.34
    m = 39.95
    n = n_side ** 3
    L = (m * n / (1374 / 1.6747)) ** (1 / 3)
    v_dev = math.sqrt(kb * T / m)
    bodies = generate_bodies_in_cell_nodes(n, m, v_dev, L)
    system = PotentialNBodySystem(bodies, {"lennard_jones": LennardJonesParameters(eps, sigma, 0.5 * L)})
    return system, L, kb, 0.5e-3


def test_lennard_jones_three_atoms_temperature_and_energy():
    # test/lennard_jones_test.jl:46-97
    kb = 8.3144598e-3
    T = 120.0
    eps, sigma = T * kb, 0.34
    m = 39.95
    L, tau = 5 * sigma, 0.5e-3
    v = math.sqrt(3 * kb * T / m)
    bodies = [MassBody([L / 3, L / 3, 2 * L / 3], [0, 0, -v], m), MassBody([L / 3, 2 * L / 3, L / 3], [0, -v, 0], m),
              MassBody([2 * L / 3, L / 3, L / 3], [-v, 0, 0], m)]
    pot = {"lennard_jones": LennardJonesParameters(eps, sigma, 2.25 * sigma)}
    t2 = 100 * tau
    for bc in (PeriodicBoundaryConditions(L), CubicPeriodicBoundaryConditions(L)):
        sim = NBodySimulation(PotentialNBodySystem(bodies, pot), (0.0, t2), bc, kb)
        sr = run_simulation(sim, VelocityVerlet(), dt=tau)
        assert temperature(sr, 0.0) == pytest.approx(T, abs=1e-6)                          # :77-79
        assert kinetic_energy(sr, 0.0) == pytest.approx(3 * m * v * v / 2, rel=1e-14)        # :81-82
        e0, e1 = total_energy(sr, 0.0), total_energy(sr, t2)
        assert abs(e1 - e0) <= 0.1 * abs(e0)                                                 # :84-97
        assert initial_energy(sim) == pytest.approx(e0, rel=1e-12)


@pytest.mark.parametrize("thermostat,tol", [("andersen", 0.5), ("berendsen", 0.1), ("nosehoover", 0.5),
                                            ("langevin", 0.5)])
def test_thermostats_reach_target_temperature(thermostat, tol):
    # test/thermostat_test.jl:25-94: 125 argon atoms, 200 steps, |T2 - T0| / T0 within the reference's bound
    system, L, kb, tau = _argon()
    T0 = 90.0
    th = {"andersen": AndersenThermostat(T0, 0.1 / tau), "berendsen": BerendsenThermostat(T0, 10 * tau),
          "nosehoover": NoseHooverThermostat(T0, 20 * tau), "langevin": LangevinThermostat(T0, 10)}[thermostat]
    t1 = 200 * tau
    sim = NBodySimulation(system, (0.0, t1), CubicPeriodicBoundaryConditions(L), th, kb)
    alg = EM() if thermostat == "langevin" else VelocityVerlet()
    sr = run_simulation(sim, alg, dt=tau, seed=1234)
    T2 = temperature(sr, t1)
    assert np.isfinite(T2)
    assert abs(T2 - T0) / T0 < tol
    if thermostat == "andersen":  # save_everystep = false: start and end frames only, unique times (:29-37)
        assert len(sr.t) == 2 and sr.t[0] != sr.t[1]


def _water_system(nside=2):
    # test/water_test.jl:5-54 with the example charges
    kb = 8.3144598e-3
    mO, mH = 15.999, 1.00794
    T = 370.0
    n = nside ** 3
    L = (n * (mO + 2 * mH) / (997.0 / 1.6747)) ** (1 / 3)
    lj = LennardJonesParameters(0.1554253 * 4.184, 0.3165492, 0.49 * L)
    el = ElectrostaticParameters(138.935458, 0.49 * L)
    sp = SPCFwParameters(0.1012, 113.24 * math.pi / 180, 1059.162 * 4.184 * 1e2, 75.9 * 4.184)
    bodies = generate_bodies_in_cell_nodes(n, mO + 2 * mH, math.sqrt(kb * T / (mO + 2 * mH)), L)
    return WaterSPCFw(bodies, mH, mO, 0.41, -0.82, lj, el, sp), L, kb


def test_water_spcfw_short_run():
    # test/water_test.jl:56-82: 10 VV steps; bonds/angle within 1 %, energy drift < 1 %
    water, L, kb = _water_system(2)
    dt = 0.5e-4
    t1 = 10 * dt
    sim = NBodySimulation(water, (0.0, t1), CubicPeriodicBoundaryConditions(L), kb)
    sr = run_simulation(sim, VelocityVerlet(), dt=dt)
    x = get_position(sr, t1)
    sp = water.scpfw_parameters
    for m in range(len(water.bodies)):
        o, h1, h2 = x[:, 3 * m], x[:, 3 * m + 1], x[:, 3 * m + 2]
        assert np.linalg.norm(h1 - o) == pytest.approx(sp.rOH, rel=0.01)
        assert np.linalg.norm(h2 - o) == pytest.approx(sp.rOH, rel=0.01)
        cosang = np.dot(h1 - o, h2 - o) / np.linalg.norm(h1 - o) / np.linalg.norm(h2 - o)
        assert math.acos(cosang) == pytest.approx(sp.aHOH, rel=0.01)
    e0, e1 = total_energy(sr, 0.0), total_energy(sr, t1)
    assert e0 == pytest.approx(initial_energy(sim), rel=1e-12)   # :72-74 (device reduction: not bit-exact)
    assert abs(e1 - e0) / abs(e0) < 0.01
    ndf = 3 * 3 * len(water.bodies) - 2 * len(water.bodies)      # :79-82
    vs = get_velocity(sr, t1)
    ms = np.tile([water.mO, water.mH, water.mH], len(water.bodies))
    assert temperature(sr, t1) == pytest.approx(np.dot(ms, (vs ** 2).sum(axis=0)) / (kb * ndf), rel=1e-12)


def test_water_langevin_thermostating():
    # test/water_test.jl:135-152 "Water thermostating": 216 SPC/Fw molecules, LangevinThermostat(275, 100), run_simulation(sim, EM(),
    # dt = 0.5e-3) to t2 = 200 x 0.5e-4 (20 steps); the reference's own (loose) criterion |T2 - T0| / T0 <= 1.  The SDE is the
    # water variant of src/nbody_to_ode.jl:600-680 (oxygen friction term, noise amplitudes divided by the mass).
    water, L, kb = _water_system(6)
    T0, tau = 275.0, 0.5e-3
    t2 = 200 * 0.5e-4
    sim = NBodySimulation(water, (0.0, t2), CubicPeriodicBoundaryConditions(L), LangevinThermostat(T0, 100.0), kb)
    sr = run_simulation(sim, EM(), dt=tau, seed=11)
    T2 = temperature(sr, t2)
    assert np.isfinite(T2) and abs(T2 - T0) / T0 <= 1.0
    x = get_position(sr, t2)
    assert np.isfinite(x).all()


def test_lennard_jones_rdf_and_msd():
    # test/lennard_jones_test.jl:120-150: 125 argon atoms, cubic PBC, R = 0.5 L, 400 VV steps; MSD grows, RDF peaks near sigma
    T, kb = 120.0, 8.3144598e-3
    eps, sigma, m = T * kb, 0.34, 39.95
    L = 5 * sigma
    N, tau = 125, 0.5e-3
    bodies = generate_bodies_in_cell_nodes(N, m, math.sqrt(kb * T / m), L, rng=np.random.Generator(np.random.Philox(125)))
    system = PotentialNBodySystem(bodies, {"lennard_jones": LennardJonesParameters(eps, sigma, 0.5 * L)})
    sim = NBodySimulation(system, (0.0, 400 * tau), CubicPeriodicBoundaryConditions(L), kb)
    result = run_simulation(sim, VelocityVerlet(), dt=tau, saveat=20)
    ts, dr2 = msd(result)
    assert dr2[0] < dr2[-1] and dr2[0] == 0.0
    rs, grs = rdf(result)
    assert rs[int(np.argmax(grs))] / sigma == pytest.approx(1.0, abs=1.0)
    # the device histogram is the reference's, count for count; the device msd is the reference's sum to rounding
    from oracle import nbody_oracle as orc

    orc.build()
    hist = sum(orc.rdf_hist(np.asfortranarray(get_position(result, t)), L) for t in result.t)
    rs_o, gr_o = orc.rdf_normalise(hist, len(result.t), N, L)
    assert np.allclose(rs, rs_o, rtol=1e-14, atol=0.0) and np.allclose(grs, gr_o, rtol=1e-12, atol=0.0)  # (same counts; the normalisation is host arithmetic)
    x0 = np.asfortranarray(get_position(result, result.t[0]))
    for t, d in zip(ts, dr2):
        assert d == pytest.approx(orc.msd(np.asfortranarray(get_position(result, t)), x0), rel=1e-12, abs=0.0)


def test_plugin_closure_contract():
    """test/shared/custom_potential_body.jl:33-38: acceleration!(dv, u, v, t, i) ADDS into a 3-vector view."""
    bodies = _figure_eight()
    sim = NBodySimulation(GravitationalSystem(bodies, 1.0), (0.0, 1.0))
    prob = SecondOrderODEProblem(sim)
    acc = get_accelerating_function(sim.system.potentials["gravitational"], sim)
    full = np.zeros((3, 3), order="F")
    prob.f(full, prob.v0, prob.u0, None, 0.0)
    for i in range(3):
        dv = np.ones(3)
        acc(dv, prob.u0, prob.v0, 0.0, i)
        assert np.allclose(dv - 1.0, full[:, i], rtol=1e-14, atol=1e-16)
