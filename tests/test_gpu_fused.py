"""The fused cutoff step (csrc/nbx_fused.cu: state in cell order, cluster lists, one kernel per velocity-Verlet
step) against the CPU oracle and against the unfused kernels.

The pair set must be the reference's (src/basic_potentials.jl:253-266, :288-297 through
src/boundary_conditions.jl:138-165): after a hot run with several on-device list rebuilds the resident
accelerations equal the reference loop evaluated at the resident positions to <= 1e-12 per body, and the
trajectory equals the unfused one up to the summation order (<= 1e-9 relative after 80 steps).
"""
import numpy as np
import pytest

import nbody_b200.workloads as wl
from tests._common import F, make_context, make_oracle
from tests.test_gpu_parity import NT, _check

pytestmark = pytest.mark.gpu


def _argon(cells, seed, hot=3.0, drift=False):
    w = wl.fcc_argon_reduced(cells)
    rng = np.random.Generator(np.random.Philox(seed))
    u = w["u"] + 0.05 * rng.standard_normal(w["u"].shape)
    if drift:  # the reference never wraps positions
        u = u + w["L"] * rng.integers(-2, 3, size=u.shape)
    return w, F(u), F(hot * w["v"])


def _run(spec, u, v, dt, steps, fused, cluster=4, calls=1, graph=1):
    ctx = make_context(spec)
    ctx.set_option("fused_step", 1 if fused else 0)
    ctx.set_option("fused_cluster", cluster)
    ctx.set_option("fused_min_steps", 2)
    ctx.set_option("graph", graph)
    ctx.upload(u, v)
    for _ in range(calls):
        ctx.step_vv(dt, steps // calls)
    out = ctx.download(want_dv=True)
    info = {k: ctx.info(k) for k in ("fused_steps", "fused_disabled", "verlet_rebuilds")}
    ctx.close()
    return out, info


@pytest.mark.parametrize("thermostat", [False, True])
@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
def test_fused_lj_step_matches_oracle_and_unfused_path(oracle, cluster, thermostat):
    w, u, v = _argon(10, 41, drift=(cluster == 2))  # 4,000 atoms
    n = u.shape[1]
    dt, steps = 2e-3, 80
    spec = dict(ms=w["ms"], bc=("cubic", w["L"]), lj=w["lj"])
    if thermostat:
        spec["thermostat"] = dict(kind="berendsen", T=90.0, tau=20 * dt, kB=w["kB"], N=n, Nc=0)
    ref, iref = _run(spec, u, v, dt, steps, fused=False)
    got, info = _run(spec, u, v, dt, steps, fused=True, cluster=cluster)
    assert iref["fused_steps"] == 0
    assert info["fused_steps"] == steps and info["fused_disabled"] == 0
    assert 3 <= info["verlet_rebuilds"] < steps  # rebuilt on the device several times, not on every step
    for a, b in zip(got, ref):
        assert np.abs(a - b).max() <= 1e-9 * np.abs(b).max()  # only the summation order differs
    # exact at the final state: no pair was missed or added.  (With Berendsen the resident dv carries the RHS term
    # of the previous velocity, src/thermostats.jl:76-83 as velocity Verlet evaluates it: compared above only.)
    if not thermostat:
        ug, vg, ag = got
        _check(ag, make_oracle(oracle, spec).rhs(ug, vg.copy(order="F"), NT))


def test_fused_runs_compose(oracle):
    """Two calls of 40 steps == one call of 80 (each call starts with a list build: order of summation only);
    eager launches == graph replay bit for bit."""
    w, u, v = _argon(8, 43)
    dt = 2e-3
    spec = dict(ms=w["ms"], bc=("cubic", w["L"]), lj=w["lj"],
                thermostat=dict(kind="berendsen", T=90.0, tau=20 * dt, kB=w["kB"], N=u.shape[1], Nc=0))
    one, _ = _run(spec, u, v, dt, 80, fused=True)
    two, info = _run(spec, u, v, dt, 80, fused=True, calls=2)
    eager, _ = _run(spec, u, v, dt, 80, fused=True, graph=0)
    assert info["fused_steps"] == 80
    for a, b in zip(two, one):
        assert np.abs(a - b).max() <= 1e-9 * np.abs(b).max()
    for a, b in zip(eager, one):
        assert np.array_equal(a, b)
    again, _ = _run(spec, u, v, dt, 80, fused=True)
    for a, b in zip(again, one):
        assert np.array_equal(a, b)  # deterministic


def test_fused_odd_step_counts(oracle):
    w, u, v = _argon(6, 47)
    dt = 2e-3
    spec = dict(ms=w["ms"], bc=("cubic", w["L"]), lj=w["lj"])
    for steps in (2, 3, 33, 34):
        ref, _ = _run(spec, u, v, dt, steps, fused=False)
        got, info = _run(spec, u, v, dt, steps, fused=True)
        assert info["fused_steps"] == steps
        for a, b in zip(got, ref):
            assert np.abs(a - b).max() <= 1e-10 * np.abs(b).max()


def test_fused_coulomb_cutoff(oracle):
    """Charged particles, cubic box, finite cutoff (src/basic_potentials.jl:274-304 with exclude = {i})."""
    rng = np.random.Generator(np.random.Philox(51))
    m = 14
    n = m ** 3
    L = 14.0
    g = (np.arange(m) + 0.5) * (L / m)
    u = np.stack(np.meshgrid(g, g, g, indexing="ij")).reshape(3, -1) + 0.1 * rng.standard_normal((3, n))
    qs = np.where(np.arange(n) % 2 == 0, 1.0, -1.0) * (0.5 + rng.random(n))
    spec = dict(ms=rng.random(n) + 1.0, qs=qs, bc=("cubic", L), coulomb=dict(k=0.7, R=3.2))
    u, v = F(u), F(0.3 * rng.standard_normal((3, n)))
    dt, steps = 1e-3, 60
    ref, _ = _run(spec, u, v, dt, steps, fused=False)
    got, info = _run(spec, u, v, dt, steps, fused=True)
    assert info["fused_steps"] == steps
    for a, b in zip(got, ref):
        assert np.abs(a - b).max() <= 1e-9 * np.abs(b).max()
    ug, vg, ag = got
    _check(ag, make_oracle(oracle, spec).rhs(ug, vg.copy(order="F"), NT))


def test_fused_list_overflow_hands_over_to_the_unfused_path(oracle):
    """Clusters far denser than the box average overflow the cluster lists (sized from the mean density): the
    state freezes at the last complete step and the unfused kernels finish the run -- same trajectory."""
    rng = np.random.Generator(np.random.Philox(77))
    L, R, n = 24.0, 2.5, 4000
    centres = rng.random((3, 5)) * L
    u = centres[:, rng.integers(0, 5, n)] + 1.1 * rng.standard_normal((3, n))
    u[:, :1000] = rng.random((3, 1000)) * L
    u, v = F(u), F(np.zeros((3, n)))
    spec = dict(ms=rng.random(n) + 0.5, bc=("cubic", L), lj=dict(eps=1e-30, sigma=0.35, R=R))
    dt, steps = 1e-3, 20
    ref, _ = _run(spec, u, v, dt, steps, fused=False)
    got, info = _run(spec, u, v, dt, steps, fused=True)
    assert info["fused_disabled"] == 1 and info["fused_steps"] < steps
    for a, b in zip(got, ref):
        assert np.abs(a - b).max() <= 1e-9 * np.abs(b).max()


@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
def test_fused_cluster_lists_cover_every_in_cutoff_pair(cluster):
    """Structure of the cluster lists after a build (nbx_debug_fetch): padded cell starts are multiples of the
    cluster size, a cluster never spans two cells, the slots hold a permutation of the particles, no padding slot is
    ever listed, and every pair the reference's predicate accepts (src/boundary_conditions.jl:138-165 +
    src/basic_potentials.jl:258) is in the union list of the target's cluster."""
    w, u, v = _argon(6, 41, hot=0.0)  # 864 atoms, 4 cells per dimension: every cell touches the periodic faces
    n = u.shape[1]
    spec = dict(ms=w["ms"], bc=("cubic", w["L"]), lj=w["lj"])
    ctx = make_context(spec)
    ctx.set_option("fused_step", 1)
    ctx.set_option("fused_cluster", cluster)
    ctx.set_option("fused_min_steps", 2)
    ctx.set_option("fused_debug", 4)  # build the lists, leave the steps to the unfused kernels
    ctx.upload(u, v)
    ctx.step_vv(0.0, 2)
    start, pid, scell, nlist = (ctx.debug_fetch(k) for k in ("start", "pid", "scell", "nlist"))
    cap = len(pid)
    lst = ctx.debug_fetch("list").reshape(-1, cap)
    x1 = ctx.debug_fetch("x", 1).reshape(-1, 4)
    ctx.close()
    ns = int(start[-1])
    assert (start % cluster == 0).all() and ns % cluster == 0
    real = pid[:ns] >= 0
    assert np.array_equal(np.sort(pid[:ns][real]), np.arange(n))
    cells = scell[:ns].reshape(-1, cluster)
    assert (cells == cells[:, :1]).all()
    assert np.array_equal(x1[:ns][real, :3], u.T[pid[:ns][real]])
    L, R = w["L"], w["lj"]["R"]
    pos = u.T[np.maximum(pid[:ns], 0)]
    for g in range(ns // cluster):
        ks = range(g * cluster, (g + 1) * cluster)
        union = set()
        for k in ks:
            union |= set(lst[:nlist[k], k].tolist())
        assert all(pid[m] >= 0 for m in union)
        for k in ks:
            if pid[k] < 0:
                continue
            d = pos[k] - pos
            d -= L * np.round(d / L)
            r2 = (d ** 2).sum(axis=1)
            inside = set(np.where((r2 < R * R) & real)[0].tolist()) - {k}
            assert inside <= union, (g, k, sorted(inside - union))
