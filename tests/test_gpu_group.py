"""Multi-GPU behind the C ABI (csrc/nbx_multi.cu, nbx_group.cu, slab_enqueue of nbx_slab.cu), exercised on ONE device:
`Context([0, 0, ...])` = nbx_create_multi with several members on the same GPU, and single contexts joined by the
nbx_group_* calls and driven from one thread each.  The members exchange everything through peer-memory stores and
flags inside their kernels -- no torch.distributed, no host in the loop -- so on one device they run as concurrent
streams.  tests/mgpu_worker.py repeats the cross-process form (CUDA IPC) when >= 2 GPUs are visible.

Parity bars: slabs without a thermostat reproduce the single-context trajectory BIT FOR BIT (same rebuild steps, cell
order ranked by global id); with Berendsen the global sum m v^2 is added per slab, hence 1e-11.  Pair sharding and
target blocks sum in a different order: <= 1e-12 per acceleration, 1e-10 after a short trajectory.
"""
import threading

import numpy as np
import pytest

import nbody_b200.workloads as wl
from nbody_b200 import _lib
from nbody_b200.parallel import join_group_local
from tests._common import F, make_context, make_oracle

pytestmark = pytest.mark.gpu
NT = 8


def _argon(cells, seed, hot=3.0, thermostat=False):
    w = wl.fcc_argon_reduced(cells)
    rng = np.random.Generator(np.random.Philox(seed))
    u = F(w["u"] + 0.05 * rng.standard_normal(w["u"].shape))
    v = F(hot * w["v"])
    spec = dict(ms=w["ms"], bc=("cubic", w["L"]), lj=w["lj"])
    if thermostat:
        spec["thermostat"] = dict(kind="berendsen", T=90.0, tau=10 * 2e-3, kB=w["kB"])
    return spec, u, v


def _group(spec, devices, **opts):
    """nbx_create_multi context; device-side waits give up after 3 s here (a protocol bug must not eat GPU minutes)."""
    grp = make_context(spec, device=devices)
    grp.set_option("spin_timeout_ms", 3000)
    for k, val in opts.items():
        grp.set_option(k, val)
    return grp


def _relmax(a, b):
    den = np.maximum(np.linalg.norm(b, axis=0), 1e-300)
    return float((np.linalg.norm(a - b, axis=0) / den).max())


def _single_run(spec, u, v, dt, nsteps, em=False, seed=0):
    ctx = make_context(spec)
    # slabs keep the plain list order (the position-ordered layout depends on the LOCAL slot numbers): the single context
    # is compared with them in that order, bit for bit
    ctx.set_option("verlet_banked", 0)
    ctx.upload(u, v)
    (ctx.step_em(dt, nsteps, seed) if em else ctx.step_vv(dt, nsteps))
    out = ctx.download(want_dv=True)
    info = ctx.info("verlet_rebuilds") if spec.get("lj") else 0
    T = ctx.energy(potential=False)[2]
    ctx.close()
    return out, info, T


@pytest.mark.parametrize("world", [1, 2, 4])
@pytest.mark.parametrize("thermostat", [False, True])
def test_slab_group_reproduces_the_single_context_trajectory(world, thermostat):
    """nbx_create_multi over x-slabs: 90 hot steps (several collective list rebuilds with migration, decided on the device
    from a peer-memory all-reduce of the displacement flags; the rest replayed as a CUDA graph with the rebuild chain
    in an IF node) against nbx_step_vv of one context."""
    spec, u, v = _argon(12, 5, thermostat=thermostat)   # 6,912 atoms, 8 layers with the skin: 2 per slab at world 4
    dt, nsteps = 2e-3, 90
    (u1, v1, a1), rebuilds1, T1 = _single_run(spec, u, v, dt, nsteps)
    grp = _group(spec, [0] * world)
    grp.upload(u, v)
    assert grp.info("group_mode") == 3 and grp.info("group_size") == world and grp.info("slab_verlet") == 1
    grp.step_vv(dt, 50)
    grp.step_vv(dt, nsteps - 50)   # a second call: two eager steps, then the cached graph
    ug, vg, ag = grp.download(want_dv=True)
    assert grp.info("verlet_rebuilds") >= 3
    assert grp.info("slab_own") == u.shape[1]
    Tg = grp.energy(potential=False)[2]
    if not thermostat:
        assert grp.info("verlet_rebuilds") in (rebuilds1, rebuilds1 + 1)   # (+1: one slab = the upload's lists AND the primed ones)
        assert np.array_equal(ug, u1) and np.array_equal(vg, v1) and np.array_equal(ag, a1)
    else:
        assert _relmax(ug, u1) < 1e-11 and _relmax(vg, v1) < 1e-9 and _relmax(ag, a1) < 1e-8
    assert abs(Tg - T1) <= 1e-12 * abs(T1)
    with pytest.raises(_lib.NbxError):
        grp.accel(u)       # slabs do not serve the RHS drop-in (group_mode 2 does)
    grp.close()


def test_slab_group_without_graph_and_with_cells_every_step():
    """The same loop launched eagerly (every kernel of the rebuild chain returns at once unless the reduced flag is set),
    and the list-less variant (verlet_skin_permille = 0: migration + cell rebuild every step)."""
    spec, u, v = _argon(12, 6)
    dt, nsteps = 2e-3, 40
    (u1, v1, a1), _, _ = _single_run(spec, u, v, dt, nsteps)
    for opts in (dict(graph=0), dict(verlet_skin_permille=0), dict(graph_if_nodes=0)):
        grp = _group(spec, [0, 0], **opts)
        grp.upload(u, v)
        grp.step_vv(dt, nsteps)
        ug, vg, ag = grp.download(want_dv=True)
        if "verlet_skin_permille" in opts:   # other cell order inside the lists: same pair set, sums in another order
            assert _relmax(ug, u1) < 1e-12 and _relmax(ag, a1) < 1e-9
        else:
            assert np.array_equal(ug, u1) and np.array_equal(vg, v1) and np.array_equal(ag, a1)
        grp.close()


def test_group_members_joined_by_hand_and_driven_from_threads():
    """The cross-process protocol inside one process: nbx_group_init / _export / _connect / _start on ordinary contexts
    (device pointers instead of IPC handles), every member stepped by its own host thread."""
    spec, u, v = _argon(12, 7)
    dt, nsteps = 2e-3, 60
    (u1, v1, a1), _, _ = _single_run(spec, u, v, dt, nsteps)
    world = 2
    ctxs = []
    for _ in range(world):
        c = make_context(spec)
        c.set_option("spin_timeout_ms", 3000)
        c.set_option("verlet_banked", 0)   # a(0) is evaluated before the context becomes a slab: same list order from the start
        c.upload(u, v)
        ctxs.append(c)
    join_group_local(ctxs)
    errs = []

    def run(c):
        try:
            c.step_vv(dt, nsteps)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=run, args=(c,)) for c in ctxs]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    n = u.shape[1]
    got = [np.zeros((3, n), order="F") for _ in range(3)]
    seen = np.zeros(n, dtype=int)
    for c in ctxs:
        gid, uu, vv, aa = c.slab_download()
        seen[gid] += 1
        for dst, src in zip(got, (uu, vv, aa)):
            dst[:, gid] = src
    assert (seen == 1).all()
    for g, ref in zip(got, (u1, v1, a1)):
        assert np.array_equal(g, ref)
    [c.close() for c in ctxs]


@pytest.mark.parametrize("world,n", [(2, 12288), (4, 20481), (3, 9000)])
def test_pair_sharded_gravity_group(oracle, world, n):
    """Unbounded gravity over a group: ring offsets of the Newton's-third-law kernel per member, partial accelerations
    pushed to their owners over peer memory and added in rank order; RHS drop-in with host buffers (own block up, own
    columns back) and the resident velocity-Verlet loop (positions all-gathered by the update kernel's peer stores)."""
    u, v, ms = wl.plummer(n, seed=n)
    spec = dict(ms=ms, gravity=dict(G=1.0))
    one = make_context(spec)
    a1 = one.accel(u).copy()
    grp = _group(spec, [0] * world)
    ag = grp.accel(u)
    assert grp.info("group_mode") == (1 if n >= 8192 else 1)
    assert _relmax(ag, a1) < 1e-12
    idx = np.random.Generator(np.random.Philox(3)).choice(n, 96, replace=False)
    ref = make_oracle(oracle, spec).accel_targets(u, idx, NT)
    assert _relmax(ag[:, idx], ref) < 1e-11
    dt, nsteps = 1e-4, 12
    one.upload(u, v)
    one.step_vv(dt, nsteps)
    u1, v1, a1 = one.download(want_dv=True)
    grp.upload(u, v)
    grp.step_vv(dt, nsteps)
    ug, vg, a2 = grp.download(want_dv=True)
    assert _relmax(ug, u1) < 1e-13 and _relmax(vg, v1) < 1e-11 and _relmax(a2, a1) < 1e-11
    ek1, _, _ = one.energy(potential=False)
    ekg, _, _ = grp.energy(potential=False)
    assert abs(ekg - ek1) <= 1e-12 * ek1
    # the RHS drop-in keeps working after the resident run (positions arbitrary again)
    assert _relmax(grp.accel(u), one.accel(u)) < 1e-12
    one.close(); grp.close()


def test_target_block_group_water(oracle):
    """SPC/Fw water (LJ on the oxygens + Coulomb cutoff + bonds + angle) over target blocks of whole molecules."""
    w = wl.water_omm(6, Rel=0.9)   # 216 molecules
    spec = dict(ms=w["ms"], qs=w["qs"], water=True, bc=("cubic", w["L"]), lj=w["lj"], coulomb=w["coulomb"], spcfw=w["spcfw"],
                thermostat=dict(kind="berendsen", T=300.0, tau=0.05, kB=w["kB"], N=3 * w["nmol"], Nc=2 * w["nmol"]))
    u, v = F(w["u"]), F(w["v"])
    one = make_context(spec)
    grp = _group(spec, [0, 0, 0])
    vv = F(v.copy())
    a1 = one.accel(u, vv).copy()
    ag = grp.accel(u, F(v.copy()))
    assert grp.info("group_mode") == 2
    assert _relmax(ag, a1) < 1e-12
    ref = make_oracle(oracle, spec).rhs(u, v, NT)
    assert _relmax(ag, ref) < 1e-11
    dt, nsteps = w["dt"], 30
    one.upload(u, v); one.step_vv(dt, nsteps)
    grp.upload(u, v); grp.step_vv(dt, nsteps)
    u1, v1, _ = one.download()
    ug, vg, _ = grp.download()
    assert _relmax(ug, u1) < 1e-11 and _relmax(vg, v1) < 1e-8
    T1, Tg = one.energy(potential=False)[2], grp.energy(potential=False)[2]
    assert abs(Tg - T1) < 1e-9 * T1
    one.close(); grp.close()


@pytest.mark.parametrize("world", [2, 3])
def test_pair_sharded_water_group_with_the_default_cutoff(oracle, world):
    """SPC/Fw water with the reference's default electrostatic cutoff of 0.49 L (no cell list can serve it): the group
    shards the UNORDERED Coulomb pairs (periodic Newton's-third-law kernel, every rank its ring offsets, partial rows summed
    at the owners), while the O-O Lennard-Jones term and the bonds / angle are evaluated by the owner of a block alone."""
    w = wl.water_omm(14)            # 2,744 molecules = 8,232 atoms: above symmetric_min_n
    spec = dict(ms=w["ms"], qs=w["qs"], water=True, bc=("cubic", w["L"]), lj=w["lj"], coulomb=w["coulomb"], spcfw=w["spcfw"],
                thermostat=dict(kind="berendsen", T=300.0, tau=0.05, kB=w["kB"], N=3 * w["nmol"], Nc=2 * w["nmol"]))
    assert w["coulomb"]["R"] > w["L"] / 3
    rng = np.random.Generator(np.random.Philox(21))
    u, v = F(w["u"] + 0.004 * rng.standard_normal(w["u"].shape)), F(w["v"])
    one = make_context(spec)
    grp = _group(spec, [0] * world)
    a1 = one.accel(u, F(v.copy())).copy()
    ag = grp.accel(u, F(v.copy()))
    assert grp.info("group_mode") == 1
    assert _relmax(ag, a1) < 1e-12
    mols = np.sort(np.random.Generator(np.random.Philox(3)).choice(w["nmol"], 40, replace=False))
    targets = (3 * mols[:, None] + np.arange(3)[None, :]).ravel()
    nothermo = {k: val for k, val in spec.items() if k != "thermostat"}
    ref = make_oracle(oracle, nothermo).accel_molecules(u, mols, NT)
    g0 = _group(nothermo, [0] * world)
    assert _relmax(g0.accel(u)[:, targets], ref) < 1e-11
    g0.close()
    dt, nsteps = w["dt"], 20
    one.upload(u, v); one.step_vv(dt, nsteps)
    grp.upload(u, v); grp.step_vv(dt, nsteps)
    u1, v1, _ = one.download()
    ug, vg, _ = grp.download()
    assert _relmax(ug, u1) < 1e-11 and _relmax(vg, v1) < 1e-8
    one.close(); grp.close()


@pytest.mark.parametrize("kind", ["coulomb", "dipole"])
def test_group_langevin_euler_maruyama(kind):
    """Config 5 over a group: the noise is keyed by (seed, step, global column), so the group walks the
    single-context trajectory; charges use pair sharding, dipoles target blocks."""
    n = 9216
    w = wl.charged_lattice(n) if kind == "coulomb" else wl.dipole_lattice(n)
    spec = dict(ms=w["ms"], qs=w.get("qs"), mm=w.get("mm"), thermostat=dict(kind="langevin", T=90.0, gamma=10.0, kB=1.38e-23))
    if kind == "coulomb":
        spec["coulomb"] = dict(k=w["coulomb"]["k"], R=np.inf)
    else:
        spec["dipole"] = w["dipole"]
    u, v = F(w["u"]), F(w["v"])
    dt, nsteps = 1e-9, 6
    (u1, v1, a1), _, _ = _single_run(spec, u, v, dt, nsteps, em=True, seed=77)
    grp = _group(spec, [0, 0])
    grp.upload(u, v)
    assert grp.info("group_mode") == (1 if kind == "coulomb" else 2)
    grp.step_em(dt, nsteps, 77)
    ug, vg, ag = grp.download(want_dv=True)
    assert _relmax(ug, u1) < 1e-13 and _relmax(vg, v1) < 1e-11 and _relmax(ag, a1) < 1e-10
    grp.close()


def test_group_andersen_matches_single_context():
    spec, u, v = _argon(8, 9, hot=1.0)
    spec["thermostat"] = dict(kind="andersen", T=90.0, nu=0.1 / 2e-3, kB=1.0 / 120.0)
    dt, nsteps = 2e-3, 25
    one = make_context(spec); one.set_seed(5); one.upload(u, v); one.step_vv(dt, nsteps)
    u1, v1, _ = one.download()
    grp = _group(spec, [0, 0], group_mode=2); grp.set_seed(5); grp.upload(u, v)
    assert grp.info("group_mode") == 2
    grp.step_vv(dt, nsteps)
    ug, vg, _ = grp.download()
    assert _relmax(ug, u1) < 1e-11 and _relmax(vg, v1) < 1e-8
    one.close(); grp.close()


@pytest.mark.parametrize("devices", [0, [0, 0]])
def test_run_vv_streams_frames_through_the_abi(devices):
    """nbx_run_vv (the saveat path of run_simulation): the frames equal what step + download give, on one context and on a
    slab group."""
    spec, u, v = _argon(8, 13, hot=2.0)
    dt = 2e-3
    a = make_context(spec, device=devices)
    a.upload(u, v)
    uf, vf = a.run_vv(dt, 25, save_every=10)
    assert uf.shape == (3, 3, u.shape[1]) and vf.shape == uf.shape
    b = make_context(spec)
    if devices != 0:
        b.set_option("verlet_banked", 0)   # the list order of the slabs
    b.upload(u, v)
    for k, n in enumerate((10, 10, 5)):
        b.step_vv(dt, n)
        ub, vb, _ = b.download()
        assert np.array_equal(uf[k], ub) and np.array_equal(vf[k], vb)
    with pytest.raises(_lib.NbxError):
        a.lib.nbx_run_vv  # noqa: B018  (symbol exists)
        a._ck(a.lib.nbx_run_vv(a.h, dt, 30, 10, uf.ctypes.data_as(_lib._dp), None, 2, None))   # capacity: 3 frames needed
    a.close(); b.close()
