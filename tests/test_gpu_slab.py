"""Slab decomposition (csrc/nbx_slab.cu + parallel.SlabStepper) on ONE GPU: K contexts play the K ranks of a
ring and the exchange is a device-to-device copy, so pack / migrate / halo / unpack are covered by the
single-GPU `-m gpu` suite; tests/test_gpu_multi.py repeats it over NCCL when >= 2 GPUs are visible.

Parity: the slabs together must reproduce the single-context trajectory.  Without a thermostat the match is
bit-exact (cell order is ranked by global particle id, so every force sum has the same order); with
Berendsen the global sum m v^2 is added in a different order (per-slab partials), hence 1e-12.
"""
import numpy as np
import pytest

import nbody_b200.workloads as wl
from nbody_b200 import _lib
from tests._common import F

pytestmark = pytest.mark.gpu


class LocalRing:
    """K slab contexts on one device; neighbour exchange by tensor copies."""

    def __init__(self, make_ctx, u, v, world, thermostat, direct=False, soft=1.0):
        import torch

        from nbody_b200.parallel import CudaEngine

        self.torch = torch
        self.world = world
        self.thermostat = thermostat
        self.soft = soft
        self.rebuilds = 0
        self.engines = []
        for r in range(world):
            ctx = make_ctx()
            eng = CudaEngine(ctx, 0)
            ctx.upload(u, v)
            eng.slab_init(r, world)
            self.engines.append(eng)
        self.direct = direct and world > 1
        if self.direct:  # same process: the neighbours' receive areas as plain device pointers
            rx = [e.ctx.slab_rx()[0] for e in self.engines]
            for r, e in enumerate(self.engines):
                e.ctx.slab_connect(left_ptr=rx[(r - 1) % world], right_ptr=rx[(r + 1) % world])
        else:
            self.bufs = [e.slab_buffers() for e in self.engines]
        for e in self.engines:
            e.slab_pack()
        self._exchange()
        self.counts = [e.slab_unpack() for e in self.engines]
        self.verlet = all(e.slab_verlet() for e in self.engines)
        self.flags = [torch.zeros(2, dtype=torch.int32, device="cuda") for _ in self.engines]
        self.force_rebuild = False
        if self.verlet:  # as parallel.SlabStepper: record the halo lists, build the Verlet lists from x(0)
            for e in self.engines:
                e.slab_mark("slab_record_halo")
                e.slab_pack()
            self._exchange()
            self.counts = [e.slab_unpack() for e in self.engines]
            for e in self.engines:
                e.slab_prime()

    def _rebuild_wanted(self):
        """The collective decision of parallel.SlabStepper, taken synchronously here (no lag): max over the ranks."""
        for e, f in zip(self.engines, self.flags):
            f.zero_()
            e.slab_verlet_check(f, self.soft)
        self.torch.cuda.synchronize()
        soft = max(int(f[0]) for f in self.flags)
        hard = max(int(f[1]) for f in self.flags)
        want = self.force_rebuild or bool(soft)
        assert want or not hard, "a list went stale without a rebuild"
        return want

    def _exchange(self):
        if self.world == 1 or self.direct:
            return
        for r in range(self.world):
            left, right = (r - 1) % self.world, (r + 1) % self.world
            self.bufs[left][3].copy_(self.bufs[r][0])   # send-to-left  -> left's recv-from-right
            self.bufs[right][2].copy_(self.bufs[r][1])  # send-to-right -> right's recv-from-left

    def step(self, dt, nsteps):
        migrated = 0
        for _ in range(nsteps):
            for e in self.engines:
                e.vv_begin(dt)
            if self.verlet and not self._rebuild_wanted():
                for e in self.engines:
                    e.slab_refresh_send()
                self._exchange()
                for e in self.engines:
                    e.slab_refresh_recv()
            else:
                for e in self.engines:
                    e.slab_pack()
                self._exchange()
                self.counts = [e.slab_unpack() for e in self.engines]
                migrated += sum(c[2] + c[3] for c in self.counts)
                if self.verlet:  # second round of a rebuild: halo including the arrivals, recorded for the refreshes
                    for e in self.engines:
                        e.slab_mark("slab_record_halo")
                        e.slab_pack()
                    self._exchange()
                    self.counts = [e.slab_unpack() for e in self.engines]
                    for e in self.engines:
                        e.slab_mark("slab_rebuild")
                    self.force_rebuild = False
                    self.rebuilds += 1
            for e in self.engines:
                e.vv_forces()
                e.vv_finish(dt)
            if self.thermostat and self.world > 1:
                total = sum(float(e.scalars()[0].item()) for e in self.engines)
                for e in self.engines:
                    e.scalars()[0] = total
        return migrated

    def gather(self, n):
        out = [np.zeros((3, n), order="F") for _ in range(3)]
        seen = np.zeros(n, dtype=int)
        for e in self.engines:
            gid, u, v, dv = e.slab_download()
            seen[gid] += 1
            for dst, src in zip(out, (u, v, dv)):
                dst[:, gid] = src
        assert (seen == 1).all(), "slab ownership is not a partition"
        return out

    def close(self):
        for e in self.engines:
            e.ctx.close()


def _argon(cells, seed, hot=3.0):
    w = wl.fcc_argon_reduced(cells)
    rng = np.random.Generator(np.random.Philox(seed))
    u = F(w["u"] + 0.05 * rng.standard_normal(w["u"].shape))
    v = F(hot * w["v"])
    return w, u, v


@pytest.mark.parametrize("direct", [False, True])
@pytest.mark.parametrize("world", [1, 2, 4])
@pytest.mark.parametrize("thermostat", [False, True])
def test_slabs_reproduce_the_single_context_trajectory(world, thermostat, direct):
    w, u, v = _argon(12, 5)  # 6,912 atoms, 9 cell layers
    n = u.shape[1]
    dt, steps = 2e-3, 60

    def make_ctx():
        ctx = _lib.Context(0)
        ctx.system(w["ms"])
        ctx.boundary(_lib.BC_CUBIC, [w["L"]])
        ctx.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"])
        ctx.set_option("verlet_skin_permille", 0)  # one summation order everywhere: slabs rescan the cells every step
        if thermostat:
            ctx.thermostat(_lib.THERMO_BERENDSEN, 90.0, 20 * dt, w["kB"], n, 0)
        return ctx

    ref = make_ctx()
    ref.upload(u, v)
    ref.step_vv(dt, steps)
    ur, vr, ar = ref.download(want_dv=True)
    ref.close()

    ring = LocalRing(make_ctx, u, v, world, thermostat, direct)
    assert sum(c[0] for c in ring.counts) == n
    if world > 1:
        assert all(c[1] > 0 for c in ring.counts)  # ghosts
    migrated = ring.step(dt, steps)
    ug, vg, ag = ring.gather(n)
    ring.close()
    if world > 1:
        assert migrated > 0, "the run was meant to move particles across slab faces"
    if not thermostat:
        assert np.array_equal(ug, ur) and np.array_equal(vg, vr) and np.array_equal(ag, ar)
    else:
        for a, b in ((ug, ur), (vg, vr), (ag, ar)):
            assert np.abs(a - b).max() <= 1e-11 * np.abs(b).max()


@pytest.mark.parametrize("direct", [False, True])
@pytest.mark.parametrize("world", [1, 2, 4])
@pytest.mark.parametrize("thermostat", [False, True])
def test_slabs_with_verlet_lists_reproduce_the_single_context_trajectory(world, thermostat, direct):
    """Verlet lists inside the slabs: local slots stay put between collective rebuilds, only the halo positions travel.
    With the single context's own criterion (rebuild as soon as a particle has moved skin/2, decided without lag) the
    rebuilds fall on the same steps, the lists hold the same partners in the same order, and the trajectory is
    bit-identical to the single-context Verlet run (one lane per target on both sides)."""
    w, u, v = _argon(12, 5)  # 6,912 atoms; 8 cell layers of edge >= R + skin
    n = u.shape[1]
    dt, steps = 2e-3, 60

    def make_ctx():
        ctx = _lib.Context(0)
        ctx.system(w["ms"])
        ctx.boundary(_lib.BC_CUBIC, [w["L"]])
        ctx.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"])
        ctx.set_option("verlet_lanes", 1)
        # the single context would otherwise order its lists by record position (slot number mod 4: local numbering),
        # which a slab never does -- the comparison is between equal summation orders
        ctx.set_option("verlet_banked", 0)
        if thermostat:
            ctx.thermostat(_lib.THERMO_BERENDSEN, 90.0, 20 * dt, w["kB"], n, 0)
        return ctx

    ref = make_ctx()
    ref.upload(u, v)
    ref.step_vv(dt, steps)
    ur, vr, ar = ref.download(want_dv=True)
    ref_rebuilds = ref.info("verlet_rebuilds")
    ref.close()

    ring = LocalRing(make_ctx, u, v, world, thermostat, direct)
    assert ring.verlet
    migrated = ring.step(dt, steps)
    ug, vg, ag = ring.gather(n)
    ring.close()
    assert 3 <= ring.rebuilds < steps // 2   # several collective rebuilds, halo refreshes in between
    assert ring.rebuilds == ref_rebuilds - 1  # (both also built once from x(0): at upload / after the distribution)
    if world > 1:
        assert migrated > 0
    if not thermostat:
        assert np.array_equal(ug, ur) and np.array_equal(vg, vr) and np.array_equal(ag, ar)
    else:
        for a, b in ((ug, ur), (vg, vr), (ag, ar)):
            assert np.abs(a - b).max() <= 1e-11 * np.abs(b).max()


def test_slab_stepper_lagged_rebuilds_single_rank():
    """parallel.SlabStepper itself (one rank, no process group): the rebuild decision is read two steps late at a soft
    limit (0.4 x skin/2 here: the system is hot, the fastest atoms cover 0.04 sigma per step), so rebuilds fall on other
    steps than in the single context -- the pair set is the same, the order of summation is not: 1e-9 after 80 hot steps."""
    from nbody_b200.parallel import CudaEngine, SlabStepper

    w, u, v = _argon(10, 7)
    dt, steps = 2e-3, 80

    def make_ctx():
        ctx = _lib.Context(0)
        ctx.system(w["ms"])
        ctx.boundary(_lib.BC_CUBIC, [w["L"]])
        ctx.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"])
        return ctx

    ref = make_ctx()
    ref.upload(u, v)
    ref.step_vv(dt, steps)
    ur, vr, _ = ref.download()
    ref.close()
    ctx = make_ctx()
    ctx.upload(u, v)
    st = SlabStepper(CudaEngine(ctx, 0), soft=0.4)
    assert st.verlet
    st.step(dt, steps)
    ug, vg, _ = st.gather(u.shape[1])
    ctx.close()
    assert 3 <= st.rebuilds < steps
    assert np.abs(ug - ur).max() <= 1e-9 * np.abs(ur).max() and np.abs(vg - vr).max() <= 1e-9 * np.abs(vr).max()


def test_slab_stepper_raises_when_a_list_goes_stale_before_the_collective_rebuild():
    """The rebuild decision is read two steps late.  Particles fast enough to cross skin/2 inside that window must not
    be integrated with a stale list silently: the hard flag of the displacement check raises."""
    from nbody_b200.parallel import CudaEngine, SlabStepper

    w, u, v = _argon(8, 11, hot=40.0)  # velocities ~ 40 sigma per time unit: 0.08 sigma per step, skin/2 = 0.11 sigma
    ctx = _lib.Context(0)
    ctx.system(w["ms"])
    ctx.boundary(_lib.BC_CUBIC, [w["L"]])
    ctx.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"])
    ctx.upload(u, v)
    st = SlabStepper(CudaEngine(ctx, 0))
    assert st.verlet
    with pytest.raises(RuntimeError, match="skin/2"):
        st.step(2e-3, 12)
    ctx.close()


def test_slab_errors_are_reported():
    w, u, v = _argon(8, 2)
    ctx = _lib.Context(0)
    ctx.system(w["ms"])
    ctx.boundary(_lib.BC_CUBIC, [w["L"]])
    ctx.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"])
    with pytest.raises(_lib.NbxError):
        ctx.slab_init(0, 2)  # nothing uploaded
    ctx.upload(u, v)
    with pytest.raises(_lib.NbxError):
        ctx.slab_init(0, 4)  # 6 layers cannot give 4 slabs of >= 2 layers
    with pytest.raises(_lib.NbxError):
        ctx.slab_pack()
    ctx.slab_init(1, 2)
    ctx.slab_pack()
    with pytest.raises(_lib.NbxError):
        ctx.slab_pack()  # the previous pack was not completed
    with pytest.raises(_lib.NbxError):
        ctx.step_vv(1e-3, 1)  # whole-system entry points are closed on a slab context
    with pytest.raises(_lib.NbxError):
        ctx.upload(u, v)
    ctx.close()
    g = _lib.Context(0)
    g.system(w["ms"])
    g.add_gravity(1.0)
    g.upload(u, v)
    with pytest.raises(_lib.NbxError):
        g.slab_init(0, 2)  # not a cutoff system in a cubic box
    g.close()
