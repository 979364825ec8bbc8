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

    def __init__(self, make_ctx, u, v, world, thermostat, direct=False):
        import torch

        from nbody_b200.parallel import CudaEngine

        self.torch = torch
        self.world = world
        self.thermostat = thermostat
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
                e.slab_pack()
            self._exchange()
            self.counts = [e.slab_unpack() for e in self.engines]
            migrated += sum(c[2] + c[3] for c in self.counts)
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
