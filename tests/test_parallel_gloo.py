"""The N>1 host path on CPU: world_size-2 gloo process group driving nbody_b200.parallel.ShardedStepper
with a NumPy engine whose forces come from the oracle (test infrastructure).  Checks the sharding plan,
the in-place and the padded all-gather of the SoA position rows, and the temperature all-reduce."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_partition_tiles_the_range():
    from nbody_b200.parallel import numpy_reference_partition_check, partition

    for n, w, m in [(10, 3, 1), (262144, 8, 1), (1000, 7, 1), (30, 4, 3), (98304, 8, 3), (5, 8, 1)]:
        assert numpy_reference_partition_check(n, w, m)
        for r in range(w):
            lo, hi, per = partition(n, w, r, m)
            assert lo % m == 0 and (hi % m == 0 or hi == n) and hi - lo <= per


def test_rebuild_schedule_of_the_slab_verlet_lists():
    """parallel.RebuildSchedule: decisions come `lag` steps late, flags taken against lists that were rebuilt since are
    ignored, and a hard flag that no rebuild covered raises."""
    from nbody_b200.parallel import RebuildSchedule

    flags = {}
    asked = []

    def flags_of(j):
        asked.append(j)
        return flags.get(j, (0, 0))

    s = RebuildSchedule(lag=2)
    rebuilt_at = []
    flags[5] = (1, 0)           # at step 5 some particle passed the soft limit
    flags[6] = (1, 0)           # ... and of course still has at 6 and 7: those checks must not trigger again
    flags[7] = (1, 1)           # by step 7 it is even past skin/2 relative to the OLD lists -- which step 7 replaced
    for k in range(14):
        assert s.k == k
        due = s.due(flags_of)
        if due:
            rebuilt_at.append(k)
        s.advance(due)
    assert rebuilt_at == [7]                       # read two steps late; the soft flag of step 6 does not trigger again
    assert asked == [0, 1, 2, 3, 4, 5, 6, 8, 9, 10, 11]  # step 7's check preceded its own rebuild: moot
    assert s.rebuilds == 1 and s.last_rebuild == 7

    s = RebuildSchedule(lag=2)
    flags.clear()
    flags[5] = (1, 0)
    flags[6] = (1, 1)           # step 6 still ran on the old lists and was already past skin/2: must not go unnoticed
    with pytest.raises(RuntimeError, match="step 6"):
        for k in range(12):
            s.advance(s.due(flags_of))

    s = RebuildSchedule(lag=2)
    s.force = True
    assert s.due(flags_of) and not s.force is False
    s.advance(True)
    assert not s.force and s.last_rebuild == 0

    s = RebuildSchedule(lag=2)
    flags.clear()
    flags[3] = (1, 1)           # past skin/2 at a step whose forces came from the lists: not recoverable
    with pytest.raises(RuntimeError, match="skin/2"):
        for k in range(8):
            s.advance(s.due(flags_of))


class NumpyEngine:
    """Same duck type as parallel.CudaEngine, on host memory.  targets mode: forces from the CPU oracle;
    pairs mode: this rank's share of the unordered pairs {i, j} ((i + j) % world == rank), partial
    accelerations of ALL particles (plain NumPy gravity)."""

    def __init__(self, spec, u, v, thermostat=False):
        import torch

        from oracle import nbody_oracle as orc
        from tests._common import make_oracle

        self.sys = make_oracle(orc, spec)
        self.G = spec["gravity"]["G"]
        self.n = u.shape[1]
        self.ld = ((self.n + 7) // 8) * 8
        self.ms = np.asarray(spec["ms"], dtype=np.float64)
        self.pos = np.zeros((3, self.ld))
        self.pos[:, :self.n] = u
        self.vel = np.array(v, dtype=np.float64)
        self.accbuf = np.zeros((3, self.ld))
        self.scal = np.zeros(16)
        self.needs_temperature = thermostat
        self._pos_t = torch.from_numpy(self.pos)
        self._acc_t = torch.from_numpy(self.accbuf)
        self._scal_t = torch.from_numpy(self.scal)
        self.lo, self.hi = 0, self.n
        self.pair = None
        self.acc_old = None

    def shard(self, lo, hi):
        self.lo, self.hi = lo, hi
        # a(0): every rank holds the whole state at start
        u = np.asfortranarray(self.pos[:, :self.n])
        self.accbuf[:, lo:hi] = self.sys.accel_targets(u, np.arange(lo, hi))

    def shard_pairs(self, rank, world):
        self.pair = (rank, world)

    def pos_rows(self):
        return self._pos_t

    def acc_rows(self):
        return self._acc_t

    def scalars(self):
        return self._scal_t

    def vv_begin(self, dt):
        s = slice(self.lo, self.hi)
        self.pos[:, s] += dt * self.vel[:, s] + 0.5 * dt * dt * self.accbuf[:, s]

    def vv_forces(self):
        self.acc_old = self.accbuf[:, self.lo:self.hi].copy()
        x = self.pos[:, :self.n]
        if self.pair is None:
            self.accbuf[:, self.lo:self.hi] = self.sys.accel_targets(np.asfortranarray(x), np.arange(self.lo, self.hi))
            return
        rank, world = self.pair
        i, j = np.triu_indices(self.n, 1)
        keep = (i + j) % world == rank
        i, j = i[keep], j[keep]
        d = x[:, j] - x[:, i]
        g = self.G / np.sum(d * d, axis=0) ** 1.5
        a = np.zeros((3, self.n))
        for c in range(3):
            np.add.at(a[c], i, g * self.ms[j] * d[c])
            np.add.at(a[c], j, -g * self.ms[i] * d[c])
        self.accbuf[:, :self.n] = a

    def vv_finish(self, dt):
        s = slice(self.lo, self.hi)
        self.vel[:, s] += 0.5 * dt * (self.acc_old + self.accbuf[:, s])
        self.scal[0] = float(np.dot(self.ms[s], (self.vel[:, s] ** 2).sum(axis=0)))

    # RHS drop-in: whole (targets mode) and split-phase (pairs mode)
    def accel(self, u, out=None):
        out = np.zeros((3, self.n), order="F") if out is None else out
        out[:] = 0.0
        out[:, self.lo:self.hi] = self.sys.accel_targets(np.asfortranarray(u), np.arange(self.lo, self.hi))
        return out

    def accel_begin(self, u):
        self.pos[:, :self.n] = u
        self.vv_forces()

    def accel_end(self, out=None):
        out = np.zeros((3, self.n), order="F") if out is None else out
        out[:] = 0.0
        out[:, self.lo:self.hi] = self.accbuf[:, self.lo:self.hi]
        return out


def _worker(rank, world, port, n, out_dir, mode):
    import torch.distributed as dist

    import nbody_b200.workloads as wl
    from nbody_b200.parallel import ShardedStepper

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    u, v, ms = wl.plummer(n, seed=3)
    spec = dict(ms=ms, gravity=dict(G=1.0))
    eng = NumpyEngine(spec, u, v, thermostat=True)
    st = ShardedStepper(eng, mode=mode)
    st.step(1e-3, 3)
    np.save(os.path.join(out_dir, f"pos{rank}.npy"), eng.pos[:, :n])
    np.save(os.path.join(out_dir, f"scal{rank}.npy"), eng.scal[:1])
    np.save(os.path.join(out_dir, f"vel{rank}.npy"), eng.vel[:, st.lo:st.hi])
    dist.destroy_process_group()


def _accel_worker(rank, world, port, n, out_dir, mode):
    import torch.distributed as dist

    import nbody_b200.workloads as wl
    from nbody_b200.parallel import ShardedStepper

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    u, v, ms = wl.plummer(n, seed=3)
    st = ShardedStepper(NumpyEngine(dict(ms=ms, gravity=dict(G=1.0)), u, v), mode=mode)
    rng = np.random.Generator(np.random.Philox(9))
    x = np.asfortranarray(u + 0.01 * rng.standard_normal(u.shape))  # the integrator asks for other positions
    dv = st.accel(x)
    np.save(os.path.join(out_dir, f"dv{rank}.npy"), dv)
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["targets", "pairs"])
def test_two_rank_rhs_dropin_matches_serial(tmp_path, mode):
    """ShardedStepper.accel: every rank returns its own block of soode_system!'s dv (src/nbody_to_ode.jl:474-488);
    pairs mode goes through accel_begin / reduce / accel_end."""
    import torch.multiprocessing as mp

    import nbody_b200.workloads as wl
    from oracle import nbody_oracle as orc
    from tests._common import make_oracle

    n, world = 64, 2
    mp.spawn(_accel_worker, args=(world, _free_port(), n, str(tmp_path), mode), nprocs=world, join=True)
    u, v, ms = wl.plummer(n, seed=3)
    rng = np.random.Generator(np.random.Philox(9))
    x = np.asfortranarray(u + 0.01 * rng.standard_normal(u.shape))
    ref = make_oracle(orc, dict(ms=ms, gravity=dict(G=1.0))).rhs(x, v)
    total = sum(np.load(os.path.join(str(tmp_path), f"dv{r}.npy")) for r in range(world))
    assert np.abs(total - ref).max() <= 1e-12 * np.abs(ref).max()
    blocks = [np.load(os.path.join(str(tmp_path), f"dv{r}.npy")) for r in range(world)]
    assert not (blocks[0][:, n // 2:]).any() and not (blocks[1][:, :n // 2]).any()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("mode", ["targets", "pairs"])
@pytest.mark.parametrize("n", [64, 61])  # even shards -> in-place all-gather; uneven -> padded exchange
def test_two_rank_velocity_verlet_matches_serial(tmp_path, n, mode):
    import torch.multiprocessing as mp

    import nbody_b200.workloads as wl
    from oracle import nbody_oracle as orc
    from tests._common import make_oracle

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path), mode), nprocs=world, join=True)
    u, v, ms = wl.plummer(n, seed=3)
    s = make_oracle(orc, dict(ms=ms, gravity=dict(G=1.0)))
    ur, vr = orc.velocity_verlet(s, u, v, 1e-3, 3)
    p0, p1 = np.load(tmp_path / "pos0.npy"), np.load(tmp_path / "pos1.npy")
    assert np.array_equal(p0, p1)                       # every rank holds all positions after the gather
    assert np.allclose(p0, ur, rtol=1e-12, atol=1e-15)
    vel = np.concatenate([np.load(tmp_path / "vel0.npy"), np.load(tmp_path / "vel1.npy")], axis=1)
    assert np.allclose(vel, vr, rtol=1e-11, atol=1e-14)
    mv2 = float(np.dot(ms, (vr ** 2).sum(axis=0)))
    for r in range(world):                              # all-reduced sum m v^2 = the global value on every rank
        assert np.load(tmp_path / f"scal{r}.npy")[0] == pytest.approx(mv2, rel=1e-12)


# ------------------------------------------------------------------------------------------------
# slab decomposition: parallel.SlabStepper over gloo with a NumPy engine (forces from the oracle)
# ------------------------------------------------------------------------------------------------
class NumpySlabEngine:
    """Host restatement of the slab protocol of csrc/nbx_slab.cu behind the SlabStepper duck type:
    own particles of the x-layers [c0, c1) + ghost copies of the adjacent layers; one message per
    neighbour per step = [n_migrants, n_halo, migrant records (id x v a m), halo records (id x)]."""

    MIG, HALO = 11, 4

    def __init__(self, w, u, v, thermostat=None):
        import torch

        from oracle import nbody_oracle as orc

        self.torch, self.orc, self.w = torch, orc, w
        self.n_total = u.shape[1]
        self.L, self.R = w["L"], w["lj"]["R"]
        self.nc = int(np.floor(self.L / (self.R * (1 + 1e-6))))
        self.th = thermostat
        self.needs_temperature = thermostat is not None
        full = orc.System(w["ms"], bc=("cubic", self.L), lj=w["lj"])
        self.gid = np.arange(self.n_total)
        self.pos, self.vel = np.array(u), np.array(v)
        self.mass = np.array(w["ms"], dtype=np.float64)
        self.acc = full.rhs(np.asfortranarray(u), np.asfortranarray(v), 2)
        self.scal_np = np.zeros(16)
        self.scal_np[0] = float(np.dot(self.mass, (self.vel ** 2).sum(axis=0)))
        self._scal = torch.from_numpy(self.scal_np)
        if thermostat:
            self.acc += self._berendsen() * self.vel
        cap = 2 + self.n_total * (self.MIG + self.HALO)
        self.bufs = [torch.zeros(cap, dtype=torch.float64) for _ in range(4)]

    def _berendsen(self):
        T = self.scal_np[0] / (self.th["kB"] * 3 * self.n_total)
        return 0.5 / self.th["tau"] * (self.th["T"] / T - 1.0)

    def _layer(self, x):
        wv = x - self.L * np.floor(x / self.L)
        return np.clip((wv * (self.nc / self.L)).astype(int), 0, self.nc - 1)

    def scalars(self):
        return self._scal

    def slab_buffers(self):
        return self.bufs

    def slab_init(self, rank, world):
        lo = lambda r: r * self.nc // world
        self.c0, self.c1 = lo(rank), lo(rank + 1)
        self.first = True

    def slab_pack(self):
        self._pack(init=self.first)
        self.first = False

    def _pack(self, init):
        rel = (self._layer(self.pos[0]) - self.c0) % self.nc
        width = self.c1 - self.c0
        stay = rel < width
        left = ~stay & ((self.nc - rel) <= (rel - width + 1)) & (not init)
        right = ~stay & ~left & (not init)
        for buf, mig, halo in ((self.bufs[0], left, stay & (rel == 0)), (self.bufs[1], right, stay & (rel == width - 1))):
            rec = np.concatenate([self.gid[mig][None].astype(float), self.pos[:, mig], self.vel[:, mig], self.acc[:, mig],
                                  self.mass[mig][None]]).T.ravel()
            hrec = np.concatenate([self.gid[halo][None].astype(float), self.pos[:, halo]]).T.ravel()
            out = np.concatenate([[mig.sum(), halo.sum()], rec, hrec])
            buf.zero_()
            buf[:out.size] = self.torch.from_numpy(out)
        self.kept = [np.concatenate([self.gid[m][None].astype(float), self.pos[:, m]]) for m in (left, right)]
        for name in ("gid", "mass"):
            setattr(self, name, getattr(self, name)[stay])
        for name in ("pos", "vel", "acc"):
            setattr(self, name, getattr(self, name)[:, stay])

    def slab_check(self):
        return self._counts

    def slab_unpack(self, sync=True):
        nstay = self.gid.size
        ghosts = list(self.kept)
        arrivals = []
        for buf in (self.bufs[2], self.bufs[3]):
            b = buf.numpy()
            nm, nh = int(b[0]), int(b[1])
            rec = b[2:2 + nm * self.MIG].reshape(nm, self.MIG).T
            hrec = b[2 + nm * self.MIG:2 + nm * self.MIG + nh * self.HALO].reshape(nh, self.HALO).T
            arrivals.append(rec)
            ghosts.append(hrec)
        for rec in arrivals:
            self.gid = np.concatenate([self.gid, rec[0].astype(int)])
            self.pos = np.concatenate([self.pos, rec[1:4]], axis=1)
            self.vel = np.concatenate([self.vel, rec[4:7]], axis=1)
            self.acc = np.concatenate([self.acc, rec[7:10]], axis=1)
            self.mass = np.concatenate([self.mass, rec[10]])
        self.ghost_gid = np.concatenate([g[0].astype(int) for g in ghosts])
        self.ghost_pos = np.concatenate([g[1:4] for g in ghosts], axis=1)
        self._counts = [self.gid.size, self.ghost_gid.size, self.kept[0].shape[1], self.kept[1].shape[1],
                        arrivals[0].shape[1], arrivals[1].shape[1]]
        return self._counts

    def vv_begin(self, dt):
        self.pos = self.pos + dt * self.vel + 0.5 * dt * dt * self.acc

    def vv_forces(self):
        self.acc_old = self.acc
        m = self.gid.size
        x = np.asfortranarray(np.concatenate([self.pos, self.ghost_pos], axis=1))
        ms = np.concatenate([self.mass, np.ones(self.ghost_gid.size)])
        local = self.orc.System(ms, bc=("cubic", self.L), lj=self.w["lj"])
        self.acc = local.accel_targets(x, np.arange(m))

    def vv_finish(self, dt):
        if self.th:
            self.acc = self.acc + self._berendsen() * self.vel
        self.vel = self.vel + 0.5 * dt * (self.acc_old + self.acc)
        self.scal_np[0] = float(np.dot(self.mass, (self.vel ** 2).sum(axis=0)))

    def slab_download(self):
        return self.gid.copy(), self.pos.copy(), self.vel.copy(), self.acc.copy()


class NumpyVerletSlabEngine(NumpySlabEngine):
    """The Verlet-list slab protocol of csrc/nbx_slab.cu on the host: between collective rebuilds the own set and its
    numbering stay put and only the positions of the halo particles recorded at the rebuild travel; a rebuild is a
    migration round followed by a halo round that is remembered.  Forces are the oracle's over own + ghosts, i.e. exact
    as long as the ghost set recorded with the skin still covers everything within the cutoff -- which is what the
    displacement flags and the collective, lagged rebuild schedule of parallel.SlabStepper have to guarantee."""

    def __init__(self, w, u, v, thermostat=None, skin=0.2):
        super().__init__(w, u, v, thermostat)
        self.skin = skin
        self.nc = int(np.floor(self.L / ((self.R + skin) * (1 + 1e-6))))  # layers of edge >= R + skin
        self.device = self.torch.device("cpu")
        self.marks = set()
        self.ref = None
        self.halo_idx = [np.zeros(0, dtype=int), np.zeros(0, dtype=int)]
        self.refreshes = 0

    def slab_verlet(self):
        return True

    def slab_mark(self, key):
        self.marks.add(key)

    def _pack(self, init):
        record = "slab_record_halo" in self.marks
        self.marks.discard("slab_record_halo")
        if record:  # where the halo members will sit after the (order-preserving) compaction
            rel = (self._layer(self.pos[0]) - self.c0) % self.nc
            stay = rel < (self.c1 - self.c0)
            new_index = np.cumsum(stay) - 1
            self.halo_idx = [new_index[stay & (rel == 0)], new_index[stay & (rel == self.c1 - self.c0 - 1)]]
        super()._pack(init)

    def slab_prime(self):
        self.ref = self.pos.copy()

    def vv_forces(self):
        if "slab_rebuild" in self.marks:
            self.marks.discard("slab_rebuild")
            self.ref = self.pos.copy()
        super().vv_forces()

    def slab_verlet_check(self, flags, soft_fraction=0.75):
        f = flags.numpy()
        if self.ref is None or self.ref.shape != self.pos.shape:
            f[0] = 1
            return
        lim = 0.5 * self.skin * (1.0 - 1e-9)
        d2 = ((self.pos - self.ref) ** 2).sum(axis=0)
        if not (d2 <= (lim * soft_fraction) ** 2).all():
            f[0] = 1
        if not (d2 <= lim ** 2).all():
            f[1] = 1

    def slab_refresh_send(self):
        self.refreshes += 1
        for buf, idx in zip(self.bufs[:2], self.halo_idx):
            hrec = np.concatenate([self.gid[idx][None].astype(float), self.pos[:, idx]]).T.ravel()
            out = np.concatenate([[0, idx.size], hrec])
            buf.zero_()
            buf[:out.size] = self.torch.from_numpy(out)

    def slab_refresh_recv(self):
        pos, gids = [], []
        for buf in (self.bufs[2], self.bufs[3]):
            b = buf.numpy()
            assert int(b[0]) == 0
            nh = int(b[1])
            hrec = b[2:2 + nh * self.HALO].reshape(nh, self.HALO).T
            gids.append(hrec[0].astype(int))
            pos.append(hrec[1:4])
        assert np.array_equal(np.concatenate(gids), self.ghost_gid), "the halo changed without a collective rebuild"
        self.ghost_pos = np.concatenate(pos, axis=1)


def _verlet_slab_worker(rank, world, port, out_dir, thermo):
    import torch.distributed as dist

    from nbody_b200.parallel import SlabStepper

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    w, u, v, th, dt, steps = _verlet_slab_setup(thermo)
    eng = NumpyVerletSlabEngine(w, u, v, th, skin=0.2)
    st = SlabStepper(eng, soft=0.5)
    assert st.verlet and not st.merged
    rebuilt_at = []
    for k in range(steps):
        before = st.rebuilds
        st.step(dt, 1)
        if st.rebuilds > before:
            rebuilt_at.append(k)
    ug, vg, ag = st.gather(u.shape[1])
    if rank == 0:
        np.save(os.path.join(out_dir, "u.npy"), ug)
        np.save(os.path.join(out_dir, "v.npy"), vg)
    np.save(os.path.join(out_dir, f"rebuilt{rank}.npy"), np.array(rebuilt_at))
    np.save(os.path.join(out_dir, f"stat{rank}.npy"), np.array([eng.refreshes, st.counts[0], st.counts[1]]))
    dist.destroy_process_group()


def _verlet_slab_setup(thermo):
    import nbody_b200.workloads as wl

    w = wl.fcc_argon_reduced(5)  # 500 atoms, L = 8.55 sigma
    w["lj"] = dict(w["lj"], R=1.9)  # with the skin of 0.2: 4 layers of edge >= 2.1 -> 2 slabs of 2
    rng = np.random.Generator(np.random.Philox(9))
    u = np.asfortranarray(w["u"] + 0.05 * rng.standard_normal(w["u"].shape))
    v = np.asfortranarray(1.5 * w["v"])
    th = dict(kind="berendsen", T=90.0, tau=0.05, kB=w["kB"]) if thermo else None
    return w, u, v, th, 2e-3, 60


@pytest.mark.parametrize("thermo", [False, True])
def test_two_rank_slab_stepper_with_verlet_lists_matches_serial(tmp_path, thermo):
    """SlabStepper's Verlet-list protocol over gloo: both ranks take the same (lagged) rebuild decisions, only halo
    positions travel in between, and the trajectory is the serial one -- i.e. no interaction was ever missed."""
    import torch.multiprocessing as mp

    from oracle import nbody_oracle as orc
    from tests._common import make_oracle

    world = 2
    mp.spawn(_verlet_slab_worker, args=(world, _free_port(), str(tmp_path), thermo), nprocs=world, join=True)
    w, u, v, th, dt, steps = _verlet_slab_setup(thermo)
    spec = dict(ms=w["ms"], bc=("cubic", w["L"]), lj=w["lj"])
    if th:
        spec["thermostat"] = th
    ur, vr = orc.velocity_verlet(make_oracle(orc, spec), u, v, dt, steps)
    rebuilt = [np.load(tmp_path / f"rebuilt{r}.npy") for r in range(world)]
    stat = [np.load(tmp_path / f"stat{r}.npy") for r in range(world)]
    assert np.array_equal(rebuilt[0], rebuilt[1]) and 2 <= len(rebuilt[0]) < steps // 2   # collective, and not every step
    assert all(s[0] == steps - len(rebuilt[0]) for s in stat)                             # every other step only refreshed
    assert sum(s[1] for s in stat) == u.shape[1] and all(s[2] > 0 for s in stat)
    assert np.allclose(np.load(tmp_path / "u.npy"), ur, rtol=1e-11, atol=1e-13)
    assert np.allclose(np.load(tmp_path / "v.npy"), vr, rtol=1e-10, atol=1e-12)


def _slab_setup(thermo):
    import nbody_b200.workloads as wl

    w = wl.fcc_argon_reduced(5)  # 500 atoms, L = 8.55 sigma
    w["lj"] = dict(w["lj"], R=1.9)  # 4 cell layers -> 2 slabs of 2
    rng = np.random.Generator(np.random.Philox(8))
    u = np.asfortranarray(w["u"] + 0.05 * rng.standard_normal(w["u"].shape))
    v = np.asfortranarray(4.0 * w["v"])
    th = dict(kind="berendsen", T=90.0, tau=0.05, kB=w["kB"]) if thermo else None
    return w, u, v, th, 4e-3, 25


def _slab_worker(rank, world, port, out_dir, thermo):
    import torch.distributed as dist

    from nbody_b200.parallel import SlabStepper

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    w, u, v, th, dt, steps = _slab_setup(thermo)
    eng = NumpySlabEngine(w, u, v, th)
    st = SlabStepper(eng)
    moved = 0
    for _ in range(steps):
        st.step(dt, 1)
        moved += st.counts[2] + st.counts[3]
    ug, vg, ag = st.gather(u.shape[1])
    if rank == 0:
        np.save(os.path.join(out_dir, "u.npy"), ug)
        np.save(os.path.join(out_dir, "v.npy"), vg)
    np.save(os.path.join(out_dir, f"moved{rank}.npy"), np.array([moved, st.counts[0], st.counts[1]]))
    dist.destroy_process_group()


@pytest.mark.parametrize("thermo", [False, True])
def test_two_rank_slab_stepper_matches_serial(tmp_path, thermo):
    import torch.multiprocessing as mp

    from oracle import nbody_oracle as orc
    from tests._common import make_oracle

    world = 2
    mp.spawn(_slab_worker, args=(world, _free_port(), str(tmp_path), thermo), nprocs=world, join=True)
    w, u, v, th, dt, steps = _slab_setup(thermo)
    spec = dict(ms=w["ms"], bc=("cubic", w["L"]), lj=w["lj"])
    if th:
        spec["thermostat"] = th
    ur, vr = orc.velocity_verlet(make_oracle(orc, spec), u, v, dt, steps)
    moved = [np.load(tmp_path / f"moved{r}.npy") for r in range(world)]
    assert sum(m[0] for m in moved) > 0                      # particles crossed slab faces
    assert sum(m[1] for m in moved) == u.shape[1]            # ownership stays a partition
    assert all(m[2] > 0 for m in moved)                      # both slabs hold ghosts
    assert np.allclose(np.load(tmp_path / "u.npy"), ur, rtol=1e-11, atol=1e-13)
    assert np.allclose(np.load(tmp_path / "v.npy"), vr, rtol=1e-10, atol=1e-12)
