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
