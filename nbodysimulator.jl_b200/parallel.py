"""Multi-GPU driver: one process per GPU, torch.distributed for the plumbing.

The reference is strictly serial; the data-parallel axis is the target-particle
index ``i`` of ``soode_system!`` (src/nbody_to_ode.jl:475).  Two decompositions:

* ``mode="targets"`` (any potential): rank r evaluates the target columns [lo_r, hi_r) against ALL
  sources.  One exchange per velocity-Verlet step: the all-gather of the updated positions (24 B per
  particle).
* ``mode="pairs"`` (unbounded gravity / Coulomb): the unordered pair set is split over the ranks
  (Newton's-third-law kernel, ring offsets k = rank mod world), every rank ends up with a partial
  acceleration of ALL particles, and a reduce-scatter of the acceleration rows (24 B per particle)
  completes the own block.  Two exchanges per step, 1.6x less arithmetic.

* ``SlabStepper`` (cutoff Lennard-Jones / Coulomb in a cubic periodic box): spatial decomposition
  along x into slabs of whole cell layers.  A rank owns the particles of its layers and keeps ghost
  copies of the two adjacent layers; per step ONE message to each neighbour carries the particles
  that crossed the face (full state) and the boundary layer's positions (halo).  No collective on the
  data path; the state never leaves the devices (csrc/nbx_slab.cu).

All add an 8-byte all-reduce of sum m v^2 when a thermostat needs the global temperature.

The plumbing works on an *engine* (duck-typed):
    engine.n                       particle columns
    engine.shard(lo, hi)           restrict the engine to the columns it integrates
    engine.shard_pairs(rank, world)  (pairs mode) select the engine's share of the pair set
    engine.pos_rows()  -> tensor   (3, ld) view of the SoA position rows (device memory, no copy)
    engine.acc_rows()  -> tensor   (3, ld) view of the SoA acceleration rows after vv_forces
    engine.scalars()   -> tensor   (16,) view of the scalar block ([0] = shard's sum m v^2)
    engine.vv_begin(dt) / engine.vv_forces() / engine.vv_finish(dt) / engine.needs_temperature
    engine.accel(u, out) / engine.accel_begin(u) / engine.accel_end(out)   RHS drop-in, whole and split-phase
``CudaEngine`` wraps a libnbody_b200 context; the CPU tests drive the same plumbing over gloo with a
NumPy engine (tests/test_parallel_gloo.py).
"""
from __future__ import annotations

import numpy as np


def partition(n: int, world: int, rank: int, multiple: int = 1):
    """Contiguous target range of ``rank``: equal blocks of ceil(n/world) rounded up to ``multiple``
    (3 for water: molecules stay whole)."""
    per = -(-n // world)
    per = -(-per // multiple) * multiple
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi, per


class CudaEngine:
    """libnbody_b200 context + zero-copy torch views of its device state."""

    def __init__(self, ctx, device_index: int):
        import torch

        self.ctx = ctx
        self.n = ctx.n
        self.device = torch.device("cuda", device_index)
        # thermostats with an RHS term read the GLOBAL sum m v^2: the steppers all-reduce the shard's sum when this is set
        self.needs_temperature = ctx.info("thermostat") in (1, 2)  # NBX_THERMO_BERENDSEN, NBX_THERMO_NOSEHOOVER
        # run the library on torch's current stream so NCCL calls issued by torch are ordered with it
        # (torch's default stream has handle 0 == "own stream" in the C ABI; 0x1 is cudaStreamLegacy)
        ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream or 1)
        self._views = {}

    def _view(self, which, shape):
        import torch

        ptr, ld = self.ctx.device_ptr(which)
        key = (which, ptr)
        if key not in self._views:
            shape = tuple(ld if s is None else s for s in shape)

            class _Raw:
                __cuda_array_interface__ = {"shape": shape, "typestr": "<f8", "data": (ptr, False), "version": 3,
                                            "strides": None}

            self._views[key] = torch.as_tensor(_Raw(), device=self.device)
        return self._views[key]

    def shard(self, lo, hi):
        self.ctx.shard(lo, hi)

    def shard_pairs(self, rank, world):
        self.ctx.shard_pairs(rank, world)

    def pos_rows(self):
        return self._view(0, (3, None))

    def acc_rows(self):
        return self._view(2, (3, None))  # acc and acc_old swap every step: looked up by pointer each time

    def scalars(self):
        return self._view(3, (16,))

    def vv_begin(self, dt):
        self.ctx.vv_begin(dt)

    def vv_forces(self):
        self.ctx.vv_forces()

    def accel(self, u, out=None):
        return self.ctx.accel(u, out=out)

    def accel_begin(self, u):
        self.ctx.accel_begin(u)

    def accel_end(self, out=None):
        return self.ctx.accel_end(out)

    def vv_finish(self, dt):
        self.ctx.vv_finish(dt)

    # -- slab decomposition (SlabStepper) -----------------------------------------------------------
    def _raw(self, ptr, count):
        import torch

        class _Raw:
            __cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3,
                                        "strides": None}

        return torch.as_tensor(_Raw(), device=self.device)

    def slab_init(self, rank, world):
        self.ctx.slab_init(rank, world)

    def slab_connect(self, stepper):
        """Try to map the neighbours' receive areas (CUDA IPC over NVLink) so that the pack kernel stores its
        messages straight into them.  Collective over the stepper's group; returns True when EVERY rank
        connected (otherwise all ranks stay on the host-driven send/recv exchange)."""
        import torch

        dist = stepper.dist
        if stepper.world == 1:
            return True
        if dist.get_backend(stepper.group) != "nccl":
            return False
        ok = 1
        try:
            _, _, handle = self.ctx.slab_rx(want_handle=True)
        except Exception:
            handle, ok = None, 0
        handles = [None] * stepper.world
        dist.all_gather_object(handles, handle, group=stepper.group)
        if ok and all(h is not None for h in handles):
            try:
                self.ctx.slab_connect(handles[stepper.left], handles[stepper.right])
            except Exception:
                ok = 0
        else:
            ok = 0
        t = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=stepper.group)
        if int(t.item()) == 0:
            self.ctx.slab_connect()  # disconnect: every rank must use the same exchange
            return False
        dist.barrier(group=stepper.group)  # every receive area is mapped before anyone stores into one
        return True

    def slab_buffers(self):
        """[send-to-left, send-to-right, recv-from-left, recv-from-right] as flat float64 device tensors."""
        return [self._raw(*self.ctx.slab_buffer(k)) for k in range(4)]

    def slab_pack(self):
        self.ctx.slab_pack()

    def slab_unpack(self, sync=True):
        return self.ctx.slab_unpack(sync)

    def slab_check(self):
        return self.ctx.slab_check()

    # Verlet lists inside the slab (csrc/nbx_slab.cu): collective rebuilds, halo refresh in between
    def slab_verlet(self):
        return bool(self.ctx.info("slab_verlet"))

    def slab_verlet_check(self, flags, soft_fraction=0.75):
        """flags: int32 device tensor of two elements (ORed into)."""
        self.ctx.slab_verlet_check(flags.data_ptr(), soft_fraction)

    def slab_refresh_send(self):
        self.ctx.slab_refresh_send()

    def slab_refresh_recv(self):
        self.ctx.slab_refresh_recv()

    def slab_prime(self):
        self.ctx.slab_prime()

    def slab_step_begin(self, dt, soft_fraction=0.75):
        self.ctx.slab_step_begin(dt, soft_fraction)

    def slab_step_end(self, dt, refresh):
        self.ctx.slab_step_end(dt, refresh)

    def slab_mark(self, key):
        self.ctx.set_option(key, 1)  # "slab_record_halo" before the second pack of a rebuild, "slab_rebuild" before its forces

    def slab_download(self):
        return self.ctx.slab_download()


class ShardedStepper:
    """Velocity Verlet over a process group (see the module docstring for the two modes)."""

    def __init__(self, engine, group=None, multiple: int = 1, mode: str = "targets"):
        import torch.distributed as dist

        if mode not in ("targets", "pairs"):
            raise ValueError(mode)
        self.dist = dist
        self.engine = engine
        self.group = group
        self.mode = mode
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.lo, self.hi, self.per = partition(engine.n, self.world, self.rank, multiple)
        engine.shard(self.lo, self.hi)
        if mode == "pairs":
            engine.shard_pairs(self.rank, self.world)
        self.even = self.per * self.world == engine.n
        self.backend = dist.get_backend(group)
        self._tmp = None

    def _all_gather_positions(self):
        import torch

        rows = self.engine.pos_rows()
        n, per, dist = self.engine.n, self.per, self.dist
        if self.world == 1:
            return
        if self.even:
            # in place: each rank's block already sits at its final offset of the row
            for d in range(3):
                dist.all_gather_into_tensor(rows[d, :n], rows[d, self.lo:self.hi], group=self.group)
            return
        if self._tmp is None:
            self._tmp = (torch.zeros(3 * per, dtype=rows.dtype, device=rows.device),
                         torch.zeros(self.world * 3 * per, dtype=rows.dtype, device=rows.device))
        send, recv = self._tmp
        cnt = self.hi - self.lo
        for d in range(3):
            send[d * per:d * per + cnt] = rows[d, self.lo:self.hi]
        dist.all_gather_into_tensor(recv, send, group=self.group)
        recv = recv.view(self.world, 3, per)
        for r in range(self.world):
            lo, hi = min(n, r * per), min(n, r * per + per)
            if hi > lo and r != self.rank:
                rows[:, lo:hi] = recv[r, :, :hi - lo]

    def _sum_partial_accelerations(self):
        """pairs mode: acc rows hold this rank's partial sums for ALL particles -> the own block gets the
        sum over ranks (reduce-scatter in place on NCCL with equal blocks, all-reduce otherwise)."""
        rows = self.engine.acc_rows()
        n, dist = self.engine.n, self.dist
        if self.world == 1:
            return
        if self.even and self.backend == "nccl":
            for d in range(3):
                dist.reduce_scatter_tensor(rows[d, self.lo:self.hi], rows[d, :n], op=dist.ReduceOp.SUM,
                                           group=self.group)
        else:
            for d in range(3):
                dist.all_reduce(rows[d, :n], op=dist.ReduceOp.SUM, group=self.group)

    def accel(self, u, out=None):
        """The RHS drop-in (soode_system!, src/nbody_to_ode.jl:474-488) over the group: every rank passes the same
        host positions u (3, n) and gets the accelerations of its own columns [lo, hi) in ``out`` (other columns
        zero).  pairs mode: each rank evaluates its share of the unordered pairs, one reduce-scatter of the
        acceleration rows completes the own block; targets mode: the own block against all sources, no exchange."""
        e = self.engine
        if self.mode != "pairs" or self.world == 1:
            return e.accel(u, out)
        e.accel_begin(u)
        self._sum_partial_accelerations()
        return e.accel_end(out)

    def step(self, dt: float, nsteps: int = 1):
        for _ in range(nsteps):
            self.engine.vv_begin(dt)
            self._all_gather_positions()
            self.engine.vv_forces()
            if self.mode == "pairs":
                self._sum_partial_accelerations()
            self.engine.vv_finish(dt)
            if self.engine.needs_temperature and self.world > 1:
                s = self.engine.scalars()
                self.dist.all_reduce(s[0:1], op=self.dist.ReduceOp.SUM, group=self.group)


class RebuildSchedule:
    """When do the slabs rebuild their Verlet lists?  Pure host logic (no device, no process group), so it is tested on
    the CPU (tests/test_parallel_gloo.py).

    Every step k takes a displacement check (soft flag: beyond SOFT x skin/2; hard flag: beyond skin/2 or a list
    overflow), reduced over all ranks.  The decision for step k reads the flags of step k - lag, so the host never
    waits for the step it is enqueuing; every rank reads the same reduced flags of the same step, hence all decide
    alike.  A soft flag taken at or before the last rebuild compares against lists that no longer exist and is ignored.
    A hard flag of ANY step that did not itself rebuild means forces were computed from a list that was no longer a
    superset: that raises instead of going unnoticed (the check of a rebuild step precedes the rebuild and is moot)."""

    def __init__(self, lag: int = 2):
        self.lag = lag
        self.k = 0               # the step being enqueued
        self.last_rebuild = -1
        self.rebuilds = 0
        self.force = False       # rebuild at the next step whatever the flags say
        self._recent = []        # the rebuild steps that checks still in flight can refer to

    def due(self, flags_of) -> bool:
        """flags_of(j) -> (soft, hard) of the check taken at step j (may block until they have arrived)."""
        want = self.force
        j = self.k - self.lag
        if j >= 0 and j not in self._recent:
            soft, hard = flags_of(j)
            if j <= self.last_rebuild:
                soft = 0  # measured against lists that have been replaced since
            if hard:
                raise RuntimeError(f"slab Verlet lists: at step {j} a particle had moved more than skin/2 since the last "
                                   "rebuild (or a list overflowed) before the collective rebuild could happen; use a larger "
                                   "verlet_skin_permille, a smaller time step, or verlet_skin_permille = 0")
            want = want or bool(soft)
        return want

    def advance(self, rebuilt: bool):
        if rebuilt:
            self.last_rebuild = self.k
            self.force = False
            self.rebuilds += 1
            self._recent = [r for r in self._recent if r >= self.k - self.lag] + [self.k]
        self.k += 1


class SlabStepper:
    """Velocity Verlet of a cutoff system over x-slabs, one rank per slab (see the module docstring).

    Every rank must have described and uploaded the FULL system on its engine before construction; the
    engine then keeps its slab.  Engine duck type, on top of ShardedStepper's: slab_init(rank, world),
    slab_buffers() -> 4 flat tensors, slab_pack(), slab_unpack() -> counts, slab_download().
    """

    def __init__(self, engine, group=None, direct=True, soft=None):
        """soft: fraction of skin/2 at which a collective rebuild is requested (default SOFT = 0.75).  The request is read
        LAG steps late, so the fastest particle must not cover the remaining (1 - soft) x skin/2 within LAG + 1 steps:
        hot systems or long time steps want a smaller value (more frequent rebuilds); violations raise, they are never
        silent."""
        import torch.distributed as dist

        self.dist = dist
        self.engine = engine
        self.group = group
        distributed = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if distributed else 1
        self.rank = dist.get_rank(group) if distributed else 0
        self.left = (self.rank - 1) % self.world
        self.right = (self.rank + 1) % self.world
        engine.slab_init(self.rank, self.world)
        # direct = the neighbours' receive areas are mapped and the pack kernel writes into them (no host exchange)
        self.direct = bool(direct and self.world > 1 and hasattr(engine, "slab_connect") and engine.slab_connect(self))
        self.bufs = None if self.direct else engine.slab_buffers()
        engine.slab_pack()  # the first pack selects the own particles out of the full upload
        self._exchange()
        self.counts = engine.slab_unpack()
        # Verlet lists inside the slabs: every rank rebuilds on the same step.  The decision is the max over the ranks of
        # a displacement flag computed on the device after each position update; it is read LAG steps late, so the
        # host never waits for the step it is enqueuing, and taken at a soft limit (0.75 of skin/2) that leaves room
        # for those steps.  A hard flag (beyond skin/2, or a list overflow) that was not covered by a rebuild raises.
        self.verlet = bool(hasattr(engine, "slab_verlet") and engine.slab_verlet())
        self.merged = False
        self.soft = float(self.SOFT if soft is None else soft)
        self.sched = RebuildSchedule(self.LAG)
        if self.verlet:
            import torch

            # a second round records the halo index lists; the lists are built from the distributed positions right away
            engine.slab_mark("slab_record_halo")
            engine.slab_pack()
            self._exchange()
            self.counts = engine.slab_unpack()
            engine.slab_prime()

            dev = engine.device
            cuda = dev.type == "cuda"  # (a host engine -- the gloo tests -- needs neither pinned memory nor events)
            self._flags_dev = torch.zeros((self.SLOTS, 2), dtype=torch.int32, device=dev)
            self._flags_host = torch.zeros((self.SLOTS, 2), dtype=torch.int32)
            if cuda:
                self._flags_host = self._flags_host.pin_memory()
            self._events = [torch.cuda.Event() for _ in range(self.SLOTS)] if cuda else None
            # merged mode (engines with slab_step_begin / slab_step_end, direct exchange or a single slab): two library
            # calls and ONE collective per step -- the sum over the ranks of [sum m v^2, soft flag, hard flag]
            self.merged = bool(cuda and hasattr(engine, "slab_step_begin") and (self.direct or self.world == 1))
            self._flags_host_d = torch.zeros((self.SLOTS, 2), dtype=torch.float64)
            if cuda:
                self._flags_host_d = self._flags_host_d.pin_memory()
            self._T_pending = False  # slot 12 holds a local sum that has not been added over the ranks yet
            # views made once: slicing tensors every step costs more host time than the kernels they feed
            self._host_i = [self._flags_host[k] for k in range(self.SLOTS)]
            self._host_d = [self._flags_host_d[k] for k in range(self.SLOTS)]
            self._np_i = self._flags_host.numpy()
            self._np_d = self._flags_host_d.numpy()
            self._stream = torch.cuda.current_stream(dev) if cuda else None
            if self.merged:
                s = engine.scalars()
                self._s_flags, self._s_all = s[13:15], s[12:15]

    LAG, SLOTS, SOFT = 2, 8, 0.75

    @property
    def rebuilds(self):
        return self.sched.rebuilds

    def _peer(self, r):
        return r if self.group is None else self.dist.get_global_rank(self.group, r)

    def _rebuild_wanted(self):
        """Enqueue this step's displacement check (+ max over the ranks) and return the decision for THIS step from the
        check of LAG steps ago.  Identical on every rank: all read the same reduced flags of the same step."""
        e, k = self.engine, self.sched.k
        slot = k % self.SLOTS
        if self.merged:  # the position update and the check were enqueued by slab_step_begin
            if self.world > 1:
                self.dist.all_reduce(self._s_all if self._T_pending else self._s_flags, op=self.dist.ReduceOp.SUM,
                                     group=self.group)
            self._T_pending = True
            host = self._np_d
            self._host_d[slot].copy_(self._s_flags, non_blocking=True)
        else:
            buf = self._flags_dev[slot]
            buf.zero_()
            e.slab_verlet_check(buf, self.soft)
            if self.world > 1:
                self.dist.all_reduce(buf, op=self.dist.ReduceOp.MAX, group=self.group)
            host = self._np_i
            self._host_i[slot].copy_(buf, non_blocking=True)
        if self._events:
            self._events[slot].record(self._stream)

        def flags_of(j):
            jj = j % self.SLOTS
            if self._events and not self._events[jj].query():
                self._events[jj].synchronize()
            return int(host[jj, 0]), int(host[jj, 1])

        return self.sched.due(flags_of)

    def _exchange(self):
        """send-to-left -> the left neighbour's recv-from-right, send-to-right -> the right neighbour's
        recv-from-left.  Posting order (sends: left, right; receives: from right, from left) keeps the two
        messages of a 2-rank ring, where both neighbours are the same peer, matched."""
        if self.world == 1 or self.direct:
            return  # one slab: no ghosts at all; direct: the pack kernel already stored into the neighbours
        dist = self.dist
        send_l, send_r, recv_l, recv_r = self.bufs
        ops = [dist.P2POp(dist.isend, send_l, self._peer(self.left), self.group),
               dist.P2POp(dist.isend, send_r, self._peer(self.right), self.group),
               dist.P2POp(dist.irecv, recv_r, self._peer(self.right), self.group),
               dist.P2POp(dist.irecv, recv_l, self._peer(self.left), self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def _one_step(self, dt):
        e = self.engine
        if self.merged:
            e.slab_step_begin(dt, self.soft)
            if self._rebuild_wanted():
                e.slab_pack()           # migration round, then the halo round that is remembered (see below)
                self._exchange()
                e.slab_unpack(sync=False)
                e.slab_mark("slab_record_halo")
                e.slab_pack()
                self._exchange()
                e.slab_unpack(sync=False)
                e.slab_mark("slab_rebuild")
                e.slab_step_end(dt, False)
                self.sched.advance(True)
            else:
                e.slab_step_end(dt, True)
                self.sched.advance(False)
            return
        e.vv_begin(dt)
        rebuilt = False
        if self.verlet and not self._rebuild_wanted():
            e.slab_refresh_send()       # nothing migrates, nothing is renumbered: only the halo positions travel
            self._exchange()
            e.slab_refresh_recv()
        else:
            e.slab_pack()
            self._exchange()
            e.slab_unpack(sync=False)  # counts stay on the device: no host round trip inside a step
            if self.verlet:
                e.slab_mark("slab_record_halo")  # second round: the halo now includes the arrivals, and is remembered
                e.slab_pack()
                self._exchange()
                e.slab_unpack(sync=False)
                e.slab_mark("slab_rebuild")
                rebuilt = True
        self.sched.advance(rebuilt)
        e.vv_forces()
        e.vv_finish(dt)
        if e.needs_temperature and self.world > 1:
            self.dist.all_reduce(e.scalars()[0:1], op=self.dist.ReduceOp.SUM, group=self.group)

    def step(self, dt: float, nsteps: int = 1, check: bool = True):
        """nsteps velocity-Verlet steps, enqueued without host synchronisation; ``check`` then waits and raises
        if any step lost a particle or overflowed a buffer (``self.counts`` = counts of the last step)."""
        for _ in range(nsteps):
            self._one_step(dt)
        if self.merged and self._T_pending and self.world > 1 and self.engine.needs_temperature:
            # leave the scalar block as the unmerged path does: [0] = the sum over all ranks
            s = self.engine.scalars()
            self.dist.all_reduce(s[12:13], op=self.dist.ReduceOp.SUM, group=self.group)
            s[0:1].copy_(s[12:13])
            self._T_pending = False
        if check:
            self.counts = self.engine.slab_check()
            self._drain_schedule()

    def _drain_schedule(self):
        """The flags are read LAG steps late: the checks of the last LAG steps have not been looked at when step()
        returns.  A hard flag among them (a step computed from a stale list) must raise here, not go unnoticed."""
        if not self.verlet or not self._events:
            return
        host = self._np_d if self.merged else self._np_i
        k = self.sched.k
        for j in range(max(0, k - self.sched.lag), k):
            if j in self.sched._recent:   # that step rebuilt: its check preceded the rebuild and is moot
                continue
            jj = j % self.SLOTS
            self._events[jj].synchronize()
            if int(host[jj, 1]):
                raise RuntimeError(f"slab Verlet lists: at step {j} a particle had moved more than skin/2 since the last rebuild "
                                   "(or a list overflowed); the run ended before the collective rebuild")

    def gather(self, n_total: int):
        """(u, v, dv) of the whole system in the original column order, on every rank (host arrays;
        diagnostics and tests, not the hot path)."""
        gid, u, v, dv = self.engine.slab_download()
        parts = [(gid, u, v, dv)]
        if self.world > 1:
            parts = [None] * self.world
            self.dist.all_gather_object(parts, (gid, u, v, dv), group=self.group)
        out = [np.zeros((3, n_total), order="F") for _ in range(3)]
        seen = np.zeros(n_total, dtype=np.int64)
        for g, uu, vv, aa in parts:
            seen[g] += 1
            for dst, src in zip(out, (uu, vv, aa)):
                dst[:, g] = src
        if not (seen == 1).all():
            raise RuntimeError(f"slab ownership is not a partition: {int((seen == 0).sum())} particles lost, "
                               f"{int((seen > 1).sum())} duplicated")
        return out


def join_group_local(contexts, mode=0):
    """Join single-GPU contexts of THIS process into one group (what nbx_create_multi does inside the library; the tests
    use it to drive every member from its own thread).  Every context holds the full uploaded system."""
    world = len(contexts)
    for r, ctx in enumerate(contexts):
        ctx.group_init(r, world, mode)
    ptrs = [ctx.group_export(want_handles=False)[0] for ctx in contexts]
    for ctx in contexts:
        ctx.group_connect(ptrs=ptrs)
    for ctx in contexts:
        ctx.group_start()
    for ctx in contexts:
        if ctx.info("group_mode") == 3:
            ctx.slab_check()
        else:
            ctx.synchronize()


def join_group_dist(ctx, group=None, mode=0):
    """Join the contexts of a torch.distributed process group (one process per GPU) into one libnbody_b200 group:
    torch.distributed only carries the CUDA IPC handles (bootstrap) -- afterwards nbx_accel / nbx_step_vv / nbx_step_em
    exchange everything device to device over NVLink peer memory, with no collective library call per step."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ctx.group_init(rank, world, mode)
    _, blob = ctx.group_export(want_handles=True)
    blobs = [None] * world
    dist.all_gather_object(blobs, blob, group=group)
    ctx.group_connect(handles=b"".join(blobs))
    dist.barrier(group=group)   # every window is mapped before anyone stores into one
    ctx.group_start()
    if ctx.info("group_mode") == 3:
        ctx.slab_check()
    else:
        ctx.synchronize()
    dist.barrier(group=group)


def numpy_reference_partition_check(n, world, multiple=1):
    """All ranks' ranges tile [0, n) exactly (host-logic self check used by the CPU tests)."""
    covered = np.zeros(n, dtype=np.int32)
    for r in range(world):
        lo, hi, _ = partition(n, world, r, multiple)
        covered[lo:hi] += 1
    return bool((covered == 1).all())
