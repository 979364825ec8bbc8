"""torchrun worker of tests/test_gpu_multi.py: N ranks step the same system with the sharded stepper
(both decompositions); rank 0 compares positions and velocities with a single-context run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def slab_main(cells):
    """LJ argon over x-slabs (parallel.SlabStepper, NCCL send/recv) against one context stepping everything."""
    import torch
    import torch.distributed as dist

    import nbody_b200.workloads as wl
    from nbody_b200 import _lib
    from nbody_b200.parallel import CudaEngine, SlabStepper

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = wl.fcc_argon_reduced(cells)
    n = w["u"].shape[1]
    rng = np.random.Generator(np.random.Philox(5))
    u = np.asfortranarray(w["u"] + 0.05 * rng.standard_normal(w["u"].shape))
    v = np.asfortranarray(3.0 * w["v"])
    dt, steps = 2e-3, 40
    ok = True
    # skin 0: cells rescanned every step (bit-identical to one GPU); skin 100: Verlet lists inside the slabs, collective
    # rebuilds decided two steps late (other rebuild steps than on one GPU: same pair set, other summation order)
    for thermo, direct, skin in ((False, True, 0), (True, True, 0), (False, False, 0), (False, True, 100), (True, False, 100)):
        def make():
            ctx = _lib.Context(local)
            ctx.system(w["ms"])
            ctx.boundary(_lib.BC_CUBIC, [w["L"]])
            ctx.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"])
            ctx.set_option("verlet_skin_permille", skin)
            ctx.set_option("verlet_banked", 0)
            if thermo:
                ctx.thermostat(_lib.THERMO_BERENDSEN, 90.0, 20 * dt, w["kB"], n, 0)
            return ctx

        ref = make()
        ref.upload(u, v)
        ref.step_vv(dt, steps)
        ur, vr, ar = ref.download(want_dv=True)
        ref.close()
        ctx = make()
        eng = CudaEngine(ctx, local)
        eng.needs_temperature = thermo
        ctx.upload(u, v)
        st = SlabStepper(eng, direct=direct, soft=0.4)  # hot atoms: 0.04 sigma per step against skin/2 = 0.11 sigma
        moved = 0
        for _ in range(steps):
            st.step(dt, 1)
            moved += st.counts[2] + st.counts[3]
        ug, vg, ag = st.gather(n)
        t = torch.tensor([float(moved)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        if rank == 0:
            if skin:
                err = max(np.abs(a - b).max() / np.abs(b).max() for a, b in ((ug, ur), (vg, vr), (ag, ar)))
                good = err < 1e-9 and st.verlet and 2 <= st.rebuilds < steps
                t += 1.0  # (migrations are only counted on rebuild steps here)
            elif thermo:
                err = max(np.abs(a - b).max() / np.abs(b).max() for a, b in ((ug, ur), (vg, vr), (ag, ar)))
                good = err < 1e-11
            else:  # cell order is ranked by global id: bit-identical to the single-GPU sums
                err = 0.0 if (np.array_equal(ug, ur) and np.array_equal(vg, vr) and np.array_equal(ag, ar)) else 1.0
                good = err == 0.0
            print(f"slab thermo={thermo} direct={st.direct} skin={skin} rebuilds={st.rebuilds} world={world} n={n} own={st.counts[0]} "
                  f"ghosts={st.counts[1]} migrations={int(t.item())} err={err:.2e}")
            ok = ok and good and t.item() > 0 and (st.direct or not direct)
        ctx.close()
    if rank == 0:
        print("MGPU_OK" if ok else "MGPU_FAIL")
    dist.destroy_process_group()


def _group_cases():
    """(name, spec, u, v, dt, steps, em, tolerances) shared by the cross-process and the single-process group runs."""
    import nbody_b200.workloads as wl

    cases = []
    w = wl.fcc_argon_reduced(16)
    rng = np.random.Generator(np.random.Philox(5))
    u = np.asfortranarray(w["u"] + 0.05 * rng.standard_normal(w["u"].shape))
    lj = dict(ms=w["ms"], bc=("cubic", w["L"]), lj=w["lj"])
    cases.append(("slabs NVE", lj, u, np.asfortranarray(3.0 * w["v"]), 2e-3, 60, False, 0.0))
    ljt = dict(lj, thermostat=dict(kind="berendsen", T=90.0, tau=0.04, kB=w["kB"]))
    cases.append(("slabs Berendsen", ljt, u, np.asfortranarray(3.0 * w["v"]), 2e-3, 60, False, 1e-9))
    ug, vg, ms = wl.plummer(16384, seed=11)
    ms = ms * (0.5 + np.random.Generator(np.random.Philox(4)).random(16384))
    cases.append(("pairs gravity", dict(ms=ms, gravity=dict(G=1.0)), ug, vg, 1e-3, 10, False, 1e-11))
    ww = wl.water_omm(8, Rel=0.9)
    water = dict(ms=ww["ms"], qs=ww["qs"], water=True, bc=("cubic", ww["L"]), lj=ww["lj"], coulomb=ww["coulomb"], spcfw=ww["spcfw"])
    cases.append(("targets water", water, np.asfortranarray(ww["u"]), np.asfortranarray(ww["v"]), ww["dt"], 20, False, 1e-9))
    wp = wl.water_omm(14)   # default electrostatic cutoff 0.49 L, 8,232 atoms: the unordered Coulomb pairs are sharded
    waterp = dict(ms=wp["ms"], qs=wp["qs"], water=True, bc=("cubic", wp["L"]), lj=wp["lj"], coulomb=wp["coulomb"], spcfw=wp["spcfw"])
    cases.append(("pairs water", waterp, np.asfortranarray(wp["u"]), np.asfortranarray(wp["v"]), wp["dt"], 10, False, 1e-9))
    wc = wl.charged_lattice(12288)
    em = dict(ms=wc["ms"], qs=wc["qs"], coulomb=dict(k=wc["coulomb"]["k"], R=np.inf),
              thermostat=dict(kind="langevin", T=90.0, gamma=10.0, kB=1.38e-23))
    cases.append(("pairs Langevin EM", em, np.asfortranarray(wc["u"]), np.asfortranarray(wc["v"]), 1e-9, 6, True, 1e-10))
    return cases


def _err(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def group_main():
    """The C-ABI group across processes: torch.distributed only carries the CUDA IPC handles; nbx_accel / nbx_step_vv /
    nbx_step_em then run with no collective library call (nbx_multi.cu, slab_enqueue).  Every rank checks its own columns
    against a single-context run of the whole system on its own GPU."""
    import torch
    import torch.distributed as dist

    from nbody_b200.parallel import join_group_dist
    from tests._common import make_context

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for name, spec, u, v, dt, steps, em, tol in _group_cases():
        one = make_context(spec, device=local)
        one.set_option("verlet_banked", 0)   # the list order of the slabs (see tests/test_gpu_group.py::_single_run)
        one.upload(u, v)
        (one.step_em(dt, steps, 9) if em else one.step_vv(dt, steps))
        ur, vr, ar = one.download(want_dv=True)
        a_rhs = one.accel(u).copy() if "thermostat" not in spec else None
        one.close()
        ctx = make_context(spec, device=local)
        ctx.set_option("verlet_banked", 0)   # (a(0) is evaluated before the context becomes a slab)
        ctx.upload(u, v)
        join_group_dist(ctx)
        mode = ctx.info("group_mode")
        (ctx.step_em(dt, steps, 9) if em else ctx.step_vv(dt, steps))
        if mode == 3:
            gid, ug, vg, ag = ctx.slab_download()
            errs = [_err(ug, ur[:, gid]), _err(vg, vr[:, gid]), _err(ag, ar[:, gid])]
            exact = np.array_equal(ug, ur[:, gid]) and np.array_equal(vg, vr[:, gid]) and np.array_equal(ag, ar[:, gid])
            own = len(gid)
        else:
            lo, hi = ctx.info("shard_lo"), ctx.info("shard_hi")
            ug, vg, ag = ctx.download(want_dv=True)
            errs = [_err(ug, ur), _err(vg[:, lo:hi], vr[:, lo:hi]), _err(ag[:, lo:hi], ar[:, lo:hi])]
            exact = False
            own = hi - lo
            if a_rhs is not None:   # the RHS drop-in: own block of u up, all-gather over NVLink, own columns of dv back
                dv = np.zeros_like(u)
                ctx.accel(u, out=dv)
                errs.append(_err(dv[:, lo:hi], a_rhs[:, lo:hi]))
                errs.append(1.0 if (np.abs(dv[:, :lo]).max(initial=0.0) + np.abs(dv[:, hi:]).max(initial=0.0)) else 0.0)
        good = exact if tol == 0.0 else max(errs) < tol
        t = torch.tensor([1.0 if good else 0.0, max(errs), float(own)], dtype=torch.float64, device="cuda")
        tmin, tsum = t.clone(), t.clone()
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        if rank == 0:
            print(f"group[{name}] world={world} mode={mode} max err {t[1].item():.2e} bit-identical={bool(exact)} "
                  f"owned {int(tsum[2].item())} of {u.shape[1]} ok={bool(tmin[0].item())}")
        ok = ok and bool(tmin[0].item()) and int(tsum[2].item()) == u.shape[1]
        ctx.close()
    if rank == 0:
        print("MGPU_OK" if ok else "MGPU_FAIL")
    dist.destroy_process_group()


def multi_main():
    """nbx_create_multi over the visible GPUs of ONE process (no torch.distributed at all): one handle, one host thread."""
    import torch

    from tests._common import make_context

    ng = torch.cuda.device_count()
    devs = list(range(min(ng, 4)))
    ok = True
    for name, spec, u, v, dt, steps, em, tol in _group_cases():
        one = make_context(spec, device=0)
        one.set_option("verlet_banked", 0)
        one.upload(u, v)
        (one.step_em(dt, steps, 9) if em else one.step_vv(dt, steps))
        ur, vr, ar = one.download(want_dv=True)
        one.close()
        grp = make_context(spec, device=devs)
        grp.upload(u, v)
        (grp.step_em(dt, steps, 9) if em else grp.step_vv(dt, steps))
        ug, vg, ag = grp.download(want_dv=True)
        errs = [_err(ug, ur), _err(vg, vr), _err(ag, ar)]
        exact = np.array_equal(ug, ur) and np.array_equal(vg, vr) and np.array_equal(ag, ar)
        good = exact if tol == 0.0 else max(errs) < tol
        print(f"multi[{name}] devices={devs} mode={grp.info('group_mode')} max err {max(errs):.2e} bit-identical={exact} ok={good}")
        ok = ok and good
        grp.close()
    print("MGPU_OK" if ok else "MGPU_FAIL")


def main():
    import torch
    import torch.distributed as dist

    import nbody_b200.workloads as wl
    from nbody_b200 import _lib
    from nbody_b200.parallel import CudaEngine, ShardedStepper

    if sys.argv[1] == "slab":
        return slab_main(int(sys.argv[2]))
    if sys.argv[1] == "group":
        return group_main()
    if sys.argv[1] == "multi":
        return multi_main()
    n = int(sys.argv[1])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    u, v, ms = wl.plummer(n, seed=11)
    rng = np.random.Generator(np.random.Philox(4))
    ms = ms * (0.5 + rng.random(n))          # unequal masses: the general kernel variant
    dt, steps = 1e-3, 4
    th = dict(T0=0.05, tau=50 * dt, kB=1.0)

    def make(thermo):
        ctx = _lib.Context(local)
        ctx.system(ms)
        ctx.add_gravity(1.0)
        if thermo:
            ctx.thermostat(_lib.THERMO_BERENDSEN, th["T0"], th["tau"], th["kB"], n, 0)
        return ctx

    ok = True
    a_ref = None
    for thermo in (False, True):
        ref = make(thermo)
        ref.upload(u, v)
        ref.step_vv(dt, steps)
        ur, vr, _ = ref.download()
        ref.close()
        for mode in ("targets", "pairs"):
            ctx = make(thermo)
            eng = CudaEngine(ctx, local)
            eng.needs_temperature = thermo
            ctx.upload(u, v)
            st = ShardedStepper(eng, mode=mode)
            st.step(dt, steps)
            torch.cuda.synchronize()
            ug, vg, _ = ctx.download()
            lo, hi = st.lo, st.hi
            eu = np.abs(ug - ur).max() / np.abs(ur).max()          # all positions are gathered on every rank
            ev = np.abs(vg[:, lo:hi] - vr[:, lo:hi]).max() / np.abs(vr).max()
            if not thermo:  # the RHS drop-in over the group (pairs mode: accel_begin / reduce-scatter / accel_end)
                a_own = st.accel(u)
                if a_ref is None:
                    one = make(False)
                    a_ref = one.accel(u).copy()
                    one.close()
                ea = np.abs(a_own[:, lo:hi] - a_ref[:, lo:hi]).max() / np.abs(a_ref).max()
                outside = np.abs(a_own[:, :lo]).max(initial=0.0) + np.abs(a_own[:, hi:]).max(initial=0.0)
                ev = max(ev, ea, 1.0 if outside else 0.0)
            t = torch.tensor([eu, ev], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(f"thermo={thermo} mode={mode} world={world} max rel err pos {t[0].item():.2e} vel {t[1].item():.2e}")
                ok = ok and t[0].item() < 1e-11 and t[1].item() < 1e-11
            ctx.close()
    if rank == 0:
        print("MGPU_OK" if ok else "MGPU_FAIL")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
