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


def main():
    import torch
    import torch.distributed as dist

    import nbody_b200.workloads as wl
    from nbody_b200 import _lib
    from nbody_b200.parallel import CudaEngine, ShardedStepper

    if sys.argv[1] == "slab":
        return slab_main(int(sys.argv[2]))
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
