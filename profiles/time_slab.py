"""Phase timing of the slab-decomposed LJ step (development probe, run under torchrun):
  python -m torch.distributed.run --nproc-per-node N profiles/time_slab.py [cells=64] [steps=50]
CUDA events between the phases of parallel.SlabStepper._one_step on the launching stream (so CPU launch gaps
show up in the phase that follows them), averaged over the steps; plus the whole-loop time without events."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import nbody_b200.workloads as wl
    from nbody_b200 import _lib
    from nbody_b200.parallel import CudaEngine, SlabStepper

    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = wl.fcc_argon_reduced(cells)
    n = w["u"].shape[1]
    rng = np.random.Generator(np.random.Philox(2))
    u = np.asfortranarray(w["u"] + 0.05 * rng.standard_normal(w["u"].shape))
    ctx = _lib.Context(local)
    ctx.system(w["ms"])
    ctx.boundary(_lib.BC_CUBIC, [w["L"]])
    ctx.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"])
    ctx.thermostat(_lib.THERMO_BERENDSEN, 90.0, 10 * w["dt"], w["kB"], n, 0)
    eng = CudaEngine(ctx, local)
    eng.needs_temperature = True
    ctx.upload(u, w["v"])
    st = SlabStepper(eng)
    dt = w["dt"]
    st.step(dt, 5)
    names = ["vv_begin", "pack", "exchange", "unpack", "forces", "finish", "allreduce"]
    acc = np.zeros(len(names))
    for _ in range(steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        ev[0].record(); eng.vv_begin(dt)
        ev[1].record(); eng.slab_pack()
        ev[2].record(); st._exchange()
        ev[3].record(); eng.slab_unpack(sync=False)
        ev[4].record(); eng.vv_forces()
        ev[5].record(); eng.vv_finish(dt)
        ev[6].record()
        if world > 1:
            dist.all_reduce(eng.scalars()[0:1])
        ev[7].record()
        torch.cuda.synchronize()
        acc += [ev[k].elapsed_time(ev[k + 1]) for k in range(len(names))]
    acc /= steps
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st.step(dt, steps, check=False)
    e1.record()
    torch.cuda.synchronize()
    loop = e0.elapsed_time(e1) / steps
    counts = st.engine.slab_check()
    print(f"rank {rank}/{world} n={n} own={counts[0]} ghosts={counts[1]} loop {loop:.4f} ms/step | " +
          " ".join(f"{k}={v:.4f}" for k, v in zip(names, acc)) + f" | sum {acc.sum():.4f}", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
