"""Device time of one rank's share of the pair-sharded gravity kernel on ONE GPU (development probe):
  python profiles/time_share.py [n=262144] [world=8] [key=value options...]"""
import sys
sys.path.insert(0, ".")
import torch
import nbody_b200.workloads as wl
from nbody_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
opts = dict(kv.split("=") for kv in sys.argv[3:])
u, v, ms = wl.plummer(n)
for r in sorted({0, 1, world - 1}):
    ctx = _lib.Context(0)
    ctx.system(ms); ctx.add_gravity(1.0)
    for k, val in opts.items():
        ctx.set_option(k, int(val))
    ctx.upload(u, v)
    if world > 1:
        ctx.shard_pairs(r, world)
    ctx.vv_begin(0.0); ctx.vv_forces(); torch.cuda.synchronize()
    ctx.timing_reset(); ctx.timing_enable(True)
    for _ in range(5):
        ctx.vv_forces()
    torch.cuda.synchronize(); ctx.timing_enable(False)
    t, cnt = ctx.timing_get(_lib.T_PAIR_ALLPAIRS)
    print(f"n={n} rank {r} of {world} opts={opts}: kernel {t / cnt:.4f} ms (x{world} = {t / cnt * world:.3f}) grid={ctx.info('allpairs_grid')} segments={ctx.info('allpairs_chunks')}", flush=True)
    ctx.close()
    if world == 1:
        break
