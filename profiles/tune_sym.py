"""Times the symmetric all-pairs kernel variants (option sym_variant) on the 262,144-body Plummer sphere."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nbody_b200.workloads as wl  # noqa: E402
from nbody_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
u, v, ms = wl.plummer(n)
rng = np.random.Generator(np.random.Philox(1))
for uniform in (True, False):
    m = ms if uniform else ms * (0.5 + rng.random(n))
    ctx = _lib.Context(0)
    ctx.system(m)
    ctx.add_gravity(1.0)
    ref = None
    for variant in [int(x) for x in os.environ.get("VARIANTS", "0,1,2,3,4,5,6").split(",")]:
        ctx.set_option("sym_variant", variant)
        a = ctx.accel(u).copy()
        if ref is None:
            ref = a
        err = float(np.max(np.linalg.norm(a - ref, axis=0) / np.linalg.norm(ref, axis=0)))
        ctx.timing_reset(); ctx.timing_enable(True)
        for _ in range(3):
            ctx.accel(u)
        ms_k, cnt = ctx.timing_get(_lib.T_PAIR_ALLPAIRS)
        ctx.timing_enable(False)
        print(f"uniform={uniform} variant={variant} kernel_ms={ms_k / cnt:.3f} grid={ctx.info('allpairs_grid')} "
              f"segs={ctx.info('allpairs_chunks')} max_rel_diff_vs_v0={err:.2e}", flush=True)
    ctx.set_option("symmetric_pairs", 0)
    ctx.timing_reset(); ctx.timing_enable(True)
    ctx.accel(u)
    ms_k, cnt = ctx.timing_get(_lib.T_PAIR_ALLPAIRS)
    print(f"uniform={uniform} ordered kernel_ms={ms_k / cnt:.3f}")
    ctx.close()

# one rank's share under pair sharding (ring offsets k = rank mod world), timed on this device
for world in (1, 2, 4, 8):
    ctx = _lib.Context(0)
    ctx.system(ms)
    ctx.add_gravity(1.0)
    ctx.upload(u, v)
    if world > 1:
        ctx.shard_pairs(0, world)
    ctx.vv_begin(0.0)
    ctx.vv_forces()
    ctx.timing_reset(); ctx.timing_enable(True)
    for _ in range(3):
        ctx.vv_forces()
    ms_k, cnt = ctx.timing_get(_lib.T_PAIR_ALLPAIRS)
    ctx.timing_enable(False)
    ideal = 42.7 * ((0.83 + (128 // world - 1) + 0.5) / 128.33 if world > 1 else 1.0)
    print(f"rank 0 of {world}: kernel_ms={ms_k / cnt:.3f} segs={ctx.info('allpairs_chunks')} (share of the offsets x 42.7 ms = {ideal:.3f})", flush=True)
    ctx.close()
