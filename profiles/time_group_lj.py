"""Wall time of the LJ argon step on a single-process group over all visible GPUs (nbx_create_multi: x-slabs):
  python profiles/time_group_lj.py [cells=64] [steps=400] [key=value options...]   (development probe, not a bench number)"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
import nbody_b200.workloads as wl
from nbody_b200 import _lib

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
opts = dict(kv.split("=") for kv in sys.argv[3:])
ndev = int(opts.pop("ndev", torch.cuda.device_count()))
w = wl.fcc_argon_reduced(cells)
n = w["u"].shape[1]
rng = np.random.Generator(np.random.Philox(2))
u = np.asfortranarray(w["u"] + 0.05 * rng.standard_normal(w["u"].shape))
ctx = _lib.Context(list(range(ndev)))
ctx.system(w["ms"]); ctx.boundary(_lib.BC_CUBIC, [w["L"]]); ctx.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"])
ctx.thermostat(_lib.THERMO_BERENDSEN, 90.0, 10 * w["dt"], w["kB"], n, 0)
for k, val in opts.items():
    ctx.set_option(k, int(val))
ctx.upload(u, w["v"])
ctx.step_vv(w["dt"], 60)
best = 1e9
for _ in range(3):
    t0 = time.perf_counter(); ctx.step_vv(w["dt"], steps); t1 = time.perf_counter()
    best = min(best, (t1 - t0) / steps * 1e3)
print(f"n={n} devices={ndev} mode={ctx.info('group_mode')} opts={opts} wall ms/step={best:.4f} atom-steps/s={n / best * 1e3:.3e} "
      f"rebuilds={ctx.info('verlet_rebuilds')}", flush=True)
ctx.close()
