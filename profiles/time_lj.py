"""Device timing of the LJ argon step on one GPU: python profiles/time_lj.py [cells=64] [steps=400] [key=value options...]"""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
import nbody_b200.workloads as wl
from nbody_b200 import _lib

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
opts = dict(kv.split("=") for kv in sys.argv[3:])
w = wl.fcc_argon_reduced(cells)
n = w["u"].shape[1]
rng = np.random.Generator(np.random.Philox(2))
u = np.asfortranarray(w["u"] + 0.05 * rng.standard_normal(w["u"].shape))
side = torch.cuda.Stream(); torch.cuda.set_stream(side)
ctx = _lib.Context(0)
ctx.set_stream(side.cuda_stream)
ctx.system(w["ms"]); ctx.boundary(_lib.BC_CUBIC, [w["L"]]); ctx.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"])
ctx.thermostat(_lib.THERMO_BERENDSEN, 90.0, 10 * w["dt"], w["kB"], n, 0)
for k, val in opts.items():
    ctx.set_option(k, int(val))
ctx.upload(u, w["v"])
ctx.step_vv(w["dt"], 60)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ctx.step_vv(w["dt"], steps); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
ctx.timing_reset(); ctx.timing_enable(True); ctx.step_vv(w["dt"], 20); ctx.timing_enable(False)
pair, pc = ctx.timing_get(_lib.T_PAIR_CELLS); build, _ = ctx.timing_get(_lib.T_CELL_BUILD); integ, _ = ctx.timing_get(_lib.T_INTEGRATE)
print(f"n={n} opts={opts} ms/step={ms:.4f} atom-steps/s={n / ms * 1e3:.3e} eager: pair={pair / max(pc, 1):.4f} build={build / 20:.4f} "
      f"integrate={integ / 20:.4f} rebuilds={ctx.info('verlet_rebuilds')}", flush=True)
