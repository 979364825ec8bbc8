"""ncu driver: one rank's share of the 8-GPU LJ argon step as a one-slab group (no peers to wait for under the profiler's
serialised launches): python profiles/prof_slab.py [cells=32] [steps=24] [key=value options].  (ncu cannot profile the
kernel nodes of a graph that holds conditional nodes: list the launches with graph_if_nodes=0.)  Also the plain timing of the same loop."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import nbody_b200.workloads as wl
from nbody_b200 import _lib

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
opts = dict(kv.split("=") for kv in sys.argv[3:])
w = wl.fcc_argon_reduced(cells)
n = w["u"].shape[1]
rng = np.random.Generator(np.random.Philox(2))
u = np.asfortranarray(w["u"] + 0.05 * rng.standard_normal(w["u"].shape))
ctx = _lib.Context([0])
ctx.system(w["ms"]); ctx.boundary(_lib.BC_CUBIC, [w["L"]]); ctx.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"])
ctx.thermostat(_lib.THERMO_BERENDSEN, 90.0, 10 * w["dt"], w["kB"], n, 0)
for k, val in opts.items():
    ctx.set_option(k, int(val))
ctx.upload(u, w["v"])
ctx.step_vv(w["dt"], 8)
t0 = time.perf_counter(); ctx.step_vv(w["dt"], steps); t1 = time.perf_counter()
print(f"n={n} group_mode={ctx.info('group_mode')} steps={steps} wall ms/step={(t1 - t0) / steps * 1e3:.4f} rebuilds={ctx.info('verlet_rebuilds')}")
