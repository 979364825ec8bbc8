"""ncu driver: 1,048,576-atom LJ argon, fused step, eager launches.  python profiles/prof_fused.py <cluster> [steps]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nbody_b200.workloads as wl  # noqa: E402
from nbody_b200 import _lib  # noqa: E402
C = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
w = wl.fcc_argon_reduced(int(os.environ.get("PROF_CELLS", 64)))
n = w["u"].shape[1]
rng = np.random.Generator(np.random.Philox(2))
u = np.asfortranarray(w["u"] + 0.05 * rng.standard_normal(w["u"].shape))
ctx = _lib.Context(0)
ctx.system(w["ms"]); ctx.boundary(_lib.BC_CUBIC, [w["L"]])
ctx.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"])
ctx.thermostat(_lib.THERMO_BERENDSEN, 90.0, 10 * w["dt"], w["kB"], n, 0)
ctx.set_option("graph", 0); ctx.set_option("fused_step", 1); ctx.set_option("fused_cluster", C)
ctx.upload(u, w["v"])
ctx.step_vv(w["dt"], steps)
print("fused steps", ctx.info("fused_steps"))
ctx.close()
