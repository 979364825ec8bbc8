"""Development probe: 1,048,576-atom LJ argon step time vs Verlet skin (unfused path, CUDA graph)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nbody_b200.workloads as wl  # noqa: E402
from nbody_b200 import _lib  # noqa: E402
w = wl.fcc_argon_reduced(64)
n = w["u"].shape[1]
rng = np.random.Generator(np.random.Philox(2))
u = np.asfortranarray(w["u"] + 0.05 * rng.standard_normal(w["u"].shape))
side = torch.cuda.Stream(); torch.cuda.set_stream(side)
for skin in (100, 70, 50, 35):
    ctx = _lib.Context(0)
    ctx.system(w["ms"]); ctx.boundary(_lib.BC_CUBIC, [w["L"]])
    ctx.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"])
    ctx.thermostat(_lib.THERMO_BERENDSEN, 90.0, 10 * w["dt"], w["kB"], n, 0)
    ctx.set_stream(side.cuda_stream)
    ctx.set_option("verlet_skin_permille", skin)
    ctx.upload(u, w["v"])
    ctx.step_vv(w["dt"], 300)
    r0 = ctx.info("verlet_rebuilds")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    ctx.step_vv(w["dt"], 1000)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 1000
    print(f"skin={skin} ms/step={ms:.4f} atom-steps/s={n / ms * 1e3:.3e} rebuilds in 1000 steps={ctx.info('verlet_rebuilds') - r0} cells={ctx.info('cells_lj')}", flush=True)
    ctx.close()
