"""Device timing of the 1,048,576-atom LJ argon step: fused step per cluster size vs the unfused kernels.

  python profiles/time_fused.py [cells=64] [steps=400]
Development probe (CUDA events around nbx_step_vv on a dedicated stream, so the CUDA graph path is taken).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import nbody_b200.workloads as wl  # noqa: E402
from nbody_b200 import _lib  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
w = wl.fcc_argon_reduced(cells)
n = w["u"].shape[1]
rng = np.random.Generator(np.random.Philox(2))
u = np.asfortranarray(w["u"] + 0.05 * rng.standard_normal(w["u"].shape))
side = torch.cuda.Stream()
torch.cuda.set_stream(side)
for fused, C in ((0, 1), (1, 1), (1, 2), (1, 4), (1, 8)):
    ctx = _lib.Context(0)
    ctx.system(w["ms"])
    ctx.boundary(_lib.BC_CUBIC, [w["L"]])
    ctx.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"])
    ctx.thermostat(_lib.THERMO_BERENDSEN, 90.0, 10 * w["dt"], w["kB"], n, 0)
    ctx.set_stream(side.cuda_stream)
    ctx.set_option("fused_step", fused)
    ctx.set_option("fused_cluster", C)
    ctx.upload(u, w["v"])
    ctx.step_vv(w["dt"], 100)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    ctx.step_vv(w["dt"], steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ctx.timing_reset(); ctx.timing_enable(True)
    ctx.step_vv(w["dt"], 40)
    ctx.timing_enable(False)
    pair, pc = ctx.timing_get(_lib.T_PAIR_CELLS)
    build, _ = ctx.timing_get(_lib.T_CELL_BUILD)
    _, _, T = ctx.energy(potential=False)
    print(f"n={n} fused={fused} C={C} ms/step={ms:.4f} atom-steps/s={n / ms * 1e3:.3e} eager: step kernel={pair / max(pc, 1):.4f} chain={build / 40:.4f} "
          f"rebuilds={ctx.info('verlet_rebuilds')} fused_steps={ctx.info('fused_steps')} list_cap={ctx.info('fused_list_cap')} T={T:.4f}", flush=True)
    ctx.close()
