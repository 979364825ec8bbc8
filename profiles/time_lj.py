"""Device timing of the LJ argon step phases (pair kernel / cell rebuild / integrate), per option set.

  python profiles/time_lj.py [cells=64] [steps=20] [opt=val ...]     e.g. prefilter=0
Not a bench line: a development probe (CUDA-event phase timers of the library, nbx_timing_*).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import nbody_b200.workloads as wl  # noqa: E402
from nbody_b200 import _lib  # noqa: E402


def run(cells, steps, opts):
    w = wl.fcc_argon_reduced(cells)
    n = w["u"].shape[1]
    rng = np.random.Generator(np.random.Philox(2))
    u = np.asfortranarray(w["u"] + 0.05 * rng.standard_normal(w["u"].shape))
    ctx = _lib.Context(0)
    ctx.system(w["ms"])
    ctx.boundary(_lib.BC_CUBIC, [w["L"]])
    ctx.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"])
    ctx.thermostat(_lib.THERMO_BERENDSEN, 90.0, 10 * w["dt"], w["kB"], n, 0)
    for k, v in opts.items():
        ctx.set_option(k, v)
    ctx.upload(u, w["v"])
    ctx.step_vv(w["dt"], 5)
    ctx.synchronize()
    import time as _t
    t0 = _t.perf_counter()
    ctx.step_vv(w["dt"], steps)          # nbx_step_vv synchronises at the end
    clean = (_t.perf_counter() - t0) * 1e3 / steps
    ctx.timing_reset()
    ctx.timing_enable(True)
    import time
    t0 = time.perf_counter()
    ctx.step_vv(w["dt"], steps)
    ctx.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / steps
    ctx.timing_enable(False)
    pair, pc = ctx.timing_get(_lib.T_PAIR_CELLS)
    build, _ = ctx.timing_get(_lib.T_CELL_BUILD)
    integ, _ = ctx.timing_get(_lib.T_INTEGRATE)
    _, _, T = ctx.energy(potential=False)
    print(f"n={n} opts={opts} ms/step wall(no timers)={clean:.4f} wall(with timers)={wall:.4f} pair={pair / max(pc, 1):.4f} build={build / steps:.4f} "
          f"integrate={integ / steps:.4f} T={T:.6f}", flush=True)
    ctx.close()


if __name__ == "__main__":
    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    opts = {}
    for a in sys.argv[3:]:
        k, v = a.split("=")
        opts[k] = int(v)
    run(cells, steps, opts)
