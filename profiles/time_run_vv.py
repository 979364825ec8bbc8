"""nbx_run_vv with frames (the saveat path of run_simulation): ms per step for a few chunk lengths, 216 and 32,000 argon atoms."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import nbody_b200.workloads as wl
from nbody_b200 import _lib
for name, w in (("216 atoms (config 1)", wl.liquid_argon_si(216)), ("32,000 atoms", wl.fcc_argon_reduced(20))):
    for save in (7, 10, 100):
        ctx = _lib.Context(0)
        ctx.system(w["ms"]); ctx.boundary(_lib.BC_CUBIC, [w["L"]]); ctx.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"])
        ctx.upload(w["u"], w["v"])
        ctx.run_vv(w["dt"], 200, save_every=save)
        t0 = time.perf_counter(); uf, vf = ctx.run_vv(w["dt"], 2000, save_every=save); t1 = time.perf_counter()
        ref = _lib.Context(0)
        ref.system(w["ms"]); ref.boundary(_lib.BC_CUBIC, [w["L"]]); ref.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"])
        ref.upload(w["u"], w["v"]); ref.step_vv(w["dt"], 2200); ur, vr, _ = ref.download()
        print(f"{name} save_every={save}: {(t1 - t0) / 2000 * 1e6:.1f} us/step incl. {uf.shape[0]} frame downloads; last frame equals step+download: {np.array_equal(uf[-1], ur) and np.array_equal(vf[-1], vr)}", flush=True)
        ctx.close(); ref.close()
