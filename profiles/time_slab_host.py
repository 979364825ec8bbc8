"""Development probe: host cost of one slab step.  One rank (no process group), a small box (131,072 atoms), so the
step time is what the host needs to enqueue it.  python profiles/time_slab_host.py [cells=32] [steps=500]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nbody_b200.workloads as wl  # noqa: E402
from nbody_b200 import _lib  # noqa: E402
from nbody_b200.parallel import CudaEngine, SlabStepper  # noqa: E402
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 500
w = wl.fcc_argon_reduced(cells)
n = w["u"].shape[1]
rng = np.random.Generator(np.random.Philox(2))
u = np.asfortranarray(w["u"] + 0.05 * rng.standard_normal(w["u"].shape))
side = torch.cuda.Stream(); torch.cuda.set_stream(side)
ctx = _lib.Context(0)
ctx.system(w["ms"]); ctx.boundary(_lib.BC_CUBIC, [w["L"]])
ctx.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"])
ctx.thermostat(_lib.THERMO_BERENDSEN, 90.0, 10 * w["dt"], w["kB"], n, 0)
eng = CudaEngine(ctx, 0); eng.needs_temperature = True
ctx.upload(u, w["v"])
st = SlabStepper(eng)
st.step(w["dt"], 50, check=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
st.step(w["dt"], steps, check=False)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"n={n} merged={st.merged} verlet={st.verlet} host enqueue {1e3 * (t1 - t0) / steps:.4f} ms/step, with drain {1e3 * (t2 - t0) / steps:.4f} ms/step, rebuilds={st.rebuilds}")
ctx.close()

# where the host time goes: per-call wall clock of the pieces of a regular merged step
ctx = _lib.Context(0)
ctx.system(w["ms"]); ctx.boundary(_lib.BC_CUBIC, [w["L"]])
ctx.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"])
ctx.thermostat(_lib.THERMO_BERENDSEN, 90.0, 10 * w["dt"], w["kB"], n, 0)
eng = CudaEngine(ctx, 0); eng.needs_temperature = True
ctx.upload(u, w["v"])
st = SlabStepper(eng)
st.step(w["dt"], 20, check=False)
torch.cuda.synchronize()
acc = {"begin": 0.0, "decide": 0.0, "end": 0.0}
m = 200
for _ in range(m):
    a = time.perf_counter(); eng.slab_step_begin(w["dt"], st.soft)
    b = time.perf_counter(); want = st._rebuild_wanted()
    c = time.perf_counter()
    if want:
        eng.slab_pack(); eng.slab_unpack(sync=False); eng.slab_mark("slab_record_halo"); eng.slab_pack(); eng.slab_unpack(sync=False)
        eng.slab_mark("slab_rebuild")
        eng.slab_step_end(w["dt"], False)
    else:
        eng.slab_step_end(w["dt"], True)
    st.sched.advance(want)
    d = time.perf_counter()
    acc["begin"] += b - a; acc["decide"] += c - b; acc["end"] += d - c
    if _ % 4 == 3:
        torch.cuda.synchronize()  # keep the GPU from back-pressuring the host timers
print({k: f"{1e6 * v / m:.1f} us" for k, v in acc.items()})
ctx.close()
