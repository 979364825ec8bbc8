"""Device timing of the SPC/Fw water step (config 4): python profiles/time_water.py [rel=0 -> 0.49 L] [steps=5] [key=value...]"""
import sys
sys.path.insert(0, ".")
import torch
import nbody_b200.workloads as wl
from nbody_b200 import _lib

rel = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
opts = dict(kv.split("=") for kv in sys.argv[3:])
w = wl.water_omm(32, Rel=rel or None)
side = torch.cuda.Stream(); torch.cuda.set_stream(side)
ctx = _lib.Context(0); ctx.set_stream(side.cuda_stream)
ctx.system(w["ms"], qs=w["qs"], water=True); ctx.boundary(_lib.BC_CUBIC, [w["L"]])
ctx.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"]); ctx.add_coulomb(w["coulomb"]["k"], w["coulomb"]["R"])
ctx.add_spcfw(w["spcfw"]["rOH"], w["spcfw"]["aHOH"], w["spcfw"]["kb"], w["spcfw"]["ka"])
for k, val in opts.items():
    ctx.set_option(k, int(val))
ctx.upload(w["u"], w["v"]); ctx.step_vv(w["dt"], 2); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ctx.step_vv(w["dt"], steps); e1.record(); torch.cuda.synchronize()
print(f"water 32^3, Coulomb cutoff {w['coulomb']['R']:.4f} nm, opts={opts}: {e0.elapsed_time(e1) / steps:.4f} ms/step", flush=True)
