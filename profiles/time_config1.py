"""Config 1 (examples/liquid_argon.jl as shipped, 216 atoms, R = L/2): step time and, under ncu, the launch list."""
import sys
sys.path.insert(0, ".")
import torch
import nbody_b200.workloads as wl
from nbody_b200 import _lib
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
opts = dict(kv.split("=") for kv in sys.argv[2:])
w = wl.liquid_argon_si(216)
side = torch.cuda.Stream(); torch.cuda.set_stream(side)
ctx = _lib.Context(0); ctx.set_stream(side.cuda_stream)
ctx.system(w["ms"]); ctx.boundary(_lib.BC_CUBIC, [w["L"]]); ctx.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"])
for k, val in opts.items():
    ctx.set_option(k, int(val))
ctx.upload(w["u"], w["v"]); ctx.step_vv(w["dt"], 10); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ctx.step_vv(w["dt"], steps); e1.record(); torch.cuda.synchronize()
print(f"216 atoms opts={opts}: {e0.elapsed_time(e1) / steps * 1e3:.2f} us/step graph_cached={ctx.info('graph_cached')}", flush=True)
