"""Development probe: slab group on one device, leader vs threads, with timings."""
import sys, time, threading
sys.path.insert(0, ".")
import numpy as np
import nbody_b200.workloads as wl
from nbody_b200.parallel import join_group_local
from tests._common import F, make_context

w = wl.fcc_argon_reduced(12)
rng = np.random.Generator(np.random.Philox(5))
u = F(w["u"] + 0.05 * rng.standard_normal(w["u"].shape)); v = F(3.0 * w["v"])
spec = dict(ms=w["ms"], bc=("cubic", w["L"]), lj=w["lj"])
mode = sys.argv[1]
opts = dict(kv.split("=") for kv in sys.argv[2:])
t0 = time.time()
try:
    if mode == "leader":
        g = make_context(spec, device=[0, 0])
        g.set_option("spin_timeout_ms", 4000)
        for k, val in opts.items():
            g.set_option(k, int(val))
        g.upload(u, v); print("upload", time.time() - t0, flush=True)
        for n in (2, 6, 50):
            t1 = time.time(); g.step_vv(2e-3, n); print("step", n, time.time() - t1, flush=True)
    else:
        cs = []
        for _ in range(2):
            c = make_context(spec); c.set_option("spin_timeout_ms", 4000)
            for k, val in opts.items():
                c.set_option(k, int(val))
            c.upload(u, v); cs.append(c)
        join_group_local(cs); print("joined", time.time() - t0, flush=True)
        for n in (2, 6, 50):
            errs = []
            def run(c):
                try: c.step_vv(2e-3, n)
                except Exception as e: errs.append(e)
            t1 = time.time(); ts = [threading.Thread(target=run, args=(c,)) for c in cs]
            [t.start() for t in ts]; [t.join() for t in ts]
            print("step", n, time.time() - t1, errs, flush=True)
except Exception as e:
    print("FAILED after", time.time() - t0, e)
