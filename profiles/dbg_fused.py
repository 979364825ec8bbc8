import sys, numpy as np
sys.path.insert(0, '/root/repo')
import nbody_b200.workloads as wl
from tests._common import F, make_context
w = wl.fcc_argon_reduced(6)
rng = np.random.Generator(np.random.Philox(41))
u = F(w["u"] + 0.05 * rng.standard_normal(w["u"].shape)); v = F(0.0 * w["v"])
spec = dict(ms=w["ms"], bc=("cubic", w["L"]), lj=w["lj"])
C = int(sys.argv[1]); steps = int(sys.argv[2])
ctx = make_context(spec)
ref = ctx.accel(u).copy()
ctx.set_option("fused_step", 1); ctx.set_option("fused_cluster", C); ctx.set_option("fused_min_steps", -12345); ctx.set_option("graph", 0)
ctx.upload(u, v); ctx.step_vv(0.0, steps)
uu, vv, a = ctx.download(want_dv=True)
print({k: ctx.info(k) for k in ("fused_steps", "fused_disabled", "verlet_rebuilds", "fused_list_cap")}, np.isfinite(a).all(), np.array_equal(uu, u))
start = ctx.debug_fetch("start"); pid = ctx.debug_fetch("pid"); scell = ctx.debug_fetch("scell"); nlist = ctx.debug_fetch("nlist")
lst = ctx.debug_fetch("list"); x1 = ctx.debug_fetch("x", 1).reshape(-1, 4); x0 = ctx.debug_fetch("x", 0).reshape(-1, 4)
cap = len(pid); lst = lst.reshape(-1, cap)
ns = start[-1]
print("nslots", ns, "cap", cap, "start%C", np.unique(start % C), "real", (pid[:ns] >= 0).sum(), "n", u.shape[1])
print("pid perm ok", np.array_equal(np.sort(pid[:ns][pid[:ns] >= 0]), np.arange(u.shape[1])))
sc = scell[:ns].reshape(-1, C)
print("clusters in one cell", (sc == sc[:, :1]).all())
L = w["L"]; R = w["lj"]["R"]
pos = np.where(pid[:ns, None] >= 0, u.T[np.maximum(pid[:ns], 0)], np.nan)
print("x1 matches pos for real", np.array_equal(x1[:ns][pid[:ns] >= 0, :3], pos[pid[:ns] >= 0]))
missing = 0; dummy_listed = 0; listed = 0
for g in range(ns // C):
    ks = range(g * C, g * C + C)
    ent = set()
    for k in ks:
        e = lst[:nlist[k], k]
        ent |= set(e.tolist()); listed += len(e)
    dummy_listed += sum(1 for m in ent if pid[m] < 0)
    for k in ks:
        if pid[k] < 0: continue
        d = pos[k] - pos; d -= L * np.round(d / L)
        r2 = (d ** 2).sum(axis=1)
        inn = set(np.where(r2 < R * R)[0].tolist()) - {k}
        miss = inn - ent
        if miss and missing < 5:
            print("cluster", g, "slot", k, "pid", pid[k], "cell", scell[k], "missing", [(m, pid[m], scell[m]) for m in miss], "nlist", [nlist[q] for q in ks])
        missing += len(miss)
print("missing pairs", missing, "dummy listed", dummy_listed, "entries", listed, "per real", listed / u.shape[1])
ctx.close()
