"""Print the key figures of a bench.py JSON line (development helper)."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for k in ("n_gpus", "value", "ms_per_step", "e2e", "parity", "general_masses", "clocks", "gpu_launches"):
    print(k, d.get(k))
r = d.get("roofline", {})
print("roofline", {k: r.get(k) for k in ("achieved", "peak", "frac", "kernel_ms", "kernel_share_of_step", "rank0_share_of_pairs", "integrate_and_exchange_ms_per_step")})
lj = d.get("lj", {})
print("LJ", {k: lj.get(k) for k in ("value", "ms_per_step", "ms_per_step_pair_kernel", "ms_per_step_cell_build", "ms_per_step_integrate", "parity", "neighbour_structure", "graph", "rank0_own", "rank0_ghosts", "error")})
for k, v in (d.get("other_configs") or {}).items():
    print(k, v if not isinstance(v, dict) else {kk: v.get(kk) for kk in ("value", "ms_per_step", "decomposition", "n_atoms", "parity", "error")})
