"""Export the golden cases and the benchmark workloads as raw fp64 files + a small JSON manifest for the Julia
scripts in this directory.  python baseline/julia/export_inputs.py <outdir> [--bench]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import nbody_b200.workloads as wl  # noqa: E402
from tests._golden import CASES, load  # noqa: E402


def main():
    out = sys.argv[1]
    os.makedirs(out, exist_ok=True)
    manifest = {}
    for name in CASES:
        spec, z = load(name)
        wl.dump_raw(os.path.join(out, name), **{k: z[k] for k in ("u", "v", "dv", "ms", "qs", "mm") if k in z})
        manifest[name] = {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in spec.items() if k not in ("ms", "qs", "mm")}
        manifest[name]["n"] = int(len(spec["ms"]))
    if "--bench" in sys.argv:
        u, v, ms = wl.plummer(262144)
        wl.dump_raw(os.path.join(out, "bench_gravity_262144"), u=u, v=v, ms=ms)
        manifest["bench_gravity_262144"] = dict(bc=["infinite"], gravity=dict(G=1.0), n=262144)
        w = wl.liquid_argon_si(216)
        wl.dump_raw(os.path.join(out, "bench_argon_216"), u=w["u"], v=w["v"], ms=w["ms"])
        manifest["bench_argon_216"] = dict(bc=["cubic", w["L"]], lj=w["lj"], n=216, dt=w["dt"])
    with open(os.path.join(out, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("wrote", len(manifest), "cases to", out)


if __name__ == "__main__":
    main()
