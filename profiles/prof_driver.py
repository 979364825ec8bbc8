"""Short driver for ncu captures (never a timing source): a few launches of one workload's kernels.

  python profiles/prof_driver.py gravity [n]     all-pairs gravity, Plummer sphere
  python profiles/prof_driver.py lj [cells]      cell-list LJ argon, FCC cells^3 x 4 atoms
  python profiles/prof_driver.py water [side]    SPC/Fw water, side^3 molecules
  python profiles/prof_driver.py coulomb|dipole [n]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import nbody_b200.workloads as wl  # noqa: E402
from nbody_b200 import _lib  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "gravity"
    arg = int(sys.argv[2]) if len(sys.argv) > 2 else None
    steps = int(os.environ.get("PROF_STEPS", "2"))
    ctx = _lib.Context(0)
    if what == "gravity":
        u, v, ms = wl.plummer(arg or 262144)
        ctx.system(ms)
        ctx.add_gravity(1.0)
        dt = 1e-4
    elif what == "coulomb":
        w = wl.charged_lattice(arg or 65536)
        ctx.system(w["ms"], qs=w["qs"])
        ctx.add_coulomb(w["coulomb"]["k"])
        u, v, dt = w["u"], w["v"], 1e-6
    elif what == "dipole":
        w = wl.dipole_lattice(arg or 65536)
        ctx.system(w["ms"], mm=w["mm"])
        ctx.add_dipole(w["dipole"]["mu_4pi"])
        u, v, dt = w["u"], w["v"], 1e-6
    elif what == "lj":
        w = wl.fcc_argon_reduced(arg or 64)
        rng = np.random.Generator(np.random.Philox(2))
        u = np.asfortranarray(w["u"] + 0.05 * rng.standard_normal(w["u"].shape))
        v, dt = w["v"], w["dt"]
        ctx.system(w["ms"])
        ctx.boundary(_lib.BC_CUBIC, [w["L"]])
        ctx.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"])
        ctx.thermostat(_lib.THERMO_BERENDSEN, 90.0, 10 * dt, w["kB"], u.shape[1], 0)
        if "PROF_VERLET" in os.environ:
            ctx.set_option("verlet_skin_permille", int(os.environ["PROF_VERLET"]))
    elif what == "water":
        w = wl.water_omm(arg or 32, Rel=(float(os.environ.get("PROF_REL", "0.9162")) or None))
        u, v, dt = w["u"], w["v"], w["dt"]
        ctx.system(w["ms"], qs=w["qs"], water=True)
        ctx.boundary(_lib.BC_CUBIC, [w["L"]])
        ctx.add_lj(w["lj"]["eps"], w["lj"]["sigma"], w["lj"]["R"])
        ctx.add_coulomb(w["coulomb"]["k"], w["coulomb"]["R"])
        s = w["spcfw"]
        ctx.add_spcfw(s["rOH"], s["aHOH"], s["kb"], s["ka"])
    else:
        raise SystemExit(__doc__)
    for kv in filter(None, os.environ.get("PROF_OPTS", "").split(",")):  # e.g. PROF_OPTS=verlet_banked=0,verlet_lanes=4
        k, val = kv.split("=")
        ctx.set_option(k, int(val))
    ctx.upload(u, v)
    ctx.step_vv(dt, steps)
    ctx.synchronize()
    print(what, "done")


if __name__ == "__main__":
    main()
