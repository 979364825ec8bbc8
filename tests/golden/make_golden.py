"""Generates the golden input/output vectors under tests/golden/ (run from the repo root:
python tests/golden/make_golden.py).

PARITY UNPINNED -- read this before trusting the files.  The reference is a Julia package, Julia is not installed
in this image (nor on the GPU box) and the reference ships no golden force vectors, so these outputs
do NOT come from a run of the reference: they come from oracle/nbody_oracle.c, the C restatement of
src/basic_potentials.jl:240-433, src/boundary_conditions.jl:111-172, src/nbody_to_ode.jl:474-488 / :502-532 and
src/thermostats.jl:76-128, compiled with gcc -O2 -ffp-contract=off.  What the files pin is therefore: (a) the C
oracle against silent change, (b) the independent pure-Python restatement oracle/nbody_oracle_np.py against the C
oracle bit for bit (tests/test_golden.py, CPU), (c) the CUDA path against fixed numbers without running any oracle
(tests/test_gpu_golden.py).  Anyone with Julia can regenerate the `dv` arrays with baseline/julia/golden_from_reference.jl
from the same inputs and compare.

Each case is one .npz: inputs (u, v, ms, qs, mm, scalars of the spec as a JSON string) and outputs (dv, and for the
cutoff cases the CSR neighbour lists offsets/neigh of the reference predicate `r2 < R2`).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import nbody_b200.workloads as wl  # noqa: E402
from oracle import nbody_oracle as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def F(a):
    return np.asfortranarray(np.array(a, dtype=np.float64))


def spec_system(spec, arrays):
    return orc.System(arrays["ms"], qs=arrays.get("qs"), mm=arrays.get("mm"), water=spec.get("water", False),
                      bc=tuple(spec["bc"]), gravity=spec.get("gravity"), lj=spec.get("lj"), coulomb=spec.get("coulomb"),
                      dipole=spec.get("dipole"), spcfw=spec.get("spcfw"), thermostat=spec.get("thermostat"))


def emit(name, spec, arrays, neighbors_R=None, idx_stride=1, analysis=False):
    s = spec_system(spec, arrays)
    u, v = arrays["u"], arrays["v"].copy(order="F")
    out = dict(arrays)
    out["dv"] = s.rhs(u, v)
    if analysis:  # rdf pair histogram of the frame (src/nbody_simulation_result.jl:676-693) and msd against a second frame
        L = spec["bc"][1]
        out["rdf_hist"] = orc.rdf_hist(u, L, idx_stride=3 if spec.get("water") else 1)
        u1 = F(u + 0.02 * L * np.random.Generator(np.random.Philox(7)).standard_normal(u.shape))
        out["u1"] = u1
        w = arrays["ms"]
        out["msd"] = np.array(orc.msd(u1, u, water=bool(spec.get("water")), mO=w[0], mH=w[1] if len(w) > 1 else 0.0))
    out["v_after"] = v  # Nose-Hoover writes v[zeta_ind] (src/thermostats.jl:126)
    if neighbors_R is not None:
        n = len(arrays["ms"])
        lists = [s.neighbors(u, i, neighbors_R, idx_stride) if i % idx_stride == 0 else np.zeros(0, np.int32) for i in range(n)]
        out["offsets"] = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.int64)
        out["neigh"] = np.concatenate(lists).astype(np.int32)
    out["spec"] = np.array(json.dumps(spec))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: n = {len(arrays['ms'])}, |dv|max = {np.abs(out['dv']).max():.6e}")


def main():
    orc.build()
    rng = np.random.Generator(np.random.Philox(20261017))

    # 1. three-body figure-eight, G = 1 (test/gravitational_test.jl:7-18)
    u = F([[-0.995492, 0.995492, 0.0], [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]])
    v = F([[-0.347902, -0.347902, 0.695804], [-0.53393, -0.53393, 1.06786], [0.0, 0.0, 0.0]])
    emit("gravity_figure_eight", dict(bc=["infinite"], gravity=dict(G=1.0)), dict(u=u, v=v, ms=np.ones(3)))

    # 2. Plummer sphere, 512 bodies (config 2 at reduced size)
    u, v, ms = wl.plummer(512)
    emit("gravity_plummer_512", dict(bc=["infinite"], gravity=dict(G=1.0)), dict(u=u, v=v, ms=ms))

    # 3. liquid argon as shipped: 216 atoms, SI units, cubic PBC, R = 0.5 L (config 1; examples/liquid_argon.jl:34-57)
    w = wl.liquid_argon_si(216)
    x = F(w["u"] + 0.02 * w["lj"]["sigma"] * rng.standard_normal(w["u"].shape))
    emit("lj_argon_si_216", dict(bc=["cubic", w["L"]], lj=w["lj"]), dict(u=x, v=w["v"], ms=w["ms"]),
         neighbors_R=w["lj"]["R"])

    # 4. reduced-unit FCC argon, 500 atoms, R = 2.25 sigma, drifted out of the box by whole box lengths (config 3 style)
    w = wl.fcc_argon_reduced(5)
    x = w["u"] + 0.05 * rng.standard_normal(w["u"].shape) + w["L"] * rng.integers(-2, 3, size=w["u"].shape)
    emit("lj_argon_reduced_500_berendsen",
         dict(bc=["cubic", w["L"]], lj=w["lj"], thermostat=dict(kind="berendsen", T=90.0, tau=10 * w["dt"], kB=w["kB"])),
         dict(u=F(x), v=w["v"], ms=w["ms"]), neighbors_R=w["lj"]["R"], analysis=True)

    # 5. PeriodicBoundaryConditions is NOT a minimum image (src/boundary_conditions.jl:111-136): 64 atoms
    n, L = 64, 4.0
    x = F(rng.random((3, n)) * L * 1.4 - 0.2 * L)
    emit("lj_periodic6_64", dict(bc=["periodic", [0.0, L, 0.0, L, 0.0, L]], lj=dict(eps=0.8, sigma=0.9, R=1.9)),
         dict(u=x, v=F(rng.standard_normal((3, n))), ms=rng.random(n) + 0.5), neighbors_R=1.9)

    # 6. charged particles, InfiniteBox, R = inf (config 5a style) and 7. magnetic dipoles (config 5b style)
    c = wl.charged_lattice(343)
    emit("coulomb_infinite_343", dict(bc=["infinite"], coulomb=dict(k=c["coulomb"]["k"])),
         dict(u=c["u"], v=c["v"], ms=c["ms"], qs=c["qs"]))
    d = wl.dipole_lattice(216)
    emit("dipole_216", dict(bc=["infinite"], dipole=d["dipole"]), dict(u=d["u"], v=d["v"], ms=d["ms"], mm=d["mm"]))

    # 8. SPC/Fw water, 27 molecules, OMM units, both cutoffs inside the box (config 4 style; src/nbody_to_ode.jl:502-532)
    w = wl.water_omm(3, Rel=0.45)
    x = F(w["u"] + 0.005 * rng.standard_normal(w["u"].shape))
    lj = dict(w["lj"]); lj["R"] = 0.45
    emit("water_spcfw_27", dict(bc=["cubic", w["L"]], water=True, lj=lj, coulomb=w["coulomb"], spcfw=w["spcfw"]),
         dict(u=x, v=w["v"], ms=w["ms"], qs=w["qs"]), analysis=True)

    # 9. Nose-Hoover: the state carries one extra column (src/nbody_to_ode.jl:6-8, src/thermostats.jl:121-128)
    w = wl.fcc_argon_reduced(3)
    n = w["u"].shape[1]
    x = np.zeros((3, n + 1), order="F"); x[:, :n] = w["u"] + 0.05 * rng.standard_normal(w["u"].shape); x[0, n] = 0.37
    vv = np.zeros((3, n + 1), order="F"); vv[:, :n] = w["v"]
    emit("lj_argon_reduced_108_nosehoover",
         dict(bc=["cubic", w["L"]], lj=w["lj"], thermostat=dict(kind="nosehoover", T=90.0, tau=20 * w["dt"], kB=w["kB"], N=n)),
         dict(u=x, v=vv, ms=w["ms"]))


if __name__ == "__main__":
    main()
