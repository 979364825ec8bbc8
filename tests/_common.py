"""Shared helpers of the parity tests: build the CPU oracle and a CUDA context from ONE spec dict.

spec keys: ms, qs, mm, water, bc (("infinite",) | ("cubic", L) | ("periodic", six)), gravity {G},
lj {eps, sigma, R}, coulomb {k, R}, dipole {mu_4pi}, spcfw {rOH, aHOH, kb, ka},
thermostat {kind, T, tau, kB, N, Nc}.
"""
import numpy as np

BC_KIND = {"infinite": 0, "cubic": 1, "periodic": 2}
THERMO_KIND = {"berendsen": 1, "nosehoover": 2, "andersen": 3, "langevin": 4}


def F(a):
    return np.asfortranarray(np.array(a, dtype=np.float64))


def make_oracle(orc, spec):
    return orc.System(spec["ms"], qs=spec.get("qs"), mm=spec.get("mm"), water=spec.get("water", False),
                      bc=spec.get("bc", ("infinite",)), gravity=spec.get("gravity"), lj=spec.get("lj"),
                      coulomb=spec.get("coulomb"), dipole=spec.get("dipole"), spcfw=spec.get("spcfw"),
                      thermostat=spec.get("thermostat"))


def make_context(spec, device=0):
    from nbody_b200._lib import Context

    ctx = Context(device)
    th = spec.get("thermostat")
    n = len(spec["ms"])
    ctx.system(spec["ms"], qs=spec.get("qs"), mm=spec.get("mm"), water=spec.get("water", False))
    bc = spec.get("bc", ("infinite",))
    ctx.boundary(BC_KIND[bc[0]], None if bc[0] == "infinite" else bc[1])
    if spec.get("lj"):
        ctx.add_lj(spec["lj"]["eps"], spec["lj"]["sigma"], spec["lj"]["R"])
    if spec.get("coulomb"):
        ctx.add_coulomb(spec["coulomb"]["k"], spec["coulomb"].get("R", np.inf))
    if spec.get("dipole"):
        ctx.add_dipole(spec["dipole"]["mu_4pi"])
    if spec.get("gravity"):
        ctx.add_gravity(spec["gravity"]["G"])
    if spec.get("spcfw"):
        s = spec["spcfw"]
        ctx.add_spcfw(s["rOH"], s["aHOH"], s["kb"], s["ka"])
    if th:
        param = th.get("tau", th.get("nu", th.get("gamma", 0.0)))
        ctx.thermostat(THERMO_KIND[th["kind"]], th["T"], param, th["kB"], th.get("N", n), th.get("Nc", 0))
    return ctx


def rel_err_per_body(a, ref):
    """||a_i - ref_i||_2 / ||ref_i||_2 per column (the parity metric of SURVEY.md 7.2)."""
    num = np.linalg.norm(a - ref, axis=0)
    den = np.linalg.norm(ref, axis=0)
    den = np.where(den == 0.0, 1.0, den)
    return num / den
