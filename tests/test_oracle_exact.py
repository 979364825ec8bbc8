"""The CPU oracle against EXACT arithmetic: the reference's formulas (src/basic_potentials.jl:240-365, the cubic minimum image
of src/boundary_conditions.jl:138-165) written a third time, in 40-digit mpmath, on small systems.  The oracle's fp64 result
may differ from the exact one only by rounding: <= 1e-13 of a body's acceleration (1e-3 of the system RMS as floor where the
sum cancels).  This does not pin the oracle against Julia (nothing here can), it pins its arithmetic against slips that two
fp64 restatements by the same hand could share.  CPU only."""
import numpy as np
import pytest

mp = pytest.importorskip("mpmath")
mp.mp.dps = 40

from oracle import nbody_oracle as orc  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _built():
    orc.build()


def _wrap(d, L):
    """src/boundary_conditions.jl:143-160 on an exact number: while d >= L/2 subtract L, while d < -L/2 add L."""
    half = L / 2
    while d >= half:
        d -= L
    while d < -half:
        d += L
    return d


def _exact(u, ms, kind, par, qs=None, mm=None, L=None, R=None):
    n = u.shape[1]
    X = [[mp.mpf(float(u[d, i])) for d in range(3)] for i in range(n)]
    out = np.zeros((3, n))
    Lm = None if L is None else mp.mpf(float(L))
    R2 = None if R is None else mp.mpf(float(R)) ** 2
    for i in range(n):
        f = [mp.mpf(0)] * 3
        for j in range(n):
            if j == i:
                continue
            r = [X[i][d] - X[j][d] for d in range(3)]
            if Lm is not None:
                r = [_wrap(c, Lm) for c in r]
            r2 = sum(c * c for c in r)
            if R2 is not None and not (r2 < R2):
                continue
            if kind == "gravity":        # :321-325
                fac = -mp.mpf(par) * mp.mpf(float(ms[j])) / mp.sqrt(r2) ** 3
                f = [f[d] + fac * r[d] for d in range(3)]
            elif kind == "lj":           # :258-265
                s6 = (mp.mpf(par["sigma"]) ** 2 / r2) ** 3
                fac = (2 * s6 * s6 - s6) / r2
                f = [f[d] + fac * r[d] for d in range(3)]
            elif kind == "coulomb":      # :292-297
                fac = mp.mpf(float(qs[j])) / (mp.sqrt(r2) * r2)
                f = [f[d] + fac * r[d] for d in range(3)]
            else:                        # dipole :344-357
                mi = [mp.mpf(float(mm[d, i])) for d in range(3)]
                mj = [mp.mpf(float(mm[d, j])) for d in range(3)]
                rn = mp.sqrt(r2)
                rh = [c / rn for c in r]
                mir = sum(a * b for a, b in zip(mi, rh))
                mjr = sum(a * b for a, b in zip(mj, rh))
                mimj = sum(a * b for a, b in zip(mi, mj))
                f = [f[d] + (mi[d] * mjr + mj[d] * mir + rh[d] * mimj - 5 * rh[d] * mir * mjr) / r2 ** 2 for d in range(3)]
        if kind == "lj":
            coeff = 24 * mp.mpf(par["eps"]) / mp.mpf(float(ms[i]))          # :267
        elif kind == "coulomb":
            coeff = mp.mpf(par) * mp.mpf(float(qs[i])) / mp.mpf(float(ms[i]))  # :299
        elif kind == "dipole":
            coeff = 3 * mp.mpf(par) / mp.mpf(float(ms[i]))                   # :360
        else:
            coeff = mp.mpf(1)
        out[:, i] = [float(coeff * c) for c in f]
    return out


def _close(a, exact, tol=1e-13):
    norms = np.linalg.norm(exact, axis=0)
    floor = 1e-3 * np.sqrt(np.mean(norms ** 2))
    err = np.linalg.norm(a - exact, axis=0) / np.maximum(norms, floor)
    assert err.max() <= tol, err.max()


def _box(n, L, seed, min_gap=0.05):
    """Random positions in [0, L)^3 (some drifted out of the box by whole box lengths: the reference never wraps them)."""
    rng = np.random.default_rng(seed)
    u = rng.random((3, n)) * L
    u += L * rng.integers(-2, 3, size=u.shape) * (rng.random(u.shape) < 0.2)
    return np.asfortranarray(u), rng


def test_gravity_against_exact_arithmetic():
    u, rng = _box(20, 3.0, 1)
    ms = rng.random(20) + 0.5
    s = orc.System(ms, gravity=dict(G=0.7))
    _close(s.rhs(u, np.zeros_like(u)), _exact(u, ms, "gravity", 0.7))


def test_lennard_jones_cubic_box_against_exact_arithmetic():
    L, R = 7.0, 2.5
    u, rng = _box(40, L, 2)
    ms = rng.random(40) + 0.5
    lj = dict(eps=1.3, sigma=0.9, R=R)
    s = orc.System(ms, bc=("cubic", L), lj=lj)
    _close(s.rhs(u, np.zeros_like(u)), _exact(u, ms, "lj", lj, L=L, R=R))


def test_coulomb_cutoff_cubic_box_against_exact_arithmetic():
    L, R = 5.0, 0.49 * 5.0
    u, rng = _box(30, L, 3)
    ms, qs = rng.random(30) + 0.5, rng.standard_normal(30)
    s = orc.System(ms, qs=qs, bc=("cubic", L), coulomb=dict(k=2.5, R=R))
    _close(s.rhs(u, np.zeros_like(u)), _exact(u, ms, "coulomb", 2.5, qs=qs, L=L, R=R))


def test_dipoles_against_exact_arithmetic():
    rng = np.random.default_rng(4)
    u = np.asfortranarray(rng.random((3, 14)) * 2.0)
    ms = rng.random(14) + 0.5
    mm = np.asfortranarray(rng.standard_normal((3, 14)))
    s = orc.System(ms, mm=mm, dipole=dict(mu_4pi=1e-2))
    _close(s.rhs(u, np.zeros_like(u)), _exact(u, ms, "dipole", 1e-2, mm=mm))


def _cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def _unit(a):
    n = mp.sqrt(sum(c * c for c in a))
    return [c / n for c in a]


def test_spcfw_bonds_and_angle_against_exact_arithmetic():
    """harmonic_bond_potential_acceleration! (:367-393, every bond seen from both ends) and
    valence_angle_potential_acceleration! (:395-433, called with a = H1, b = O, c = H2: src/nbody_to_ode.jl:255-260)."""
    rng = np.random.default_rng(6)
    nm = 5
    mO, mH = 15.999, 1.00794
    rOH, aHOH, kb, ka = 0.1012, 113.24 * np.pi / 180, 1059.162 * 4.184 * 1e2, 75.9 * 4.184
    o = rng.random((3, nm)) * 2.0
    u = np.zeros((3, 3 * nm), order="F")
    u[:, 0::3] = o
    u[:, 1::3] = o + np.array([[rOH], [0.0], [0.0]]) + 0.01 * rng.standard_normal((3, nm))
    u[:, 2::3] = o + np.array([[np.cos(aHOH) * rOH], [0.0], [np.sin(aHOH) * rOH]]) + 0.01 * rng.standard_normal((3, nm))
    ms = np.tile([mO, mH, mH], nm)
    s = orc.System(ms, qs=np.zeros(3 * nm), water=True, spcfw=dict(rOH=rOH, aHOH=aHOH, kb=kb, ka=ka))
    got = s.rhs(u, np.zeros_like(u))
    exact = np.zeros_like(u)
    M = [mp.mpf(float(m)) for m in ms]
    for m in range(nm):
        O, H1, H2 = 3 * m, 3 * m + 1, 3 * m + 2
        X = {k: [mp.mpf(float(u[d, k])) for d in range(3)] for k in (O, H1, H2)}
        acc = {k: [mp.mpf(0)] * 3 for k in (O, H1, H2)}
        for i, j in ((O, H1), (O, H2), (H1, O), (H2, O)):
            rij = [X[i][d] - X[j][d] for d in range(3)]
            r = mp.sqrt(sum(c * c for c in rij))
            fac = -(r - mp.mpf(rOH)) * mp.mpf(kb) / r
            acc[i] = [acc[i][d] + fac * rij[d] / M[i] for d in range(3)]
        rba = [X[H1][d] - X[O][d] for d in range(3)]
        rbc = [X[H2][d] - X[O][d] for d in range(3)]
        rcb = [-c for c in rbc]
        x = _cross(rba, rbc)
        pa, pc = _unit(_cross(rba, x)), _unit(_cross(rcb, x))
        nba, nbc = mp.sqrt(sum(c * c for c in rba)), mp.sqrt(sum(c * c for c in rbc))
        ang = mp.acos(sum(a * b for a, b in zip(rba, rbc)) / (nba * nbc))
        force = -mp.mpf(ka) * (ang - mp.mpf(aHOH))
        fa = [c * force / nba for c in pa]
        fc = [c * force / nbc for c in pc]
        fb = [-(a + c) for a, c in zip(fa, fc)]
        for k, f in ((H1, fa), (O, fb), (H2, fc)):
            acc[k] = [acc[k][d] + f[d] / M[k] for d in range(3)]
        for k in (O, H1, H2):
            exact[:, k] = [float(c) for c in acc[k]]
    _close(got, exact, tol=1e-12)   # (acos near 113 degrees and the stiff bond constant amplify the last bits)
