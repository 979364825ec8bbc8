"""The CUDA path against EXACT arithmetic, without the oracle in between: the 40-digit mpmath restatements of the reference's
formulas (tests/test_oracle_exact.py) on small systems, through the C ABI.  Tolerance 1e-12 per body (BASELINE.json north_star;
1e-3 of the system RMS as floor where a sum cancels).  Small systems take the warp-per-target kernel of csrc/nbx_allpairs.cu."""
import numpy as np
import pytest

pytest.importorskip("mpmath")

from tests._common import make_context  # noqa: E402
from tests.test_oracle_exact import _box, _close, _exact  # noqa: E402

pytestmark = pytest.mark.gpu


def test_gravity_gpu_against_exact_arithmetic():
    u, rng = _box(20, 3.0, 1)
    ms = rng.random(20) + 0.5
    ctx = make_context(dict(ms=ms, gravity=dict(G=0.7)))
    _close(ctx.accel(u), _exact(u, ms, "gravity", 0.7), tol=1e-12)
    ctx.close()


def test_lennard_jones_gpu_against_exact_arithmetic():
    L, R = 7.0, 2.5
    u, rng = _box(40, L, 2)
    ms = rng.random(40) + 0.5
    lj = dict(eps=1.3, sigma=0.9, R=R)
    ctx = make_context(dict(ms=ms, bc=("cubic", L), lj=lj))
    _close(ctx.accel(u), _exact(u, ms, "lj", lj, L=L, R=R), tol=1e-12)
    ctx.close()


def test_lennard_jones_cell_list_gpu_against_exact_arithmetic():
    """Large enough for the cell / Verlet-list path (L / (R + skin) >= 3), small enough for the exact evaluation."""
    L, R = 9.0, 2.5
    u, rng = _box(160, L, 7)
    ms = rng.random(160) + 0.5
    lj = dict(eps=1.3, sigma=0.9, R=R)
    ctx = make_context(dict(ms=ms, bc=("cubic", L), lj=lj))
    a = ctx.accel(u)
    assert ctx.info("cells_lj") > 0
    _close(a, _exact(u, ms, "lj", lj, L=L, R=R), tol=1e-12)
    ctx.close()


def test_coulomb_cutoff_gpu_against_exact_arithmetic():
    L, R = 5.0, 0.49 * 5.0
    u, rng = _box(30, L, 3)
    ms, qs = rng.random(30) + 0.5, rng.standard_normal(30)
    ctx = make_context(dict(ms=ms, qs=qs, bc=("cubic", L), coulomb=dict(k=2.5, R=R)))
    _close(ctx.accel(u), _exact(u, ms, "coulomb", 2.5, qs=qs, L=L, R=R), tol=1e-12)
    ctx.close()


def test_dipoles_gpu_against_exact_arithmetic():
    rng = np.random.default_rng(4)
    u = np.asfortranarray(rng.random((3, 14)) * 2.0)
    ms = rng.random(14) + 0.5
    mm = np.asfortranarray(rng.standard_normal((3, 14)))
    ctx = make_context(dict(ms=ms, mm=mm, dipole=dict(mu_4pi=1e-2)))
    _close(ctx.accel(u), _exact(u, ms, "dipole", 1e-2, mm=mm), tol=1e-12)
    ctx.close()
