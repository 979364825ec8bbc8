"""GPU: the CUDA path (through the C ABI) against the committed golden vectors -- no oracle code runs here.

Tolerance: per-body accelerations <= 1e-12 relative (BASELINE.json north_star), neighbour lists bit-exact."""
import numpy as np
import pytest

from tests._common import make_context
from tests._golden import CASES, load
from tests.test_gpu_parity import _check

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", CASES)
def test_cuda_path_reproduces_golden(name):
    spec, z = load(name)
    n = len(spec["ms"])
    ctx = make_context(spec)
    v = z["v"].copy(order="F")
    a = ctx.accel(z["u"], v)
    _check(a[:, :n], z["dv"][:, :n])
    if "nosehoover" in name:
        assert not a[:, n].any()
        assert v[0, n] == pytest.approx(z["v_after"][0, n], rel=1e-13)
    if "rdf_hist" in z:  # frame analysis: integer histogram count for count, msd to rounding
        ctx.rdf_reset(1000)
        ctx.rdf_add(z["u"])
        hist, frames = ctx.rdf_get()
        assert frames == 1 and np.array_equal(hist, z["rdf_hist"])
        assert ctx.msd(z["u"], z["u1"]) == pytest.approx(float(z["msd"]), rel=1e-13)
    if "offsets" in z:
        ctx.upload(z["u"], z["v"])
        off, lst = ctx.neighbors()
        assert np.array_equal(off, z["offsets"]) and np.array_equal(lst, z["neigh"])
    ctx.close()
