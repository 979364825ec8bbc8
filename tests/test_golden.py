"""CPU: the oracle (C) and the independent pure-Python restatement against the committed golden vectors.

The golden `dv` arrays were produced by the C oracle (tests/golden/make_golden.py; parity unpinned: no Julia here),
so the C comparison guards the oracle against silent change (compiler flags, edits) and the Python comparison is a
second, independently written reading of src/basic_potentials.jl:240-433 / src/boundary_conditions.jl:111-172 /
src/nbody_to_ode.jl:474-532 that must land on the same bits.
"""
import numpy as np
import pytest

from oracle import nbody_oracle as orc
from oracle import nbody_oracle_np as onp
from tests._common import make_oracle
from tests._golden import CASES, load


def test_golden_cases_are_present():
    assert len(CASES) >= 9


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_reproduces_golden_bit_for_bit(name):
    orc.build()
    spec, z = load(name)
    v = z["v"].copy(order="F")
    dv = make_oracle(orc, spec).rhs(z["u"], v, 1)
    assert np.array_equal(dv, z["dv"])
    assert np.array_equal(v, z["v_after"])
    dv8 = make_oracle(orc, spec).rhs(z["u"], z["v"].copy(order="F"), 8)  # threads split targets, not sums
    assert np.array_equal(dv8, z["dv"])


@pytest.mark.parametrize("name", [c for c in CASES if "nosehoover" not in c and "plummer" not in c])
def test_python_restatement_reproduces_golden_bit_for_bit(name):
    """(Nose-Hoover is not in the Python restatement; the 512-body sphere is too slow for pure-Python loops.)"""
    spec, z = load(name)
    dv = onp.rhs(spec, z["u"], z["v"])
    assert np.array_equal(dv, z["dv"])


@pytest.mark.parametrize("name", [c for c in CASES if c.startswith("lj_")])
def test_golden_neighbor_lists_match_the_predicate(name):
    """The CSR lists are the reference predicate's pair set (src/basic_potentials.jl:258 on the distance of
    src/boundary_conditions.jl:111-165), here re-derived with the Python restatement of that distance."""
    spec, z = load(name)
    if "offsets" not in z:
        pytest.skip("no lists in this case")
    u, off, lst = z["u"], z["offsets"], z["neigh"]
    n = len(spec["ms"])
    R2 = spec["lj"]["R"] ** 2
    for i in range(0, n, max(1, n // 12)):
        ri = [u[k, i] for k in range(3)]
        mine = [j for j in range(n) if j != i and onp.distance(ri, [u[k, j] for k in range(3)], spec["bc"])[2] < R2]
        assert mine == lst[off[i]:off[i + 1]].tolist()


@pytest.mark.parametrize("name", [c for c in CASES if c in ("lj_argon_reduced_500_berendsen", "water_spcfw_27")])
def test_golden_rdf_and_msd(name):
    """rdf pair histogram and msd of the golden frames: the C oracle reproduces them bit for bit, and so does the
    pure-Python restatement (histogram; atomic msd)."""
    orc.build()
    spec, z = load(name)
    water = bool(spec.get("water"))
    L = spec["bc"][1]
    assert np.array_equal(orc.rdf_hist(z["u"], L, idx_stride=3 if water else 1), z["rdf_hist"])
    assert np.array_equal(np.array(onp.rdf_hist(z["u"], L, idx_stride=3 if water else 1)), z["rdf_hist"])
    ms = spec["ms"]
    assert orc.msd(z["u1"], z["u"], water=water, mO=ms[0], mH=ms[1]) == float(z["msd"])
    if not water:
        assert onp.msd(z["u1"], z["u"]) == float(z["msd"])
