"""CPU-side checks: the C-ABI library loads and exports every symbol include/nbody_b200.h declares (no
compute calls without a GPU), the product path fails loudly without a device, and the host-side logic
of the reference-shaped API (gathers, degrees of freedom, constructors) follows the reference."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from nbody_b200 import build as b
    from nbody_b200 import _lib

    b.build()  # nvcc cross-compiles sm_100a without a GPU
    return _lib


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nbody_b200.h")).read()
    return sorted(set(re.findall(r"NBX_API\s+[\w\s\*]*?\b(nbx_\w+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    names = _declared_symbols()
    assert len(names) >= 30
    cdll = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(cdll, n), f"{n} declared in include/nbody_b200.h but not exported"
    assert set(names) == set(lib.SIGNATURES), (set(names) ^ set(lib.SIGNATURES))
    assert lib.load().nbx_version() >= 100


def test_library_holds_sm100a_code_only(lib):
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_tma_and_rsqrt_seed_in_sass(lib):
    """The all-pairs kernel stages sources with TMA bulk copies (UBLKCP) and seeds 1/sqrt with MUFU.RSQ64H."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "MUFU.RSQ64H" in sass and "SYNCS.ARRIVE.TRANS64" in sass


def test_no_cpu_fallback_without_device(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lib.NbxError) as e:
        lib.Context(0)
    assert e.value.code == lib.ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "nbodysimulator.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "nbody_oracle" not in text and "from oracle" not in text, f


# ---- host logic of the reference-shaped API ----------------------------------------------------
def test_gather_coordinates_and_dof_follow_the_reference():
    from nbody_b200 import api

    sp = api.SPCFwParameters(0.1012, 113.24 * math.pi / 180, 1.0, 1.0)
    bodies = [api.MassBody([1.0, 2.0, 3.0], [0.1, 0.2, 0.3], 18.0), api.MassBody([4.0, 5.0, 6.0], [0, 0, 0], 18.0)]
    water = api.WaterSPCFw(bodies, 1.0, 16.0, 0.41, -0.82, api.LennardJonesParameters(),
                           api.ElectrostaticParameters(), sp)
    sim = api.NBodySimulation(water, (0.0, 1.0), api.CubicPeriodicBoundaryConditions(10.0),
                              api.NoseHooverThermostat(300.0, 0.1), 1.0)
    u0, v0, n = api.gather_bodies_initial_coordinates(sim)
    assert n == 2 and u0.shape == (3, 7)          # 3n + 1 columns with Nose-Hoover (nbody_to_ode.jl:27-31)
    assert np.allclose(u0[:, 1], [1.0 + sp.rOH, 2.0, 3.0])                                    # :47
    assert np.allclose(u0[:, 2], [1.0 + math.cos(sp.aHOH) * sp.rOH, 2.0, 3.0 + math.sin(sp.aHOH) * sp.rOH])
    assert np.array_equal(v0[:, 0], v0[:, 1]) and np.array_equal(v0[:, 0], v0[:, 2])
    assert api.get_degrees_of_freedom(water) == (6, 4, 14)   # test/water_test.jl:95-134: n=6, nc=4
    assert list(api.get_masses(water)) == [16.0, 1.0, 1.0, 16.0, 1.0, 1.0]


def test_constructors_and_defaults():
    from nbody_b200 import api

    assert api.LennardJonesParameters().R == 2.5 and api.LennardJonesParameters(1, 2, 3).σ2 == 4
    assert api.GravitationalParameters().G == 6.67408e-11
    assert api.ElectrostaticParameters().k == 9e9 and math.isinf(api.ElectrostaticParameters().R2)
    assert api.MagnetostaticParameters().μ_4π == 1e-7
    assert api.BerendsenThermostat(300.0, 0.1).γ == 5.0
    assert api.PeriodicBoundaryConditions(2.0).boundary == (0.0, 2.0, 0.0, 2.0, 0.0, 2.0)
    b = [api.MassBody([0, 0, 0], [0, 0, 0], 1.0)]
    s = api.PotentialNBodySystem(b, potentials=["gravitational", "lennard_jones"])
    assert set(s.potentials) == {"gravitational", "lennard_jones"}
    assert str(s).startswith("Potentials: \nLennard-Jones:")     # fixed print order (nbody_system.jl:124-134)
    sim = api.NBodySimulation(api.GravitationalSystem(b, 2.0), (0.0, 1.0))
    assert isinstance(sim.system, api.PotentialNBodySystem) and sim.system.potentials["gravitational"].G == 2.0
    assert isinstance(sim.boundary_conditions, api.InfiniteBox) and sim.kb == api.kb_SI
    sim2 = api.NBodySimulation(s, (0.0, 1.0), api.CubicPeriodicBoundaryConditions(3.0), 2.5)  # (sys, tspan, bc, kb)
    assert sim2.kb == 2.5 and isinstance(sim2.thermostat, api.NullThermostat)


def test_cell_node_lattice_matches_reference_rule():
    import nbody_b200.workloads as wl

    pos = wl.cell_node_positions(10, 3.0)       # ceil(cbrt(10)) = 3 nodes per edge, dL = 1, z fastest
    assert pos.shape == (3, 10)
    assert np.allclose(pos[:, 0], [0.5, 0.5, 0.5]) and np.allclose(pos[:, 1], [0.5, 0.5, 1.5])
    assert np.allclose(pos[:, 3], [0.5, 1.5, 0.5]) and np.allclose(pos[:, 9], [1.5, 0.5, 0.5])
