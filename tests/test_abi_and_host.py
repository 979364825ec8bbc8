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


def _split_args(text):
    out, depth, cur = [], 0, ""
    for ch in text:
        if ch in "({[":
            depth += 1
        if ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _c_kind(decl):
    """Class of one C parameter / return type of include/nbody_b200.h."""
    d = decl.strip()
    if d in ("void", ""):
        return None
    if "*" in d:
        return "cstr" if re.match(r"const\s+char\s*\*", d) else "ptr"
    base = re.sub(r"\b(const|unsigned)\b", lambda m: "u" if m.group(1) == "unsigned" else "", d)
    base = re.sub(r"\s+\w+$", "", base.strip()) if " " in base.strip() else base.strip()  # drop the parameter name
    return {"int": "i32", "int32_t": "i32", "int64_t": "i64", "uint64_t": "u64", "double": "f64"}[base.strip()]


def _header_prototypes():
    text = open(os.path.join(ROOT, "include", "nbody_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"NBX_API\s+([\w\s\*]*?)\b(nbx_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        protos[name] = (_c_kind(ret), [k for k in (_c_kind(a) for a in _split_args(args)) if k is not None])
    return protos


def _ctypes_kind(t):
    if t is None:
        return None
    if t in (ctypes.c_char_p,):
        return "cstr"
    if t is ctypes.c_void_p or hasattr(t, "contents") or issubclass(t, ctypes._Pointer):
        return "ptr"
    return {ctypes.c_int: "i32", ctypes.c_int32: "i32", ctypes.c_int64: "i64", ctypes.c_uint64: "u64", ctypes.c_double: "f64"}[t]


def test_ctypes_signatures_match_the_header(lib):
    """Every prototype of include/nbody_b200.h against the ctypes binding: same arity, same class (pointer / 32- or 64-bit
    integer / double / C string) of every parameter and of the return value."""
    protos = _header_prototypes()
    assert set(protos) == set(lib.SIGNATURES)
    for name, (ret, args) in protos.items():
        cres, cargs = lib.SIGNATURES[name]
        assert _ctypes_kind(cres) == ret, name
        assert [_ctypes_kind(a) for a in cargs] == args, (name, args)


def test_julia_shim_ccalls_match_the_header():
    """julia/NBodySimulatorB200.jl cannot be executed here (no Julia): every ccall in it is checked statically against the
    header -- the symbol exists, the arity is right and every argument has the class the C prototype wants."""
    protos = _header_prototypes()
    text = open(os.path.join(ROOT, "julia", "NBodySimulatorB200.jl")).read()
    jl = {"Cint": "i32", "Int64": "i64", "UInt64": "u64", "Float64": "f64", "Cstring": "cstr"}

    def kind(t):
        t = t.strip()
        return "ptr" if t.startswith(("Ptr{", "Ref{")) else jl[t]

    calls = re.findall(r"ccall\(\(:(nbx_\w+), LIB\),\s*(\w+),\s*\(([^)]*)\)", text)
    assert len(calls) >= 25
    seen = set()
    for name, ret, args in calls:
        assert name in protos, f"{name} is not in include/nbody_b200.h"
        want_ret, want_args = protos[name]
        assert kind(ret) == want_ret, name
        got = [kind(a) for a in _split_args(args)]
        want = ["ptr" if k == "cstr" else k for k in want_args]
        assert got == want, (name, got, want)
        seen.add(name)
    for must in ("nbx_create", "nbx_create_multi", "nbx_system", "nbx_boundary", "nbx_thermostat", "nbx_accel", "nbx_upload",
                 "nbx_run_vv", "nbx_step_em", "nbx_download", "nbx_rdf_add", "nbx_msd"):
        assert must in seen


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


# ------------------------------------------------------------------------------------------------
# trajectory text: the PDB writer / loader of the reference-shaped host API (no device involved)
# ------------------------------------------------------------------------------------------------
# The fixture of test/water_test.jl:196-219 (the reference's own known-answer input for extract_from_pdb).
_PDB_FIXTURE = """MODEL     1
REMARK 250 time=0.0000 picoseconds
HETATM    1  O   HOH     1      18.642  18.642   0.000  1.00  0.00
HETATM    2  H1  HOH     1      17.685  18.642   0.000  1.00  0.00
HETATM    3  H2  HOH     1      18.882  17.715   0.000  1.00  0.00
HETATM    4  O   HOH     2       0.000   0.000   3.107  1.00  0.00
HETATM    5  H1  HOH     2       0.957   0.000   3.107  1.00  0.00
HETATM    6  H2  HOH     2      -0.240   0.927   3.107  1.00  0.00
HETATM    7  O   HOH     3      18.642  18.642   6.214  1.00  0.00
HETATM    8  H1  HOH     3      17.685  18.642   6.214  1.00  0.00
HETATM    9  H2  HOH     3      18.882  17.715   6.214  1.00  0.00
ENDMDL
MODEL     2
REMARK 250 time=0.0005 picoseconds
HETATM    1  O   HOH     1      18.642  18.642  -0.000  1.00  0.00
HETATM    2  H1  HOH     1      17.685  18.642   0.000  1.00  0.00
HETATM    3  H2  HOH     1      18.882  17.715   0.000  1.00  0.00
HETATM    4  O   HOH     2      -0.000  -0.000   3.107  1.00  0.00
HETATM    5  H1  HOH     2       0.958  -0.000   3.107  1.00  0.00
HETATM    6  H2  HOH     2      -0.240   0.928   3.107  1.00  0.00
HETATM    7  O   HOH     3      18.642  18.642   6.214  1.00  0.00
HETATM    8  H1  HOH     3      17.685  18.642   6.214  1.00  0.00
HETATM    9  H2  HOH     3      18.881  17.715   6.214  1.00  0.00
ENDMDL"""


def test_extract_from_pdb_known_answers():
    """test/water_test.jl:195-248: positions of the first frame / 10, zero velocities."""
    import io

    from nbody_b200.api import extract_from_pdb

    bodies = extract_from_pdb(io.StringIO(_PDB_FIXTURE))
    eps = 0.00005
    assert len(bodies) == 3
    assert np.allclose(bodies[0].O.r, [1.8642, 1.8642, 0.0], atol=eps)
    assert np.allclose(bodies[1].H1.r, [0.0957, 0.0, 0.3107], atol=eps)
    assert np.allclose(bodies[2].H2.r, [1.8882, 1.7715, 0.6214], atol=eps)
    assert not np.any(bodies[2].H2.v) and bodies[0].O.m == 15.999 and bodies[0].H1.m == 1.00794


def test_pdb_writer_layout_and_round_trip():
    """write_pdb_data (src/nbody_simulation_result.jl:792-865): one MODEL / REMARK 250 / ENDMDL per saved time, one
    HETATM per atom (test/lennard_jones_test.jl:99-117, test/water_test.jl:173-192), fixed columns, coordinates in
    Angstrom wrapped into the box with molecules kept whole; the loader reads the writer's output back."""
    import io

    from nbody_b200.api import (CubicPeriodicBoundaryConditions, ElectrostaticParameters, LennardJonesParameters, MassBody,
                                NBodySimulation, PotentialNBodySystem, SimulationResult, SPCFwParameters, WaterSPCFw,
                                extract_from_pdb, write_pdb_data)

    L = 1.5
    pbc = CubicPeriodicBoundaryConditions(L)
    atoms = [MassBody([0.1, 0.2, 0.3], [0, 0, 0], 1.0), MassBody([1.6, -0.1, 0.7], [0, 0, 0], 1.0)]
    sim = NBodySimulation(PotentialNBodySystem(atoms, {"lennard_jones": LennardJonesParameters(1.0, 0.3, 0.7)}), (0.0, 1.0), pbc, 1.0)
    frames = [np.asfortranarray([[0.1, 1.6], [0.2, -0.1], [0.3, 0.7]]), np.asfortranarray([[0.15, 1.7], [0.2, -0.2], [0.3, 0.7]])]
    sr = SimulationResult(sim, [0.0, 5e-5], frames, [np.zeros((3, 2))] * 2, 1)
    out = io.StringIO()
    write_pdb_data(out, sr)
    lines = out.getvalue().split("\n")
    assert sum(ln.startswith("REMARK 250") for ln in lines) == 2 and sum(ln.startswith("HETATM") for ln in lines) == 4
    assert lines[0] == "MODEL     1" and lines[1] == "REMARK 250 time=0.0 steps" and lines[6] == "REMARK 250 time=5.0e-5 steps"
    assert lines[2] == "HETATM    1  Ar  Ar     1       1.000   2.000   3.000  1.00  0.00          Ar"
    assert lines[3] == "HETATM    2  Ar  Ar     2       1.000  14.000   7.000  1.00  0.00          Ar"  # wrapped into [0, 15)

    mols = [MassBody([1.45, 0.5, 0.5], [0, 0, 0], 18.0), MassBody([0.2, 0.9, 1.2], [0, 0, 0], 18.0)]
    water = WaterSPCFw(mols, 1.00794, 15.999, 0.41, -0.82, LennardJonesParameters(0.65, 0.3165, 0.7),
                       ElectrostaticParameters(138.9, 0.7), SPCFwParameters(0.1012, 1.976, 443153.0, 317.5))
    wsim = NBodySimulation(water, (0.0, 1.0), pbc, 1.0)
    from nbody_b200.api import gather_bodies_initial_coordinates

    u0, v0, _ = gather_bodies_initial_coordinates(wsim)
    wsr = SimulationResult(wsim, [0.0, 0.0005], [u0, u0 + 0.001], [v0, v0], 1)
    out = io.StringIO()
    write_pdb_data(out, wsr)
    text = out.getvalue()
    assert text.count("HETATM") == 12 and text.count("picoseconds") == 2
    first = text.split("\n")[2]
    assert first == "HETATM    1  O   HOH     1      14.500   5.000   5.000  1.00  0.00           O"
    back = extract_from_pdb(io.StringIO(text))
    assert len(back) == 2
    # H1 of molecule 1 sits at x = 1.45 + 0.1012 = 1.5512 > L: written next to its oxygen (15.512 A), not wrapped away from it
    assert np.allclose(back[0].H1.r, [1.5512, 0.5, 0.5], atol=5e-4) and np.allclose(back[0].O.r, [1.45, 0.5, 0.5], atol=5e-4)
    assert np.allclose(back[1].O.r, u0[:, 3], atol=5e-4)


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference runs on the host cores only (the CPU restatement of the reference's loop): one JSON line
    with the keys the driver reads, here with one short step."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "pair-interactions/s" and d["value"] > 1e7
    assert d["config"]["workload"] == "gravity_plummer_262144" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["gpu_launches"] == 0


def test_bench_own_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback anywhere: without a CUDA device the bench's own arm exits with an error instead of timing
    something else."""
    import subprocess
    import sys

    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
