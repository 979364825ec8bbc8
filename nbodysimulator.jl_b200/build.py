"""In-tree build of libnbody_b200.so (hand-written sm_100a CUDA behind the C ABI of include/nbody_b200.h).

``python -m``-free on purpose: ``__graft_entry__.build()`` and the ctypes loader call :func:`build`.
nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB = os.path.join(HERE, "libnbody_b200.so")
SOURCES = ["nbx_api.cu", "nbx_allpairs.cu", "nbx_sympairs.cu", "nbx_cells.cu", "nbx_slab.cu", "nbx_multi.cu", "nbx_group.cu", "nbx_bonded.cu", "nbx_integrate.cu", "nbx_energy.cu", "nbx_analysis.cu"]
HEADERS = [os.path.join(CSRC, "nbx_internal.cuh"), os.path.join(CSRC, "nbx_graph.inl"), os.path.join(HERE, "..", "include", "nbody_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall", "-Xptxas", "-v",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libnbody_b200.so cannot be built (there is no CPU fallback)")
    return exe


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src: str, force: bool) -> tuple[str, str]:
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    log = ""
    if force or _stale(obj, [path] + HEADERS):
        r = subprocess.run([nvcc(), *NVCC_FLAGS, "-c", path, "-o", obj], capture_output=True, text=True)
        log = r.stdout + r.stderr
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{log}")
        with open(obj + ".log", "w") as f:
            f.write(log)
    return obj, log


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a and link libnbody_b200.so in-tree.  Returns the library path."""
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(lambda s: _compile(s, force), SOURCES))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log)
    if force or _stale(LIB, objs):
        r = subprocess.run([nvcc(), "-shared", "-o", LIB, *objs, "-cudart", "static",
                            "-gencode", "arch=compute_100a,code=sm_100a"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose=True))
