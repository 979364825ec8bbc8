"""nbodysimulator.jl_b200 -- B200-native acceleration hot path of NBodySimulator.jl.

Holds only what the hot path needs: ``csrc/`` (hand-written sm_100a CUDA kernels behind the
C ABI declared in ``include/nbody_b200.h``), the ctypes loader (``_lib``), the host-side mirror
of the reference's plugin interface (``api``) and the seeded workload generators.
Import as ``nbody_b200`` through the shim at the repository root.
"""
__version__ = "0.1.0"
