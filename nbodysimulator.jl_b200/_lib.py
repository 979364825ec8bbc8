"""ctypes binding of libnbody_b200.so -- exactly the symbols include/nbody_b200.h declares.

No CPU fallback: :func:`load` raises if the library is missing, and ``nbx_create`` fails without a
CUDA device.  :class:`Context` is a thin object wrapper over one ``nbx_ctx``; the reference-shaped
host API lives in ``api.py``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnbody_b200.so")

NBX_OK = 0
ERR_INVALID, ERR_CUDA, ERR_NONFINITE, ERR_UNSUPPORTED, ERR_CAPACITY = -1, -2, -3, -4, -5
BC_INFINITE, BC_CUBIC, BC_PERIODIC = 0, 1, 2
THERMO_NONE, THERMO_BERENDSEN, THERMO_NOSEHOOVER, THERMO_ANDERSEN, THERMO_LANGEVIN = 0, 1, 2, 3, 4
T_PAIR_ALLPAIRS, T_CELL_BUILD, T_PAIR_CELLS, T_BONDED, T_INTEGRATE, T_TRANSPOSE = 0, 1, 2, 3, 4, 5

_dp = C.POINTER(C.c_double)
_vp = C.c_void_p
_i64 = C.c_int64

# name -> (restype, argtypes); must list every NBX_API symbol of include/nbody_b200.h
SIGNATURES = {
    "nbx_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "nbx_create_multi": (C.c_int, [C.POINTER(_vp), C.c_int, C.POINTER(C.c_int)]),
    "nbx_group_init": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    "nbx_group_export": (C.c_int, [_vp, C.c_int, C.POINTER(_vp), _vp]),
    "nbx_group_connect": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "nbx_group_start": (C.c_int, [_vp]),
    "nbx_destroy": (C.c_int, [_vp]),
    "nbx_last_error": (C.c_char_p, [_vp]),
    "nbx_version": (C.c_int, []),
    "nbx_system": (C.c_int, [_vp, _i64, _dp, _dp, _dp, C.c_int]),
    "nbx_boundary": (C.c_int, [_vp, C.c_int, _dp]),
    "nbx_add_gravity": (C.c_int, [_vp, C.c_double]),
    "nbx_add_lj": (C.c_int, [_vp, C.c_double, C.c_double, C.c_double]),
    "nbx_add_coulomb": (C.c_int, [_vp, C.c_double, C.c_double]),
    "nbx_add_dipole": (C.c_int, [_vp, C.c_double]),
    "nbx_add_spcfw": (C.c_int, [_vp, C.c_double, C.c_double, C.c_double, C.c_double]),
    "nbx_clear_potentials": (C.c_int, [_vp]),
    "nbx_thermostat": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, C.c_double, _i64, _i64]),
    "nbx_shard": (C.c_int, [_vp, _i64, _i64]),
    "nbx_shard_pairs": (C.c_int, [_vp, C.c_int, C.c_int]),
    "nbx_vv_forces": (C.c_int, [_vp]),
    "nbx_accel": (C.c_int, [_vp, _dp, _dp, C.c_double, _dp]),
    "nbx_upload": (C.c_int, [_vp, _dp, _dp]),
    "nbx_step_vv": (C.c_int, [_vp, C.c_double, _i64]),
    "nbx_step_em": (C.c_int, [_vp, C.c_double, _i64, C.c_uint64]),
    "nbx_download": (C.c_int, [_vp, _dp, _dp, _dp]),
    "nbx_run_vv": (C.c_int, [_vp, C.c_double, _i64, _i64, _dp, _dp, _i64, C.POINTER(_i64)]),
    "nbx_set_seed": (C.c_int, [_vp, C.c_uint64]),
    "nbx_vv_begin": (C.c_int, [_vp, C.c_double]),
    "nbx_vv_finish": (C.c_int, [_vp, C.c_double]),
    "nbx_eval_resident": (C.c_int, [_vp]),
    "nbx_energy": (C.c_int, [_vp, _dp, _dp, _dp]),
    "nbx_neighbors": (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(C.c_int32), _i64]),
    "nbx_slab_init": (C.c_int, [_vp, C.c_int, C.c_int]),
    "nbx_slab_pack": (C.c_int, [_vp]),
    "nbx_slab_unpack": (C.c_int, [_vp, C.POINTER(_i64)]),
    "nbx_slab_prime": (C.c_int, [_vp]),
    "nbx_slab_step_begin": (C.c_int, [_vp, C.c_double, C.c_double]),
    "nbx_slab_step_end": (C.c_int, [_vp, C.c_double, C.c_int]),
    "nbx_slab_refresh_send": (C.c_int, [_vp]),
    "nbx_slab_refresh_recv": (C.c_int, [_vp]),
    "nbx_slab_verlet_check": (C.c_int, [_vp, C.c_double, _vp]),
    "nbx_slab_rx": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_i64), _vp]),
    "nbx_slab_connect": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "nbx_slab_check": (C.c_int, [_vp, C.POINTER(_i64)]),
    "nbx_slab_buffer": (C.c_int, [_vp, C.c_int, C.POINTER(_vp), C.POINTER(_i64)]),
    "nbx_slab_download": (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(C.c_int32), _dp, _dp, _dp]),
    "nbx_set_stream": (C.c_int, [_vp, _vp]),
    "nbx_synchronize": (C.c_int, [_vp]),
    "nbx_device_ptr": (C.c_int, [_vp, C.c_int, C.POINTER(_vp), C.POINTER(_i64)]),
    "nbx_accel_device": (C.c_int, [_vp, _vp, _vp, C.c_double, _vp]),
    "nbx_timing_enable": (C.c_int, [_vp, C.c_int]),
    "nbx_timing_get": (C.c_int, [_vp, C.c_int, _dp, C.POINTER(_i64)]),
    "nbx_timing_reset": (C.c_int, [_vp]),
    "nbx_set_option": (C.c_int, [_vp, C.c_char_p, _i64]),
    "nbx_get_info": (C.c_int, [_vp, C.c_char_p, C.POINTER(_i64)]),
    "nbx_measure_fp64_peak": (C.c_int, [_vp, _dp, _dp]),
    "nbx_measure_hbm_peak": (C.c_int, [_vp, _dp]),
    "nbx_accel_begin": (C.c_int, [_vp, _dp]),
    "nbx_accel_end": (C.c_int, [_vp, _dp]),
    "nbx_rdf_reset": (C.c_int, [_vp, C.c_int]),
    "nbx_rdf_add": (C.c_int, [_vp, _dp]),
    "nbx_rdf_get": (C.c_int, [_vp, C.POINTER(_i64), _i64, C.POINTER(_i64)]),
    "nbx_msd": (C.c_int, [_vp, _dp, _dp, _dp]),
}

_lib = None


class NbxError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"nbx error {code}: {msg}")
        self.code = code


def load(build_if_missing: bool = False):
    """dlopen libnbody_b200.so and set the prototypes.  Raises if it is not there (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            if build_if_missing:
                from . import build as _b

                _b.build()
            else:
                raise OSError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`; "
                              "there is no CPU fallback for the acceleration path")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _f(a, ncols=None):
    """float64 (3, ncols) array with Julia's Matrix{Float64} bytes (Fortran order)."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim != 2 or a.shape[0] != 3 or (ncols is not None and a.shape[1] != ncols):
        raise ValueError(f"expected a (3, {ncols if ncols is not None else 'n'}) array, got {a.shape}")
    return a if a.flags.f_contiguous else np.asfortranarray(a)


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


GROUP_AUTO, GROUP_PAIRS, GROUP_TARGETS, GROUP_SLABS = 0, 1, 2, 3


class Context:
    """One nbx_ctx: one simulation on one GPU, or -- ``device`` a list -- on several GPUs of this process
    (nbx_create_multi: the same calls, fanned out by the library; a device may repeat)."""

    def __init__(self, device=0):
        self.lib = load()
        h = _vp()
        if isinstance(device, (list, tuple)):
            devs = (C.c_int * len(device))(*[int(d) for d in device])
            rc = self.lib.nbx_create_multi(C.byref(h), len(device), devs)
            self.group_size = len(device)
        else:
            rc = self.lib.nbx_create(C.byref(h), int(device))
            self.group_size = 1
        if rc != NBX_OK:
            raise NbxError(rc, self.lib.nbx_last_error(None).decode())
        self.h = h
        self.n = 0
        self.ncols = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.nbx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != NBX_OK:
            raise NbxError(rc, self.lib.nbx_last_error(self.h).decode())

    def _refresh(self):
        self.ncols = self.info("ncols")

    # -- description -------------------------------------------------------------------
    def system(self, ms, qs=None, mm=None, water=False):
        ms = np.ascontiguousarray(ms, dtype=np.float64)
        qs = None if qs is None else np.ascontiguousarray(qs, dtype=np.float64)
        mm = None if mm is None else _f(mm, ms.shape[0])
        if qs is not None and qs.shape != ms.shape:
            raise ValueError("charges must match masses")
        self._ck(self.lib.nbx_system(self.h, ms.shape[0], _p(ms), _p(qs), _p(mm), int(bool(water))))
        self.n = int(ms.shape[0])
        self._refresh()

    def boundary(self, kind, box=None):
        b = None if box is None else np.ascontiguousarray(np.atleast_1d(box), dtype=np.float64)
        self._ck(self.lib.nbx_boundary(self.h, int(kind), _p(b)))

    def add_gravity(self, G):
        self._ck(self.lib.nbx_add_gravity(self.h, float(G)))

    def add_lj(self, eps, sigma, R):
        self._ck(self.lib.nbx_add_lj(self.h, float(eps), float(sigma), float(R)))

    def add_coulomb(self, k, R=float("inf")):
        self._ck(self.lib.nbx_add_coulomb(self.h, float(k), float(R)))

    def add_dipole(self, mu_4pi):
        self._ck(self.lib.nbx_add_dipole(self.h, float(mu_4pi)))

    def add_spcfw(self, rOH, aHOH, kb, ka):
        self._ck(self.lib.nbx_add_spcfw(self.h, float(rOH), float(aHOH), float(kb), float(ka)))

    def clear_potentials(self):
        self._ck(self.lib.nbx_clear_potentials(self.h))

    def thermostat(self, kind, T0=0.0, param=0.0, kB=1.0, N=None, Nc=0):
        self._ck(self.lib.nbx_thermostat(self.h, int(kind), float(T0), float(param), float(kB),
                                         int(self.n if N is None else N), int(Nc)))
        self._refresh()

    def shard(self, lo, hi):
        self._ck(self.lib.nbx_shard(self.h, int(lo), int(hi)))

    # -- RHS drop-in -------------------------------------------------------------------------
    def accel(self, u, v=None, t=0.0, out=None):
        """soode_system!(dv, v, u, p, t): returns dv (3, ncols).  ``v`` is mutated for Nose-Hoover."""
        u = _f(u, self.ncols)
        if v is not None:
            if not (isinstance(v, np.ndarray) and v.dtype == np.float64 and v.flags.f_contiguous
                    and v.shape == (3, self.ncols)):
                raise ValueError("v must be a float64 Fortran-ordered (3, ncols) array (it may be written)")
        dv = np.empty((3, self.ncols), order="F") if out is None else out
        self._ck(self.lib.nbx_accel(self.h, _p(u), _p(v), float(t), _p(dv)))
        return dv

    def accel_begin(self, u):
        """Split-phase RHS drop-in of a pair-sharded context: partial accelerations of all bodies stay on the device."""
        self._ck(self.lib.nbx_accel_begin(self.h, _p(_f(u, self.ncols))))

    def accel_end(self, out=None):
        dv = np.empty((3, self.ncols), order="F") if out is None else out
        self._ck(self.lib.nbx_accel_end(self.h, _p(dv)))
        return dv

    # -- resident stepping ---------------------------------------------------------------------
    def upload(self, u, v):
        self._ck(self.lib.nbx_upload(self.h, _p(_f(u, self.ncols)), _p(_f(v, self.ncols))))

    def step_vv(self, dt, nsteps=1):
        self._ck(self.lib.nbx_step_vv(self.h, float(dt), int(nsteps)))

    def run_vv(self, dt, nsteps, save_every=1, want_v=True):
        """nsteps velocity-Verlet steps with the state saved every ``save_every`` steps: (u_frames, v_frames), each
        (frames, 3, ncols) with the frames' arrays in Julia's column-major layout."""
        frames = -(-int(nsteps) // int(save_every))
        u = np.zeros((frames, self.ncols, 3))   # frame-major; each frame is 3 x ncols column-major = (ncols, 3) C-order
        v = np.zeros((frames, self.ncols, 3)) if want_v else None
        got = _i64()
        self._ck(self.lib.nbx_run_vv(self.h, float(dt), int(nsteps), int(save_every), _p(u), _p(v), frames, C.byref(got)))
        u = u[: got.value].transpose(0, 2, 1)
        return u, (v[: got.value].transpose(0, 2, 1) if want_v else None)

    def step_em(self, dt, nsteps=1, seed=0):
        self._ck(self.lib.nbx_step_em(self.h, float(dt), int(nsteps), int(seed)))

    def vv_begin(self, dt):
        self._ck(self.lib.nbx_vv_begin(self.h, float(dt)))

    def vv_forces(self):
        self._ck(self.lib.nbx_vv_forces(self.h))

    def vv_finish(self, dt):
        self._ck(self.lib.nbx_vv_finish(self.h, float(dt)))

    def shard_pairs(self, rank, nranks):
        self._ck(self.lib.nbx_shard_pairs(self.h, int(rank), int(nranks)))

    def eval_resident(self):
        self._ck(self.lib.nbx_eval_resident(self.h))

    def set_seed(self, seed):
        self._ck(self.lib.nbx_set_seed(self.h, int(seed)))

    def download(self, want_u=True, want_v=True, want_dv=False):
        # (a group member fills only its own columns of v and dv: the others stay zero)
        u = np.zeros((3, self.ncols), order="F") if want_u else None
        v = np.zeros((3, self.ncols), order="F") if want_v else None
        dv = np.zeros((3, self.ncols), order="F") if want_dv else None
        self._ck(self.lib.nbx_download(self.h, _p(u), _p(v), _p(dv)))
        return u, v, dv

    def energy(self, potential=True):
        ek, ep, T = C.c_double(), C.c_double(), C.c_double()
        self._ck(self.lib.nbx_energy(self.h, C.byref(ek), C.byref(ep) if potential else None, C.byref(T)))
        return ek.value, (ep.value if potential else None), T.value

    def neighbors(self, cap=None):
        n = self.n // 3 if self.info("water") else self.n
        offsets = np.zeros(n + 1, dtype=np.int64)
        cap = int(cap if cap is not None else 256 * n)
        lst = np.zeros(max(cap, 1), dtype=np.int32)
        self._ck(self.lib.nbx_neighbors(self.h, offsets.ctypes.data_as(C.POINTER(_i64)),
                                        lst.ctypes.data_as(C.POINTER(C.c_int32)), cap))
        return offsets, lst[: offsets[-1]]

    # -- groups across processes / contexts (nbx_group_*) ------------------------------------------
    def group_init(self, rank, nranks, mode=GROUP_AUTO):
        self._ck(self.lib.nbx_group_init(self.h, int(rank), int(nranks), int(mode)))

    def group_export(self, want_handles=True):
        """(pointers[4], handles: 4 x 64 bytes or None) of the memory the peers map: window, slab receive area,
        position rows, staging area."""
        ptrs, blob = [], b""
        for kind in range(4):
            p = _vp()
            h = C.create_string_buffer(64) if want_handles else None
            self._ck(self.lib.nbx_group_export(self.h, kind, C.byref(p), h))
            ptrs.append(int(p.value or 0))
            if want_handles:
                blob += h.raw
        return ptrs, (blob if want_handles else None)

    def group_connect(self, handles=None, ptrs=None):
        """handles: bytes of nranks x 4 x 64 (other processes); ptrs: nranks x 4 device pointers (this process)."""
        hb = C.create_string_buffer(handles, len(handles)) if handles is not None else None
        pa = None
        if ptrs is not None:
            flat = [int(x) for row in ptrs for x in row]
            pa = (_vp * len(flat))(*[_vp(x) if x else None for x in flat])
        self._ck(self.lib.nbx_group_connect(self.h, hb, pa))

    def group_start(self):
        self._ck(self.lib.nbx_group_start(self.h))

    # -- slab decomposition ------------------------------------------------------------------------
    def slab_init(self, rank, nranks):
        self._ck(self.lib.nbx_slab_init(self.h, int(rank), int(nranks)))

    def slab_pack(self):
        self._ck(self.lib.nbx_slab_pack(self.h))

    def slab_unpack(self, sync=True):
        """sync=False: asynchronous (no host round trip); errors surface at the next slab_check()."""
        if not sync:
            self._ck(self.lib.nbx_slab_unpack(self.h, None))
            return None
        counts = (_i64 * 6)()
        self._ck(self.lib.nbx_slab_unpack(self.h, counts))
        return [int(x) for x in counts]

    def slab_check(self):
        counts = (_i64 * 6)()
        self._ck(self.lib.nbx_slab_check(self.h, counts))
        return [int(x) for x in counts]

    def slab_prime(self):
        self._ck(self.lib.nbx_slab_prime(self.h))

    def slab_step_begin(self, dt, soft_fraction=0.75):
        self._ck(self.lib.nbx_slab_step_begin(self.h, float(dt), float(soft_fraction)))

    def slab_step_end(self, dt, refresh):
        self._ck(self.lib.nbx_slab_step_end(self.h, float(dt), int(bool(refresh))))

    def slab_refresh_send(self):
        self._ck(self.lib.nbx_slab_refresh_send(self.h))

    def slab_refresh_recv(self):
        self._ck(self.lib.nbx_slab_refresh_recv(self.h))

    def slab_verlet_check(self, out_ptr, soft_fraction=0.75):
        """Enqueues the displacement check of the own particles; out_ptr: device pointer to two int32."""
        self._ck(self.lib.nbx_slab_verlet_check(self.h, float(soft_fraction), _vp(out_ptr)))

    def slab_rx(self, want_handle=False):
        """(device pointer, doubles, 64-byte CUDA IPC handle or None) of this slab's receive area."""
        p, nd = _vp(), _i64()
        h = C.create_string_buffer(64) if want_handle else None
        self._ck(self.lib.nbx_slab_rx(self.h, C.byref(p), C.byref(nd), h))
        return int(p.value), int(nd.value), (h.raw if want_handle else None)

    def slab_connect(self, left_handle=None, right_handle=None, left_ptr=None, right_ptr=None):
        self._ck(self.lib.nbx_slab_connect(self.h, left_handle, right_handle, _vp(left_ptr) if left_ptr else None,
                                           _vp(right_ptr) if right_ptr else None))

    def slab_buffer(self, which):
        p, nd = _vp(), _i64()
        self._ck(self.lib.nbx_slab_buffer(self.h, int(which), C.byref(p), C.byref(nd)))
        return int(p.value), int(nd.value)

    def slab_download(self):
        """(gid, u, v, dv) of the own particles of this slab."""
        cap = self.n
        m = _i64()
        gid = np.zeros(cap, dtype=np.int32)
        u, v, dv = (np.zeros((3, cap), order="F") for _ in range(3))
        self._ck(self.lib.nbx_slab_download(self.h, C.byref(m), gid.ctypes.data_as(C.POINTER(C.c_int32)), _p(u), _p(v),
                                            _p(dv)))
        k = int(m.value)
        return gid[:k].copy(), np.asfortranarray(u[:, :k]), np.asfortranarray(v[:, :k]), np.asfortranarray(dv[:, :k])

    # -- plumbing ----------------------------------------------------------------------------------
    def set_stream(self, stream_ptr):
        self._ck(self.lib.nbx_set_stream(self.h, _vp(stream_ptr) if stream_ptr else None))

    def synchronize(self):
        self._ck(self.lib.nbx_synchronize(self.h))

    def device_ptr(self, which):
        p, ld = _vp(), _i64()
        self._ck(self.lib.nbx_device_ptr(self.h, int(which), C.byref(p), C.byref(ld)))
        return int(p.value), int(ld.value)

    def accel_device(self, u_ptr, v_ptr, dv_ptr, t=0.0):
        self._ck(self.lib.nbx_accel_device(self.h, _vp(u_ptr), _vp(v_ptr) if v_ptr else None, float(t), _vp(dv_ptr)))

    def timing_enable(self, on=True):
        self._ck(self.lib.nbx_timing_enable(self.h, int(bool(on))))

    def timing_get(self, phase):
        ms, cnt = C.c_double(), _i64()
        self._ck(self.lib.nbx_timing_get(self.h, int(phase), C.byref(ms), C.byref(cnt)))
        return ms.value, int(cnt.value)

    def timing_reset(self):
        self._ck(self.lib.nbx_timing_reset(self.h))

    def set_option(self, key, value):
        self._ck(self.lib.nbx_set_option(self.h, key.encode(), int(value)))

    def info(self, key):
        v = _i64()
        self._ck(self.lib.nbx_get_info(self.h, key.encode(), C.byref(v)))
        return int(v.value)

    def measure_fp64_peak(self):
        tf, mhz = C.c_double(), C.c_double()
        self._ck(self.lib.nbx_measure_fp64_peak(self.h, C.byref(tf), C.byref(mhz)))
        return tf.value, mhz.value

    # -- analysis of frames (rdf / msd of the reference's result accessors) ------------------------
    def rdf_reset(self, maxbin=1000):
        self._ck(self.lib.nbx_rdf_reset(self.h, int(maxbin)))
        self._rdf_bins = int(maxbin)

    def rdf_add(self, u=None):
        """Adds one frame's pair distances to the device histogram (u None: the resident positions)."""
        self._ck(self.lib.nbx_rdf_add(self.h, None if u is None else _p(_f(u, self.ncols))))

    def rdf_get(self):
        """(histogram as int64, hist[b - 1] = the reference's hist[b]; frames added)."""
        bins = getattr(self, "_rdf_bins", 1000)
        hist = np.zeros(bins, dtype=np.int64)
        frames = _i64()
        self._ck(self.lib.nbx_rdf_get(self.h, hist.ctypes.data_as(C.POINTER(_i64)), hist.size, C.byref(frames)))
        return hist, int(frames.value)

    def msd(self, u0, u=None):
        out = C.c_double()
        self._ck(self.lib.nbx_msd(self.h, _p(_f(u0, self.ncols)), None if u is None else _p(_f(u, self.ncols)), C.byref(out)))
        return out.value

    def measure_hbm_peak(self):
        g = C.c_double()
        self._ck(self.lib.nbx_measure_hbm_peak(self.h, C.byref(g)))
        return g.value
