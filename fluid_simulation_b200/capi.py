"""ctypes binding of include/fsb.h (one method per C entry point).

Method names follow the reference's stages (see the citations in fsb.h).  No
numpy arithmetic happens here: arrays are only handed to / filled by the
library.  Raises RuntimeError with the library's message on any error, which
is what the reference's step* functions throw (src/FluidSolver.cpp:101-107).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfsb.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
        "(nvcc, sm_100a).  There is no CPU fallback.")

_lib = C.CDLL(LIB_PATH)

U_FRONT, V_FRONT, U_BACK, V_BACK, U_PREV, V_PREV, U_DIFF, V_DIFF = range(8)
LIQUID, AIR, SOLID = 0, 1, 2
G2P_PIC, G2P_FLIP, G2P_PICFLIP = range(3)
STEP_SL, STEP_PIC, STEP_FLIP, STEP_PICFLIP = range(4)
INTEGRATOR_RK3, INTEGRATOR_EULER = 0, 1
SHARD_BLOB_BYTES = 512
PROF_NAMES = ["classify", "sort", "p2g", "grid_pre", "extend", "rhs", "cg", "patch", "g2p",
              "advect_sl", "advect_part"]

_f, _i, _p, _l = C.c_float, C.c_int, C.c_void_p, C.c_int64

# name -> (restype, argtypes); every symbol include/fsb.h declares
SIGNATURES = {
    "fsb_create": (_i, [C.POINTER(_p), _i, _i, _f, _f, _f, _f, _i]),
    "fsb_destroy": (None, [_p]),
    "fsb_last_error": (C.c_char_p, [_p]),
    "fsb_version": (C.c_char_p, []),
    "fsb_set_stream": (_i, [_p, _p]),
    "fsb_synchronize": (_i, [_p]),
    "fsb_size_x": (_i, [_p]),
    "fsb_size_y": (_i, [_p]),
    "fsb_delta_x": (_f, [_p]),
    "fsb_delta_y": (_f, [_p]),
    "fsb_set_cg": (_i, [_p, _i, _f]),
    "fsb_get_cg_info": (_i, [_p, C.POINTER(_i), C.POINTER(_f)]),
    "fsb_set_preconditioner": (_i, [_p, _i]),
    "fsb_set_pic_ratio": (_i, [_p, _f]),
    "fsb_set_density": (_i, [_p, _f]),
    "fsb_set_integrator": (_i, [_p, _i]),
    "fsb_set_gravity": (_i, [_p, _f, _f]),
    "fsb_set_pool": (_i, [_p, _i, _i, _f, _f]),
    "fsb_clear_cell_types": (_i, [_p]),
    "fsb_swap_velocity_buffers": (_i, [_p]),
    "fsb_set_particles": (_i, [_p, _p, _l]),
    "fsb_append_particles": (_i, [_p, _p, _l]),
    "fsb_num_particles": (_l, [_p]),
    "fsb_get_particles": (_i, [_p, _p]),
    "fsb_emit_source": (_i, [_p, _f, _f, _f, _f, _f, _f, _f, _f, C.POINTER(_l)]),
    "fsb_set_grid": (_i, [_p, _i, _p]),
    "fsb_get_grid": (_i, [_p, _i, _p]),
    "fsb_set_cell_types": (_i, [_p, _p]),
    "fsb_get_cell_types": (_i, [_p, _p]),
    "fsb_get_pressure": (_i, [_p, _p]),
    "fsb_classify_cells": (_i, [_p]),
    "fsb_p2g_spread": (_i, [_p]),
    "fsb_save_previous": (_i, [_p]),
    "fsb_add_acceleration": (_i, [_p, _f, _f, _f]),
    "fsb_enforce_dirichlet": (_i, [_p]),
    "fsb_extend_velocity": (_i, [_p, _i]),
    "fsb_extend_velocity_averaging": (_i, [_p, _i]),
    "fsb_pressure_solve": (_i, [_p, _f, _f]),
    "fsb_update_diff": (_i, [_p]),
    "fsb_g2p": (_i, [_p, _i, _f]),
    "fsb_advect_particles": (_i, [_p, _f, _i]),
    "fsb_advect_velocity_sl": (_i, [_p, _f]),
    "fsb_advect_particles_grid": (_i, [_p, _f]),
    "fsb_add_external_force": (_i, [_p, _f, _f, _f]),
    "fsb_p2g_gather": (_i, [_p]),
    "fsb_step": (_i, [_p, _i, _f]),
    "fsb_render_rgb": (_i, [_p, _i, _i, _f, _f, _f, _f, _p]),
    "fsb_write_ppm": (_i, [_p, C.c_char_p, _i, _i, _f, _f, _f, _f]),
    "fsb_save_state": (_i, [_p, C.c_char_p]),
    "fsb_load_state": (_i, [_p, C.c_char_p]),
    "fsb_shard_export": (_i, [_p, _p]),
    "fsb_shard_connect": (_i, [_p, _i, _i, _p]),
    "fsb_shard_disconnect": (_i, [_p]),
    "fsb_shard_rows": (_i, [_p, C.POINTER(_i), C.POINTER(_i)]),
    "fsb_slab_configure": (_i, [_p, _i, _i]),
    "fsb_slab_rows": (_i, [_p, C.POINTER(_i), C.POINTER(_i)]),
    "fsb_slab_add": (_i, [_p, _p, _p, _l]),
    "fsb_slab_sort_out": (_i, [_p, C.POINTER(_l)]),
    "fsb_slab_take": (_i, [_p, _i, _p, _p]),
    "fsb_slab_keep_own": (_i, [_p]),
    "fsb_slab_boundary": (_i, [_p, _i, C.POINTER(_l)]),
    "fsb_slab_boundary_take": (_i, [_p, _p, _p]),
    "fsb_slab_get": (_i, [_p, _p, _p]),
    "fsb_get_rows": (_i, [_p, _i, _i, _i, _p]),
    "fsb_set_rows": (_i, [_p, _i, _i, _i, _p]),
    "fsb_slab_step_a": (_i, [_p, _i]),
    "fsb_slab_step_b": (_i, [_p, _i, _f]),
    "fsb_slab_step_c": (_i, [_p, _i, _f]),
    "fsb_profile_enable": (_i, [_p, _i]),
    "fsb_profile_read": (_i, [_p, _p, _p]),
    "fsb_launch_count": (_l, [_p]),
    "fsb_cg_launch_mode": (_i, [_p]),
    "fsb_cg_swept_cells": (C.c_int64, [_p]),
    "fsb_timer_start": (_i, [_p]),
    "fsb_timer_stop": (_i, [_p, C.POINTER(_f)]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(_lib, _name)  # AttributeError here = the library lacks a declared symbol
    _fn.restype = _res
    _fn.argtypes = _args


PRECOND_JACOBI, PRECOND_MULTIGRID = 0, 1
ROWS_LABELS = 8


def version():
    return _lib.fsb_version().decode()


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Sim:
    """One simulation domain in the HBM of one B200 (FluidDomain + FluidSolver)."""

    def __init__(self, nx, ny, lx=1.0, ly=1.0, density=0.01, pic_ratio=0.05, device=0):
        h = _p()
        rc = _lib.fsb_create(C.byref(h), nx, ny, lx, ly, density, pic_ratio, device)
        if rc != 0:
            raise RuntimeError(f"fsb_create failed ({rc}): {_lib.fsb_last_error(None).decode()}")
        self.h = h
        self.nx, self.ny = nx, ny
        self.density, self.pic_ratio = density, pic_ratio
        self.dx = _lib.fsb_delta_x(h)
        self.dy = _lib.fsb_delta_y(h)

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(f"libfsb error {rc}: {_lib.fsb_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            _lib.fsb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # parameters
    def set_preconditioner(self, kind):
        """PRECOND_JACOBI (the reference's, default) or PRECOND_MULTIGRID (opt-in)."""
        self._ck(_lib.fsb_set_preconditioner(self.h, kind))

    def set_cg(self, max_iters, tol):
        self._ck(_lib.fsb_set_cg(self.h, max_iters, tol))

    def cg_info(self):
        it, err = _i(), _f()
        self._ck(_lib.fsb_get_cg_info(self.h, C.byref(it), C.byref(err)))
        return it.value, err.value

    def set_pic_ratio(self, r):
        self._ck(_lib.fsb_set_pic_ratio(self.h, r))
        self.pic_ratio = min(max(r, 0.0), 1.0)

    def set_integrator(self, k):
        self._ck(_lib.fsb_set_integrator(self.h, k))

    def set_stream(self, cuda_stream_ptr):
        self._ck(_lib.fsb_set_stream(self.h, _p(cuda_stream_ptr)))

    def set_pool(self, nx, ny, dx, dy):
        self._ck(_lib.fsb_set_pool(self.h, nx, ny, dx, dy))

    def clear_cell_types(self):
        self._ck(_lib.fsb_clear_cell_types(self.h))

    def swap_velocity_buffers(self):
        self._ck(_lib.fsb_swap_velocity_buffers(self.h))

    def synchronize(self):
        self._ck(_lib.fsb_synchronize(self.h))

    # state
    def set_particles(self, a):
        a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 4)
        self._ck(_lib.fsb_set_particles(self.h, _ptr(a), a.shape[0]))

    def set_particles_ptr(self, host_ptr, n):
        """host_ptr: address of n*4 floats (e.g. a pinned torch tensor's data_ptr())."""
        self._ck(_lib.fsb_set_particles(self.h, _p(host_ptr), n))

    def get_particles_ptr(self, host_ptr):
        self._ck(_lib.fsb_get_particles(self.h, _p(host_ptr)))

    def append_particles(self, a):
        a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 4)
        self._ck(_lib.fsb_append_particles(self.h, _ptr(a), a.shape[0]))

    def num_particles(self):
        return _lib.fsb_num_particles(self.h)

    def get_particles(self):
        a = np.empty((self.num_particles(), 4), dtype=np.float32)
        self._ck(_lib.fsb_get_particles(self.h, _ptr(a)))
        return a

    def emit_source(self, x_min, x_max, y_min, y_max, dx=None, dy=None, vx=0.0, vy=0.0):
        n = _l()
        self._ck(_lib.fsb_emit_source(self.h, x_min, x_max, y_min, y_max,
                                      self.dx if dx is None else dx,
                                      self.dy if dy is None else dy, vx, vy, C.byref(n)))
        return n.value

    def set_grid(self, which, a):
        a = np.ascontiguousarray(a, dtype=np.float32).reshape(self.ny, self.nx)
        self._ck(_lib.fsb_set_grid(self.h, which, _ptr(a)))

    def set_grid_ptr(self, which, host_ptr):
        """host_ptr: address of ny*nx dense floats (e.g. a pinned torch tensor's data_ptr())."""
        self._ck(_lib.fsb_set_grid(self.h, which, _p(host_ptr)))

    def get_pressure_ptr(self, host_ptr):
        self._ck(_lib.fsb_get_pressure(self.h, _p(host_ptr)))

    def get_grid(self, which):
        a = np.empty((self.ny, self.nx), dtype=np.float32)
        self._ck(_lib.fsb_get_grid(self.h, which, _ptr(a)))
        return a

    def set_cell_types(self, a):
        a = np.ascontiguousarray(a, dtype=np.uint8).reshape(self.ny, self.nx)
        self._ck(_lib.fsb_set_cell_types(self.h, _ptr(a)))

    def get_cell_types(self):
        a = np.empty((self.ny, self.nx), dtype=np.uint8)
        self._ck(_lib.fsb_get_cell_types(self.h, _ptr(a)))
        return a

    def get_pressure(self):
        a = np.empty((self.ny, self.nx), dtype=np.float32)
        self._ck(_lib.fsb_get_pressure(self.h, _ptr(a)))
        return a

    # stages
    def classify_cells(self):
        self._ck(_lib.fsb_classify_cells(self.h))

    def p2g_spread(self):
        self._ck(_lib.fsb_p2g_spread(self.h))

    def save_previous(self):
        self._ck(_lib.fsb_save_previous(self.h))

    def add_acceleration(self, ax, ay, dt):
        self._ck(_lib.fsb_add_acceleration(self.h, ax, ay, dt))

    def enforce_dirichlet(self):
        self._ck(_lib.fsb_enforce_dirichlet(self.h))

    def extend_velocity(self, n_iter=2):
        self._ck(_lib.fsb_extend_velocity(self.h, n_iter))

    def extend_velocity_avg(self, n_iter=2):
        self._ck(_lib.fsb_extend_velocity_averaging(self.h, n_iter))

    def pressure_solve(self, density=None, dt=0.01):
        self._ck(_lib.fsb_pressure_solve(self.h, self.density if density is None else density, dt))

    def update_diff(self):
        self._ck(_lib.fsb_update_diff(self.h))

    def g2p(self, mode, pic_ratio=None):
        self._ck(_lib.fsb_g2p(self.h, mode, self.pic_ratio if pic_ratio is None else pic_ratio))

    def advect_particles(self, dt, ensure_outside=True):
        self._ck(_lib.fsb_advect_particles(self.h, dt, 1 if ensure_outside else 0))

    def advect_velocity_sl(self, dt):
        self._ck(_lib.fsb_advect_velocity_sl(self.h, dt))

    def advect_particles_grid(self, dt):
        self._ck(_lib.fsb_advect_particles_grid(self.h, dt))

    def step(self, kind, dt):
        self._ck(_lib.fsb_step(self.h, kind, dt))

    # multi-GPU row-slab sharding of the pressure solve
    def shard_export(self):
        blob = np.zeros(SHARD_BLOB_BYTES, dtype=np.uint8)
        self._ck(_lib.fsb_shard_export(self.h, _ptr(blob)))
        return blob

    def shard_connect(self, rank, world, all_blobs):
        a = np.ascontiguousarray(all_blobs, dtype=np.uint8).reshape(world, SHARD_BLOB_BYTES)
        self._ck(_lib.fsb_shard_connect(self.h, rank, world, _ptr(a)))

    def shard_disconnect(self):
        self._ck(_lib.fsb_shard_disconnect(self.h))

    def shard_rows(self):
        lo, hi = _i(), _i()
        self._ck(_lib.fsb_shard_rows(self.h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    # measurement
    def profile_enable(self, on=True):
        self._ck(_lib.fsb_profile_enable(self.h, 1 if on else 0))

    def profile_read(self):
        ms = (C.c_float * len(PROF_NAMES))()
        calls = (C.c_int * len(PROF_NAMES))()
        self._ck(_lib.fsb_profile_read(self.h, ms, calls))
        return {n: (ms[k], calls[k]) for k, n in enumerate(PROF_NAMES)}

    def launch_count(self):
        return _lib.fsb_launch_count(self.h)

    def cg_swept_cells(self):
        """Cells the CG sweeps of the last solve visited per iteration on this rank."""
        return int(_lib.fsb_cg_swept_cells(self.h))

    def cg_launch_mode(self):
        """0 not configured yet, 1 two kernels per iteration (CUDA graph), 2 persistent kernel."""
        return _lib.fsb_cg_launch_mode(self.h)

    # routines of the reference's solver that no step calls
    def add_external_force(self, fx, fy, dt):
        self._ck(_lib.fsb_add_external_force(self.h, fx, fy, dt))

    def p2g_gather(self):
        self._ck(_lib.fsb_p2g_gather(self.h))

    # frames
    def render_rgb(self, width, height, area=(0.0, 1.0, 0.0, 1.0)):
        a = np.zeros((height, width, 3), dtype=np.uint8)
        self._ck(_lib.fsb_render_rgb(self.h, width, height, *[float(v) for v in area], _ptr(a)))
        return a

    def write_ppm(self, path, width, height, area=(0.0, 1.0, 0.0, 1.0)):
        self._ck(_lib.fsb_write_ppm(self.h, os.fsencode(path), width, height,
                                    *[float(v) for v in area]))

    # particle slabs (one rank's view; see sharding.py for the exchanges)
    def slab_configure(self, rank, world):
        self._ck(_lib.fsb_slab_configure(self.h, rank, world))
        lo, hi = _i(), _i()
        self._ck(_lib.fsb_slab_rows(self.h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def slab_add(self, parts, ids):
        parts = np.ascontiguousarray(parts, dtype=np.float32).reshape(-1, 4)
        ids = np.ascontiguousarray(ids, dtype=np.int32).reshape(-1)
        assert parts.shape[0] == ids.shape[0]
        self._ck(_lib.fsb_slab_add(self.h, _ptr(parts), _ptr(ids), parts.shape[0]))

    def slab_sort_out(self, world):
        counts = (_l * world)()
        self._ck(_lib.fsb_slab_sort_out(self.h, counts))
        return [int(v) for v in counts]

    def slab_take(self, dest, n):
        parts, ids = np.empty((n, 4), dtype=np.float32), np.empty(n, dtype=np.int32)
        self._ck(_lib.fsb_slab_take(self.h, dest, _ptr(parts), _ptr(ids)))
        return parts, ids

    def slab_keep_own(self):
        self._ck(_lib.fsb_slab_keep_own(self.h))

    def slab_boundary(self, side):
        n = _l()
        self._ck(_lib.fsb_slab_boundary(self.h, side, C.byref(n)))
        parts, ids = np.empty((n.value, 4), dtype=np.float32), np.empty(n.value, dtype=np.int32)
        self._ck(_lib.fsb_slab_boundary_take(self.h, _ptr(parts), _ptr(ids)))
        return parts, ids

    # the same calls with raw pointers (host or device memory: the library copies with
    # cudaMemcpyDefault and returns after the copy has completed)
    def slab_add_ptr(self, parts_ptr, ids_ptr, n):
        if n:
            self._ck(_lib.fsb_slab_add(self.h, C.c_void_p(parts_ptr), C.c_void_p(ids_ptr), int(n)))

    def slab_take_ptr(self, dest, parts_ptr, ids_ptr):
        self._ck(_lib.fsb_slab_take(self.h, dest, C.c_void_p(parts_ptr), C.c_void_p(ids_ptr)))

    def slab_boundary_count(self, side):
        n = _l()
        self._ck(_lib.fsb_slab_boundary(self.h, side, C.byref(n)))
        return int(n.value)

    def slab_boundary_take_ptr(self, parts_ptr, ids_ptr):
        self._ck(_lib.fsb_slab_boundary_take(self.h, C.c_void_p(parts_ptr), C.c_void_p(ids_ptr)))

    def slab_get_ptr(self, parts_ptr, ids_ptr):
        self._ck(_lib.fsb_slab_get(self.h, C.c_void_p(parts_ptr), C.c_void_p(ids_ptr)))

    def get_rows_ptr(self, which, lo, hi, ptr):
        self._ck(_lib.fsb_get_rows(self.h, which, lo, hi, C.c_void_p(ptr)))

    def set_rows_ptr(self, which, lo, hi, ptr):
        self._ck(_lib.fsb_set_rows(self.h, which, lo, hi, C.c_void_p(ptr)))

    def slab_get(self):
        n = self.num_particles()
        parts, ids = np.empty((n, 4), dtype=np.float32), np.empty(n, dtype=np.int32)
        self._ck(_lib.fsb_slab_get(self.h, _ptr(parts), _ptr(ids)))
        return parts, ids

    def get_rows(self, which, lo, hi):
        a = np.empty((hi - lo, self.nx), dtype=np.uint8 if which == ROWS_LABELS else np.float32)
        self._ck(_lib.fsb_get_rows(self.h, which, lo, hi, _ptr(a)))
        return a

    def set_rows(self, which, lo, hi, a):
        a = np.ascontiguousarray(a, dtype=np.uint8 if which == ROWS_LABELS else np.float32)
        assert a.shape == (hi - lo, self.nx)
        self._ck(_lib.fsb_set_rows(self.h, which, lo, hi, _ptr(a)))

    def slab_step_a(self, kind):
        self._ck(_lib.fsb_slab_step_a(self.h, kind))

    def slab_step_b(self, kind, dt):
        self._ck(_lib.fsb_slab_step_b(self.h, kind, dt))

    def slab_step_c(self, kind, dt):
        self._ck(_lib.fsb_slab_step_c(self.h, kind, dt))

    # state files
    def save_state(self, path):
        self._ck(_lib.fsb_save_state(self.h, os.fsencode(path)))

    def load_state(self, path):
        self._ck(_lib.fsb_load_state(self.h, os.fsencode(path)))

    def timer_start(self):
        self._ck(_lib.fsb_timer_start(self.h))

    def timer_stop(self):
        ms = _f()
        self._ck(_lib.fsb_timer_stop(self.h, C.byref(ms)))
        return ms.value
