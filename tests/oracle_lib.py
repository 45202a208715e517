"""ctypes bindings for the CPU checkers (TEST INFRASTRUCTURE ONLY).

`OracleLib("fso")` loads oracle/libfsoracle.so (the plain-C restatement),
`OracleLib("fsr")` loads oracle/_ref/libfsref.so (the reference's own sources,
compiled unchanged against oracle/eigen_shim).  Both export the API declared in
oracle/fluid_oracle_api.h.  The product package never imports this module.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATHS = {
    "fso": os.path.join(ROOT, "oracle", "libfsoracle.so"),
    "fsr": os.path.join(ROOT, "oracle", "_ref", "libfsref.so"),
}

U_FRONT, V_FRONT, U_BACK, V_BACK, U_PREV, V_PREV, U_DIFF, V_DIFF = range(8)
G2P_PIC, G2P_FLIP, G2P_PICFLIP = range(3)
STEP_SL, STEP_PIC, STEP_FLIP, STEP_PICFLIP = range(4)
LIQUID, AIR, SOLID = 0, 1, 2

_f = C.c_float
_i = C.c_int
_p = C.c_void_p
_l = C.c_int64

_SIGS = {
    "create": (_p, [_i, _i, _f, _f, _f, _f]),
    "destroy": (None, [_p]),
    "delta_x": (_f, [_p]),
    "delta_y": (_f, [_p]),
    "set_cg": (None, [_p, _i, _f]),
    "set_particles": (None, [_p, _p, _l]),
    "append_particles": (None, [_p, _p, _l]),
    "num_particles": (_l, [_p]),
    "get_particles": (None, [_p, _p]),
    "emit_source": (_l, [_p, _f, _f, _f, _f, _f, _f, _f, _f]),
    "set_grid": (None, [_p, _i, _p]),
    "get_grid": (None, [_p, _i, _p]),
    "set_cell_types": (None, [_p, _p]),
    "get_cell_types": (None, [_p, _p]),
    "classify_cells": (None, [_p]),
    "p2g_spread": (None, [_p]),
    "save_previous": (None, [_p]),
    "add_acceleration": (None, [_p, _f, _f, _f]),
    "enforce_dirichlet": (None, [_p]),
    "extend_velocity": (None, [_p, _i]),
    "pressure_solve": (None, [_p, _f, _f]),
    "get_pressure": (None, [_p, _p]),
    "cg_iterations": (_i, [_p]),
    "cg_error": (_f, [_p]),
    "update_diff": (None, [_p]),
    "g2p": (None, [_p, _i, _f]),
    "advect_particles": (None, [_p, _f, _i]),
    "advect_velocity_sl": (None, [_p, _f]),
    "advect_particles_grid": (None, [_p, _f]),
    "step": (_i, [_p, _i, _f]),
    "add_external_force": (None, [_p, _f, _f, _f]),
    "p2g_gather": (None, [_p]),
    "extend_velocity_avg": (None, [_p, _i]),
    "render_rgb": (None, [_p, _i, _i, _f, _f, _f, _f, _p]),
}


def available(prefix):
    return os.path.exists(PATHS[prefix])


class OracleLib:
    def __init__(self, prefix):
        self.prefix = prefix
        self.lib = C.CDLL(PATHS[prefix])
        for name, (res, args) in _SIGS.items():
            fn = getattr(self.lib, f"{prefix}_{name}")
            fn.restype = res
            fn.argtypes = args
            setattr(self, "_" + name, fn)

    def sim(self, nx, ny, lx=1.0, ly=1.0, density=0.01, pic_ratio=0.05):
        return OracleSim(self, nx, ny, lx, ly, density, pic_ratio)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleSim:
    """Same method names as fluid_simulation_b200.capi.Sim so tests can drive
    the CPU checker and the CUDA path with the same code."""

    def __init__(self, lib, nx, ny, lx, ly, density, pic_ratio):
        self.L = lib
        self.nx, self.ny = nx, ny
        self.density, self.pic_ratio = density, pic_ratio
        self.h = lib._create(nx, ny, lx, ly, density, pic_ratio)
        self.dx = lib._delta_x(self.h)
        self.dy = lib._delta_y(self.h)

    def close(self):
        if self.h:
            self.L._destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_cg(self, max_iters, tol):
        self.L._set_cg(self.h, max_iters, tol)

    def set_particles(self, a):
        a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 4)
        self.L._set_particles(self.h, _ptr(a), a.shape[0])

    def append_particles(self, a):
        a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 4)
        self.L._append_particles(self.h, _ptr(a), a.shape[0])

    def num_particles(self):
        return self.L._num_particles(self.h)

    def get_particles(self):
        a = np.empty((self.num_particles(), 4), dtype=np.float32)
        self.L._get_particles(self.h, _ptr(a))
        return a

    def emit_source(self, x_min, x_max, y_min, y_max, dx=None, dy=None, vx=0.0, vy=0.0):
        return self.L._emit_source(self.h, x_min, x_max, y_min, y_max,
                                   self.dx if dx is None else dx,
                                   self.dy if dy is None else dy, vx, vy)

    def set_grid(self, which, a):
        a = np.ascontiguousarray(a, dtype=np.float32).reshape(self.ny, self.nx)
        self.L._set_grid(self.h, which, _ptr(a))

    def get_grid(self, which):
        a = np.empty((self.ny, self.nx), dtype=np.float32)
        self.L._get_grid(self.h, which, _ptr(a))
        return a

    def set_cell_types(self, a):
        a = np.ascontiguousarray(a, dtype=np.uint8).reshape(self.ny, self.nx)
        self.L._set_cell_types(self.h, _ptr(a))

    def get_cell_types(self):
        a = np.empty((self.ny, self.nx), dtype=np.uint8)
        self.L._get_cell_types(self.h, _ptr(a))
        return a

    def classify_cells(self):
        self.L._classify_cells(self.h)

    def p2g_spread(self):
        self.L._p2g_spread(self.h)

    def save_previous(self):
        self.L._save_previous(self.h)

    def add_acceleration(self, ax, ay, dt):
        self.L._add_acceleration(self.h, ax, ay, dt)

    def enforce_dirichlet(self):
        self.L._enforce_dirichlet(self.h)

    def extend_velocity(self, n_iter=2):
        self.L._extend_velocity(self.h, n_iter)

    def pressure_solve(self, density=None, dt=0.01):
        self.L._pressure_solve(self.h, self.density if density is None else density, dt)

    def get_pressure(self):
        a = np.empty((self.ny, self.nx), dtype=np.float32)
        self.L._get_pressure(self.h, _ptr(a))
        return a

    def cg_info(self):
        return self.L._cg_iterations(self.h), self.L._cg_error(self.h)

    def update_diff(self):
        self.L._update_diff(self.h)

    def g2p(self, mode, pic_ratio=None):
        self.L._g2p(self.h, mode, self.pic_ratio if pic_ratio is None else pic_ratio)

    def advect_particles(self, dt, ensure_outside=True):
        self.L._advect_particles(self.h, dt, 1 if ensure_outside else 0)

    def advect_velocity_sl(self, dt):
        self.L._advect_velocity_sl(self.h, dt)

    def advect_particles_grid(self, dt):
        self.L._advect_particles_grid(self.h, dt)

    def add_external_force(self, fx, fy, dt):
        self.L._add_external_force(self.h, fx, fy, dt)

    def p2g_gather(self):
        self.L._p2g_gather(self.h)

    def extend_velocity_avg(self, n_iter=2):
        self.L._extend_velocity_avg(self.h, n_iter)

    def render_rgb(self, width, height, area=(0.0, 1.0, 0.0, 1.0)):
        a = np.zeros((height, width, 3), dtype=np.uint8)
        self.L._render_rgb(self.h, width, height, *[float(v) for v in area], _ptr(a))
        return a

    def step(self, kind, dt):
        rc = self.L._step(self.h, kind, dt)
        if rc != 0:
            raise RuntimeError("Memory pool and fluid domain does not match.")
