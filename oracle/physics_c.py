"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/physics_port.c (the compiled float64 statement of K8; PARITY
UNPINNED, see physics_port.py).  Used by tests and by bench.py's CPU arm."""
from __future__ import annotations

import ctypes as C

import numpy as np

from mobileposer_b200.config import FLOOR_Y, SMPL_J_ZERO, amass, datasets

from . import build as _build

STATE_DOUBLES = 76


class _Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in ('w_vel', 'w_contact', 'damping', 'damping_abs', 'fps', 'vel_scale', 'floor_y')]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.mp_oracle_physics_optimize.restype = C.c_int
        _lib.mp_oracle_physics_optimize.argtypes = [C.c_void_p] * 5 + [C.c_int32, C.c_int32, C.POINTER(_Params)] + [C.c_void_p] * 3
        _lib.mp_oracle_set_threads.restype = None
        _lib.mp_oracle_set_threads.argtypes = [C.c_int32]
    return _lib


def set_threads(n: int):
    """OpenMP threads of the C port (torchrun sets OMP_NUM_THREADS=1 for its children)."""
    lib().mp_oracle_set_threads(int(n))


class PhysicsOptimizerC:
    def __init__(self, B=1, w_vel=1.0, w_contact=10.0, damping=1.0, damping_abs=1e-2, fps=datasets.fps, floor_y=FLOOR_Y):
        self.B = B
        self.prm = _Params(w_vel, w_contact, damping, damping_abs, fps, amass.vel_scale, floor_y)
        self.j_zero = np.ascontiguousarray(np.asarray(SMPL_J_ZERO, np.float64))
        self.reset_states()

    def reset_states(self):
        self.state = np.zeros((self.B, STATE_DOUBLES), np.float64)

    def optimize_sequences(self, pose, vel72, contact, lengths=None, vel_scale=amass.vel_scale):
        """Same contract as PhysicsOptimizerPort.optimize_sequences; float32 in / out."""
        pose = np.ascontiguousarray(pose, np.float32)
        B, T = pose.shape[:2]
        assert B == self.B
        vel72 = np.ascontiguousarray(vel72, np.float32).reshape(B, T, 72)
        contact = np.ascontiguousarray(contact, np.float32).reshape(B, T, 2)
        lens = None if lengths is None else np.ascontiguousarray(lengths, np.int32)
        out = np.empty((B, T, 24, 3, 3), np.float32)
        tran = np.empty((B, T, 3), np.float32)
        self.prm.vel_scale = vel_scale
        st = lib().mp_oracle_physics_optimize(pose.ctypes.data, vel72.ctypes.data, contact.ctypes.data,
                                              None if lens is None else lens.ctypes.data, self.state.ctypes.data, B, T,
                                              C.byref(self.prm), self.j_zero.ctypes.data, out.ctypes.data, tran.ctypes.data)
        assert st == 0
        return out, tran
