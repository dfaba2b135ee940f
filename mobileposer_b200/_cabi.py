"""ctypes binding of include/mobileposer_b200.h (the only way the package reaches the GPU kernels).

There is deliberately no fallback: if the shared library is missing, or the device is not an
sm_100 part, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH

_lib = None

c_float_p = C.c_void_p     # device / host pointers travel as integers (tensor.data_ptr())
c_int_p = C.c_void_p
c_stream = C.c_void_p


class RnnWeights(C.Structure):
    """struct mp_rnn_weights (include/mobileposer_b200.h)."""
    _fields_ = [
        ('n_input', C.c_int32), ('n_output', C.c_int32), ('n_hidden', C.c_int32),
        ('n_layers', C.c_int32), ('bidirectional', C.c_int32),
        ('linear1_w', C.c_void_p), ('linear1_b', C.c_void_p),
        ('linear2_w', C.c_void_p), ('linear2_b', C.c_void_p),
        ('w_ih', (C.c_void_p * 2) * 2), ('w_hh', (C.c_void_p * 2) * 2),
        ('b_ih', (C.c_void_p * 2) * 2), ('b_hh', (C.c_void_p * 2) * 2),
    ]


class RnnGrads(C.Structure):
    """struct mp_rnn_grads."""
    _fields_ = [('linear1_w', C.c_void_p), ('linear1_b', C.c_void_p), ('linear2_w', C.c_void_p), ('linear2_b', C.c_void_p),
                ('w_ih', (C.c_void_p * 2) * 2), ('w_hh', (C.c_void_p * 2) * 2), ('b_ih', (C.c_void_p * 2) * 2), ('b_hh', (C.c_void_p * 2) * 2)]


class ProfileEntry(C.Structure):
    """struct mp_profile_entry."""
    _fields_ = [('name', C.c_char * 32), ('launches', C.c_int64), ('total_ms', C.c_double),
                ('algorithmic_bytes', C.c_double)]


ONLINE_STATE_BYTES = 64   # sizeof(mp_online_state_t)


class PhysicsParams(C.Structure):
    """struct mp_physics_params."""
    _fields_ = [('w_vel', C.c_float), ('w_contact', C.c_float), ('damping', C.c_float), ('damping_abs', C.c_float), ('fps', C.c_float),
                ('vel_scale', C.c_float), ('floor_y', C.c_float)]


PHYSICS_STATE_FLOATS = 80   # MP_PHYSICS_STATE_FLOATS
ABI_VERSION = 1             # MP_ABI_VERSION this binding was written against


class Constants(C.Structure):
    """struct mp_constants."""
    _fields_ = [('parent', C.c_int32 * 24), ('reduced', C.c_int32 * 16), ('ignored', C.c_int32 * 9), ('reduced_slot', C.c_int32 * 24),
                ('j_zero', (C.c_float * 3) * 24), ('feet', C.c_float * 6), ('gravity_velocity', C.c_float), ('vel_div', C.c_float),
                ('prob_lo', C.c_float), ('prob_hi', C.c_float), ('floor_y', C.c_double)]


# name -> (restype, argtypes); must list every symbol the header declares (tests/test_cabi.py)
SIGNATURES = {
    'mp_abi_version': (C.c_int, []),
    'mp_last_error': (C.c_char_p, []),
    'mp_constants': (C.c_int, [C.POINTER(Constants)]),
    'mp_device_check': (C.c_int, []),
    'mp_launch_count': (C.c_int64, []),
    'mp_profile_enable': (C.c_int, [C.c_int32]),
    'mp_profile_collect': (C.c_int, [C.POINTER(ProfileEntry), C.c_int32, C.POINTER(C.c_int32)]),
    'mp_rnn_create': (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(RnnWeights), c_stream]),
    'mp_rnn_destroy': (None, [C.c_void_p]),
    'mp_rnn_workspace_bytes': (C.c_size_t, [C.c_void_p, C.c_int32, C.c_int32]),
    'mp_rnn_forward': (C.c_int, [C.c_void_p, c_float_p, C.c_int32, c_float_p, C.c_int32, C.c_int32, C.c_int32,
                                 c_int_p, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p,
                                 C.c_void_p, C.c_size_t, c_stream]),
    'mp_rnn_train_workspace_bytes': (C.c_size_t, [C.POINTER(RnnWeights), C.c_int32, C.c_int32]),
    'mp_rnn_train_forward': (C.c_int, [C.POINTER(RnnWeights), c_float_p, C.c_int32, C.c_int32, c_int_p, c_float_p, c_float_p, C.c_void_p,
                                       C.c_size_t, c_stream]),
    'mp_rnn_train_backward': (C.c_int, [C.POINTER(RnnWeights), c_float_p, C.c_int32, C.c_int32, c_int_p, c_float_p, c_float_p,
                                        C.POINTER(RnnGrads), C.c_void_p, C.c_size_t, c_stream]),
    'mp_joints_loss': (C.c_int, [c_float_p, c_float_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p, c_float_p, c_stream]),
    'mp_poser_loss': (C.c_int, [c_float_p, c_float_p, c_float_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p, c_float_p, c_stream]),
    'mp_footcontact_loss': (C.c_int, [c_float_p, c_float_p, C.c_int32, C.c_int32, C.c_void_p, c_float_p, c_stream]),
    'mp_velocity_loss': (C.c_int, [c_float_p, c_float_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, c_float_p, c_stream]),
    'mp_grad_sq_norm': (C.c_int, [c_float_p, C.c_size_t, C.c_void_p, c_stream]),
    'mp_adamw_step': (C.c_int, [c_float_p, c_float_p, c_float_p, c_float_p, C.c_size_t, C.c_float, C.c_float, C.c_float, C.c_float,
                                C.c_float, C.c_int32, C.c_void_p, C.c_float, C.c_float, c_stream]),
    'mp_gemm_bias': (C.c_int, [c_float_p, c_float_p, c_float_p, c_float_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                               C.c_int32, c_stream]),
    'mp_pose_reduced_global_to_full': (C.c_int, [c_float_p, C.c_int64, c_float_p, c_stream]),
    'mp_tran_offline': (C.c_int, [c_float_p, c_float_p, c_float_p, c_int_p, C.c_int32, C.c_int32, c_float_p, c_stream]),
    'mp_imu_assemble': (C.c_int, [c_float_p, c_float_p, C.c_int64, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_float, C.c_int32,
                                  c_float_p, c_stream]),
    'mp_imu_live_normalize': (C.c_int, [c_float_p, c_float_p, C.c_int64, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                        C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_float, c_float_p, c_stream]),
    'mp_physics_optimize': (C.c_int, [c_float_p, c_float_p, c_float_p, c_int_p, c_float_p, C.c_int32, C.c_int32,
                                      C.POINTER(PhysicsParams), c_float_p, c_float_p, c_stream]),
    'mp_physics_optimize_debug': (C.c_int, [c_float_p, c_float_p, c_float_p, c_int_p, c_float_p, C.c_int32, C.c_int32,
                                            C.POINTER(PhysicsParams), c_float_p, c_float_p, c_float_p, C.c_int32, c_stream]),
    'mp_eval_frame_errors': (C.c_int, [c_float_p, c_float_p, c_float_p, c_float_p, C.c_int64, c_float_p, c_float_p, c_float_p,
                                       c_float_p, c_float_p, c_stream]),
    'mp_eval_vertex_errors': (C.c_int, [c_float_p, c_float_p, C.c_int64, c_float_p, c_float_p, C.c_int32, C.c_void_p, C.c_void_p, c_stream]),
    'mp_eval_motion_rows': (C.c_int, [c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, C.c_int64, C.c_int32, C.c_uint32, c_float_p, c_stream]),
    'mp_eval_motion_rows_batch': (C.c_int, [c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, C.c_void_p, C.c_int32, C.c_int32, C.c_uint32,
                                            c_float_p, c_stream]),
    'mp_eval_tran_windows': (C.c_int, [c_float_p, c_float_p, c_int_p, C.c_int32, C.c_int32, c_float_p, c_int_p, c_stream]),
    'mp_physics_fk': (C.c_int, [c_float_p, C.c_int64, c_float_p, c_float_p, c_stream]),
    'mp_online_update': (C.c_int, [C.c_void_p, c_float_p, c_float_p, c_float_p, c_float_p, C.c_int32, C.c_int32,
                                   C.c_int32, c_float_p, c_float_p, c_float_p, c_stream]),
    'mp_online_push_frame': (C.c_int, [c_float_p, c_float_p, c_float_p, C.c_int32, C.c_int32, C.c_int32, c_stream]),
    'mp_online_reset': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_stream]),
    'mp_net_create': (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'mp_net_destroy': (None, [C.c_void_p]),
    'mp_net_workspace_bytes': (C.c_size_t, [C.c_void_p, C.c_int32, C.c_int32]),
    'mp_net_set_graph': (C.c_int, [C.c_void_p, C.c_int32]),
    'mp_net_set_rec_tile': (C.c_int, [C.c_void_p, C.c_int32]),
    'mp_net_set_physics': (C.c_int, [C.c_void_p, C.POINTER(PhysicsParams)]),
    'mp_net_forward': (C.c_int, [C.c_void_p, c_float_p, C.c_int32, C.c_int32, c_int_p, c_float_p, c_float_p,
                                 c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p,
                                 C.c_void_p, C.c_size_t, c_stream]),
    'mp_net_host_staging_bytes': (C.c_size_t, [C.c_int32, C.c_int32]),
    'mp_net_enqueue_offline_host': (C.c_int, [C.c_void_p, c_float_p, C.c_int32, C.c_int32, c_int_p, c_float_p,
                                              c_float_p, c_float_p, c_float_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                              c_stream]),
    'mp_net_enqueue_offline_host_compact': (C.c_int, [C.c_void_p, c_float_p, C.c_int32, C.c_int32, c_int_p, c_float_p,
                                                      c_float_p, c_float_p, c_float_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                                      c_stream]),
    'mp_pose_full_to_local6d': (C.c_int, [c_float_p, C.c_int64, c_float_p, c_stream]),
    'mp_net_forward_offline_host': (C.c_int, [C.c_void_p, c_float_p, C.c_int32, C.c_int32, c_int_p, c_float_p,
                                              c_float_p, c_float_p, c_float_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                              c_stream]),
}


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: build it with `python -m mobileposer_b200.build` '
                '(mobileposer_b200 has no CPU or eager fallback).')
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.mp_abi_version() != ABI_VERSION:
            raise RuntimeError(f'{LIB_PATH} has ABI version {handle.mp_abi_version()}, this package binds version {ABI_VERSION}: '
                               'rebuild it with `python -m mobileposer_b200.build --force`')
        _lib = handle
    return _lib


def check(status: int, what: str = '') -> None:
    if status != 0:
        msg = lib().mp_last_error().decode('utf-8', 'replace')
        raise RuntimeError(f'mobileposer_b200 {what} failed (status {status}): {msg}')


def profile_collect():
    """-> {kernel name: dict(launches, total_ms, algorithmic_bytes)} since mp_profile_enable(1)."""
    arr = (ProfileEntry * 32)()
    n = C.c_int32(0)
    check(lib().mp_profile_collect(arr, 32, C.byref(n)), 'mp_profile_collect')
    return {arr[i].name.decode(): dict(launches=int(arr[i].launches), total_ms=float(arr[i].total_ms),
                                       algorithmic_bytes=float(arr[i].algorithmic_bytes)) for i in range(n.value)}
