"""The numbers baked into the kernels (csrc/mp_constants.cuh, read back through the C ABI's mp_constants()) against the
host-side mirror of the reference's config (mobileposer_b200/config.py, itself re-checked against the reference's config and
SMPL pickle by oracle/make_golden.py).  Runs without a GPU: mp_constants is a host call."""
import ctypes as C

import numpy as np

from mobileposer_b200 import _cabi, config


def _constants():
    c = _cabi.Constants()
    _cabi.check(_cabi.lib().mp_constants(C.byref(c)), 'mp_constants')
    return c


def test_kernel_constants_equal_config_bit_for_bit():
    c = _constants()
    assert list(c.parent) == config.SMPL_PARENT
    assert list(c.reduced) == config.joint_set.reduced and list(c.ignored) == config.joint_set.ignored
    slot = [-1] * 24
    for i, j in enumerate(config.joint_set.reduced):
        slot[j] = i
    assert list(c.reduced_slot) == slot
    j0 = np.ctypeslib.as_array(c.j_zero).reshape(24, 3)
    want = np.asarray(config.SMPL_J_ZERO, dtype=np.float32)
    assert j0.dtype == np.float32 and np.array_equal(j0.view(np.uint32), want.view(np.uint32))
    feet = np.asarray(list(c.feet), dtype=np.float32).reshape(2, 3)
    assert np.array_equal(feet, want[10:12])
    assert c.floor_y == config.FLOOR_Y == float(min(want[10, 1], want[11, 1]))
    assert np.float32(c.gravity_velocity) == np.float32(config.joint_set.gravity_velocity)
    assert c.vel_div == config.datasets.fps / config.amass.vel_scale
    assert (np.float32(c.prob_lo), np.float32(c.prob_hi)) == tuple(np.float32(v) for v in config.PROB_THRESHOLD)


def test_abi_version_is_checked_on_load():
    assert _cabi.lib().mp_abi_version() == _cabi.ABI_VERSION
