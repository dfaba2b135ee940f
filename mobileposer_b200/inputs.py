"""Input assembly on the device: raw per-slot IMU streams -> the [*, T, 60] frames the heads consume.

Mirrors `PoseDataset._process_file_data/_process_combo_data` (mobileposer/data.py:60-61,69-76) and
`DataLoader._get_imu` (mobileposer/loader.py:39-49) through `mp_imu_assemble`; all requested device combos come out of
one launch (the reference loops over its 12 combos in Python, data.py:70).  `LiveCalibration` / `normalize_live` are the
live demo's calibration and per-tick normalisation (mobileposer/live_demo.py:160-177, 210-234) through `mp_imu_live_normalize`."""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi
from .config import amass
from .modules import _f32c, _require_cuda, current_stream_ptr


def combo_mask(combo) -> int:
    """'lw_rp' or [0, 3] -> bit mask over the 5 IMU slots (config.py:60-73)."""
    slots = amass.combos[combo] if isinstance(combo, str) else combo
    return sum(1 << int(s) for s in slots)


@torch.no_grad()
def assemble_imu(acc, ori, combos=None, smooth: bool = False, acc_scale: float = amass.acc_scale):
    """acc [T, S, 3], ori [T, S, 3, 3] (S >= 5 slots, CUDA) -> imu [n_combos, T, 60].
    combos: names / slot lists (default: all 12 of amass.combos, in the reference's order); smooth=True is the viewer's
    `_get_imu` (3-tap moving average of the scaled accelerations), False the dataset's."""
    _require_cuda(acc, 'acc')
    acc, ori = _f32c(acc), _f32c(ori)
    if acc.dim() != 3 or acc.shape[2] != 3 or ori.shape[:2] != acc.shape[:2] or tuple(ori.shape[2:]) != (3, 3):
        raise ValueError(f'expected acc [T, S, 3] and ori [T, S, 3, 3], got {tuple(acc.shape)} and {tuple(ori.shape)}')
    combos = list(amass.combos.keys()) if combos is None else list(combos)
    masks = (C.c_int32 * len(combos))(*[combo_mask(c) for c in combos])
    T, S = acc.shape[0], acc.shape[1]
    out = torch.empty(len(combos), T, 60, device=acc.device, dtype=torch.float32)
    with torch.cuda.device(acc.device):
        _cabi.check(_cabi.lib().mp_imu_assemble(acc.data_ptr(), ori.data_ptr(), T, S, masks, len(combos), float(acc_scale),
                                                int(bool(smooth)), out.data_ptr(), current_stream_ptr(acc.device)), 'mp_imu_assemble')
    return out


LIVE_SLOT_ORDER = (1, 4, 3, 0, 2)      # live_demo.py:216-217: sensor index feeding each of the 5 model slots


class LiveCalibration:
    """The live demo's calibration (live_demo.py:160-177): `smpl2imu` from the reading of sensor 0 aligned with the body
    frame, `device2bone` and `acc_offsets` from a T-pose reading of all sensors.  A cold path (once per session): plain
    torch on the host, 16 small matrices."""

    def __init__(self, smpl2imu, device2bone, acc_offsets):
        self.smpl2imu = torch.as_tensor(smpl2imu, dtype=torch.float32).reshape(3, 3).cpu().contiguous()
        self.device2bone = torch.as_tensor(device2bone, dtype=torch.float32).reshape(5, 3, 3).cpu().contiguous()
        self.acc_offsets = torch.as_tensor(acc_offsets, dtype=torch.float32).reshape(5, 3).cpu().contiguous()

    @staticmethod
    def quaternion_to_rotation_matrix(q):
        """articulate/math/angular.py:224-236."""
        q = torch.as_tensor(q, dtype=torch.float32).reshape(-1, 4)
        q = q / q.norm(dim=1, keepdim=True)
        a, b, c, d = q[:, 0:1], q[:, 1:2], q[:, 2:3], q[:, 3:4]
        r = torch.cat((-2 * c * c - 2 * d * d + 1, 2 * b * c - 2 * a * d, 2 * a * c + 2 * b * d,
                       2 * b * c + 2 * a * d, -2 * b * b - 2 * d * d + 1, 2 * c * d - 2 * a * b,
                       2 * b * d - 2 * a * c, 2 * a * b + 2 * c * d, -2 * b * b - 2 * c * c + 1), dim=1)
        return r.view(-1, 3, 3)

    @classmethod
    def from_readings(cls, align_quat, tpose_quats, tpose_accs):
        """align_quat [4]: mean reading of sensor 0 held in the body frame; tpose_quats [5, 4], tpose_accs [5, 3]: mean
        readings in T-pose (live_demo.py:163-177)."""
        s2i = cls.quaternion_to_rotation_matrix(align_quat).view(3, 3).t()
        oris = cls.quaternion_to_rotation_matrix(tpose_quats)
        d2b = s2i.matmul(oris).transpose(1, 2).matmul(torch.eye(3))
        off = s2i.matmul(torch.as_tensor(tpose_accs, dtype=torch.float32).reshape(5, 3, 1)).squeeze(-1)
        return cls(s2i, d2b, off)


@torch.no_grad()
def normalize_live(quat, acc_raw, cal: LiveCalibration, combo='lw_rp', phone_as_watch: bool = False,
                   acc_scale: float = amass.acc_scale):
    """Sensor readings of one tick or a buffer of ticks -- quat [n, 5, 4] (wxyz), acc_raw [n, 5, 3], CUDA -> imu [n, 60]
    (live_demo.py:210-234), ready for `MobilePoserNet.forward_online`."""
    _require_cuda(quat, 'quat')
    quat, acc_raw = _f32c(quat).view(-1, 5, 4), _f32c(acc_raw).view(-1, 5, 3)
    n = quat.shape[0]
    if acc_raw.shape[0] != n:
        raise ValueError(f'{n} quaternion ticks, {acc_raw.shape[0]} acceleration ticks')
    out = torch.empty(n, 60, device=quat.device, dtype=torch.float32)
    f = C.c_float
    s2i = (f * 9)(*cal.smpl2imu.flatten().tolist())
    d2b = (f * 45)(*cal.device2bone.flatten().tolist())
    off = (f * 15)(*cal.acc_offsets.flatten().tolist())
    perm = (C.c_int32 * 5)(*LIVE_SLOT_ORDER)
    with torch.cuda.device(quat.device):
        _cabi.check(_cabi.lib().mp_imu_live_normalize(quat.data_ptr(), acc_raw.data_ptr(), n, s2i, d2b, off, perm, combo_mask(combo),
                                                      int(bool(phone_as_watch)), float(acc_scale), out.data_ptr(),
                                                      current_stream_ptr(quat.device)), 'mp_imu_live_normalize')
    return out
