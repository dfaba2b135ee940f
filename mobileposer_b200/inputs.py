"""Input assembly on the device: raw per-slot IMU streams -> the [*, T, 60] frames the heads consume.

Mirrors `PoseDataset._process_file_data/_process_combo_data` (mobileposer/data.py:60-61,69-76) and
`DataLoader._get_imu` (mobileposer/loader.py:39-49) through `mp_imu_assemble`; all requested device combos come out of
one launch (the reference loops over its 12 combos in Python, data.py:70)."""
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
