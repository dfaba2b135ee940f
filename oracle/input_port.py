"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's input assembly.

Pinned to the live reference through tests/golden/input_assembly.npz (oracle/make_golden_inputs.py).

  assemble_dataset  mobileposer/data.py:60-61 (`acc[:, :5] / amass.acc_scale, ori[:, :5]`) and :69-76 (per combo: zero the
                    slots that are not worn, `cat(acc.flatten(1), ori.flatten(1))` -> [T, 60])
  assemble_loader   mobileposer/loader.py:39-49 (same masking, then `smooth_avg` over the scaled accelerations)
  smooth_avg        mobileposer/utils/model_utils.py:28-37 (3-tap moving average, `nanmean` over the taps that exist)
"""
from __future__ import annotations

import numpy as np

from mobileposer_b200.config import amass


def smooth_avg(acc, s=3):
    acc = np.asarray(acc, np.float32)
    T = acc.shape[0]
    out = np.empty_like(acc)
    h = s // 2
    for t in range(T):
        taps = [acc[u] for u in range(t - h, t + h + 1) if 0 <= u < T]
        tot = np.zeros_like(acc[0])
        for x in taps:                      # nansum adds the taps in order, then one division by their count
            tot = (tot + x).astype(np.float32)
        out[t] = tot / np.float32(len(taps))
    return out


def _masked(raw_acc, raw_ori, combo, acc_scale):
    acc = np.zeros((raw_acc.shape[0], 5, 3), np.float32)
    ori = np.zeros((raw_ori.shape[0], 5, 3, 3), np.float32)
    acc[:, combo] = np.asarray(raw_acc, np.float32)[:, combo] / np.float32(acc_scale)
    ori[:, combo] = np.asarray(raw_ori, np.float32)[:, combo]
    return acc, ori


def assemble_dataset(raw_acc, raw_ori, combos=None, acc_scale=amass.acc_scale):
    """raw_acc [T, >=5, 3], raw_ori [T, >=5, 3, 3] -> [n_combos, T, 60] (combos: list of slot lists; default all 12)."""
    combos = list(amass.combos.values()) if combos is None else combos
    out = []
    for c in combos:
        acc, ori = _masked(raw_acc, raw_ori, c, acc_scale)
        out.append(np.concatenate([acc.reshape(len(acc), -1), ori.reshape(len(ori), -1)], axis=1))
    return np.stack(out)


def assemble_loader(raw_acc, raw_ori, combo, acc_scale=amass.acc_scale):
    """loader.py:39-49 -> [T, 60] with the accelerations smoothed."""
    acc, ori = _masked(raw_acc, raw_ori, combo, acc_scale)
    acc = smooth_avg(acc)
    return np.concatenate([acc.reshape(len(acc), -1), ori.reshape(len(ori), -1)], axis=1)


def quaternion_to_rotation_matrix(q):
    """articulate/math/angular.py:224-236: (unnormalised) wxyz quaternions -> rotation matrices, float32 like the reference."""
    q = np.asarray(q, np.float32).reshape(-1, 4)
    q = q / np.sqrt((q * q).sum(axis=1, keepdims=True, dtype=np.float32))
    a, b, c, d = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    r = np.stack([-2 * c * c - 2 * d * d + 1, 2 * b * c - 2 * a * d, 2 * a * c + 2 * b * d,
                  2 * b * c + 2 * a * d, -2 * b * b - 2 * d * d + 1, 2 * c * d - 2 * a * b,
                  2 * b * d - 2 * a * c, 2 * a * b + 2 * c * d, -2 * b * b - 2 * c * c + 1], axis=1)
    return r.reshape(-1, 3, 3).astype(np.float32)


def live_normalize(ori_q, acc_raw, smpl2imu, device2bone, acc_offsets, combo=None, phone_as_watch=False,
                   perm=(1, 4, 3, 0, 2), acc_scale=amass.acc_scale):
    """live_demo.py:210-234: sensor quaternions [n,5,4] + accelerations [n,5,3] of one tick -> the [n,60] frame.
    Calibration (live_demo.py:160-177): smpl2imu [3,3], device2bone [5,3,3], acc_offsets [5,3]."""
    ori_q, acc_raw = np.asarray(ori_q, np.float32), np.asarray(acc_raw, np.float32)
    n = ori_q.shape[0]
    s2i, d2b, off = (np.asarray(x, np.float32) for x in (smpl2imu, device2bone, acc_offsets))
    ori_raw = quaternion_to_rotation_matrix(ori_q).reshape(n, 5, 3, 3)
    glb_acc = np.einsum('ab,nsb->nsa', s2i, acc_raw) - off[None]
    glb_ori = np.einsum('ab,nsbc,scd->nsad', s2i, ori_raw, d2b)
    _acc = glb_acc[:, list(perm)] / np.float32(acc_scale)
    _ori = glb_ori[:, list(perm)]
    acc, ori = np.zeros_like(_acc), np.zeros_like(_ori)
    if phone_as_watch:
        acc[:, 0], ori[:, 0] = _acc[:, 3], _ori[:, 3]
    else:
        acc[:, combo], ori[:, combo] = _acc[:, combo], _ori[:, combo]
    return np.concatenate([acc.reshape(n, -1), ori.reshape(n, -1)], axis=1).astype(np.float32)
