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
