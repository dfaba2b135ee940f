"""TEST INFRASTRUCTURE ONLY -- golden vectors of the reference's INPUT ASSEMBLY from the live reference.

    python oracle/make_golden_inputs.py        # writes tests/golden/input_assembly.npz

Runs only in the build container (needs /root/reference).  Calls, unmodified and through the shims of make_golden.py:
  * PoseDataset._process_combo_data   (mobileposer/data.py:69-86)   -- all 12 combos of one raw stream, evaluate mode
  * DataLoader._get_imu               (mobileposer/loader.py:39-49) -- one combo, acc smoothed by smooth_avg
  * smooth_avg                        (mobileposer/utils/model_utils.py:28-37)
"""
from __future__ import annotations

import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import make_golden as MG          # sets sys.path for the reference + shims

import numpy as np
import torch


@torch.no_grad()
def main():
    cwd = os.getcwd()
    os.chdir(os.path.join(MG.REF, 'mobileposer'))
    try:
        import mobileposer.config as RC
        from mobileposer.data import PoseDataset
        from mobileposer.loader import DataLoader
        from mobileposer.utils.model_utils import smooth_avg
    finally:
        os.chdir(cwd)
    g = torch.Generator().manual_seed(2024)
    T = 37
    raw_acc = torch.randn(T, 6, 3, generator=g) * 4.0          # processed files carry 6 slots; the reference keeps [:, :5]
    raw_ori = torch.randn(T, 6, 3, 3, generator=g)
    # data.py:61: acc[:, :5] / acc_scale, ori[:, :5]; then _process_combo_data for every combo (evaluate mode: whole sequence)
    acc, ori = raw_acc[:, :5] / RC.amass.acc_scale, raw_ori[:, :5]
    fake = types.SimpleNamespace(combos=list(RC.amass.combos.items()), evaluate='dip', finetune=None)
    data = {k: [] for k in ['imu_inputs', 'pose_outputs', 'joint_outputs', 'tran_outputs']}
    dummy = torch.zeros(T, 1)
    PoseDataset._process_combo_data(fake, acc, ori, dummy, dummy, dummy, None, data)
    dataset_imu = torch.stack(data['imu_inputs'])              # [12, T, 60]
    # loader.py:39-49 for two combos
    loader = {}
    for name in ('lw_rp', 'rw_lp_h'):
        fake_l = types.SimpleNamespace(combo=RC.amass.combos[name])
        d, a, o = DataLoader._get_imu(fake_l, raw_acc[:, :5].clone(), raw_ori[:, :5].clone())
        loader[name] = d
    MG.save('input_assembly', raw_acc=raw_acc, raw_ori=raw_ori,
            combo_slots=np.array([sum(1 << s for s in c) for c in RC.amass.combos.values()], np.int32),
            dataset_imu=dataset_imu, loader_lw_rp=loader['lw_rp'], loader_rw_lp_h=loader['rw_lp_h'],
            smooth3=smooth_avg(raw_acc[:, :5].clone()), acc_scale=np.float32(RC.amass.acc_scale))


if __name__ == '__main__':
    main()
