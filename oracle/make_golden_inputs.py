"""TEST INFRASTRUCTURE ONLY -- golden vectors of the reference's INPUT ASSEMBLY from the live reference.

    python oracle/make_golden_inputs.py        # writes tests/golden/input_assembly.npz and tests/golden/live_normalize.npz

Runs only in the build container (needs /root/reference).  Calls, unmodified and through the shims of make_golden.py:
  * PoseDataset._process_combo_data   (mobileposer/data.py:69-86)   -- all 12 combos of one raw stream, evaluate mode
  * DataLoader._get_imu               (mobileposer/loader.py:39-49) -- one combo, acc smoothed by smooth_avg
  * smooth_avg                        (mobileposer/utils/model_utils.py:28-37)
  * live_demo.py:160-177,210-234      -- calibration + per-tick normalisation of the live demo.  Those lines sit inside the
    script's `__main__` loop (sockets, pygame clock) and cannot be called; live_golden() evaluates the same torch expressions
    on synthetic sensor readings with the reference's own `quaternion_to_rotation_matrix` and `amass` config.
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


@torch.no_grad()
def live_golden():
    cwd = os.getcwd()
    os.chdir(os.path.join(MG.REF, 'mobileposer'))
    try:
        import mobileposer.config as RC
        from mobileposer.articulate.math import quaternion_to_rotation_matrix
    finally:
        os.chdir(cwd)
    g = torch.Generator().manual_seed(77)
    n, n_imus = 9, 5
    # calibration (live_demo.py:160-177): sensor 0 aligned with the body frame, then a T-pose reading of all sensors
    oris0 = torch.randn(4, generator=g)
    smpl2imu = quaternion_to_rotation_matrix(oris0).view(3, 3).t()
    oris_t, accs_t = torch.randn(n_imus, 4, generator=g), torch.randn(n_imus, 3, generator=g) * 9.8
    device2bone = smpl2imu.matmul(quaternion_to_rotation_matrix(oris_t)).transpose(1, 2).matmul(torch.eye(3))
    acc_offsets = smpl2imu.matmul(accs_t.unsqueeze(-1))
    # per tick (live_demo.py:210-234); the sensor quaternions arrive unnormalised
    ori_q = torch.randn(n, n_imus, 4, generator=g) * 1.7
    acc_raw = torch.randn(n, n_imus, 3, generator=g) * 12.0
    ori_raw = quaternion_to_rotation_matrix(ori_q).view(-1, n_imus, 3, 3)
    glb_acc = (smpl2imu.matmul(acc_raw.view(-1, n_imus, 3, 1)) - acc_offsets).view(-1, n_imus, 3)
    glb_ori = smpl2imu.matmul(ori_raw).matmul(device2bone)
    _acc = glb_acc.view(-1, 5, 3)[:, [1, 4, 3, 0, 2]] / RC.amass.acc_scale
    _ori = glb_ori.view(-1, 5, 3, 3)[:, [1, 4, 3, 0, 2]]
    outs = {}
    for name in ('lw_rp', 'rw_rp_h', 'phone_as_watch'):
        acc, ori = torch.zeros_like(_acc), torch.zeros_like(_ori)
        if name == 'phone_as_watch':
            acc[:, [0]] = _acc[:, [3]]
            ori[:, [0]] = _ori[:, [3]]
        else:
            c = RC.amass.combos[name]
            acc[:, c] = _acc[:, c]
            ori[:, c] = _ori[:, c]
        outs[name] = torch.cat([acc.flatten(1), ori.flatten(1)], dim=1)
    MG.save('live_normalize', ori_q=ori_q, acc_raw=acc_raw, smpl2imu=smpl2imu, device2bone=device2bone,
            acc_offsets=acc_offsets.squeeze(-1), perm=np.array([1, 4, 3, 0, 2], np.int32), ori_raw=ori_raw,
            acc_scale=np.float32(RC.amass.acc_scale), imu_lw_rp=outs['lw_rp'], imu_rw_rp_h=outs['rw_rp_h'],
            imu_phone_as_watch=outs['phone_as_watch'])


if __name__ == '__main__':
    main()
    live_golden()
