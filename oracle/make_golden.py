"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the LIVE reference.

Runs only in the build container (needs /root/reference, which does not exist
on the GPU box).  Imports the unmodified reference `MobilePoserNet` through the
two import shims in oracle/shims/ (lightning, chumpy -- SURVEY.md section 8c),
seeds it, feeds it the synthetic IMU windows of mobileposer_b200.synthetic and
dumps inputs + outputs as small fixtures.  Also re-checks the SMPL constants
mirrored in mobileposer_b200/config.py against the reference's pickle.

    python oracle/make_golden.py            # rewrites tests/golden/

torch version that produced the committed fixtures is recorded in
tests/golden/MANIFEST.json.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'shims'))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import numpy as np
import torch

from mobileposer_b200 import config as C
from mobileposer_b200.synthetic import synthetic_imu, synthetic_imu_batch, well_conditioned_state_dict

OUT = os.path.join(ROOT, 'tests', 'golden')


def load_reference(seed=0):
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, 'mobileposer'))     # config.py:28-30 resolves smpl/ from cwd
    try:
        from mobileposer.models import MobilePoserNet
        torch.manual_seed(seed)
        net = MobilePoserNet().eval()
    finally:
        os.chdir(cwd)
    return net


def tensor_sha(t):
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()


def save(name, **arrs):
    np.savez_compressed(os.path.join(OUT, name + '.npz'),
                        **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                           for k, v in arrs.items()})
    print('wrote', name, {k: tuple(np.asarray(v.detach().cpu() if torch.is_tensor(v) else v).shape)
                          for k, v in arrs.items()})


@torch.no_grad()
def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    net = load_reference(0)
    sd = net.state_dict()

    # --- constants mirrored in config.py must equal the reference's -------------------------
    parent = [-1 if p is None else int(p) for p in net.bodymodel.parent]
    assert parent == C.SMPL_PARENT, parent
    assert np.array_equal(net.j.numpy(), np.asarray(C.SMPL_J_ZERO, np.float32))
    assert net.floor_y == C.FLOOR_Y
    import mobileposer.config as RC
    assert RC.joint_set.reduced == C.joint_set.reduced and RC.joint_set.ignored == C.joint_set.ignored
    assert RC.amass.combos == C.amass.combos and RC.datasets.fps == C.datasets.fps
    assert RC.model_config.past_frames == C.model_config.past_frames
    assert RC.model_config.future_frames == C.model_config.future_frames

    manifest = {
        'torch': torch.__version__,
        'reference_commit': 'eddfe2a (SURVEY.md header)',
        'weights': 'torch.manual_seed(0); MobilePoserNet() default init',
        'weight_sha256': {k: tensor_sha(v) for k, v in sd.items()},
        'weight_shapes': {k: list(v.shape) for k, v in sd.items()},
    }

    # --- A: cfg1, joints head only, one 300-frame sequence ---------------------------------
    x = synthetic_imu(0, 300)
    save('cfg1_joints_T300', imu=x, joints=net.joints(x[None], [300])[0])

    # --- B: cfg2, full forward + forward_offline, B=1, T=300 -------------------------------
    net.velocity.rnn_state = None
    pose, joints, vel, contact = net.forward(x[None], [300])
    hn, cn = net.velocity.rnn_state
    net.velocity.rnn_state = None
    net.reset()
    pose_o, joints_o, tran_o, contact_o = net.forward_offline(x[None], [300])
    assert torch.equal(pose, pose_o) and torch.equal(joints, joints_o)
    save('cfg2_forward_T300', imu=x, pose=pose, joints=joints[0], vel=vel, contact=contact[0],
         tran=tran_o, vel_hn=hn, vel_cn=cn)

    # --- C: ragged batched forward ----------------------------------------------------------
    lens = [50, 17, 33]
    xb = synthetic_imu_batch([1, 2, 3], 50)
    for b, L in enumerate(lens):
        xb[b, L:] = 0
    net.velocity.rnn_state = None
    pose, joints, vel, contact = net.forward(xb, lens)
    hn, cn = net.velocity.rnn_state
    save('ragged_forward_B3', imu=xb, lengths=np.asarray(lens), pose=pose, joints=joints, vel=vel,
         contact=contact, vel_hn=hn, vel_cn=cn)

    # --- D: velocity state leak across sequences -----------------------------
    xa, xc = synthetic_imu(10, 100), synthetic_imu(11, 100)
    net.velocity.rnn_state = None
    net.reset()
    _, _, tran_a, _ = net.forward_offline(xa[None], [100])
    net.reset()                                   # does NOT clear velocity.rnn_state
    _, _, tran_leak, _ = net.forward_offline(xc[None], [100])
    net.velocity.rnn_state = None
    net.reset()
    _, _, tran_clean, _ = net.forward_offline(xc[None], [100])
    save('state_leak_T100', imu_a=xa, imu_b=xc, tran_a=tran_a, tran_b_leak=tran_leak, tran_b_clean=tran_clean)

    # --- E: online, 60 ticks, reference window 45 -------------------------------------------
    xo = synthetic_imu(20, 60)
    net2 = load_reference(0)                       # fresh foot/root state
    net2.velocity.rnn_state = None
    poses, roots, contacts, lastj = [], [], [], None
    for f in xo:
        p, j, r, c = net2.forward_online(f)
        poses.append(p.clone()); roots.append(r.clone()); contacts.append(c.clone()); lastj = j
    hn, cn = net2.velocity.rnn_state
    save('online_60ticks', imu=xo, pose=torch.stack(poses), root=torch.stack(roots),
         contact=torch.stack(contacts), last_joints=lastj, vel_hn=hn, vel_cn=cn)

    # --- F: tiny lengths -----------------------------------------------------------------------
    for T in (1, 3):
        xt = synthetic_imu(30 + T, T)
        net.velocity.rnn_state = None
        net.reset()
        p, j, t, c = net.forward_offline(xt[None], [T])
        save(f'edge_T{T}', imu=xt, pose=p, joints=j[0], tran=t, contact=c)

    # --- G: K5 unit incl. degenerate r6d (NaN -> 0 rows, angular.py:181) --------------------
    g = torch.Generator().manual_seed(77)
    r6d = torch.randn(40, 96, generator=g)
    r6d[3, 0:3] = 0                                # zero first column of joint 0
    r6d[5, 6 + 3:6 + 6] = r6d[5, 6:6 + 3] * 2      # colinear columns of joint 1
    r6d[7] = 0                                     # everything degenerate
    r6d[9, 90:96] = 0
    save('k5_unit', r6d=r6d, pose=net._reduced_global_to_full(r6d.clone()))

    # --- H: K6 unit: crafted head outputs that exercise ties and the floor clamp -----------
    T = 96
    g = torch.Generator().manual_seed(78)
    jh = torch.randn(1, T, 72, generator=g) * 0.05
    t = torch.arange(T, dtype=torch.float32)
    jh[0, :, 10 * 3 + 1] = -0.93 - 0.05 * torch.sin(t / 7)          # feet near the floor
    jh[0, :, 11 * 3 + 1] = -0.94 - 0.05 * torch.cos(t / 5)
    vh = torch.randn(T, 72, generator=g)
    vh[:, 1] = -1.5 + torch.randn(T, generator=g)                     # mostly falling
    ch = torch.randn(1, T, 2, generator=g) * 3
    ch[0, 5] = torch.tensor([0.7, 0.7])                               # tie -> index 0 (left)
    ch[0, 6] = torch.tensor([-4.0, -5.0])                             # low prob -> weight 0
    ch[0, 7] = torch.tensor([6.0, 1.0])                               # high prob -> weight 1
    dummy_pose = torch.eye(3).repeat(T, 24, 1, 1)
    real_forward = net.forward
    net.forward = lambda imu, lens=None: (dummy_pose, jh, vh, ch)
    net.reset()
    _, _, tran_h, _ = net.forward_offline(torch.zeros(1, T, 60), [T])
    net.forward = real_forward
    save('k6_unit', joints=jh[0], vel=vh, contact=ch[0], tran=tran_h)

    # --- I: K7 unit: online state machine on crafted head outputs ---------------------------
    net3 = load_reference(0)
    W, P = 45, 40
    g = torch.Generator().manual_seed(79)
    n_ticks = 24
    J = torch.randn(n_ticks, 72, generator=g) * 0.05
    J[:, 31] = -0.93 - 0.04 * torch.sin(torch.arange(n_ticks) / 3.0)
    J[:, 34] = -0.95 + 0.04 * torch.cos(torch.arange(n_ticks) / 2.0)
    V = torch.randn(n_ticks, 72, generator=g)
    V[:, 1] = -2.0 + torch.randn(n_ticks, generator=g)
    Cn = torch.randn(n_ticks, 2, generator=g) * 1.5
    Cn[4] = torch.tensor([0.3, 0.3])                                   # tie -> right foot online (strict >)
    Cn[5] = torch.tensor([0.95, -1.0])
    R6 = torch.randn(n_ticks, 96, generator=g)
    tick = {'i': 0}

    def fake_forward(imu, lens=None):
        i = tick['i']
        pose = net3._reduced_global_to_full(R6[i].repeat(W, 1))
        return pose, J[i].repeat(1, W, 1), V[i].repeat(W, 1), Cn[i].repeat(1, W, 1)

    net3.forward = fake_forward
    roots, poses = [], []
    for i in range(n_ticks):
        tick['i'] = i
        p, _, r, _ = net3.forward_online(torch.zeros(60))
        roots.append(r.clone()); poses.append(p.clone())
    save('k7_unit', joints=J, vel=V, contact=Cn, r6d=R6, root=torch.stack(roots), pose=torch.stack(poses))

    # --- J: batch of equal-length sequences == B independent forward_offline calls (F6) -----
    ids = list(range(40, 48))
    xj = synthetic_imu_batch(ids, 64)
    poses, joints, trans, contacts = [], [], [], []
    for b in range(len(ids)):
        net.velocity.rnn_state = None
        net.reset()
        p, j, t, c = net.forward_offline(xj[b:b + 1], [64])
        poses.append(p); joints.append(j[0]); trans.append(t); contacts.append(c)
    save('batch8_T64', imu=xj, pose=torch.stack(poses), joints=torch.stack(joints),
         tran=torch.stack(trans), contact=torch.stack(contacts))

    # --- K: metric side: SMPL FK and the FullMotionEvaluator rows on a small random motion ---
    import mobileposer.articulate as art
    g = torch.Generator().manual_seed(80)
    n = 40
    def rand_rot(k):
        q = torch.randn(k, 4, generator=g)
        q = q / q.norm(dim=1, keepdim=True)
        w, x, y, z = q.unbind(1)
        return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                            2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                            2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], dim=1).view(k, 3, 3)
    pose_a = rand_rot(n * 24).view(n, 24, 3, 3)
    pose_b = rand_rot(n * 24).view(n, 24, 3, 3)
    tran_a = torch.randn(n, 3, generator=g) * 0.1
    tran_b = torch.randn(n, 3, generator=g) * 0.1
    glb_a, joint_a = net.bodymodel.forward_kinematics(pose_a, tran=tran_a)
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, 'mobileposer'))
    ev = art.FullMotionEvaluator(str(RC.paths.smpl_file), joint_mask=torch.tensor([2, 5, 16, 20]), fps=RC.datasets.fps)
    os.chdir(cwd)
    errs = ev(pose_a, pose_b, tran_p=tran_a, tran_t=tran_b)
    save('metrics_unit', pose_a=pose_a, pose_b=pose_b, tran_a=tran_a, tran_b=tran_b, glb_a=glb_a, joint_a=joint_a, errs=errs)

    # --- L: WELL-CONDITIONED weights (mobileposer_b200.synthetic.well_conditioned_state_dict): the seeded init with the pose
    # head's linear2 re-centred on (1,0,0 | 0,1,0), so the r6d -> rotation step has trained-model conditioning and the flat
    # 1e-4 rad of BASELINE.json:north_star can be asserted on every (frame, joint) -----------------------------------------
    net_wc = load_reference(0)
    wc = well_conditioned_state_dict(net_wc.state_dict())
    net_wc.load_state_dict(wc)
    manifest['wc_weights'] = 'seed-0 init; pose.pose.linear2.bias = [1,0,0,0,1,0]x16, pose.pose.linear2.weight *= 0.1'
    manifest['wc_weight_sha256'] = {k: tensor_sha(wc[k]) for k in ('pose.pose.linear2.bias', 'pose.pose.linear2.weight')}
    x = synthetic_imu(0, 300)
    net_wc.velocity.rnn_state = None
    net_wc.reset()
    pose_o, joints_o, tran_o, contact_o = net_wc.forward_offline(x[None], [300])
    save('wc_cfg2_offline_T300', imu=x, pose=pose_o, joints=joints_o[0], tran=tran_o, contact=contact_o)
    lens = [50, 17, 33]
    xb = synthetic_imu_batch([1, 2, 3], 50)
    for b, L in enumerate(lens):
        xb[b, L:] = 0
    net_wc.velocity.rnn_state = None
    pose, joints, vel, contact = net_wc.forward(xb, lens)
    save('wc_ragged_forward_B3', imu=xb, lengths=np.asarray(lens), pose=pose, joints=joints, vel=vel, contact=contact)
    xo = synthetic_imu(20, 50)
    net_wc2 = load_reference(0)
    net_wc2.load_state_dict(wc)
    net_wc2.velocity.rnn_state = None
    poses, roots, contacts = [], [], []
    for f in xo:
        p, j, r, c = net_wc2.forward_online(f)
        poses.append(p.clone()); roots.append(r.clone()); contacts.append(c.clone())
    save('wc_online_50ticks', imu=xo, pose=torch.stack(poses), root=torch.stack(roots), contact=torch.stack(contacts))

    with open(os.path.join(OUT, 'MANIFEST.json'), 'w') as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print('manifest written')


if __name__ == '__main__':
    main()
