"""GPU suite of K8, the kinematic-physics optimizer behind the PHYSICS hook (csrc/physics.cu through the C ABI).

PARITY UNPINNED against the reference (its `dynamics` module is absent): the checker is the float64
statement of this repository's definition (oracle/physics_port.py), which uses a different algebra (explicit Jacobian,
dense LAPACK solve) than the kernel (subtree moments, envelope Cholesky).  Pinned to the reference: the forward
kinematics (golden metrics_unit.npz from ParametricModel.forward_kinematics).
Tolerances: the north star's 1e-4 rad / 1e-4 m."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import physics_port as pp
from parity import ANGLE_TOL, TRAN_TOL, max_abs, max_angle
from physics_inputs import synthetic_motion

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'

# elimination order of the kernel's unknowns (csrc/physics.cu:kOrd) -> permutation of the oracle's [45 + 3]
ORD = [15, 12, 18, 16, 13, 19, 17, 14, 9, 6, 3, 4, 1, 5, 2]
PERM = [3 * pp.OPT_JOINTS.index(k) + a for k in ORD for a in range(3)] + [45, 46, 47]


def _gpu(*arrs):
    return [torch.from_numpy(np.ascontiguousarray(a)).to(DEV) for a in arrs]


def test_fk_matches_reference_forward_kinematics():
    from mobileposer_b200.dynamics import forward_kinematics
    g = load_golden('metrics_unit')
    glb, pos = forward_kinematics(g['pose_a'].to(DEV))
    assert max_abs(glb, g['glb_a']) < 2e-6
    assert max_abs(pos, g['joint_a'] - g['tran_a'][:, None, :]) < 2e-6


def test_normal_equations_and_solution_match_the_oracle():
    from mobileposer_b200.dynamics import PhysicsOptimizer
    B = 5
    R, vel, contact = synthetic_motion(B, 3, seed=21)
    opt = PhysicsOptimizer()
    dbg = torch.zeros(B, 49 * 49 + 48, device=DEV)
    tR, tv, tc = _gpu(R.reshape(B, 3, 24, 9), vel, contact)
    opt._run(tR, tv, tc, None, 2.0, dbg=dbg, dbg_frame=2)
    torch.cuda.synchronize()
    port = pp.PhysicsOptimizerPort(B=B)
    jv = vel.reshape(B, 3, 24, 3).astype(np.float64) * 2.0
    for t in range(2):
        port.optimize_frames(R[:, t], jv[:, t], contact[:, t])
    H, g = port.normal_equations(R[:, 2].astype(np.float64), jv[:, 2], contact[:, 2])
    x = np.linalg.solve(H, g[..., None])[..., 0]
    Hp, gp, xp = H[:, PERM][:, :, PERM], g[:, PERM], x[:, PERM]
    d = dbg.cpu().numpy().astype(np.float64)
    Hk = d[:, :49 * 49].reshape(B, 49, 49)
    low = np.tril(np.ones((48, 48), bool))
    scale = np.abs(Hp).max()
    assert np.abs(Hk[:, :48, :48][:, low] - Hp[:, low]).max() < 2e-5 * scale
    assert np.abs(Hk[:, 48, :48] - gp).max() < 2e-5 * max(1.0, np.abs(gp).max())
    assert np.abs(d[:, 49 * 49:] - xp).max() < 2e-5


@pytest.mark.parametrize('B,T,lengths', [(4, 64, None), (3, 40, [40, 23, 1])])
def test_sequences_match_the_oracle(B, T, lengths):
    from mobileposer_b200.dynamics import PhysicsOptimizer
    R, vel, contact = synthetic_motion(B, T, seed=31 + B)
    opt = PhysicsOptimizer()
    tR, tv, tc = _gpu(R, vel, contact)
    lens = torch.tensor(lengths, dtype=torch.int32, device=DEV) if lengths else None
    pose, tran = opt.optimize_sequences(tR, tv, tc, lens)
    ref_pose, ref_tran = pp.PhysicsOptimizerPort(B=B).optimize_sequences(R, vel, contact, lengths)
    a, t = max_angle(pose, torch.from_numpy(ref_pose)), max_abs(tran, torch.from_numpy(ref_tran))
    print(f'K8 B={B} T={T}: max angle err {a:.2e} rad, max tran err {t:.2e} m')
    assert a <= ANGLE_TOL and t <= TRAN_TOL
    if lengths:
        for b, L in enumerate(lengths):
            assert torch.equal(pose[b, L:].cpu(), torch.from_numpy(R[b, L:]))


def test_frame_by_frame_equals_one_launch():
    """optimize_frame called per frame (the reference's loop, net.py:165-168) == the batched launch, bit for bit."""
    from mobileposer_b200.dynamics import PhysicsOptimizer
    R, vel, contact = synthetic_motion(1, 20, seed=41)
    tR, tv, tc = _gpu(R, vel, contact)
    pose, tran = PhysicsOptimizer().optimize_sequences(tR, tv, tc)
    opt = PhysicsOptimizer()
    opt.reset_states()
    for t in range(20):
        p, tr = opt.optimize_frame(tR[0, t], tv[0, t].view(24, 3) * 2.0, tc[0, t], torch.zeros(5, 3))
        assert torch.equal(p, pose[0, t]) and torch.equal(tr, tran[0, t])


def test_limits_damping_and_contact():
    from mobileposer_b200.dynamics import PhysicsOptimizer, forward_kinematics
    R, vel, contact = synthetic_motion(2, 60, seed=51)
    tR, tv, tc = _gpu(R, vel, contact)
    pose, _ = PhysicsOptimizer(damping=1e9, damping_abs=1e9).optimize_sequences(tR, tv, tc)
    assert max_abs(pose, tR) < 1e-6                       # infinite damping: the network pose passes through
    tc[..., 0], tc[..., 1] = 6.0, -6.0                    # left foot planted
    pose, tran = PhysicsOptimizer(w_contact=1e4, floor_y=-10.0).optimize_sequences(tR, tv, tc)
    _, pos = forward_kinematics(pose.view(-1, 24, 3, 3))
    foot = pos.view(2, 60, 24, 3)[:, :, 10] + tran
    assert (foot[:, 1:] - foot[:, :1]).abs().max().item() < 2e-3


def test_baseline_size_properties():
    """B = 256 x T = 300 (BASELINE config 3): orthonormal output, batch invariance, floor respected."""
    from mobileposer_b200.dynamics import PhysicsOptimizer, forward_kinematics
    R, vel, contact = synthetic_motion(256, 300, seed=61)
    tR, tv, tc = _gpu(R, vel, contact)
    pose, tran = PhysicsOptimizer().optimize_sequences(tR, tv, tc)
    m = pose.view(-1, 3, 3).double()
    assert (m @ m.transpose(1, 2) - torch.eye(3, device=DEV, dtype=torch.float64)).abs().max().item() < 5e-6
    sub, subt = PhysicsOptimizer().optimize_sequences(tR[:3], tv[:3], tc[:3])
    assert torch.equal(sub, pose[:3]) and torch.equal(subt, tran[:3])
    _, pos = forward_kinematics(pose.view(-1, 24, 3, 3))
    pos = pos.view(256, 300, 24, 3)
    feet_y = torch.minimum(pos[:, :, 10, 1], pos[:, :, 11, 1]) + tran[:, :, 1]
    assert feet_y.min().item() >= pp.FLOOR_Y - 1e-5


def test_forward_offline_with_physics_hook(seeded_state_dict, oracle):
    """MobilePoserNet.forward_offline with the hook on == oracle forward + float64 optimizer on the oracle's outputs."""
    import mobileposer_b200 as mp
    from mobileposer_b200.synthetic import synthetic_imu_batch
    from oracle.torch_port import offline_batched
    net = mp.MobilePoserNet()
    net.load_state_dict(seeded_state_dict)
    net = net.to(DEV).eval()
    net.enable_physics()
    lens = [48, 31]
    x = synthetic_imu_batch([5, 6], 48)
    x[1, 31:] = 0
    pose, joints, tran, contact = net.forward_offline(x.to(DEV), lens)
    pose = pose.view(2, 48, 24, 3, 3)
    # the optimizer's inputs from the CUDA net itself (its parity with the oracle is test_gpu_parity's subject)
    net.enable_physics(False)
    p0, _, t0, c0 = net.forward_offline(x.to(DEV), lens)
    assert torch.equal(t0, tran) and torch.equal(c0, contact)          # tran / contact untouched by the hook (net.py:169)
    _, _, vel, _ = net.forward(x.to(DEV), lens)
    ref_pose, _ = pp.PhysicsOptimizerPort(B=2).optimize_sequences(
        p0.view(2, 48, 24, 3, 3).cpu().numpy(), vel.cpu().numpy(), c0.cpu().numpy(), lens)
    for b, L in enumerate(lens):
        assert max_angle(pose[b, :L], torch.from_numpy(ref_pose[b, :L])) <= ANGLE_TOL
    assert max_angle(pose[0], p0.view(2, 48, 24, 3, 3)[0]) > 1e-3       # the hook did something


def test_forward_offline_b1_carries_the_optimizer_state(seeded_state_dict):
    """B == 1 is the reference's own call pattern: reset_states() only runs in the constructor (net.py:69), so the
    optimizer state of one forward_offline call is the initial state of the next."""
    import mobileposer_b200 as mp
    from mobileposer_b200.synthetic import synthetic_imu_batch
    net = mp.MobilePoserNet()
    net.load_state_dict(seeded_state_dict)
    net = net.to(DEV).eval()
    xs = [synthetic_imu_batch([9], 30).to(DEV), synthetic_imu_batch([10], 24).to(DEV)]
    net.enable_physics()
    got = []
    for x in xs:
        net.velocity.rnn_state = None
        got.append(net.forward_offline(x, [x.shape[1]])[0])
    net.enable_physics(False)
    port = pp.PhysicsOptimizerPort(B=1)
    for x, pose in zip(xs, got):
        T = x.shape[1]
        net.velocity.rnn_state = None
        p0, _, _, c0 = net.forward_offline(x, [T])
        net.velocity.rnn_state = None
        vel = net.forward(x, [T])[2]
        ref, _ = port.optimize_sequences(p0.view(1, T, 24, 3, 3).cpu().numpy(), vel.view(1, T, 72).cpu().numpy(),
                                         c0.view(1, T, 2).cpu().numpy())
        assert max_angle(pose.view(T, 24, 3, 3), torch.from_numpy(ref[0])) <= ANGLE_TOL


def test_forward_online_with_the_physics_hook(seeded_state_dict):
    """net.py:211-217: the online tick ends with optimize_frame on the tick's frame (T = 1 launches with the optimizer state
    kept on the device between ticks): [24, 3, 3] rotations, equal to the optimizer applied to the hook-less online outputs."""
    import mobileposer_b200 as mp
    from mobileposer_b200.synthetic import synthetic_imu
    x = synthetic_imu(31, 12).to(DEV)

    def run(physics):
        net = mp.MobilePoserNet()
        net.load_state_dict(seeded_state_dict)
        net = net.to(DEV).eval()
        net.enable_physics(physics)
        out = []
        for t in range(x.shape[0]):
            pose, _, root, contact = net.forward_online(x[t])
            vel_p = net._slots[next(iter(net._slots))].vel[0, net.num_past_frames].clone()
            out.append((pose.clone().view(24, 3, 3), root.clone(), contact.clone(), vel_p))
        return out

    plain, hooked = run(False), run(True)
    port = pp.PhysicsOptimizerPort(B=1)
    for (p0, r0, c0, v0), (p1, r1, c1, _) in zip(plain, hooked):
        assert torch.equal(r0, r1) and torch.equal(c0, c1)              # translation and contact are not touched (net.py:215)
        ref, _ = port.optimize_sequences(p0.view(1, 1, 24, 3, 3).cpu().numpy(), v0.view(1, 1, 72).cpu().numpy(),
                                         c0.view(1, 1, 2).cpu().numpy())
        assert max_angle(p1, torch.from_numpy(ref[0, 0])) <= ANGLE_TOL
    assert max_angle(hooked[-1][0], plain[-1][0]) > 1e-4                # the hook did something by the last tick
