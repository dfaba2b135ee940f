"""CPU checks of oracle/physics_port.py -- the float64 statement of the K8 optimizer (PARITY UNPINNED: the reference's
`dynamics` module is absent).  What can be pinned is pinned: its forward kinematics against the live
reference's ParametricModel.forward_kinematics (tests/golden/metrics_unit.npz); the rest is self-consistency
(Jacobian vs finite differences, optimality of the solve, behaviour in limiting cases)."""
import numpy as np

from conftest import load_golden
from oracle import physics_port as pp
from physics_inputs import synthetic_motion


def test_fk_matches_reference_forward_kinematics():
    g = load_golden('metrics_unit')
    G, P = pp.forward_kinematics(g['pose_a'].numpy(), g['tran_a'].numpy())
    assert np.abs(G - g['glb_a'].numpy()).max() < 2e-6
    assert np.abs(P - g['joint_a'].numpy()).max() < 2e-6


def test_position_jacobian_matches_finite_differences():
    R, _, _ = synthetic_motion(2, 1, seed=3)
    R = R[:, 0].astype(np.float64)
    G, P = pp.forward_kinematics(R)
    J = pp.position_jacobian(G, P)
    eps = 1e-6
    for ci, k in enumerate(pp.OPT_JOINTS):
        for a in range(3):
            w = np.zeros(3); w[a] = eps
            R2 = R.copy(); R2[:, k] = R[:, k] @ pp.exp_so3(w)
            _, P2 = pp.forward_kinematics(R2)
            assert np.abs((P2 - P) / eps - J[..., 3 * ci + a]).max() < 5e-6


def test_solution_is_the_minimiser_of_the_frame_cost():
    R, vel, contact = synthetic_motion(3, 2, seed=5)
    opt = pp.PhysicsOptimizerPort(B=3)
    opt.optimize_frames(R[:, 0], vel[:, 0].reshape(3, 24, 3) * 2, contact[:, 0])
    H, g = opt.normal_equations(R[:, 1].astype(np.float64), vel[:, 1].reshape(3, 24, 3) * 2.0, contact[:, 1])
    x = np.linalg.solve(H, g[..., None])[..., 0]
    assert np.linalg.eigvalsh(H).min() > 0.5e-2        # SPD, bounded below by the absolute damping
    cost = lambda v: 0.5 * np.einsum('bi,bij,bj->b', v, H, v) - np.einsum('bi,bi->b', g, v)
    rng = np.random.default_rng(0)
    for _ in range(8):
        assert np.all(cost(x) <= cost(x + rng.normal(size=x.shape) * 1e-3) + 1e-15)


def test_consistent_motion_is_left_alone():
    """If the velocity head says exactly what the poses do and no foot is in contact, nothing needs optimising."""
    R, _, _ = synthetic_motion(2, 6, seed=7, amp=0.2)
    _, P = pp.forward_kinematics(R.astype(np.float64))
    P = P + 0.5        # keep the feet above the floor: the clamp stays out of the way
    opt = pp.PhysicsOptimizerPort(B=2, floor_y=-10.0)
    vel = np.zeros((2, 6, 24, 3))
    vel[:, 1:] = (P[:, 1:] - P[:, :-1]) * 30.0          # m/s at 30 fps
    contact = np.full((2, 6, 2), -5.0)
    pose, tran = opt.optimize_sequences(R, vel.reshape(2, 6, 72), contact, vel_scale=1.0)
    assert np.abs(pose - R).max() < 1e-9 and np.abs(tran).max() < 1e-9


def test_stance_foot_stays_put_when_contact_dominates():
    R, vel, contact = synthetic_motion(2, 40, seed=9)
    contact[..., 0], contact[..., 1] = 6.0, -6.0        # left foot planted throughout
    opt = pp.PhysicsOptimizerPort(B=2, w_contact=1e4, floor_y=-10.0)
    pose, tran = opt.optimize_sequences(R, vel, contact)
    _, P = pp.forward_kinematics(pose)
    foot = P[:, :, 10] + tran
    assert np.abs(foot[:, 1:] - foot[:, :1]).max() < 2e-3
    free = P[:, :, 11] + tran
    assert np.abs(free[:, 1:] - free[:, :1]).max() > 2e-2


def test_floor_clamp_and_ragged_lengths():
    R, vel, contact = synthetic_motion(3, 30, seed=11)
    opt = pp.PhysicsOptimizerPort(B=3)
    pose, tran = opt.optimize_sequences(R, vel, contact, lengths=[30, 17, 1])
    _, P = pp.forward_kinematics(pose)
    feet_y = np.minimum(P[:, :, 10, 1], P[:, :, 11, 1]) + tran[:, :, 1]
    for b, L in enumerate([30, 17, 1]):
        assert feet_y[b, :L].min() >= pp.FLOOR_Y - 1e-12
        assert np.array_equal(pose[b, L:], R[b, L:].astype(np.float64))      # frames past the length pass through
    assert np.linalg.norm(pose.reshape(-1, 3, 3) @ pose.reshape(-1, 3, 3).transpose(0, 2, 1) - np.eye(3), axis=(1, 2)).max() < 1e-6    # the inputs are float32 rotations


def test_c_port_matches_numpy_port():
    """oracle/physics_port.c (what bench.py's CPU arm times) == oracle/physics_port.py, ragged batch included."""
    from oracle.physics_c import PhysicsOptimizerC
    R, vel, contact = synthetic_motion(4, 50, seed=13)
    lens = [50, 50, 33, 2]
    ref_pose, ref_tran = pp.PhysicsOptimizerPort(B=4).optimize_sequences(R, vel, contact, lens)
    pose, tran = PhysicsOptimizerC(B=4).optimize_sequences(R, vel, contact, lens)
    assert np.abs(pose - ref_pose).max() < 5e-7 and np.abs(tran - ref_tran).max() < 5e-7
