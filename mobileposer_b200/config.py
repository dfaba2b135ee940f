"""Constants of the MobilePoser hot path, mirrored from the reference.

Every value here is a *number* the reference defines (mobileposer/config.py and
the SMPL zero-pose skeleton it loads at mobileposer/models/net.py:47-49); the
kernels in csrc/ bake the same numbers in (csrc/mp_constants.cuh, exported by
the C ABI's mp_constants()) and tests/test_constants.py checks the two copies
agree bit for bit.

Reference citations (relative to /root/reference):
  model_config  mobileposer/config.py:40-54
  amass         mobileposer/config.py:57-83
  datasets      mobileposer/config.py:86-126
  joint_set     mobileposer/config.py:129-142
"""
from __future__ import annotations


class model_config:
    """mobileposer/config.py:40-54 (device is resolved lazily -- see net.py)."""
    n_joints = 5
    n_imu = 12 * n_joints            # 60 = 5 x (3 acc + 9 ori)
    n_output_joints = 24
    n_pose_output = n_output_joints * 6
    past_frames = 40
    future_frames = 5
    total_frames = past_frames + future_frames


class amass:
    """mobileposer/config.py:57-83."""
    combos = {
        'lw_rp_h': [0, 3, 4],
        'rw_rp_h': [1, 3, 4],
        'lw_lp_h': [0, 2, 4],
        'rw_lp_h': [1, 2, 4],
        'lw_lp': [0, 2],
        'lw_rp': [0, 3],
        'rw_lp': [1, 2],
        'rw_rp': [1, 3],
        'lp_h': [2, 4],
        'rp_h': [3, 4],
        'lp': [2],
        'rp': [3],
    }
    acc_scale = 30
    vel_scale = 2


class datasets:
    """mobileposer/config.py:86-126 (only what the hot path reads)."""
    fps = 30
    window_length = 125


class joint_set:
    """mobileposer/config.py:129-142."""
    gravity_velocity = -0.018
    full = list(range(24))
    reduced = [0, 1, 2, 3, 4, 5, 6, 9, 12, 13, 14, 15, 16, 17, 18, 19]
    ignored = [0, 7, 8, 10, 11, 20, 21, 22, 23]
    n_full = len(full)
    n_ignored = len(ignored)
    n_reduced = len(reduced)


# Contact-probability -> lerp-weight clamp (mobileposer/models/net.py:53,90-91).
PROB_THRESHOLD = (0.5, 0.9)

# SMPL kinematic tree (articulate/model.py:36-37 reading kintree_table of
# smpl/basicmodel_m.pkl); -1 marks the root.
SMPL_PARENT = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]

# Zero-pose joint positions J - J[0] as float32 (articulate/model.py:77-92 with
# shape=None); shortest round-trip reprs of the float32 values, extracted by
# oracle/make_golden.py (which re-checks them against the pickle).
SMPL_J_ZERO = [
    [0.0, 0.0, 0.0],
    [0.058581352, -0.08228004, -0.017664082],
    [-0.060309727, -0.09051329, -0.013542531],
    [0.004439451, 0.12440355, -0.03838522],
    [0.10203278, -0.46874952, -0.009627081],
    [-0.103566356, -0.47420114, -0.018385574],
    [0.008927891, 0.26235995, -0.0115648955],
    [0.08724245, -0.8956239, -0.047055073],
    [-0.08451081, -0.8942467, -0.052947246],
    [0.0066633024, 0.31839234, -0.008709848],
    [0.1282968, -0.95590985, 0.074987344],
    [-0.11935069, -0.95635235, 0.07737604],
    [-0.006726882, 0.53002787, -0.042177428],
    [0.07836577, 0.43239203, -0.02760802],
    [-0.076290354, 0.4308647, -0.032417234],
    [0.0033863292, 0.61896527, 0.008232435],
    [0.20128717, 0.47759712, -0.04665402],
    [-0.18951866, 0.47771794, -0.040889304],
    [0.45661905, 0.4619481, -0.06960051],
    [-0.44964615, 0.46334866, -0.07215803],
    [0.7223283, 0.4746462, -0.07697524],
    [-0.7187546, 0.47014236, -0.0781848],
    [0.80901885, 0.46401018, -0.09256954],
    [-0.80750835, 0.4614908, -0.088291876],
]

LFOOT, RFOOT = 10, 11
# floor_y = min(j[10:12, 1]) as a float32 value widened to double
# (mobileposer/models/net.py:49: `.item()`).
FLOOR_Y = -0.9563523530960083

# Head shapes: (n_input, n_output, n_hidden, bidirectional)
#   joints.py:29, poser.py:32, footcontact.py:28, velocity.py:29
HEAD_SHAPES = {
    'joints': (60, 72, 256, True),
    'pose': (132, 96, 256, True),
    'foot_contact': (132, 2, 64, True),
    'velocity': (132, 72, 256, False),
}
# state_dict prefixes.
HEAD_PREFIX = {
    'joints': 'joints.joints.',
    'pose': 'pose.pose.',
    'foot_contact': 'foot_contact.footcontact.',
    'velocity': 'velocity.vel.',
}
