"""TEST INFRASTRUCTURE ONLY -- float64 numpy statement of the kinematic-physics optimizer behind the PHYSICS hook.

PARITY UNPINNED.  The reference calls `dynamics.PhysicsOptimizer(debug=False)`, `.reset_states()` and
`.optimize_frame(pose, jvel, contact, acc) -> (pose, tran)` at mobileposer/models/net.py:66-69,157-169,
211-217, but the module `dynamics` is not in the reference tree and its dependency `rbdl` is neither
vendored nor pinned.  There is nothing of the reference's to restate, so the
algorithm below is DEFINED by this repository behind the reference's hook signature,
and this file is its checker: the straightforward formulation (explicit Jacobian, dense normal equations,
LAPACK solve), deliberately different in algebra from the CUDA kernel (subtree moments + in-warp Cholesky).
What IS pinned to the reference inside it: `forward_kinematics` against
`ParametricModel.forward_kinematics` (articulate/model.py:208-232) through tests/golden/metrics_unit.npz.

Per frame and skeleton, with theta the network's local rotations (K5 output), p the optimizer's root position
and q_prev the previous frame's optimized world joint positions:

    e_j  = q_prev_j + jvel_j / fps - (p + FK_j(theta))          velocity-consistent target displacement, 24 joints
    ec_f = q_prev_f - (p + FK_f(theta))                          stance foot stays where it was, f in {10, 11}
    wc_f = w_contact * prob_to_weight(sigmoid(contact_f))        same clamp as net.py:90-91
    min over (dtheta [15 x 3], d [3]):
        w_vel * sum_j |J_j dtheta + d - e_j|^2 + sum_f wc_f |J_f dtheta + d - ec_f|^2
            + sum_i (damping * H_ii + damping_abs) dtheta_i^2        (Marquardt scaling: H = J^T W J of the rows above)
    R_k <- R_k exp([dtheta_k]x)  for the 15 optimised joints (joint_set.reduced without the root)
    floor: if min_f (p + d + FK_f(theta'))_y < floor_y: d_y += floor_y - that minimum        (net.py:148-153's clamp)
    p <- p + d ;  q_prev <- p + FK(theta')

The first frame after `reset_states()` has no q_prev: pose passes through, p stays, q_prev is initialised.
`acc` is accepted and ignored (the reference passes zeros offline, net.py:159).
"""
from __future__ import annotations

import numpy as np

from mobileposer_b200.config import FLOOR_Y, SMPL_J_ZERO, SMPL_PARENT, datasets, joint_set, PROB_THRESHOLD

PARENT = list(SMPL_PARENT)
J0 = np.asarray(SMPL_J_ZERO, np.float64)
BONE = np.stack([J0[j] - (J0[PARENT[j]] if PARENT[j] >= 0 else 0.0) for j in range(24)])
OPT_JOINTS = [j for j in joint_set.reduced if j != 0]           # 15 joints, 45 rotational unknowns
FEET = (10, 11)


def _ancestors(j):
    out = []
    while PARENT[j] >= 0:
        j = PARENT[j]
        out.append(j)
    return out


def forward_kinematics(R_local, tran=None):
    """articulate/model.py:208-232 with shape=None, calc_mesh=False.  R_local [..., 24, 3, 3] ->
    (global rotations [..., 24, 3, 3], joint positions [..., 24, 3])."""
    R_local = np.asarray(R_local, np.float64)
    G = np.empty_like(R_local)
    P = np.empty(R_local.shape[:-2] + (3,), np.float64)
    G[..., 0, :, :] = R_local[..., 0, :, :]
    P[..., 0, :] = BONE[0]
    for j in range(1, 24):
        p = PARENT[j]
        G[..., j, :, :] = G[..., p, :, :] @ R_local[..., j, :, :]
        P[..., j, :] = P[..., p, :] + G[..., p, :, :] @ BONE[j]
    if tran is not None:
        P = P + np.asarray(tran, np.float64)[..., None, :]
    return G, P


def position_jacobian(G, P):
    """d P_j / d dtheta_k for right-multiplied increments R_k <- R_k exp([dtheta]x): column (k, a) of joint j is
    G_k[:, a] x (P_j - P_k) when k is a strict ancestor-or-self of j.  -> [..., 24, 3, 45]."""
    J = np.zeros(P.shape[:-2] + (24, 3, 3 * len(OPT_JOINTS)), np.float64)
    for j in range(24):
        chain = [j] + _ancestors(j)
        for ci, k in enumerate(OPT_JOINTS):
            if k not in chain:
                continue
            r = P[..., j, :] - P[..., k, :]
            for a in range(3):
                J[..., j, :, 3 * ci + a] = np.cross(G[..., k, :, a], r)
    return J


def exp_so3(w):
    """Rodrigues: [..., 3] -> [..., 3, 3]."""
    w = np.asarray(w, np.float64)
    th = np.linalg.norm(w, axis=-1)[..., None, None]
    K = np.zeros(w.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -w[..., 2], w[..., 1]
    K[..., 1, 0], K[..., 1, 2] = w[..., 2], -w[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -w[..., 1], w[..., 0]
    small = th < 1e-8
    ths = np.where(small, 1.0, th)
    A = np.where(small, 1.0 - th * th / 6.0, np.sin(ths) / ths)
    Bc = np.where(small, 0.5 - th * th / 24.0, (1.0 - np.cos(ths)) / (ths * ths))
    return np.eye(3) + A * K + Bc * (K @ K)


def prob_to_weight(p):
    lo, hi = PROB_THRESHOLD
    return (np.clip(p, lo, hi) - lo) / (hi - lo)


class PhysicsOptimizerPort:
    """Batched (B skeletons side by side, frames in sequence) float64 statement of the optimizer."""

    def __init__(self, B=1, w_vel=1.0, w_contact=10.0, damping=1.0, damping_abs=1e-2, fps=datasets.fps, floor_y=FLOOR_Y):
        self.B, self.w_vel, self.w_contact, self.damping, self.fps, self.floor_y = B, w_vel, w_contact, damping, fps, floor_y
        self.damping_abs = damping_abs
        self.reset_states()

    def reset_states(self):
        self.p = np.zeros((self.B, 3))
        self.q_prev = np.zeros((self.B, 24, 3))
        self.started = np.zeros(self.B, bool)

    def normal_equations(self, R, jvel, contact):
        """-> (H [B,48,48], g [B,48]) of the frame's least-squares problem (unknowns: 45 rotations then d)."""
        G, P = forward_kinematics(R)
        J = position_jacobian(G, P)                                          # [B,24,3,45]
        world = self.p[:, None, :] + P
        e = self.q_prev + np.asarray(jvel, np.float64) / self.fps - world    # [B,24,3]
        wc = self.w_contact * prob_to_weight(1.0 / (1.0 + np.exp(-np.asarray(contact, np.float64))))   # [B,2]
        rows_A, rows_b, rows_w = [], [], []
        eye = np.broadcast_to(np.eye(3), (self.B, 3, 3))
        for j in range(24):
            rows_A.append(np.concatenate([J[:, j], eye], axis=-1)); rows_b.append(e[:, j])
            rows_w.append(np.full((self.B, 3), self.w_vel))
        for f, jf in enumerate(FEET):
            rows_A.append(np.concatenate([J[:, jf], eye], axis=-1)); rows_b.append(self.q_prev[:, jf] - world[:, jf])
            rows_w.append(np.repeat(wc[:, f:f + 1], 3, axis=1))
        A = np.concatenate(rows_A, axis=1)                                   # [B,78,48]
        b = np.concatenate(rows_b, axis=1)                                   # [B,78]
        w = np.concatenate(rows_w, axis=1)
        H = np.einsum('bri,br,brj->bij', A, w, A)
        idx = np.arange(45)
        H[:, idx, idx] = H[:, idx, idx] * (1.0 + self.damping) + self.damping_abs
        g = np.einsum('bri,br,br->bi', A, w, b)
        return H, g

    def optimize_frames(self, R, jvel, contact, active=None):
        """One frame for all B skeletons: R [B,24,3,3], jvel [B,24,3] (m/s), contact [B,2] logits ->
        (pose_opt [B,24,3,3], tran [B,3]).  `active` [B] bool: skeletons past their length are left untouched."""
        R = np.asarray(R, np.float64).reshape(self.B, 24, 3, 3)
        jvel = np.asarray(jvel, np.float64).reshape(self.B, 24, 3)
        contact = np.asarray(contact, np.float64).reshape(self.B, 2)
        active = np.ones(self.B, bool) if active is None else np.asarray(active, bool)
        H, g = self.normal_equations(R, jvel, contact)
        x = np.linalg.solve(H, g[..., None])[..., 0]
        x = np.where(self.started[:, None], x, 0.0)                          # first frame: pass through
        R_new = R.copy()
        for ci, k in enumerate(OPT_JOINTS):
            R_new[:, k] = R[:, k] @ exp_so3(x[:, 3 * ci:3 * ci + 3])
        d = x[:, 45:48].copy()
        _, P_new = forward_kinematics(R_new)
        foot_y = np.minimum(P_new[:, 10, 1], P_new[:, 11, 1]) + self.p[:, 1] + d[:, 1]
        d[:, 1] += np.where(foot_y < self.floor_y, self.floor_y - foot_y, 0.0)
        p_new = self.p + d
        q_new = p_new[:, None, :] + P_new
        self.p = np.where(active[:, None], p_new, self.p)
        self.q_prev = np.where(active[:, None, None], q_new, self.q_prev)
        self.started = self.started | active
        return R_new, self.p.copy()

    def optimize_sequences(self, pose, vel72, contact, lengths=None, vel_scale=2.0):
        """pose [B,T,24,3,3], vel72 [B,T,72] RAW velocity-head output (scaled by vel_scale here, net.py:162),
        contact [B,T,2] -> (pose_opt [B,T,24,3,3], tran [B,T,3]); frames past a sequence's length pass through."""
        pose = np.asarray(pose, np.float64)
        B, T = pose.shape[:2]
        assert B == self.B
        lengths = np.full(B, T) if lengths is None else np.asarray(lengths)
        out_pose = pose.reshape(B, T, 24, 3, 3).copy()
        out_tran = np.zeros((B, T, 3))
        jv = np.asarray(vel72, np.float64).reshape(B, T, 24, 3) * vel_scale
        for t in range(T):
            act = t < lengths
            Rn, p = self.optimize_frames(out_pose[:, t], jv[:, t], np.asarray(contact)[:, t], act)
            out_pose[:, t] = np.where(act[:, None, None, None], Rn, out_pose[:, t])
            out_tran[:, t] = p
        return out_pose, out_tran
