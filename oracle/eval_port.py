"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the translation-error windows of the reference evaluator
(mobileposer/evaluate.py:66-92) and of its mesh row (linear blend skinning, articulate/model.py:208-240;
vertex error, articulate/evaluator.py:319-323,336; row N1).  Pinned to the live reference by oracle/make_golden_eval.py ->
tests/golden/tran_windows.npz, tests/golden/mesh_unit.npz (tests/test_evaluate.py).  Never imported by the product path.

For one sequence with ground-truth root translation `tran_t` [T, 3] and predicted `tran_p` [T, 3]:
  * `move[j+1] = move[j] + |tran_t[j+1] - tran_t[j]|`, accumulated sequentially in fp32             (evaluate.py:68-71)
  * for each window w = 1..7 m a two-pointer sweep collects the (start, end) pairs over which the ground truth moves
    at least w metres -- for every distinct `end` only the first `start` that reaches it            (evaluate.py:73-83)
  * per pair `|dt - dp| / (move[end] - move[start]) * w`, and the mean over the pairs (none -> the window is skipped
    for this sequence)                                                                              (evaluate.py:85-92)
"""
from __future__ import annotations

import numpy as np

WINDOWS = (1, 2, 3, 4, 5, 6, 7)


def move_distance(tran_t: np.ndarray) -> np.ndarray:
    t = np.asarray(tran_t, np.float32)
    d = t[1:] - t[:-1]
    v = np.sqrt((d * d).sum(axis=1, dtype=np.float32)).astype(np.float32)
    move = np.zeros(t.shape[0], np.float32)
    for j in range(v.shape[0]):
        move[j + 1] = np.float32(move[j] + v[j])
    return move


def frame_pairs(move: np.ndarray, window: int):
    pairs, start, end, n = [], 0, 1, move.shape[0]
    while end < n:
        if np.float32(move[end] - move[start]) < window:
            end += 1
        else:
            if not pairs or pairs[-1][1] != end:
                pairs.append((start, end))
            start += 1
    return pairs


def tran_window_errors(tran_p: np.ndarray, tran_t: np.ndarray, windows=WINDOWS):
    """-> (mean error per window [len(windows)] float32, NaN where the sequence has no pair; pair counts)."""
    tran_p, tran_t = np.asarray(tran_p, np.float32), np.asarray(tran_t, np.float32)
    move = move_distance(tran_t)
    out = np.full(len(windows), np.nan, np.float32)
    cnt = np.zeros(len(windows), np.int32)
    for k, w in enumerate(windows):
        pairs = frame_pairs(move, w)
        if not pairs:
            continue
        tot = np.float32(0)
        for s, e in pairs:
            d = (tran_t[e] - tran_t[s]) - (tran_p[e] - tran_p[s])
            err = np.float32(np.sqrt(np.float32((d * d).sum(dtype=np.float32)))) / np.float32(move[e] - move[s]) * np.float32(w)
            tot = np.float32(tot + np.float32(err))
        out[k] = tot / np.float32(len(pairs))
        cnt[k] = len(pairs)
    return out, cnt


# ---- mesh row -------------------------------------------------------------------------------------------------------
SMPL_PARENT = (-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21)


def skinned_vertices(pose, tran, rest, weights, joints_zero, parent=SMPL_PARENT):
    """ParametricModel.forward_kinematics(pose, None, tran, calc_mesh=True) (articulate/model.py:208-240), float64, the
    reference's formulation: 4x4 global joint transforms, `T[..., -1:] -= T @ [j; 0]` (model.py:233), per-vertex blend
    `T_vertex = sum_j w[v][j] T[j]` (model.py:234), vertices = T_vertex [rest; 1] (+ tran) (model.py:238-239).
    pose [n,24,3,3] local rotations; rest [V,3] = v_template - J[0]; joints_zero [24,3] = J - J[0].
    -> (joint positions [n,24,3], vertices [n,V,3])."""
    pose = np.asarray(pose, np.float64).reshape(-1, 24, 3, 3)
    n = pose.shape[0]
    j0 = np.asarray(joints_zero, np.float64)
    rest = np.asarray(rest, np.float64)
    w = np.asarray(weights, np.float64)
    T = np.zeros((n, 24, 4, 4))
    T[:, :, 3, 3] = 1.0
    for i in range(24):
        local = np.zeros((n, 4, 4))
        local[:, :3, :3] = pose[:, i]
        local[:, :3, 3] = j0[i] - (j0[parent[i]] if parent[i] >= 0 else 0.0)      # bone vector (model.py:226)
        local[:, 3, 3] = 1.0
        T[:, i] = local if parent[i] < 0 else T[:, parent[i]] @ local
    joint = T[:, :, :3, 3].copy()
    T[:, :, :3, 3] -= np.einsum('njab,jb->nja', T[:, :, :3, :3], j0)
    Tv = np.einsum('njab,vj->nvab', T, w)
    vert = np.einsum('nvab,vb->nva', Tv[:, :, :3, :3], rest) + Tv[:, :, :3, 3]
    if tran is not None:
        t = np.asarray(tran, np.float64).reshape(-1, 1, 3)
        joint, vert = joint + t, vert + t
    return joint, vert


def vertex_error_row(pose_p, pose_t, tran_p, tran_t, rest, weights, joints_zero, align_joint=0):
    """Row 1 of FullMotionEvaluator.__call__ (evaluator.py:319-323,336): [mean, mean over vertices of the std over frames]
    of the root-aligned vertex position error."""
    jp, vp = skinned_vertices(pose_p, tran_p, rest, weights, joints_zero)
    jt, vt = skinned_vertices(pose_t, tran_t, rest, weights, joints_zero)
    off = (jt[:, align_joint] - jp[:, align_joint])[:, None]
    ve = np.linalg.norm(vp + off - vt, axis=2)
    return np.array([ve.mean(), ve.std(axis=0, ddof=1).mean()])
