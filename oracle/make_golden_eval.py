"""TEST INFRASTRUCTURE ONLY -- golden vectors of the reference's TRANSLATION-ERROR WINDOWS and of the
evaluator's MESH ROW (row N1; see mesh_golden below).

    python oracle/make_golden_eval.py        # writes tests/golden/tran_windows.npz and tests/golden/mesh_unit.npz

Runs only in the build container (needs /root/reference).  Calls the UNMODIFIED `evaluate_pose(model, dataset,
evaluate_tran=True)` of mobileposer/evaluate.py:39-107 through the shims of make_golden.py, with a stand-in model
whose `forward_offline` returns prescribed translations (the windows only look at `tran_p_offline` and the dataset's
`tran_t`, evaluate.py:66-92).  The function only prints its result (`print([0] + [mean over sequences per window])`,
evaluate.py:106); the module-level `print` is replaced by a recorder so the full-precision tensors are captured.
It is run once per single-sequence dataset (per-sequence rows) and once over all sequences (the printed means).
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import make_golden as MG          # sets sys.path for the reference + shims

import numpy as np
import torch


def walk(seed, T, step, noise):
    """A ground-truth root path that covers `step` metres per frame on average, and a prediction that drifts off it."""
    g = torch.Generator().manual_seed(seed)
    heading = torch.cumsum(torch.randn(T, generator=g) * 0.05, 0)
    speed = step * (1.0 + 0.5 * torch.sin(torch.arange(T) / 17.0)) * (torch.rand(T, generator=g) > 0.1)   # some still frames
    vel = torch.stack([speed * torch.cos(heading), 0.01 * torch.randn(T, generator=g), speed * torch.sin(heading)], 1)
    tran_t = torch.cumsum(vel, 0)
    tran_p = tran_t * (1.0 + 0.1 * torch.randn(1, generator=g)) + torch.cumsum(torch.randn(T, 3, generator=g) * noise, 0)
    return tran_t.float(), tran_p.float()


class StandIn:
    """What evaluate_pose touches of a model: eval(), reset(), forward_offline(x, lengths)."""

    def __init__(self, trans):
        self.trans, self.k = trans, 0

    def eval(self):
        return self

    def reset(self):
        pass

    def forward_offline(self, x, lengths):
        T = x.shape[1]
        tran_p = self.trans[self.k]
        self.k += 1
        eye = torch.eye(3).repeat(T, 24, 1, 1)
        return eye, torch.zeros(1, T, 72), tran_p, torch.zeros(T, 2)


@torch.no_grad()
def main():
    cwd = os.getcwd()
    os.chdir(os.path.join(MG.REF, 'mobileposer'))
    try:
        import mobileposer.evaluate as RE
        cases = [walk(1, 400, 0.03, 0.004), walk(2, 523, 0.012, 0.002), walk(3, 260, 0.008, 0.003), walk(4, 90, 0.002, 0.001)]
        eye6 = torch.tensor([1., 0., 0., 0., 1., 0.]).repeat(24)
        items = [(torch.zeros(t.shape[0], 60), eye6.repeat(t.shape[0], 1), torch.zeros(t.shape[0], 24, 3), t) for t, _ in cases]
        record = []
        RE.print = lambda *a, **k: record.append(a)          # evaluate.py only prints its results
        RE.tqdm.tqdm = lambda it, *a, **k: it

        def run(idx):
            record.clear()
            RE.evaluate_pose(StandIn([cases[i][1] for i in idx]), [items[i] for i in idx], evaluate_tran=True)
            lists = [a[0] for a in record if len(a) == 1 and isinstance(a[0], list)]
            assert len(lists) == 1 and len(lists[0]) == 8 and lists[0][0] == 0
            return torch.stack([torch.as_tensor(v, dtype=torch.float32) for v in lists[0][1:]])

        per_seq = torch.stack([run([i]) for i in range(len(cases))])      # [S, 7]; NaN = no pair of that window (mean of an empty list)
        overall = run(list(range(len(cases))))                             # [7]
    finally:
        os.chdir(cwd)
    T = max(t.shape[0] for t, _ in cases)
    tran_t = torch.zeros(len(cases), T, 3)
    tran_p = torch.zeros(len(cases), T, 3)
    for i, (t, p) in enumerate(cases):
        tran_t[i, :t.shape[0]], tran_p[i, :p.shape[0]] = t, p
    print(per_seq)
    print(overall)
    MG.save('tran_windows', tran_t=tran_t, tran_p=tran_p, lengths=np.array([t.shape[0] for t, _ in cases], np.int32),
            per_sequence=per_seq, overall=overall)
    mesh_golden()


@torch.no_grad()
def mesh_golden():
    """Mesh row (evaluator.py row 1): the reference's own FullMotionEvaluator and
    ParametricModel.forward_kinematics(calc_mesh=True), unmodified, over a SYNTHETIC template -- the SMPL vertices and
    skinning weights cannot be committed, the skinning code does not care which mesh it skins.  Only `_v_template` and
    `_skinning_weights` of the loaded model are replaced (700 random vertices around the skeleton, weights with 1-4 and a
    few with 24 non-zeros); joints, parents and every line of code are the reference's."""
    cwd = os.getcwd()
    os.chdir(os.path.join(MG.REF, 'mobileposer'))
    try:
        import mobileposer.articulate as art
        import mobileposer.config as RC
        ev = art.FullMotionEvaluator(str(RC.paths.smpl_file), joint_mask=torch.tensor([2, 5, 16, 20]), fps=RC.datasets.fps)
    finally:
        os.chdir(cwd)
    g = torch.Generator().manual_seed(4242)
    V, n = 700, 75
    J = ev.model._J.clone()
    near = torch.randint(0, 24, (V,), generator=g)
    v_template = J[near] + torch.randn(V, 3, generator=g) * 0.06
    w = torch.zeros(V, 24)
    for v in range(V):
        k = 24 if v % 50 == 0 else int(torch.randint(1, 5, (1,), generator=g))
        idx = torch.randperm(24, generator=g)[:k]
        idx[0] = near[v]
        w[v, idx] = torch.rand(k, generator=g) + 0.05
    w = w / w.sum(dim=1, keepdim=True)
    ev.model._v_template = v_template
    ev.model._skinning_weights = w

    def rand_rot(k, scale):
        a = torch.randn(k, 3, generator=g) * scale
        return art.math.axis_angle_to_rotation_matrix(a).view(k, 3, 3)
    pose_t = rand_rot(n * 24, 0.6).view(n, 24, 3, 3)
    pose_p = torch.matmul(pose_t, rand_rot(n * 24, 0.15).view(n, 24, 3, 3))
    tran_t = torch.cumsum(torch.randn(n, 3, generator=g) * 0.02, 0)
    tran_p = tran_t + torch.randn(n, 3, generator=g) * 0.05
    errs = ev(pose_p, pose_t, tran_p=tran_p, tran_t=tran_t)
    _, joint, vertex = ev.model.forward_kinematics(pose_p[:3], None, tran_p[:3], calc_mesh=True)
    print(errs[:2])
    MG.save('mesh_unit', pose_p=pose_p, pose_t=pose_t, tran_p=tran_p, tran_t=tran_t, rest=v_template - J[:1], weights=w,
            joints_zero=J - J[:1], errs=errs, vertex3=vertex, joint3=joint)


if __name__ == '__main__':
    main()
