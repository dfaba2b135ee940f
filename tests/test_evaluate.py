"""CPU suite: the metric side of evaluate.py (SMPL forward kinematics + FullMotionEvaluator rows) against a fixture
produced by the reference's own evaluator (oracle/make_golden.py, case K)."""
import math

import pytest
import torch

from conftest import load_golden
from mobileposer_b200.evaluate import (PoseEvaluator, angle_between, forward_kinematics, full_motion_errors,
                                       r6d_to_rotation_matrix, synthetic_dip)


def test_forward_kinematics_matches_reference():
    g = load_golden('metrics_unit')
    glb, joint = forward_kinematics(g['pose_a'], g['tran_a'])
    assert (glb - g['glb_a']).abs().max() < 1e-5
    assert (joint - g['joint_a']).abs().max() < 1e-5
    # identity pose -> zero-pose skeleton
    from mobileposer_b200.config import SMPL_J_ZERO
    _, j0 = forward_kinematics(torch.eye(3).repeat(1, 24, 1, 1))
    assert (j0[0] - torch.tensor(SMPL_J_ZERO)).abs().max() < 1e-6


def test_full_motion_rows_match_reference_evaluator():
    g = load_golden('metrics_unit')
    errs = full_motion_errors(g['pose_a'], g['pose_b'], g['tran_a'], g['tran_b'])
    ref = g['errs']
    rows = [0, 2, 3, 4, 5, 6, 7, 8, 9]            # row 1 is the mesh-vertex error (needs the SMPL template)
    assert torch.isnan(errs[1]).all()
    rel = ((errs[rows] - ref[rows]).abs() / ref[rows].abs().clamp_min(1e-6)).max().item()
    assert rel < 2e-4, rel


def test_angle_between_and_r6d():
    a = torch.tensor(0.3)
    rz = torch.tensor([[math.cos(a), -math.sin(a), 0.], [math.sin(a), math.cos(a), 0.], [0., 0., 1.]])
    assert abs(angle_between(torch.eye(3)[None], rz[None]).item() - 0.3) < 1e-6
    r = r6d_to_rotation_matrix(torch.tensor([[1., 0., 0., 0., 2., 0.]]))
    assert torch.allclose(r[0], torch.eye(3))


def test_pose_evaluator_rows_and_synthetic_set():
    items = synthetic_dip(n_subjects=1, n_seq=2, frames=64)
    assert len(items) == 2 and items[0][0].shape == (64, 60) and items[0][1].shape == (64, 144)
    pose_t = r6d_to_rotation_matrix(items[0][1]).view(-1, 24, 3, 3)
    rows = PoseEvaluator().eval(pose_t, items[0][1].new_tensor(pose_t), tran_p=items[0][3], tran_t=items[0][3])
    assert rows.shape == (8, 2)
    ok = [0, 1, 2, 3, 4, 7]
    assert rows[ok, 0].abs().max() < 1e-3          # identical motions -> zero errors (mesh row NaN)


def test_tran_window_oracle_matches_reference_evaluate_pose():
    """oracle/eval_port.py (restatement of evaluate.py:66-92) against what the reference's own evaluate_pose(...,
    evaluate_tran=True) printed for the same translations (oracle/make_golden_eval.py)."""
    import numpy as np
    from oracle.eval_port import frame_pairs, move_distance, tran_window_errors
    g = load_golden('tran_windows')
    per_seq = []
    for i, n in enumerate(g['lengths'].tolist()):
        err, cnt = tran_window_errors(g['tran_p'][i, :n].numpy(), g['tran_t'][i, :n].numpy())
        ref = g['per_sequence'][i].numpy()
        assert np.array_equal(np.isnan(err), np.isnan(ref)), (i, err, ref)
        assert np.array_equal(cnt > 0, ~np.isnan(ref))
        ok = ~np.isnan(ref)
        if ok.any():
            assert (np.abs(err[ok] - ref[ok]) / np.abs(ref[ok])).max() < 1e-6
        per_seq.append(err)
    per_seq = np.stack(per_seq)
    # the printed list: per window the mean over the sequences that have a pair (evaluate.py:92,106)
    overall = np.array([np.nanmean(per_seq[:, k]) for k in range(7)], np.float32)
    assert (np.abs(overall - g['overall'].numpy()) / g['overall'].numpy()).max() < 1e-6
    # sweep properties: ends strictly increase, every pair spans at least the window, the shorter span does not
    mv = move_distance(g['tran_t'][0, :g['lengths'][0]].numpy())
    for w in (1, 4):
        pairs = frame_pairs(mv, w)
        assert all(b[1] > a[1] for a, b in zip(pairs, pairs[1:]))
        assert all(mv[e] - mv[s] >= w and mv[e - 1] - mv[s] < w for s, e in pairs)


def test_mesh_oracle_matches_reference_skinning_and_evaluator_row():
    """oracle/eval_port.py (restatement of model.py:208-240 and evaluator.py:319-323) against the reference's own
    forward_kinematics(calc_mesh=True) / FullMotionEvaluator over the synthetic template of oracle/make_golden_eval.py."""
    import numpy as np
    from oracle.eval_port import skinned_vertices, vertex_error_row
    g = {k: v.numpy() for k, v in load_golden('mesh_unit').items()}
    joint, vert = skinned_vertices(g['pose_p'][:3], g['tran_p'][:3], g['rest'], g['weights'], g['joints_zero'])
    assert np.abs(joint - g['joint3']).max() < 1e-6 and np.abs(vert - g['vertex3']).max() < 2e-6
    row = vertex_error_row(g['pose_p'], g['pose_t'], g['tran_p'], g['tran_t'], g['rest'], g['weights'], g['joints_zero'])
    assert (np.abs(row - g['errs'][1]) / g['errs'][1]).max() < 1e-5
    # the template's joints are the SMPL constants the library carries
    from mobileposer_b200.config import SMPL_J_ZERO
    assert np.abs(g['joints_zero'] - np.asarray(SMPL_J_ZERO, np.float32)).max() < 1e-7


def test_smpl_file_loader_reads_a_model_file_without_chumpy(tmp_path):
    """load_smpl_mesh: a pickle shaped like the official model file (a chumpy object inside) -> rest vertices and weights."""
    import pickle
    import sys
    import types

    import numpy as np
    from mobileposer_b200.evaluate import load_smpl_mesh
    mod = types.ModuleType('chumpy')

    class Ch:                                    # what the official file pickles `shapedirs` as
        def __init__(self, x=None):
            self.x = x
    Ch.__module__, Ch.__qualname__ = 'chumpy', 'Ch'
    mod.Ch = Ch
    rng = np.random.default_rng(0)
    data = {'v_template': rng.normal(size=(50, 3)), 'J': rng.normal(size=(24, 3)), 'weights': rng.random((50, 24)),
            'shapedirs': Ch(rng.normal(size=(50, 3, 10)))}
    sys.modules['chumpy'] = mod
    try:
        blob = pickle.dumps(data, protocol=2)
    finally:
        del sys.modules['chumpy']
    f = tmp_path / 'model.pkl'
    f.write_bytes(blob)
    rest, w = load_smpl_mesh(str(f), device='cpu')
    assert rest.shape == (50, 3) and w.shape == (50, 24) and rest.dtype == torch.float32
    assert np.allclose(rest.numpy(), (data['v_template'] - data['J'][:1]).astype(np.float32))


class _StandInNet(torch.nn.Module):
    """What evaluate_pose touches of a model, on the CPU: a pose / translation that depend on the sequence's own frames only
    (so a padded batch and a loop over single sequences must agree), with MobilePoserNet.forward_offline's return shapes."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))
        self.calls = []

    def reset(self):
        pass

    @staticmethod
    def _one(x):
        a = 0.3 * torch.tanh(x[:, :24])                              # one angle per joint, about z
        c, s_, z, o = torch.cos(a), torch.sin(a), torch.zeros_like(a), torch.ones_like(a)
        pose = torch.stack([c, -s_, z, s_, c, z, z, z, o], dim=-1).view(-1, 24, 3, 3)
        return pose, torch.cumsum(0.01 * x[:, 24:27], dim=0)

    def forward_offline(self, x, lengths):
        B, T = x.shape[0], x.shape[1]
        self.calls.append(list(lengths))
        assert T == max(lengths)
        pose, tran = torch.zeros(B, T, 24, 3, 3), torch.zeros(B, T, 3)
        for b, n in enumerate(lengths):
            pose[b, :n], tran[b, :n] = self._one(x[b, :n])
            assert (x[b, n:] == 0).all()                              # padding is zero-filled
        if B == 1:
            return pose[0], torch.zeros(1, T, 72), tran[0], torch.zeros(T, 2)
        return pose.view(B * T, 24, 3, 3), torch.zeros(B, T, 72), tran, torch.zeros(B, T, 2)


def test_evaluate_pose_batched_groups_equal_the_sequence_loop():
    """Host logic of evaluate_pose(batch_size=...): grouping in dataset order, padding to the longest, true lengths,
    per-sequence slices -- against the reference-shaped loop, with a stand-in model on the CPU."""
    from mobileposer_b200.evaluate import evaluate_pose
    g = torch.Generator().manual_seed(9)
    items = []
    for n in (70, 64, 90, 75, 66, 81, 64):
        pose_r6d = torch.tensor([1., 0., 0., 0., 1., 0.]).repeat(24) + 0.2 * torch.randn(n, 144, generator=g)
        items.append((torch.randn(n, 60, generator=g), pose_r6d, torch.zeros(n, 24, 3), torch.cumsum(0.01 * torch.randn(n, 3, generator=g), 0)))
    loop_net, batch_net = _StandInNet(), _StandInNet()
    ref = evaluate_pose(loop_net, items, verbose=False)
    out = evaluate_pose(batch_net, items, verbose=False, batch_size=3)
    assert loop_net.calls == [[n] for n in (70, 64, 90, 75, 66, 81, 64)]
    assert batch_net.calls == [[70, 64, 90], [75, 66, 81], [64]]
    ok = ~torch.isnan(ref)
    assert torch.equal(torch.isnan(out), torch.isnan(ref))
    assert torch.allclose(out[ok], ref[ok], rtol=1e-5, atol=1e-7)
    with pytest.raises(ValueError):
        evaluate_pose(batch_net, items, verbose=False, batch_size=0)
