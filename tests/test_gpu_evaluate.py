"""GPU suite: the evaluator rows that run on the device beyond the per-frame errors -- the translation-error windows of
`evaluate_pose(..., evaluate_tran=True)` (mp_eval_tran_windows) against the reference's own output
(tests/golden/tran_windows.npz) and the CPU restatement (oracle/eval_port.py)."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

DEV = 'cuda:0'


def test_tran_windows_match_reference_evaluate_pose():
    from mobileposer_b200.evaluate import tran_window_errors
    g = load_golden('tran_windows')
    lengths = g['lengths'].tolist()
    ref = g['per_sequence']
    # one launch over the padded batch ...
    err, cnt = tran_window_errors(g['tran_p'].to(DEV), g['tran_t'].to(DEV), lengths)
    err, cnt = err.cpu(), cnt.cpu()
    assert torch.equal(torch.isnan(err), torch.isnan(ref))
    assert torch.equal(cnt > 0, ~torch.isnan(ref))
    ok = ~torch.isnan(ref)
    assert ((err[ok] - ref[ok]).abs() / ref[ok].abs()).max() < 2e-6
    # ... equals one launch per sequence (what evaluate_pose does), bit for bit
    for i, n in enumerate(lengths):
        e1, c1 = tran_window_errors(g['tran_p'][i, :n].to(DEV), g['tran_t'][i, :n])     # tran_t arrives on the CPU like the dataset's
        assert torch.equal(torch.nan_to_num(e1[0].cpu(), nan=-1.0), torch.nan_to_num(err[i], nan=-1.0))
        assert torch.equal(c1[0].cpu(), cnt[i])


@pytest.mark.parametrize('T,step', [(3000, 0.004), (3000, 0.02), (2, 5.0), (1, 1.0), (20000, 0.0007)])
def test_tran_windows_match_oracle_pair_for_pair(T, step):
    """DIP-sized sequences: the pair COUNTS equal the oracle's (the sequential fp32 distance reproduces every comparison of
    the two-pointer sweep) and the means agree to fp32 round-off."""
    from mobileposer_b200.evaluate import tran_window_errors
    from oracle.eval_port import tran_window_errors as oracle_windows
    g = torch.Generator().manual_seed(T + int(step * 1e4))
    tran_t = torch.cumsum(torch.randn(T, 3, generator=g).abs() * step, 0)
    tran_p = tran_t + torch.cumsum(torch.randn(T, 3, generator=g) * step * 0.2, 0)
    err, cnt = tran_window_errors(tran_p.to(DEV), tran_t.to(DEV))
    o_err, o_cnt = oracle_windows(tran_p.numpy(), tran_t.numpy())
    assert np.array_equal(cnt[0].cpu().numpy(), o_cnt)
    e = err[0].cpu().numpy()
    assert np.array_equal(np.isnan(e), np.isnan(o_err))
    ok = ~np.isnan(o_err)
    if ok.any():
        assert (np.abs(e[ok] - o_err[ok]) / np.abs(o_err[ok])).max() < 5e-6


def test_mesh_row_matches_reference_evaluator():
    """mp_eval_vertex_errors behind full_motion_errors(mesh=...) against the reference's FullMotionEvaluator row 1 over
    the synthetic template (tests/golden/mesh_unit.npz), and the other rows next to it."""
    from mobileposer_b200.evaluate import full_motion_errors, vertex_error_row
    g = load_golden('mesh_unit')
    mesh = (g['rest'].to(DEV), g['weights'].to(DEV))
    errs = full_motion_errors(g['pose_p'].to(DEV), g['pose_t'].to(DEV), g['tran_p'].to(DEV), g['tran_t'].to(DEV), mesh=mesh).cpu()
    ref = g['errs']
    assert ((errs - ref).abs() / ref.abs().clamp_min(1e-6)).max() < 2e-4, (errs, ref)
    assert ((errs[1] - ref[1]).abs() / ref[1]).max() < 2e-5, (errs[1], ref[1])
    # identical motions -> exactly zero; a single frame -> std 0 like the other rows of this mirror
    z = vertex_error_row(g['pose_p'].to(DEV), g['pose_p'].to(DEV), mesh).cpu()
    assert z.abs().max() == 0.0
    one = vertex_error_row(g['pose_p'][:1].to(DEV), g['pose_t'][:1].to(DEV), mesh).cpu()
    assert one[0] > 0 and one[1] == 0


@pytest.mark.parametrize('n,V', [(3000, 6890), (33, 513), (1, 1)])
def test_mesh_row_matches_oracle_at_smpl_size(n, V):
    """SMPL-sized template (6890 vertices, <= 4 bones per vertex) x a DIP-sized sequence: the fused kernel against the
    float64 restatement of the reference's skinning on a sample of frames (exact per-frame equality of the formulation)
    and against its own chunking (frame chunks / vertex tiles change nothing but the summation order)."""
    from mobileposer_b200.config import SMPL_J_ZERO
    from mobileposer_b200.evaluate import vertex_error_row
    from oracle.eval_port import vertex_error_row as oracle_row
    g = torch.Generator().manual_seed(n + V)
    jz = torch.tensor(SMPL_J_ZERO)
    near = torch.randint(0, 24, (V,), generator=g)
    rest = jz[near] + torch.randn(V, 3, generator=g) * 0.05
    w = torch.zeros(V, 24)
    w[torch.arange(V), near] = 1.0
    extra = torch.randint(0, 24, (V, 3), generator=g)
    w.scatter_add_(1, extra, torch.rand(V, 3, generator=g) * 0.5)
    w = w / w.sum(1, keepdim=True)

    def rots(k, scale):
        a = torch.randn(k, 3, generator=g) * scale
        th = a.norm(dim=1, keepdim=True).clamp_min(1e-8)
        u = a / th
        K = torch.zeros(k, 3, 3)
        K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -u[:, 2], u[:, 1], u[:, 2], -u[:, 0], -u[:, 1], u[:, 0]
        return torch.eye(3) + torch.sin(th)[:, :, None] * K + (1 - torch.cos(th))[:, :, None] * (K @ K)
    pose_t = rots(n * 24, 0.5).view(n, 24, 3, 3)
    pose_p = pose_t @ rots(n * 24, 0.1).view(n, 24, 3, 3)
    row = vertex_error_row(pose_p.to(DEV), pose_t.to(DEV), (rest, w)).cpu().double().numpy()
    sub = slice(0, n, max(1, n // 40))                       # the oracle materialises [n, V, 3]: a sample of frames
    o_sub = oracle_row(pose_p[sub].numpy(), pose_t[sub].numpy(), None, None, rest.numpy(), w.numpy(), jz.numpy())
    k_sub = vertex_error_row(pose_p[sub].contiguous().to(DEV), pose_t[sub].contiguous().to(DEV), (rest, w)).cpu().double().numpy()
    assert abs(k_sub[0] - o_sub[0]) / o_sub[0] < 2e-5
    if pose_p[sub].shape[0] > 1 and V > 1:
        assert abs(k_sub[1] - o_sub[1]) / max(o_sub[1], 1e-9) < 2e-4
    assert np.isfinite(row).all() and row[0] > 0
    if n > 1000:                                             # the full sequence is statistically the same motion as its sample
        assert abs(row[0] - o_sub[0]) / o_sub[0] < 0.1


def test_tran_windows_reject_cpu_tensors_and_oversized_sequences():
    from mobileposer_b200.evaluate import tran_window_errors
    with pytest.raises(RuntimeError):
        tran_window_errors(torch.zeros(10, 3), torch.zeros(10, 3))
    with pytest.raises(RuntimeError):
        tran_window_errors(torch.zeros(60000, 3, device=DEV), torch.zeros(60000, 3, device=DEV))


def test_evaluate_pose_with_translation_windows(capsys):
    from mobileposer_b200.evaluate import evaluate_pose, synthetic_dip, tran_window_errors
    from mobileposer_b200.net import MobilePoserNet
    torch.manual_seed(0)
    net = MobilePoserNet().to(DEV).eval()
    items = synthetic_dip(n_subjects=1, n_seq=3, frames=200)
    # make the ground truth travel: 4 m, 1.5 m and 0 m over the three sequences
    for k, dist in enumerate((4.0, 1.5, 0.0)):
        imu, pose, joint, tran = items[k]
        items[k] = (imu, pose, joint, torch.linspace(0, dist, 200).view(-1, 1) * torch.tensor([[1.0, 0.0, 0.5]]))
    table, windows = evaluate_pose(net, items, evaluate_tran=True)
    assert table.shape == (3, 8, 2) and windows.shape == (3, 7)
    assert torch.isfinite(windows[0, :3]).all() and torch.isnan(windows[0, 6])
    assert torch.isfinite(windows[1, 0]) and torch.isnan(windows[1, 1:]).all() and torch.isnan(windows[2]).all()
    out = capsys.readouterr().out
    assert '============== offline ================' in out and '[0, tensor(' in out
    # the rows are what a direct call on the same forward gives
    net.reset()
    net.velocity.rnn_state = None          # sequence 0 was the first of the loop: fresh velocity state (F5)
    _, _, tran_p, _ = net.forward_offline(items[0][0].to(DEV).unsqueeze(0), [200])
    e, _ = tran_window_errors(tran_p, items[0][3])
    assert torch.equal(torch.nan_to_num(e[0], nan=-1.0), torch.nan_to_num(windows[0], nan=-1.0))
