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


class _OracleNet(torch.nn.Module):
    """The CPU oracle behind the surface evaluate_pose drives (reset / forward_offline / forward_online), so the very same
    evaluate_pose code walks the reference's restatement and the CUDA net."""

    def __init__(self, state_dict):
        super().__init__()
        from oracle.torch_port import OraclePoser
        self.w = torch.nn.Parameter(torch.zeros(1))
        self.o = o = OraclePoser(state_dict)

        class _Velocity:                     # evaluate_pose clears `model.velocity.rnn_state` in sharded / batched runs
            rnn_state = property(lambda self: o.vel_state, lambda self, v: setattr(o, 'vel_state', v))
        self.velocity = _Velocity()

    def reset(self):
        self.o.reset()

    def forward_offline(self, x, lengths):
        return self.o.forward_offline(x, lengths)

    def forward_online(self, f):
        return self.o.forward_online(f)


def test_evaluate_pose_online_against_the_oracle(wc_state_dict, monkeypatch):
    """ONLINE=1 (evaluate.py:62-64,97-99): per-tick poses are kept across ticks and stacked -- every tick must hand out its own
    tensor (the reference does) -- and the offline + online tables equal the ones the same loop produces over the CPU oracle,
    including the velocity-state chain from the offline call into the ticks and across sequences."""
    import mobileposer_b200 as mp
    from mobileposer_b200.evaluate import evaluate_pose, synthetic_dip
    monkeypatch.setenv('ONLINE', '1')
    items = synthetic_dip(n_subjects=1, n_seq=2, frames=48)
    net = mp.MobilePoserNet()
    net.load_state_dict(wc_state_dict)
    net = net.to(DEV).eval()
    table, online = evaluate_pose(net, items, verbose=False, return_online=True)
    o_table, o_online = evaluate_pose(_OracleNet(wc_state_dict), items, verbose=False, return_online=True)
    rows = [0, 1, 2, 3, 4, 7]                      # all but the mesh row (NaN without a template) and jitter (below)
    for got, want, what in ((table, o_table, 'offline'), (online, o_online, 'online')):
        got = got.cpu()
        assert torch.isfinite(got[:, rows]).all(), what
        assert torch.allclose(got[:, rows], want[:, rows], rtol=2e-3, atol=2e-3), (what, (got[:, rows] - want[:, rows]).abs().max())
        # jitter = third differences x fps^3: amplifies fp32 round-off of the joint positions by 2.7e4
        assert torch.allclose(got[:, 6], want[:, 6], rtol=5e-2, atol=1e-3), what
    # the online rows are not the last tick repeated: poses differ along the sequence
    ticks = [net.forward_online(f)[0] for f in items[0][0][:6].to(DEV)]
    assert not torch.equal(ticks[0], ticks[-1]) and ticks[0].data_ptr() != ticks[-1].data_ptr()


def test_evaluate_pose_batched_on_the_gpu_equals_the_loop_with_fresh_state(wc_state_dict):
    """evaluate_pose(batch_size=3) on the device: ragged groups through batched forward_offline (throughput kernels) give the
    rows of independent sequences -- the loop with the velocity state cleared before every call."""
    import mobileposer_b200 as mp
    from mobileposer_b200.evaluate import evaluate_pose, synthetic_dip
    items = [it for n in (70, 64, 90, 75) for it in synthetic_dip(n_subjects=1, n_seq=1, frames=n)]
    items = [(imu + 0.001 * k, *rest) for k, (imu, *rest) in enumerate(items)]
    net = mp.MobilePoserNet()
    net.load_state_dict(wc_state_dict)
    net = net.to(DEV).eval()
    batched = evaluate_pose(net, items, verbose=False, batch_size=3).cpu()
    rows = []
    for it in items:
        net.velocity.rnn_state = None
        rows.append(evaluate_pose(net, [it], verbose=False)[0].cpu())
    net.velocity.rnn_state = None
    loop = torch.stack(rows)
    keep = [0, 1, 2, 3, 4, 7]
    assert torch.allclose(batched[:, keep], loop[:, keep], rtol=2e-3, atol=2e-3)
    assert torch.allclose(batched[:, 6], loop[:, 6], rtol=5e-2, atol=1e-3)


@pytest.mark.parametrize('wrapped', [False, True])
def test_load_model_reads_both_on_disk_formats(tmp_path, wc_state_dict, wrapped):
    """utils/model_utils.py:6-15: a plain state_dict `.pth` (what combine_weights.py writes) and a Lightning-style checkpoint
    with the weights under 'state_dict'; the loaded model computes with those weights."""
    import mobileposer_b200 as mp
    from mobileposer_b200.model_utils import load_model
    from mobileposer_b200.synthetic import synthetic_imu
    path = str(tmp_path / ('ckpt.pth' if wrapped else 'weights.pth'))
    torch.save({'state_dict': wc_state_dict, 'epoch': 3} if wrapped else wc_state_dict, path)
    model = load_model(path)
    assert isinstance(model, mp.MobilePoserNet) and next(model.parameters()).is_cuda
    for k, v in model.state_dict().items():
        assert torch.equal(v.cpu(), wc_state_dict[k]), k
    ref = mp.MobilePoserNet()
    ref.load_state_dict(wc_state_dict)
    ref = ref.to(DEV).eval()
    x = synthetic_imu(5, 40)[None].to(DEV)
    a, b = model.forward_offline(x, [40]), ref.forward_offline(x, [40])
    assert all(torch.equal(p, q) for p, q in zip(a, b))


def test_weights_changed_in_place_invalidate_the_packed_net(seeded_state_dict, wc_state_dict):
    """load_state_dict after a forward: the packed heads are rebuilt (a new mp_rnn may land at the freed handle's address),
    and the mp_net with its captured graphs must be rebuilt with them -- keyed on (handle, generation)."""
    import mobileposer_b200 as mp
    from mobileposer_b200.synthetic import synthetic_imu
    x = synthetic_imu(6, 32)[None].to(DEV)
    net = mp.MobilePoserNet()
    net.load_state_dict(seeded_state_dict)
    net = net.to(DEV).eval()
    for _ in range(3):                               # eager, capture, replay
        net.velocity.rnn_state = None
        before = [t.clone() for t in net.forward_offline(x, [32])]
    net.load_state_dict(wc_state_dict)
    for _ in range(3):
        net.velocity.rnn_state = None
        after = [t.clone() for t in net.forward_offline(x, [32])]
    fresh = mp.MobilePoserNet()
    fresh.load_state_dict(wc_state_dict)
    fresh = fresh.to(DEV).eval()
    want = fresh.forward_offline(x, [32])
    assert not torch.equal(before[0], after[0])
    assert all(torch.equal(p, q) for p, q in zip(after, want))


@pytest.mark.parametrize('n', [1, 2, 3, 4, 30, 31, 90])
def test_motion_rows_kernel_equals_the_torch_statement(n):
    """mp_eval_motion_rows (one launch for the ten (mean, std) rows) against the torch statement of the same rows on the CPU
    (itself pinned to the reference's evaluator in tests/test_evaluate.py), including the empty cases: fewer than 4 frames (no
    jitter), no more than fps frames (no one-second translation error), a single frame (std = 0)."""
    from mobileposer_b200.evaluate import full_motion_errors, r6d_to_rotation_matrix
    g = torch.Generator().manual_seed(100 + n)
    eye6 = torch.tensor([1., 0., 0., 0., 1., 0.]).repeat(24)
    pose_p = r6d_to_rotation_matrix(eye6 + 0.3 * torch.randn(n, 144, generator=g)).view(n, 24, 3, 3)
    pose_t = r6d_to_rotation_matrix(eye6 + 0.3 * torch.randn(n, 144, generator=g)).view(n, 24, 3, 3)
    tran_p, tran_t = torch.randn(n, 3, generator=g) * 0.1, torch.randn(n, 3, generator=g) * 0.1
    want = full_motion_errors(pose_p, pose_t, tran_p, tran_t)
    got = full_motion_errors(pose_p.to(DEV), pose_t.to(DEV), tran_p.to(DEV), tran_t.to(DEV)).cpu()
    assert torch.equal(torch.isnan(got), torch.isnan(want)), (got, want)
    ok = ~torch.isnan(want)
    assert ((got[ok] - want[ok]).abs() / want[ok].abs().clamp_min(1e-3)).max() < 2e-4


def test_eval_group_equals_per_sequence_eval_bit_for_bit():
    """PoseEvaluator.eval_group (one pass over the concatenated frames of a group + mp_eval_motion_rows_batch, one CTA per sequence)
    returns exactly the rows PoseEvaluator.eval returns for each sequence alone -- ragged lengths, a sequence shorter than one second."""
    from mobileposer_b200.evaluate import PoseEvaluator, r6d_to_rotation_matrix
    gen = torch.Generator().manual_seed(8)
    lens = [211, 40, 3000, 17, 333]
    eye6 = torch.tensor([1., 0., 0., 0., 1., 0.]).repeat(24)
    ev = PoseEvaluator()
    pose_p, gt, tran_p, tran_t = [], [], [], []
    for n in lens:
        pose_p.append(r6d_to_rotation_matrix(eye6 + 0.3 * torch.randn(n, 144, generator=gen)).view(n, 24, 3, 3).to(DEV))
        gt.append(eye6 + 0.3 * torch.randn(n, 144, generator=gen))
        tran_p.append(torch.cumsum(0.01 * torch.randn(n, 3, generator=gen), 0).to(DEV))
        tran_t.append(torch.cumsum(0.01 * torch.randn(n, 3, generator=gen), 0))
    single = torch.stack([ev.eval(pose_p[i], r6d_to_rotation_matrix(gt[i].to(DEV)).view(-1, 24, 3, 3), tran_p=tran_p[i], tran_t=tran_t[i])
                          for i in range(len(lens))])
    group = ev.eval_group(torch.cat(pose_p), torch.cat(gt).to(DEV), torch.cat(tran_p), torch.cat(tran_t).to(DEV), lens)
    assert group.shape == (len(lens), 8, 2)
    assert torch.equal(torch.nan_to_num(group, nan=-1.0), torch.nan_to_num(single, nan=-1.0))
