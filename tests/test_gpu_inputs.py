"""GPU: mp_imu_assemble against the live reference's golden vectors (bit-exact) and the CPU restatement at a larger size."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from mobileposer_b200.config import amass
from oracle import input_port as ip

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def test_dataset_and_loader_assembly_match_reference_bit_for_bit():
    from mobileposer_b200.inputs import assemble_imu
    g = load_golden('input_assembly')
    acc, ori = g['raw_acc'].to(DEV), g['raw_ori'].to(DEV)
    assert torch.equal(assemble_imu(acc, ori).cpu(), g['dataset_imu'])               # all 12 combos, one launch
    for name in ('lw_rp', 'rw_lp_h'):
        assert torch.equal(assemble_imu(acc, ori, [name], smooth=True)[0].cpu(), g['loader_' + name])
    one = assemble_imu(acc[:1], ori[:1], ['lw_rp'], smooth=True)[0].cpu()             # T = 1: a frame is its own average
    assert torch.equal(one, torch.from_numpy(ip.assemble_loader(g['raw_acc'][:1].numpy(), g['raw_ori'][:1].numpy(), amass.combos['lw_rp'])))


def test_large_stream_matches_cpu_restatement():
    from mobileposer_b200.inputs import assemble_imu
    g = torch.Generator().manual_seed(5)
    acc, ori = torch.randn(3000, 5, 3, generator=g) * 5, torch.randn(3000, 5, 3, 3, generator=g)
    out = assemble_imu(acc.to(DEV), ori.to(DEV), ['lw_rp', 'rp', 'lw_lp_h'])
    ref = ip.assemble_dataset(acc.numpy(), ori.numpy(), [amass.combos[c] for c in ('lw_rp', 'rp', 'lw_lp_h')])
    assert np.array_equal(out.cpu().numpy(), ref)
    with pytest.raises(ValueError):
        assemble_imu(acc[:, :, :2].to(DEV), ori.to(DEV))


def test_metric_rows_on_the_device_match_the_reference_evaluator():
    """mp_eval_frame_errors (N1, the part that needs no mesh) behind evaluate.full_motion_errors: the rows the reference's
    FullMotionEvaluator produced on the same motions (golden metrics_unit.npz), and the per-frame planes vs the torch statement."""
    from mobileposer_b200.evaluate import angle_between, forward_kinematics, frame_errors_cuda, full_motion_errors
    g = load_golden('metrics_unit')
    args = [g[k].to(DEV) for k in ('pose_a', 'pose_b', 'tran_a', 'tran_b')]
    errs = full_motion_errors(*args).cpu()
    rows = [0, 2, 3, 4, 5, 6, 7, 8, 9]
    rel = ((errs[rows] - g['errs'][rows]).abs() / g['errs'][rows].abs().clamp_min(1e-6)).max().item()
    assert rel < 2e-4, rel
    jp, jt, je, lae, gae = (t.cpu() for t in frame_errors_cuda(*args))
    gp, jp_ref = forward_kinematics(g['pose_a'], g['tran_a'])
    gt, jt_ref = forward_kinematics(g['pose_b'], g['tran_b'])
    assert (jp - jp_ref).abs().max() < 2e-6 and (jt - jt_ref).abs().max() < 2e-6 and (jp - g['joint_a']).abs().max() < 2e-6
    assert (lae - torch.rad2deg(angle_between(g['pose_a'], g['pose_b']))).abs().max() < 2e-3      # degrees
    assert (gae - torch.rad2deg(angle_between(gp, gt))).abs().max() < 2e-3


def test_live_normalisation_matches_reference_expressions():
    """mp_imu_live_normalize (third part of SURVEY.md 8f row N2) against the live demo's expressions evaluated with the
    reference's own functions (tests/golden/live_normalize.npz), and on a long buffer against the CPU restatement."""
    import numpy as np
    from mobileposer_b200.inputs import LiveCalibration, normalize_live
    from oracle.input_port import live_normalize
    g = load_golden('live_normalize')
    cal = LiveCalibration(g['smpl2imu'], g['device2bone'], g['acc_offsets'])
    q, a = g['ori_q'].to('cuda:0'), g['acc_raw'].to('cuda:0')
    for name, kw in (('imu_lw_rp', dict(combo='lw_rp')), ('imu_rw_rp_h', dict(combo='rw_rp_h')),
                     ('imu_phone_as_watch', dict(phone_as_watch=True))):
        out = normalize_live(q, a, cal, **kw).cpu()
        assert out.shape == g[name].shape
        assert torch.equal(out == 0, g[name] == 0)
        assert (out - g[name]).abs().max() < 2e-6, name
    gen = torch.Generator().manual_seed(3)
    qn, an = torch.randn(5000, 5, 4, generator=gen), torch.randn(5000, 5, 3, generator=gen) * 10
    out = normalize_live(qn.to('cuda:0'), an.to('cuda:0'), cal, combo='lw_rp_h').cpu().numpy()
    from mobileposer_b200.config import amass
    ref = live_normalize(qn.numpy(), an.numpy(), g['smpl2imu'].numpy(), g['device2bone'].numpy(), g['acc_offsets'].numpy(),
                         combo=amass.combos['lw_rp_h'])
    assert np.abs(out - ref).max() < 5e-6
    # one tick feeds forward_online directly
    tick = normalize_live(q[:1], a[:1], cal)
    assert tick.shape == (1, 60)
