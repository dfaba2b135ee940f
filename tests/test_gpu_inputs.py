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
