"""Input assembly: the CPU restatement against the live reference's golden vectors."""
import numpy as np

from conftest import load_golden
from mobileposer_b200.config import amass
from oracle import input_port as ip


def test_dataset_assembly_matches_reference():
    g = load_golden('input_assembly')
    assert list(g['combo_slots'].numpy()) == [sum(1 << s for s in c) for c in amass.combos.values()]
    out = ip.assemble_dataset(g['raw_acc'].numpy(), g['raw_ori'].numpy())
    assert np.array_equal(out, g['dataset_imu'].numpy())


def test_loader_assembly_and_smoothing_match_reference():
    g = load_golden('input_assembly')
    acc, ori = g['raw_acc'].numpy(), g['raw_ori'].numpy()
    assert np.array_equal(ip.smooth_avg(acc[:, :5]), g['smooth3'].numpy())
    for name in ('lw_rp', 'rw_lp_h'):
        assert np.array_equal(ip.assemble_loader(acc, ori, amass.combos[name]), g['loader_' + name].numpy())
    assert np.array_equal(ip.smooth_avg(acc[:1, :5]), acc[:1, :5])          # a single frame is its own average


def test_live_normalisation_matches_reference_expressions():
    """oracle/input_port.py:live_normalize against the live demo's calibration + normalisation evaluated with the
    reference's own quaternion_to_rotation_matrix and config (oracle/make_golden_inputs.py:live_golden)."""
    import numpy as np
    from oracle.input_port import live_normalize, quaternion_to_rotation_matrix
    from mobileposer_b200.config import amass
    from mobileposer_b200.inputs import LIVE_SLOT_ORDER, LiveCalibration
    g = {k: v.numpy() for k, v in load_golden('live_normalize').items()}
    assert tuple(g['perm']) == LIVE_SLOT_ORDER and float(g['acc_scale']) == amass.acc_scale
    assert np.abs(quaternion_to_rotation_matrix(g['ori_q']).reshape(g['ori_raw'].shape) - g['ori_raw']).max() < 5e-7
    for name, kw in (('imu_lw_rp', dict(combo=amass.combos['lw_rp'])), ('imu_rw_rp_h', dict(combo=amass.combos['rw_rp_h'])),
                     ('imu_phone_as_watch', dict(phone_as_watch=True))):
        out = live_normalize(g['ori_q'], g['acc_raw'], g['smpl2imu'], g['device2bone'], g['acc_offsets'], **kw)
        assert out.shape == g[name].shape
        assert np.array_equal(out == 0, g[name] == 0)
        assert np.abs(out - g[name]).max() < 2e-6, name
    # the host-side calibration helper reproduces the matrices the fixture was made with from the same kind of readings
    cal = LiveCalibration(g['smpl2imu'], g['device2bone'], g['acc_offsets'])
    assert cal.smpl2imu.shape == (3, 3) and cal.device2bone.shape == (5, 3, 3) and cal.acc_offsets.shape == (5, 3)
    import torch
    q0 = torch.tensor([0.9, 0.1, -0.2, 0.3])
    c2 = LiveCalibration.from_readings(q0, torch.randn(5, 4), torch.randn(5, 3))
    r = c2.smpl2imu
    assert (r @ r.t() - torch.eye(3)).abs().max() < 1e-6
