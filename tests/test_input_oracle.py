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
