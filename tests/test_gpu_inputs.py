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
