"""A/B of the 128-sequence four-sub-tile recurrence (lstm_rec_f16w.cu, MP_REC_WIDE=1) against the 64-sequence kernel: results of a
bidirectional and a unidirectional head (outputs and final states), then isolated per-kernel times at cfg3 size."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mobileposer_b200 as mp
from mobileposer_b200 import _cabi
from mobileposer_b200.synthetic import synthetic_imu_batch

torch.manual_seed(0)
net = mp.MobilePoserNet().eval().to('cuda:0')
lib = _cabi.lib()
for B, T in ((256, 37), (128, 5), (384, 64)):
    x = synthetic_imu_batch(list(range(B)), T).to('cuda:0')
    xv = torch.cat((torch.randn(B, T, 72, device='cuda:0') * 0.3, x), -1)
    lens = [T] * B
    outs = {}
    for wide in ('0', '1'):
        os.environ['MP_REC_WIDE'] = wide
        yj, _, (hj, cj) = net.joints.joints(x, lens)
        yv, _, (hv, cv) = net.velocity.vel(xv, lens)
        torch.cuda.synchronize()
        outs[wide] = [t.clone() for t in (yj, hj, cj, yv, hv, cv)]
    errs = [(a - b).abs().max().item() for a, b in zip(outs['0'], outs['1'])]
    print(f'[wide ab] B={B} T={T}: max |wide - 64| joints y/hn/cn {errs[0]:.2e} {errs[1]:.2e} {errs[2]:.2e}  velocity y/hn/cn {errs[3]:.2e} {errs[4]:.2e} {errs[5]:.2e}')
    assert max(errs) < 5e-6, errs
B, T = 256, 300
x = synthetic_imu_batch(list(range(B)), T).to('cuda:0')
lens = [T] * B
for wide in ('0', '1'):
    os.environ['MP_REC_WIDE'] = wide
    for _ in range(2):
        net.joints(x, lens)
    torch.cuda.synchronize()
    _cabi.check(lib.mp_profile_enable(1))
    for _ in range(5):
        net.joints(x, lens)
    prof = _cabi.profile_collect()
    _cabi.check(lib.mp_profile_enable(0))
    v = prof['lstm_rec_f16_h256']
    print(f'[wide ab] MP_REC_WIDE={wide}: lstm_rec_f16_h256 {v["total_ms"] / v["launches"]:.4f} ms per bidirectional layer launch ({v["launches"]} launches)')
