"""Isolated per-kernel device times (mp_profile_*) of the joints head at cfg3 size (256 x 300)."""
import os, sys, torch
sys.path.insert(0, '.')
import mobileposer_b200 as mp
from mobileposer_b200 import _cabi
from mobileposer_b200.synthetic import synthetic_imu_batch
torch.manual_seed(0)
net = mp.MobilePoserNet().eval().to('cuda:0')
x = synthetic_imu_batch(list(range(256)), 300).to('cuda:0')
lens = [300] * 256
for _ in range(2):
    net.joints(x, lens)
torch.cuda.synchronize()
lib = _cabi.lib()
_cabi.check(lib.mp_profile_enable(1))
for _ in range(5):
    net.joints(x, lens)
prof = _cabi.profile_collect()
_cabi.check(lib.mp_profile_enable(0))
for k, v in prof.items():
    print(k, v['launches'], 'avg ms', v['total_ms'] / v['launches'])
