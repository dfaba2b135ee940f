"""Ablation of lstm_rec_f16w_kernel (cost attribution only; results are WRONG with a skip bit set): needs a library built with
-DMP_RECW_ABLATION (add it to NVCC_FLAGS in mobileposer_b200/build.py and rebuild with --force), MP_RECW_NO_TMA_GIN=1 for the LDG path the
recorded run used.  Recorded in profiles/r02_rec_wide_ab.txt."""
import os, sys, torch
sys.path.insert(0, '.')
import mobileposer_b200 as mp
from mobileposer_b200 import _cabi
from mobileposer_b200.synthetic import synthetic_imu_batch
torch.manual_seed(0)
net = mp.MobilePoserNet().eval().to('cuda:0')
lib = _cabi.lib()
B, T = 256, 300
x = synthetic_imu_batch(list(range(B)), T).to('cuda:0')
os.environ['MP_REC_WIDE'] = '1'
for skip in ('0', '1', '2', '3', '4', '7'):
    os.environ['MP_RECW_SKIP'] = skip
    for _ in range(2):
        net.joints(x, [T] * B)
    torch.cuda.synchronize()
    _cabi.check(lib.mp_profile_enable(1))
    for _ in range(5):
        net.joints(x, [T] * B)
    prof = _cabi.profile_collect()
    _cabi.check(lib.mp_profile_enable(0))
    v = prof['lstm_rec_f16_h256']
    print(f'[skip {skip}] lstm_rec_f16w {v["total_ms"] / v["launches"]:.4f} ms per launch')
