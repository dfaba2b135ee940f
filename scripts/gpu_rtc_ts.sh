# per-step clock64 stamps of the tcgen05 recurrence (block (0,0)) + isolated launch time of one bidirectional layer
set -x
mkdir -p gpurun_out
MP_RTC_TS=1 timeout 300 python scripts/rtc_debug.py 256 40 > gpurun_out/rtc_ts.log 2>&1; grep "rtc ts" gpurun_out/rtc_ts.log | head -14
timeout 300 python - > gpurun_out/rtc_time.log 2>&1 <<'PY'
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
PY
cat gpurun_out/rtc_time.log | tail -5
