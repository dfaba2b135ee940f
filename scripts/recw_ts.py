import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mobileposer_b200 as mp
from mobileposer_b200.synthetic import synthetic_imu_batch
torch.manual_seed(0)
net = mp.MobilePoserNet().eval().to('cuda:0')
x = synthetic_imu_batch(list(range(256)), 40).to('cuda:0')
os.environ['MP_REC_WIDE'] = '1'
net.joints(x, [40] * 256)
torch.cuda.synchronize()
os.environ['MP_RECW_TS'] = '1'
net.joints(x, [40] * 256)
torch.cuda.synchronize()
