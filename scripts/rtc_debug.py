"""Tiny tcgen05-recurrence run for compute-sanitizer: joints head, B sequences x T frames, TC vs FFMA path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mobileposer_b200 as mp
from mobileposer_b200.synthetic import synthetic_imu_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
T = int(sys.argv[2]) if len(sys.argv) > 2 else 6
torch.manual_seed(0)
net = mp.MobilePoserNet().eval().to('cuda:0')
x = synthetic_imu_batch(list(range(B)), T).to('cuda:0')
lens = [T] * B
os.environ['MP_REC_IMPL'] = 'ffma'
ref = net.joints(x, lens).clone()
torch.cuda.synchronize()
os.environ['MP_REC_IMPL'] = sys.argv[3] if len(sys.argv) > 3 else 'tc'      # tc / f16 = the fp16-split kernel, tf32 = the first-generation one
out = net.joints(x, lens)
torch.cuda.synchronize()
print('B', B, 'T', T, os.environ['MP_REC_IMPL'], 'max |tc - ffma|', (out - ref).abs().max().item(), 'ref max', ref.abs().max().item())
