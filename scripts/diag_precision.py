"""Error budget of the CUDA path and of the fp32 CPU reference port against a float64 restatement
(oracle/np_port.py) on one long sequence.  Diagnostic only (test infrastructure)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import torch

import mobileposer_b200 as mp
from mobileposer_b200 import config as C
from mobileposer_b200.synthetic import synthetic_imu
from oracle import np_port
from oracle.torch_port import OraclePoser, reduced_global_to_full
from parity import geodesic

def k5_f64(r6d):
    """K5 in float64 (same algebra as oracle.torch_port.reduced_global_to_full)."""
    v = r6d.double().reshape(-1, 6)
    a, b = v[:, :3], v[:, 3:]
    c0 = a / a.norm(dim=1, keepdim=True)
    b = b - (c0 * b).sum(dim=1, keepdim=True) * c0
    c1 = b / b.norm(dim=1, keepdim=True)
    c2 = torch.linalg.cross(c0, c1, dim=1)
    g16 = torch.stack((c0, c1, c2), dim=-1).view(-1, 16, 3, 3)
    n = g16.shape[0]
    glb = torch.eye(3, dtype=torch.float64).repeat(n, 24, 1, 1)
    glb[:, C.joint_set.reduced] = g16
    loc = glb.clone()
    for i in range(1, 24):
        loc[:, i] = glb[:, C.SMPL_PARENT[i]].transpose(1, 2) @ glb[:, i]
    loc[:, C.joint_set.ignored] = torch.eye(3, dtype=torch.float64)
    loc[:, 0] = glb[:, 0]
    return loc


T = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
torch.manual_seed(0)
net = mp.MobilePoserNet().eval()
sd = {k: v.clone() for k, v in net.state_dict().items()}
net = net.to('cuda:0')
x = synthetic_imu(4242, T)

# float64 truth for joints -> r6d
sdn = {k: v.numpy() for k, v in sd.items()}
j64, _ = np_port.rnn_head(sdn, C.HEAD_PREFIX['joints'], x[None].numpy(), [T], True)
feat64 = np.concatenate([j64, x[None].numpy().astype(np.float64)], axis=2)
r64, _ = np_port.rnn_head(sdn, C.HEAD_PREFIX['pose'], feat64, [T], True)
pose64 = k5_f64(torch.from_numpy(r64[0]))
# CPU fp32 reference port
o = OraclePoser(sd)
jo = o.heads['joints'](x[None], [T])[0]
ro = o.heads['pose'](torch.cat((jo, x[None]), -1), [T])[0]
po = reduced_global_to_full(ro)
# CUDA
jg = net.joints(x[None].cuda(), [T])
rg = net.pose.pose(jg, [T], None, x2=x[None].cuda())[0]
pg = net._reduced_global_to_full(rg)

def mx(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())

print(f'T={T}')
print('joints  |cuda-f64| %.3e  |cpu32-f64| %.3e  |cuda-cpu32| %.3e' % (mx(jg.cpu(), j64), mx(jo, j64), mx(jg.cpu(), jo)))
print('r6d     |cuda-f64| %.3e  |cpu32-f64| %.3e  |cuda-cpu32| %.3e' % (mx(rg.cpu(), r64), mx(ro, r64), mx(rg.cpu(), ro)))
r = torch.from_numpy(r64[0]).view(T, 16, 6)
n0 = r[..., :3].norm(dim=-1)
print('min |r6d column 0| over frames/joints: %.4f ; median %.4f' % (n0.min(), n0.median()))
ag = geodesic(pg.cpu(), pose64)
ao = geodesic(po, pose64)
ago = geodesic(pg.cpu(), po)
print('angle   |cuda-f64| %.3e  |cpu32-f64| %.3e  |cuda-cpu32| %.3e' % (ag.max(), ao.max(), ago.max()))
idx = ago.argmax()
f, jt = divmod(int(idx), 24)
print('worst cuda-vs-cpu32 at frame %d joint %d: cuda-f64 %.3e cpu32-f64 %.3e' % (f, jt, ag.view(-1)[idx], ao.view(-1)[idx]))
