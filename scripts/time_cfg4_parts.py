"""Where one cfg4 pass (evaluate_pose over the synthetic DIP set, batch_size = 50, one GPU) spends its wall time."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mobileposer_b200 as mp
from mobileposer_b200.evaluate import PoseEvaluator, evaluate_pose, synthetic_dip

dev = 'cuda:0'
torch.manual_seed(0)
net = mp.MobilePoserNet().eval().to(dev)
items = synthetic_dip()
lens = [it[0].shape[0] for it in items]
for _ in range(3):
    evaluate_pose(net, items, verbose=False, batch_size=50)
torch.cuda.synchronize()


def t(fn, n=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, r


ms_all, _ = t(lambda: evaluate_pose(net, items, verbose=False, batch_size=50))


def assemble():
    xb = torch.zeros(50, 3000, 60, device=dev)
    for r, it in enumerate(items):
        xb[r, :lens[r]] = it[0].to(dev)
    return xb


ms_in, xb = t(assemble)
ms_fwd, out = t(lambda: net.forward_offline(xb, lens))
pose_b, _, tran_b, _ = out
pose_b = pose_b.view(50, 3000, 24, 3, 3)
tran_b = tran_b.view(50, 3000, 3)


def gt():
    g = torch.empty(sum(lens), 144, device=dev)
    tt = torch.empty(sum(lens), 3, device=dev)
    off = 0
    for r, it in enumerate(items):
        g[off:off + lens[r]].copy_(it[1].reshape(lens[r], 144), non_blocking=True)
        tt[off:off + lens[r]].copy_(it[3].reshape(lens[r], 3), non_blocking=True)
        off += lens[r]
    return g, tt


ms_gt, (g, tt) = t(gt)
ev = PoseEvaluator()
ms_ev, rows = t(lambda: ev.eval_group(torch.cat([pose_b[r, :lens[r]] for r in range(50)]), g, torch.cat([tran_b[r, :lens[r]] for r in range(50)]), tt, lens))
print(f'[cfg4 parts] whole pass {ms_all:.1f} ms | input assembly (50 H2D copies + pad) {ms_in:.1f} | forward_offline B=50 T=3000 {ms_fwd:.1f} | '
      f'ground truth to device {ms_gt:.1f} | eval_group {ms_ev:.1f}')
from mobileposer_b200 import _cabi
lib = _cabi.lib()
_cabi.check(lib.mp_profile_enable(1))
net.forward_offline(xb, lens)
prof = _cabi.profile_collect()
_cabi.check(lib.mp_profile_enable(0))
print('[cfg4 parts] forward kernels (ms):', {k: round(v['total_ms'], 2) for k, v in prof.items()})
