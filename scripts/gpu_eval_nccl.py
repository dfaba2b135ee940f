"""Under torchrun (NCCL, one rank per GPU): the real MobilePoserNet through evaluate_pose, sequences sharded over the ranks, one
all-gather of the metric rows -- against the table rank 0 computes alone for the same set with fresh per-sequence state.
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/gpu_eval_nccl.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import mobileposer_b200 as mp
from mobileposer_b200.evaluate import evaluate_pose, synthetic_dip
from mobileposer_b200.synthetic import well_conditioned_state_dict

local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
torch.manual_seed(0)
net = mp.MobilePoserNet()
net.load_state_dict(well_conditioned_state_dict(net.state_dict()))
net = net.to(dev).eval()
items = [it for n in (300, 240, 300, 180, 270, 300, 210) for it in synthetic_dip(n_subjects=1, n_seq=1, frames=n)]
items = [(imu + 0.001 * k, *rest) for k, (imu, *rest) in enumerate(items)]
ok = True
for bs in (1, 2, 4):
    table, windows = evaluate_pose(net, items, verbose=False, batch_size=bs, evaluate_tran=True)
    # the single-process answer with fresh state per sequence, computed by every rank on its own GPU without collectives
    rows = []
    for it in items:
        net.velocity.rnn_state = None
        net.reset()
        pose_p, _, tran_p, _ = net.forward_offline(it[0].to(dev)[None], [it[0].shape[0]])
        from mobileposer_b200.evaluate import PoseEvaluator, r6d_to_rotation_matrix
        rows.append(PoseEvaluator().eval(pose_p, r6d_to_rotation_matrix(it[1].to(dev)).view(-1, 24, 3, 3), tran_p=tran_p, tran_t=it[3]))
    net.velocity.rnn_state = None
    want = torch.stack(rows)
    keep = [0, 1, 2, 3, 4, 7]
    good = torch.allclose(table[:, keep], want[:, keep], rtol=2e-3, atol=2e-3) and torch.allclose(table[:, 6], want[:, 6], rtol=5e-2, atol=1e-3)
    # identical on every rank
    ref = table.clone()
    dist.broadcast(ref, 0)
    same = torch.equal(torch.nan_to_num(ref), torch.nan_to_num(table))
    flag = torch.tensor([float(good and same)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f'world {world} batch_size {bs}: sharded table == single-process fresh-state table: {bool(flag.item())}; '
              f'gathered {list(table.shape)} + windows {list(windows.shape)}; max |diff| {(table[:, keep] - want[:, keep]).abs().max().item():.2e}')
    ok = ok and bool(flag.item())
dist.destroy_process_group()
if rank == 0:
    print('NCCL EVALUATE OK' if ok else 'NCCL EVALUATE MISMATCH')
sys.exit(0 if ok else 1)
