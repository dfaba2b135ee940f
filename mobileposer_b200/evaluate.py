"""`evaluate.py` of the reference (mobileposer/evaluate.py:16-126) for the B200 path.

Same entry (`evaluate_pose(model, dataset, ...)`, CLI `--model --dataset`), same hot loop
(`model.reset(); model.forward_offline(x[None], [T])`, evaluate.py:56-58) and the same printed table
(evaluate.py:31-36).  Differences:
  * sequences are sharded over the ranks of `torch.distributed` when it is initialised (sharding.py) and the
    per-sequence [8, 2] rows are all-gathered once at the end;
  * the metric rows are computed on the device from SMPL forward kinematics of the 24 joints
    (evaluator.py:292-343: rows 0, 2-9).  The mesh-vertex row (evaluator.py row 1) needs the 6890-vertex SMPL
    template, which cannot ship: it is NaN unless `PoseEvaluator(smpl_file=...)` / `--smpl` names the official model
    file (or a `mesh=(rest_vertices, weights)` pair is given); then it runs on the device too
    (`vertex_error_row` -> mp_eval_vertex_errors, pinned to the reference's evaluator over a synthetic template,
    tests/golden/mesh_unit.npz);
  * the translation-error windows of `evaluate_tran=True` (evaluate.py:66-92) run on the
    device too (`tran_window_errors` -> mp_eval_tran_windows), pinned to the reference's own `evaluate_pose` output
    (tests/golden/tran_windows.npz);
  * `--dataset synthetic_dip` evaluates a synthetic DIP-shaped set (10 subjects x 5 sequences x 3000 frames,
    BASELINE.json config 4) because the real datasets cannot ship.
"""
from __future__ import annotations

import math
from argparse import ArgumentParser

import torch

from .config import SMPL_J_ZERO, SMPL_PARENT, datasets, joint_set
from .model_utils import default_device, load_model
from .net import getenv
from .sharding import gather_rows, shard_sequences
from .synthetic import synthetic_imu


def r6d_to_rotation_matrix(r6d: torch.Tensor) -> torch.Tensor:
    """articulate/math/angular.py:167-182 (ground-truth conversion at evaluate.py:60; not the hot path)."""
    v = r6d.reshape(-1, 6)
    c0 = v[:, :3] / v[:, :3].norm(dim=1, keepdim=True)
    b = v[:, 3:] - (c0 * v[:, 3:]).sum(dim=1, keepdim=True) * c0
    c1 = b / b.norm(dim=1, keepdim=True)
    r = torch.stack((c0, c1, torch.linalg.cross(c0, c1, dim=1)), dim=-1)
    return torch.nan_to_num(r, nan=0.0)


def forward_kinematics(pose_local: torch.Tensor, tran: torch.Tensor = None):
    """Local rotations [N, 24, 3, 3] -> (global rotations [N, 24, 3, 3], joint positions [N, 24, 3]) for the mean
    shape (articulate/model.py:208-232 without the mesh)."""
    n, dev = pose_local.shape[0], pose_local.device
    j = torch.tensor(SMPL_J_ZERO, dtype=pose_local.dtype, device=dev)
    bone = j.clone()
    for i in range(1, 24):
        bone[i] = j[i] - j[SMPL_PARENT[i]]
    glb, pos = [pose_local[:, 0]], [bone[0].expand(n, 3)]
    for i in range(1, 24):
        p = SMPL_PARENT[i]
        glb.append(glb[p] @ pose_local[:, i])
        pos.append(pos[p] + (glb[p] @ bone[i]))
    glb, pos = torch.stack(glb, dim=1), torch.stack(pos, dim=1)
    if tran is not None:
        pos = pos + tran.view(-1, 1, 3)
    return glb, pos


def angle_between(r1: torch.Tensor, r2: torch.Tensor) -> torch.Tensor:
    """Rotation angle of r1^T r2 in radians (angular.py:86-99), robust near 0 and pi."""
    d = r1.transpose(-1, -2) @ r2
    skew = torch.stack((d[..., 2, 1] - d[..., 1, 2], d[..., 0, 2] - d[..., 2, 0], d[..., 1, 0] - d[..., 0, 1]), dim=-1)
    tr = d[..., 0, 0] + d[..., 1, 1] + d[..., 2, 2]
    return torch.atan2(skew.norm(dim=-1), tr - 1.0)


def frame_errors_cuda(pose_p, pose_t, tran_p, tran_t):
    """mp_eval_frame_errors: -> (joint_p [n,24,3], joint_t [n,24,3], je [n,24], lae [n,24], gae [n,24])."""
    from . import _cabi
    from .modules import _f32c, current_stream_ptr
    pose_p, pose_t = _f32c(pose_p).view(-1, 24, 3, 3), _f32c(pose_t).view(-1, 24, 3, 3)
    n, dev = pose_p.shape[0], pose_p.device
    tp = _f32c(tran_p).view(n, 3) if tran_p is not None else None
    tt = _f32c(tran_t).view(n, 3) if tran_t is not None else None
    jp, jt = torch.empty(n, 24, 3, device=dev), torch.empty(n, 24, 3, device=dev)
    je, lae, gae = (torch.empty(n, 24, device=dev) for _ in range(3))
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().mp_eval_frame_errors(pose_p.data_ptr(), pose_t.data_ptr(), tp.data_ptr() if tp is not None else None,
                                                     tt.data_ptr() if tt is not None else None, n, jp.data_ptr(), jt.data_ptr(),
                                                     je.data_ptr(), lae.data_ptr(), gae.data_ptr(), current_stream_ptr(dev)),
                    'mp_eval_frame_errors')
    return jp, jt, je, lae, gae


class _Opaque:
    """Stand-in for pickled classes that are not installed (chumpy: only `shapedirs` of the official SMPL file)."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.state = state


def load_smpl_mesh(smpl_file, device=None):
    """(rest_vertices [V,3] = v_template - J[0], weights [V,24]) of an official SMPL model file, the two arrays
    ParametricModel.__init__ / get_zero_pose_joint_and_vertex (articulate/model.py:28-39,86-87) feed the skinning with."""
    import pickle

    import numpy as np

    class Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if module.split('.')[0] == 'chumpy':
                return _Opaque
            return super().find_class(module, name)

    with open(smpl_file, 'rb') as f:
        data = Unpickler(f, encoding='latin1').load()
    v = torch.from_numpy(np.asarray(data['v_template'])).float()
    j = torch.from_numpy(np.asarray(data['J'])).float()
    w = torch.from_numpy(np.asarray(data['weights'])).float()
    if w.shape != (v.shape[0], 24) or j.shape != (24, 3):
        raise ValueError(f'{smpl_file}: not a 24-joint SMPL model (weights {tuple(w.shape)}, J {tuple(j.shape)})')
    dev = device if device is not None else default_device()
    return (v - j[:1]).contiguous().to(dev), w.contiguous().to(dev)


def vertex_error_row(pose_p, pose_t, mesh):
    """Row 1 of FullMotionEvaluator.__call__ (evaluator.py:319-323,336): [mean, mean over vertices of the std over
    frames] of the root-aligned vertex position error, from mp_eval_vertex_errors (per-vertex sum and sum of squares in
    float64; the skinned vertex sets are never materialised).  mesh = (rest_vertices [V,3], weights [V,24]).  CUDA only."""
    from . import _cabi
    from .modules import _f32c, current_stream_ptr
    if not pose_p.is_cuda:
        raise RuntimeError('vertex_error_row runs on the GPU (mp_eval_vertex_errors); got a CPU tensor')
    pose_p, pose_t = _f32c(pose_p).view(-1, 24, 3, 3), _f32c(pose_t).view(-1, 24, 3, 3)
    n, dev = pose_p.shape[0], pose_p.device
    rest, w = _f32c(mesh[0].to(dev)), _f32c(mesh[1].to(dev))
    V = rest.shape[0]
    if rest.shape != (V, 3) or w.shape != (V, 24) or pose_t.shape != pose_p.shape:
        raise ValueError(f'vertex_error_row: rest {tuple(rest.shape)}, weights {tuple(w.shape)}, poses {tuple(pose_p.shape)} / {tuple(pose_t.shape)}')
    vsum = torch.empty(V, device=dev, dtype=torch.float64)
    vsq = torch.empty(V, device=dev, dtype=torch.float64)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().mp_eval_vertex_errors(pose_p.data_ptr(), pose_t.data_ptr(), n, rest.data_ptr(), w.data_ptr(), V,
                                                      vsum.data_ptr(), vsq.data_ptr(), current_stream_ptr(dev)),
                    'mp_eval_vertex_errors')
    mean = vsum.sum() / (n * V)
    if n > 1:
        std = ((vsq - vsum * vsum / n).clamp_min(0.0) / (n - 1)).sqrt().mean()
    else:
        std = vsum.new_zeros(())
    return torch.stack((mean, std)).float()


def tran_window_errors(tran_p, tran_t, lengths=None):
    """Translation-error windows of `evaluate_pose(..., evaluate_tran=True)` (evaluate.py:66-92) on the device.

    tran_p / tran_t: [T, 3] (one sequence) or [S, T, 3] (padded batch, `lengths` = valid frames per sequence) ->
    (err [S, 7], count [S, 7]): mean relative drift over the frame pairs across which the ground truth moves at least
    1..7 m (NaN where a sequence has no such pair) and the number of pairs.  CUDA only (mp_eval_tran_windows)."""
    from . import _cabi
    from .modules import _f32c, current_stream_ptr
    if not tran_p.is_cuda:
        raise RuntimeError('tran_window_errors runs on the GPU (mp_eval_tran_windows); got a CPU tensor')
    tp = _f32c(tran_p).view(-1, tran_p.shape[-2], 3)
    tt = _f32c(tran_t.to(tp.device)).view(-1, tran_p.shape[-2], 3)
    S, T, dev = tp.shape[0], tp.shape[1], tp.device
    if tt.shape != tp.shape:
        raise ValueError(f'tran_p {tuple(tp.shape)} and tran_t {tuple(tt.shape)} differ')
    ln = None
    if lengths is not None:
        ln = torch.as_tensor(lengths, dtype=torch.int32).to(dev).contiguous()
        if ln.numel() != S:
            raise ValueError(f'{ln.numel()} lengths for {S} sequences')
    err = torch.empty(S, 7, device=dev)
    cnt = torch.empty(S, 7, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().mp_eval_tran_windows(tp.data_ptr(), tt.data_ptr(), ln.data_ptr() if ln is not None else None,
                                                     S, T, err.data_ptr(), cnt.data_ptr(), current_stream_ptr(dev)),
                    'mp_eval_tran_windows')
    return err, cnt


def full_motion_errors(pose_p, pose_t, tran_p, tran_t, fps=datasets.fps, joint_mask=(2, 5, 16, 20), mesh=None):
    """[10, 2] mean/std rows of FullMotionEvaluator.__call__ (evaluator.py:292-343); row 1 (mesh) is NaN unless
    `mesh = (rest_vertices, weights)` is given (CUDA tensors only)."""
    f = fps
    if pose_p.is_cuda:
        # the per-frame part (two forward kinematics, joint / local-angle / global-angle errors) is one kernel, the reduction of
        # its outputs to the ten (mean, std) rows a second one (mp_eval_motion_rows)
        from . import _cabi
        from .modules import current_stream_ptr
        jp, jt, je, lae, gae = frame_errors_cuda(pose_p, pose_t, tran_p, tran_t)
        rows = torch.empty(10, 2, device=pose_p.device, dtype=torch.float32)
        bits = 0
        for j in joint_mask:
            bits |= 1 << int(j)
        with torch.cuda.device(pose_p.device):
            _cabi.check(_cabi.lib().mp_eval_motion_rows(jp.data_ptr(), jt.data_ptr(), je.data_ptr(), lae.data_ptr(), gae.data_ptr(),
                                                        jp.shape[0], int(f), bits, rows.data_ptr(), current_stream_ptr(pose_p.device)),
                        'mp_eval_motion_rows')
        if mesh is not None:
            rows[1] = vertex_error_row(pose_p, pose_t, mesh)
        return rows
    else:
        # CPU tensors: the torch statement of the same rows (what tests/test_evaluate.py pins to the reference's evaluator)
        gp, jp = forward_kinematics(pose_p, tran_p)
        gt, jt = forward_kinematics(pose_t, tran_t)
        off = (jt[:, 0] - jp[:, 0]).unsqueeze(1)
        je = (jp + off - jt).norm(dim=2)
        lae = torch.rad2deg(angle_between(pose_p, pose_t))
        gae = torch.rad2deg(angle_between(gp, gt))
    jkp = ((jp[3:] - 3 * jp[2:-1] + 3 * jp[1:-2] - jp[:-3]) * (f ** 3)).norm(dim=2)
    jkt = ((jt[3:] - 3 * jt[2:-1] + 3 * jt[1:-2] - jt[:-3]) * (f ** 3)).norm(dim=2)
    te = ((jp[f:, :1] - jp[:-f, :1]) - (jt[f:, :1] - jt[:-f, :1])).norm(dim=2) * 100
    m = list(joint_mask)
    nan = torch.full((2,), float('nan'), device=pose_p.device)

    def row(x):
        if x.numel() == 0:
            return nan
        return torch.stack((x.mean(), x.std(dim=0).mean() if x.shape[0] > 1 else x.new_zeros(())))

    mesh_row = vertex_error_row(pose_p, pose_t, mesh) if mesh is not None else nan      # row 1
    return torch.stack([row(je), mesh_row, row(lae), row(gae), row(jkp), row(jkt), row(te), row(je[:, m]), row(lae[:, m]),
                        row(gae[:, m])])


class PoseEvaluator:
    """evaluate.py:16-36."""
    names = ['SIP Error (deg)', 'Angular Error (deg)', 'Masked Angular Error (deg)', 'Positional Error (cm)',
             'Masked Positional Error (cm)', 'Mesh Error (cm)', 'Jitter Error (100m/s^3)', 'Distance Error (cm)']

    def __init__(self, smpl_file=None, mesh=None):
        """`smpl_file`: official SMPL model file (evaluate.py:18 reads paths.smpl_file) or `mesh` = (rest_vertices, weights)
        for the Mesh Error row; without either that row is NaN."""
        self.mesh = mesh if mesh is not None else (load_smpl_mesh(smpl_file) if smpl_file else None)

    def eval(self, pose_p, pose_t, joint_p=None, tran_p=None, tran_t=None):
        pose_p = pose_p.clone().view(-1, 24, 3, 3)
        pose_t = pose_t.clone().view(-1, 24, 3, 3).to(pose_p.device)
        tran_p = tran_p.clone().view(-1, 3)
        tran_t = tran_t.clone().view(-1, 3).to(pose_p.device)
        eye = torch.eye(3, device=pose_p.device)
        pose_p[:, joint_set.ignored] = eye
        pose_t[:, joint_set.ignored] = eye
        errs = full_motion_errors(pose_p, pose_t, tran_p, tran_t, mesh=self.mesh if pose_p.is_cuda else None)
        return torch.stack([errs[9], errs[3], errs[9], errs[0] * 100, errs[7] * 100, errs[1] * 100, errs[4] / 100, errs[6]])

    def eval_group(self, pose_p, pose_t_r6d, tran_p, tran_t, lens):
        """The rows of `eval` for a GROUP of sequences whose valid frames are concatenated (pose_p [sum(lens), 24, 3, 3] on the device,
        pose_t_r6d [sum(lens), 144] ground truth as the dataset stores it, tran_* [sum(lens), 3]): the per-frame work -- ground-truth
        r6d -> rotation (evaluate.py:60), both forward kinematics, the three per-joint errors -- is one pass over all frames instead of
        one per sequence (~40 small launches each), and the (mean, std) rows are one `mp_eval_motion_rows_batch` launch (one CTA per
        sequence over its slice, the accumulation order of a launch of its own), so every row equals what `eval` returns for that
        sequence alone.  -> [len(lens), 8, 2].  CUDA only."""
        from . import _cabi
        from .modules import current_stream_ptr
        dev = pose_p.device
        pose_p = pose_p.clone().view(-1, 24, 3, 3)
        pose_t = r6d_to_rotation_matrix(pose_t_r6d.to(dev)).view(-1, 24, 3, 3)
        eye = torch.eye(3, device=dev)
        pose_p[:, joint_set.ignored] = eye
        pose_t[:, joint_set.ignored] = eye
        tran_p, tran_t = tran_p.reshape(-1, 3), tran_t.to(dev).reshape(-1, 3)
        jp, jt, je, lae, gae = frame_errors_cuda(pose_p, pose_t, tran_p, tran_t)
        rows = torch.empty(len(lens), 10, 2, device=dev, dtype=torch.float32)
        bits = 0
        for j in (2, 5, 16, 20):
            bits |= 1 << j
        offsets = torch.tensor([0] + [int(n) for n in lens], dtype=torch.int64).cumsum(0).to(dev)
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().mp_eval_motion_rows_batch(jp.data_ptr(), jt.data_ptr(), je.data_ptr(), lae.data_ptr(), gae.data_ptr(),
                                                              offsets.data_ptr(), len(lens), int(datasets.fps), bits, rows.data_ptr(),
                                                              current_stream_ptr(dev)), 'mp_eval_motion_rows_batch')
        if self.mesh is not None:
            off = 0
            for r, n in enumerate(lens):
                rows[r, 1] = vertex_error_row(pose_p[off:off + n], pose_t[off:off + n], self.mesh)
                off += n
        scale = torch.tensor([1.0, 1.0, 1.0, 100.0, 100.0, 100.0, 0.01, 1.0], device=dev).view(1, 8, 1)
        return rows[:, [9, 3, 9, 0, 7, 1, 4, 6]] * scale

    @classmethod
    def print(cls, errors):
        for i, name in enumerate(cls.names):
            print('%s: %.2f (+/- %.2f)' % (name, errors[i, 0], errors[i, 1]))


def synthetic_dip(n_subjects=10, n_seq=5, frames=3000, combo='lw_rp'):
    """A DIP-shaped synthetic test set: items (imu [T,60], pose r6d [T,144], joint [T,24,3], tran [T,3]) like
    PoseDataset.__getitem__ in evaluation mode (data.py:96-107).  Ground truth is a smooth random motion; it only
    gives the metric code something to chew on -- accuracy numbers on it are meaningless."""
    items = []
    for k in range(n_subjects * n_seq):
        g = torch.Generator().manual_seed(50_000 + k)
        imu = synthetic_imu(10_000 + k, frames, combo)
        t = torch.arange(frames, dtype=torch.float32).view(-1, 1)
        phase = torch.rand(144, generator=g) * 6.28
        eye6 = torch.tensor([1., 0., 0., 0., 1., 0.]).repeat(24)
        pose = eye6 + 0.2 * torch.sin(t / 40.0 + phase)
        tran = torch.cumsum(0.01 * torch.sin(t / 60.0 + torch.rand(3, generator=g) * 6.28), dim=0)
        items.append((imu, pose, torch.zeros(frames, 24, 3), tran))
    return items


_COPY_STREAMS = {}


def _copy_stream(device):
    """One side stream per device for the ground-truth upload of evaluate_pose's batched path."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device)
    return _COPY_STREAMS[key]


@torch.no_grad()
def evaluate_pose(model, dataset, num_past_frame=20, num_future_frame=5, evaluate_tran=False, verbose=True, smpl_file=None,
                  mesh=None, batch_size=1, return_online=False):
    """evaluate.py:39-107.  Returns the [n_sequences, 8, 2] offline rows in dataset order (on every rank); with
    `evaluate_tran=True` (evaluate.py:66-92, 105-106) a pair (rows, windows [n_sequences, 7]) and prints the reference's
    `[0, mean drift at 1 m, ..., at 7 m]` list.

    With `ONLINE=1` in the environment (evaluate.py:62-64,97-99) the per-tick path is evaluated too and printed;
    `return_online=True` appends its [n_sequences, 8, 2] rows to the returned tuple.

    `batch_size` = 1 is the reference's loop, one `forward_offline(x[None], [T])` per sequence (evaluate.py:56-58), the
    velocity head's state carried from sequence to sequence exactly as there.  `batch_size` > 1 (extension)
    runs that many sequences of the rank's shard per `forward_offline` call -- padded to the longest, true lengths passed --
    on the throughput kernels, with a fresh velocity state per sequence (SURVEY.md 8e option 1): pose, joint and contact rows
    are the same, the two rows that see the translation (jitter, distance) are those of independent sequences.  Under
    torch.distributed (world size > 1) every sequence starts from a fresh state too, whatever the batch size, so the table is
    invariant to the number of ranks; the sequence-to-sequence carry is kept only for one process with `batch_size` = 1."""
    import torch.distributed as dist
    device = next(model.parameters()).device
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    items = list(dataset)
    lengths = [it[0].shape[0] for it in items]
    shards = shard_sequences(lengths, world)
    mine = shards[rank]
    evaluator = PoseEvaluator(smpl_file=smpl_file, mesh=mesh)
    model.eval()
    rows = []
    online_rows = []
    window_rows = []
    if batch_size < 1:
        raise ValueError(f'batch_size must be positive, got {batch_size}')
    # The reference's loop carries the velocity head's (and the optimizer's) state from one sequence into the next (SURVEY.md
    # F5); that chain only exists for ONE process walking the dataset in order.  Sharded or batched runs start every sequence
    # from a fresh state, so the table does not depend on the world size, the shard order or the batch size.
    fresh = world > 1 or batch_size > 1

    def fresh_state():
        if fresh:
            if getattr(model, 'velocity', None) is not None:
                model.velocity.rnn_state = None
            opt = getattr(model, 'dynamics_optimizer', None)
            if opt is not None:
                opt.reset_states()

    batched = {}                                   # sequence index -> (pose_p [T,24,3,3], tran_p [T,3]) of the current group
    group_rows = {}                                # sequence index -> [8, 2] rows of the current group (device path)
    for pos, i in enumerate(mine):
        imu, pose_t, _, tran_t = items[i]
        x = imu.to(device) if batch_size == 1 or getenv("ONLINE") else None      # batched: the group's input is assembled below
        if batch_size == 1:
            model.reset()
            fresh_state()
            pose_p, _, tran_p, _ = model.forward_offline(x.unsqueeze(0), [x.shape[0]])
        else:
            if i not in batched:                   # first sequence of a group: one forward for the whole group
                group = mine[pos:pos + batch_size]
                lens = [lengths[k] for k in group]
                xb = torch.zeros(len(group), max(lens), imu.shape[1], device=device, dtype=torch.float32)
                for r, k in enumerate(group):
                    xb[r, :lens[r]] = items[k][0].to(device)
                model.reset()
                fresh_state()                      # also covers a trailing group of one, which takes the B == 1 path
                pose_b, _, tran_b, _ = model.forward_offline(xb, lens)
                group_rows = {}
                use_group = pose_b.is_cuda and not getenv("ONLINE")
                if use_group:
                    # the group's ground truth goes to the device while the forward (enqueued above, asynchronous) runs: slice by
                    # slice like the per-sequence path copies it, on a side stream the metric pass then waits for
                    tot = sum(lens)
                    gt = torch.empty(tot, 144, device=device, dtype=torch.float32)
                    tt = torch.empty(tot, 3, device=device, dtype=torch.float32)
                    main, side = torch.cuda.current_stream(device), _copy_stream(device)
                    side.wait_stream(main)
                    with torch.cuda.stream(side):
                        off = 0
                        for r, k in enumerate(group):
                            gt[off:off + lens[r]].copy_(items[k][1].reshape(lens[r], 144), non_blocking=True)
                            tt[off:off + lens[r]].copy_(items[k][3].reshape(lens[r], 3), non_blocking=True)
                            off += lens[r]
                    main.wait_stream(side)
                pose_b = pose_b.view(len(group), max(lens), 24, 3, 3)
                tran_b = tran_b.view(len(group), max(lens), 3)
                # the rows of the group are formed before the next forward (the net may reuse its output buffers)
                batched = {k: (pose_b[r, :lens[r]], tran_b[r, :lens[r]]) for r, k in enumerate(group)}
                if use_group:
                    # the whole group's metric rows in one pass over its frames (PoseEvaluator.eval_group)
                    g_rows = evaluator.eval_group(torch.cat([batched[k][0] for k in group]), gt, torch.cat([batched[k][1] for k in group]),
                                                  tt, lens)
                    group_rows = {k: g_rows[r] for r, k in enumerate(group)}
            pose_p, tran_p = batched[i]
            if i in group_rows:
                rows.append(group_rows[i])
                if evaluate_tran:
                    window_rows.append(tran_window_errors(tran_p, tran_t)[0][0])
                continue
        pose_t = r6d_to_rotation_matrix(pose_t.to(device)).view(-1, 24, 3, 3)
        rows.append(evaluator.eval(pose_p, pose_t, tran_p=tran_p, tran_t=tran_t))
        if evaluate_tran:
            window_rows.append(tran_window_errors(tran_p, tran_t)[0][0])
        if getenv("ONLINE"):
            if fresh:
                # batch_size == 1 and one process: the reset before forward_offline above is the reference's (evaluate.py:56);
                # otherwise every sequence's ticks start from a reset window / root and a fresh velocity state
                model.reset()
                fresh_state()
            outs = [model.forward_online(f) for f in torch.cat((x, x[-1].repeat(num_future_frame, 1)))]
            pose_o = torch.stack([o[0] for o in outs])[num_future_frame:]
            tran_o = torch.stack([o[2] for o in outs])[num_future_frame:]
            online_rows.append(evaluator.eval(pose_o, pose_t, tran_p=tran_o, tran_t=tran_t))
    # ONE all-gather for everything this call reports: [offline 8x2 | online 8x2 | windows 7] per sequence
    n_local = len(mine)
    parts = [torch.stack(rows).reshape(n_local, 16) if rows else torch.zeros(0, 16, device=device)]
    if getenv("ONLINE"):
        parts.append(torch.stack(online_rows).reshape(n_local, 16) if online_rows else torch.zeros(0, 16, device=device))
    if evaluate_tran:
        parts.append(torch.stack(window_rows).reshape(n_local, 7) if window_rows else torch.zeros(0, 7, device=device))
    gathered = gather_rows(torch.cat([p.to(device=device, dtype=torch.float32) for p in parts], dim=1), shards, len(items))
    table = gathered[:, :16].reshape(-1, 8, 2)
    col = 16
    if verbose and rank == 0:
        print('============== offline ================')
        PoseEvaluator.print(table.mean(dim=0) if table.numel() else table)       # evaluate.py:100 (mean, not nanmean)
    if getenv("ONLINE"):
        online = gathered[:, col:col + 16].reshape(-1, 8, 2)
        col += 16
        if verbose and rank == 0:
            print('============== online ================')
            PoseEvaluator.print(online.mean(dim=0))
    if evaluate_tran:
        windows = gathered[:, col:col + 7]
        if verbose and rank == 0:
            # evaluate.py:106 -- per window the mean over the sequences that have at least one pair
            print([0] + [windows[:, k][~torch.isnan(windows[:, k])].mean() for k in range(7)])
    out = [table] + ([windows] if evaluate_tran else []) + ([online] if (return_online and getenv("ONLINE")) else [])
    return out[0] if len(out) == 1 else tuple(out)


if __name__ == '__main__':
    parser = ArgumentParser()
    parser.add_argument('--model', type=str, default=None, help='state_dict .pth; default: seeded random init')
    parser.add_argument('--dataset', type=str, default='synthetic_dip')
    parser.add_argument('--frames', type=int, default=3000)
    parser.add_argument('--smpl', type=str, default=None, help='official SMPL model file: enables the Mesh Error row')
    parser.add_argument('--batch', type=int, default=1, help='sequences per forward_offline call (1 = the reference loop)')
    parser.add_argument('--tran', action='store_true', help='also the translation-error windows (evaluate_tran=True)')
    args = parser.parse_args()
    if args.dataset != 'synthetic_dip':
        raise ValueError(f'Test dataset: {args.dataset} not found.')
    if args.model:
        net = load_model(args.model)
    else:
        from .net import MobilePoserNet
        torch.manual_seed(0)
        net = MobilePoserNet().to(default_device())
    print(f'Starting evaluation: {args.dataset.capitalize()}')
    evaluate_pose(net, synthetic_dip(frames=args.frames), evaluate_tran=args.tran, smpl_file=args.smpl, batch_size=args.batch)
