"""`MobilePoserNet` drop-in: same ctor, attributes, state_dict and methods as the reference class
(mobileposer/models/net.py:22-219), with every forward executed by the sm_100a kernels behind the
C ABI (include/mobileposer_b200.h).

Differences that are deliberate and documented (DESIGN.md section "boundary"):
  * `forward_offline` also accepts B > 1: per-sequence semantics equal B independent reference calls
    with a fresh velocity state per sequence; B == 1 keeps the reference's shapes and
    its velocity-state carry across calls.
  * the online state (feet, root, current_root_y) lives on the device; `forward_online_batch` runs S
    independent streams (one state block each) in one launch sequence.
  * inference only; dropout is the identity.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.nn as nn

from . import _cabi
from .config import (FLOOR_Y, SMPL_J_ZERO, SMPL_PARENT, amass, joint_set, model_config)
from .modules import FootContact, Joints, Poser, Velocity, _require_cuda, _f32c, current_stream_ptr


def getenv(key: str, default=0):
    """mobileposer/helpers.py:4-5."""
    return type(default)(os.getenv(key, default))


class _Slot:
    """Static device buffers for one (B, T, device): stable pointers make the C-side CUDA graph reusable."""

    def __init__(self, lib, net_handle, B, T, dev):
        f32 = dict(device=dev, dtype=torch.float32)
        self.B, self.T = B, T
        self.imu = torch.empty(B, T, 60, **f32)
        self.lengths = torch.empty(B, device=dev, dtype=torch.int32)
        self.pose = torch.empty(B * T, 24, 3, 3, **f32)
        self.joints = torch.empty(B, T, 72, **f32)
        self.vel = torch.empty(B, T, 72, **f32)
        self.contact = torch.empty(B, T, 2, **f32)
        self.tran = torch.empty(B, T, 3, **f32)
        self.h0 = torch.empty(2, B, 256, **f32)
        self.c0 = torch.empty(2, B, 256, **f32)
        self.hn = torch.empty(2, B, 256, **f32)
        self.cn = torch.empty(2, B, 256, **f32)
        self.ws_bytes = lib.mp_net_workspace_bytes(net_handle, B, T)
        self.ws = torch.empty(self.ws_bytes, device=dev, dtype=torch.uint8)


class MobilePoserNet(nn.Module):
    """
    Inputs: N IMUs.  Outputs: SMPL pose (local rotation matrices) and root translation.
    """

    def __init__(self, poser: Poser = None, joints: Joints = None, foot_contact: FootContact = None,
                 velocity: Velocity = None, finetune: bool = False):
        super().__init__()
        self.C = model_config
        self.finetune = finetune

        # same construction order as net.py:40-43 (seeded default init is then identical)
        self.pose = poser if poser else Poser()
        self.joints = joints if joints else Joints()
        self.foot_contact = foot_contact if foot_contact else FootContact()
        self.velocity = velocity if velocity else Velocity()

        # constants (net.py:46-56)
        self.parent = list(SMPL_PARENT)
        self.j = torch.tensor(SMPL_J_ZERO, dtype=torch.float32)
        self.feet_pos = self.j[10:12].clone()
        self.floor_y = FLOOR_Y
        self.gravity_velocity = torch.tensor([0, joint_set.gravity_velocity, 0])
        self.prob_threshold = (0.5, 0.9)
        self.num_past_frames = model_config.past_frames
        self.num_future_frames = model_config.future_frames
        self.num_total_frames = self.num_past_frames + self.num_future_frames

        # variables (net.py:58-64); the online ones are mirrored by device state blocks
        self.rnn_state = None
        self.imu = None
        self.reuse_outputs = False        # True: return views of the static buffers (no clone)
        self.rec_tile = int(os.environ.get('MP_NET_TILE', 0))   # mp_net_set_rec_tile policy of this module's own net handle

        # net.py:66-69.  The reference imports `dynamics.PhysicsOptimizer`, a module that is not in its tree
        # (SURVEY.md F2); here the hook is served by mobileposer_b200.dynamics (parity unpinned).
        self.dynamics_optimizer = None
        self.last_physics_tran = None
        if getenv("PHYSICS"):
            self.enable_physics()

        self._net = None
        self._net_key = None
        self._net_physics = None
        self._slots = {}
        self._online = None
        self.last_launches = 0

    # ---- handles -------------------------------------------------------------------------------
    def _device(self):
        return next(self.parameters()).device

    def _net_handle(self):
        heads = (self.joints.joints, self.pose.pose, self.foot_contact.footcontact, self.velocity.vel)
        keys = tuple(h.packed_key() for h in heads)      # (handle, generation): see PackedHead.generation
        if self._net is None or keys != self._net_key:
            self._close_net()
            out = C.c_void_p()
            with torch.cuda.device(self._device()):
                _cabi.check(_cabi.lib().mp_net_create(C.byref(out), *[k[0] for k in keys]), 'mp_net_create')
            self._net, self._net_key = out.value, keys
            self._slots = {}
            self._net_physics = None
            _cabi.check(_cabi.lib().mp_net_set_rec_tile(self._net, int(self.rec_tile)), 'mp_net_set_rec_tile')
        return self._net

    def _sync_net_physics(self, inline: bool):
        """Switch the in-graph K8 tail of mp_net_forward on (batched forward_offline) or off."""
        net = self._net_handle()
        want = None
        if inline and self.dynamics_optimizer is not None:
            prm = _cabi.PhysicsParams.from_buffer_copy(self.dynamics_optimizer.params)
            prm.vel_scale = amass.vel_scale
            want = bytes(prm)
        if want != self._net_physics:
            arg = C.byref(_cabi.PhysicsParams.from_buffer_copy(want)) if want is not None else None
            with torch.cuda.device(self._device()):
                _cabi.check(_cabi.lib().mp_net_set_physics(net, arg), 'mp_net_set_physics')
            self._net_physics = want

    def _close_net(self):
        if self._net is not None:
            _cabi.lib().mp_net_destroy(self._net)
            self._net = None
            self._net_key = None

    def __del__(self):
        try:
            self._close_net()
        except Exception:
            pass

    def enable_physics(self, on: bool = True, **params):
        """PHYSICS=1 without the environment variable; `params` go to dynamics.PhysicsOptimizer."""
        if on:
            from .dynamics import PhysicsOptimizer
            self.dynamics_optimizer = PhysicsOptimizer(debug=False, **params)
            self.dynamics_optimizer.reset_states()
        else:
            self.dynamics_optimizer = None

    def set_graph(self, enabled: bool):
        _cabi.check(_cabi.lib().mp_net_set_graph(self._net_handle(), int(enabled)), 'mp_net_set_graph')

    def _slot(self, B, T, dev):
        key = (B, T, dev.index)
        if key not in self._slots:
            if len(self._slots) >= 8:
                self._slots.pop(next(iter(self._slots)))
            self._slots[key] = _Slot(_cabi.lib(), self._net_handle(), B, T, dev)
        return self._slots[key]

    # ---- reference surface -----------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, model_path):
        model = cls()
        model.load_state_dict(torch.load(model_path, map_location='cpu'))
        model.finetune = True
        return model

    def reset(self):
        """net.py:84-88: clears imu / root position / current_root_y; NOT the feet, NOT velocity.rnn_state."""
        self.rnn_state = None
        self.imu = None
        if self._online is not None:
            self._online.reset(full=False)

    def _prob_to_weight(self, p):
        lo, hi = self.prob_threshold
        return (p.clamp(lo, hi) - lo) / (hi - lo)

    @torch.no_grad()
    def _reduced_global_to_full(self, reduced_pose):
        """net.py:93-99 through mp_pose_reduced_global_to_full."""
        _require_cuda(reduced_pose, 'reduced pose')
        r6d = _f32c(reduced_pose).view(-1, 96)
        out = torch.empty(r6d.shape[0], 24, 3, 3, device=r6d.device, dtype=torch.float32)
        with torch.cuda.device(r6d.device):
            _cabi.check(_cabi.lib().mp_pose_reduced_global_to_full(r6d.data_ptr(), r6d.shape[0], out.data_ptr(),
                                                                   current_stream_ptr(r6d.device)), 'K5')
        return out

    def _lengths(self, input_lengths, B, T):
        lens = [int(v) for v in (input_lengths.tolist() if torch.is_tensor(input_lengths) else input_lengths)]
        if len(lens) != B:
            raise ValueError(f'input_lengths has {len(lens)} entries for a batch of {B}')
        if min(lens) <= 0 or max(lens) > T:
            raise RuntimeError('Length of all samples has to be greater than 0 and at most the padded length')
        return lens

    @torch.no_grad()
    def _run(self, batch, input_lengths, want_tran, carry_velocity, physics_inline=False):
        """One fused mp_net_forward; returns the slot holding the results."""
        _require_cuda(batch, 'input batch')
        lib = _cabi.lib()
        x = _f32c(batch)
        if x.dim() != 3 or x.shape[2] != 60:
            raise ValueError(f'expected [B, T, 60] IMU input, got {tuple(x.shape)}')
        B, T = x.shape[0], x.shape[1]
        dev = x.device
        lens = self._lengths(input_lengths, B, T)
        net = self._net_handle()
        self._sync_net_physics(physics_inline)
        s = self._slot(B, T, dev)
        s.imu.copy_(x)
        ragged = min(lens) < T
        if ragged:
            s.lengths.copy_(torch.tensor(lens, dtype=torch.int32), non_blocking=True)
        h0 = c0 = None
        if carry_velocity and self.velocity.rnn_state is not None:
            sh, sc = self.velocity.rnn_state
            if tuple(sh.shape) != (2, B, 256):
                raise RuntimeError(f'Expected hidden[0] size (2, {B}, 256), got {list(sh.shape)}')
            s.h0.copy_(sh)
            s.c0.copy_(sc)
            h0, c0 = s.h0, s.c0
        with torch.cuda.device(dev):
            _cabi.check(lib.mp_net_forward(
                net, s.imu.data_ptr(), B, T, s.lengths.data_ptr() if ragged else None,
                h0.data_ptr() if h0 is not None else None, c0.data_ptr() if c0 is not None else None,
                s.hn.data_ptr(), s.cn.data_ptr(), s.pose.data_ptr(), s.joints.data_ptr(), s.vel.data_ptr(),
                s.contact.data_ptr(), s.tran.data_ptr() if want_tran else None, s.ws.data_ptr(), s.ws_bytes,
                current_stream_ptr(dev)), 'mp_net_forward')
        self.last_launches = int(lib.mp_launch_count())
        if carry_velocity:
            self.velocity.rnn_state = (s.hn.clone(), s.cn.clone())
        return s, max(lens)

    def _out(self, t):
        return t if self.reuse_outputs else t.clone()

    @torch.no_grad()
    def forward(self, batch, input_lengths=None):
        """net.py:101-119 -> (pose [B*T,24,3,3], joints [B,T,72], vel ([T,72] if B == 1 else [B,T,72]), contact [B,T,2])."""
        if input_lengths is None:
            return self._forward_unfused(batch)
        s, tmax = self._run(batch, input_lengths, want_tran=False, carry_velocity=True)
        if tmax < s.T:
            raise RuntimeError('padded length must equal max(input_lengths) (torch.cat in net.py:106 requires it)')
        return self._out(s.pose), self._out(s.joints), self._out(s.vel).squeeze(0), self._out(s.contact)

    def _forward_unfused(self, batch):
        """input_lengths=None: the reference then runs every LSTM sequence-first (rnn.py:15); head by head."""
        pred_joints = self.joints(batch, None)
        pred_pose = self._reduced_global_to_full(self.pose.pose(pred_joints, None, None, x2=batch)[0])
        contact = self.foot_contact.footcontact(pred_joints, None, None, x2=batch)[0]
        vel, _, self.velocity.rnn_state = self.velocity.vel(pred_joints, None, self.velocity.rnn_state, x2=batch)
        return pred_pose, pred_joints, vel.squeeze(0), contact

    @torch.no_grad()
    def forward_offline(self, imu, input_lengths=None):
        """net.py:121-171.  B == 1: (pose [T,24,3,3], joints [1,T,72], tran [T,3], contact [T,2]);
        B > 1: (pose [B*T,24,3,3], joints [B,T,72], tran [B,T,3], contact [B,T,2]), fresh velocity state per sequence."""
        if input_lengths is None:
            raise ValueError('forward_offline needs input_lengths (every reference caller passes them)')
        B = imu.shape[0]
        # batched sequences: K8 runs inside the net's graph (fresh optimizer state per sequence, F6 convention)
        s, tmax = self._run(imu, input_lengths, want_tran=True, carry_velocity=(B == 1), physics_inline=(B > 1))
        if tmax < s.T:
            raise RuntimeError('padded length must equal max(input_lengths)')
        pose, joints, tran, contact = self._out(s.pose), self._out(s.joints), self._out(s.tran), self._out(s.contact)
        if self.dynamics_optimizer is not None and B == 1:
            # net.py:157-169: per-frame optimize_frame over the sequence, here one launch; the optimizer's translation is
            # discarded like the reference does (`pose, _ = ...`).  B == 1 keeps the reference's state carry between calls
            # (reset_states() is only called by the constructor, net.py:69).
            pose_opt, self.last_physics_tran = self.dynamics_optimizer.optimize_sequences(
                s.pose.view(1, s.T, 24, 3, 3), s.vel, s.contact, None, out=pose.view(1, s.T, 24, 9))
            pose = pose_opt.view(s.T, 24, 3, 3)
        if B == 1:
            return pose, joints, tran[0], contact[0]
        return pose, joints, tran, contact

    # ---- online ------------------------------------------------------------------------------------
    def _online_state(self, S, dev):
        if self._online is None or self._online.S != S or self._online.dev != dev:
            self._online = OnlineStreams(self, S, dev)
        return self._online

    @torch.no_grad()
    def forward_online(self, data, input_lengths=None):
        """net.py:173-219 for one frame [60] -> (pose [24,9], joints [W,72], root [3], contact [2])."""
        _require_cuda(data, 'input frame')
        st = self._online_state(1, data.device)
        pose, joints, root, contact = st.step(data.reshape(1, 60))
        self.imu = st.window[0]
        pose = self._out(pose[0])             # a fresh tensor per tick like the reference's (the state's buffer is rewritten by the next tick)
        if self.dynamics_optimizer is not None:
            return pose.view(24, 3, 3), joints[0], root[0].clone(), contact[0]      # net.py:216-217
        return pose.view(24, 9), joints[0], root[0].clone(), contact[0]

    @torch.no_grad()
    def forward_online_batch(self, frames):
        """S independent live streams: frames [S,60] -> (pose [S,24,9], joints [S,W,72], root [S,3], contact [S,2])."""
        _require_cuda(frames, 'input frames')
        st = self._online_state(frames.shape[0], frames.device)
        pose, joints, root, contact = st.step(frames)
        return self._out(pose).view(-1, 24, 9), joints, root.clone(), contact

    @property
    def last_root_pos(self):
        if self._online is None:
            return torch.zeros(3)
        return self._online.state_floats()[0, 6:9].clone()

    @property
    def current_root_y(self):
        if self._online is None:
            return 0
        return float(self._online.state.view(torch.float64)[0, 6].item())


class OnlineStreams:
    """Device-resident state of S `forward_online` streams (net.py:58-64,84-88,173-219)."""

    def __init__(self, net: MobilePoserNet, S: int, dev, window: int = None):
        self.net, self.S, self.dev = net, S, dev
        self.W = window or net.num_total_frames
        self.P = net.num_past_frames if window is None else window - net.num_future_frames
        lib = _cabi.lib()
        self.state = torch.zeros(S, _cabi.ONLINE_STATE_BYTES, device=dev, dtype=torch.uint8)
        self.win = [torch.zeros(S, self.W, 60, device=dev), torch.zeros(S, self.W, 60, device=dev)]
        self.cur = 0
        self.cold = True
        self.pose = torch.empty(S, 24, 3, 3, device=dev)
        self.root = torch.empty(S, 3, device=dev)
        self.contact = torch.empty(S, 2, device=dev)
        self.lengths = [self.W] * S
        with torch.cuda.device(dev):
            _cabi.check(lib.mp_online_reset(self.state.data_ptr(), S, 1, current_stream_ptr(dev)), 'mp_online_reset')

    @property
    def window(self):
        return self.win[self.cur]

    def state_floats(self):
        return self.state.view(torch.float32)

    def reset(self, full=False):
        with torch.cuda.device(self.dev):
            _cabi.check(_cabi.lib().mp_online_reset(self.state.data_ptr(), self.S, int(full), current_stream_ptr(self.dev)),
                        'mp_online_reset')
        self.cold = True

    def step(self, frames):
        lib = _cabi.lib()
        net = self.net
        frames = _f32c(frames).view(self.S, 60)
        nxt = self.cur ^ 1
        with torch.cuda.device(self.dev):
            stream = current_stream_ptr(self.dev)
            _cabi.check(lib.mp_online_push_frame(self.win[self.cur].data_ptr(), self.win[nxt].data_ptr(), frames.data_ptr(),
                                                 self.S, self.W, int(self.cold), stream), 'mp_online_push_frame')
            self.cur, self.cold = nxt, False
            # the four heads over the whole window (velocity carries its state)
            s, _ = net._run(self.win[self.cur], self.lengths, want_tran=False, carry_velocity=True)
            # state machine on frame P of the window outputs (K7)
            _cabi.check(lib.mp_online_update(self.state.data_ptr(), s.pose.data_ptr(), s.joints.data_ptr(),
                                             s.vel.data_ptr(), s.contact.data_ptr(), self.S, self.W, self.P,
                                             self.pose.data_ptr(), self.root.data_ptr(), self.contact.data_ptr(), stream),
                        'mp_online_update')
            if net.dynamics_optimizer is not None:
                # net.py:211-217: optimize the tick's frame (joint velocity of frame P times vel_scale); tran discarded
                net.dynamics_optimizer.optimize_sequences(self.pose.view(self.S, 1, 24, 3, 3), s.vel[:, self.P:self.P + 1],
                                                          self.contact.view(self.S, 1, 2), None,
                                                          out=self.pose.view(self.S, 1, 24, 9))
        return self.pose, s.joints if net.reuse_outputs else s.joints.clone(), self.root, self.contact.clone()


class HostOffline:
    """forward_offline through HOST buffers: H2D of the IMU batch, the whole forward, D2H of (pose, joints, tran,
    contact), all enqueued by one C-ABI call (mp_net_enqueue_offline_host) on this object's own stream.  Fresh velocity
    (and optimizer) state per call.  Every HostOffline owns its `mp_net` handle (side streams, graph cache), device
    staging, workspace and pinned outputs for a fixed (B, T), so two of them form a depth-2 pipeline over batches:

        a.submit(x0); b.submit(x1); a.wait() -> a.pose ...; a.submit(x2); b.wait() ...

    `run()` = submit + wait (one batch at a time, what evaluate.py's loop does)."""

    def __init__(self, net: MobilePoserNet, B: int, T: int, device=None, rec_tile: int = 0, compact: bool = False):
        """rec_tile: mp_net_set_rec_tile policy of this slot's net handle (0 = auto).
        compact: transfer the pose as [B*T, 16, 6] (first two columns of the 16 non-ignored joints' local rotations, 384 B per
        frame instead of 864 B) -- `self.pose` is then that array and `model_utils.local6d_to_pose` rebuilds [B*T, 24, 3, 3]."""
        lib = _cabi.lib()
        self.net, self.B, self.T = net, B, T
        self.dev = device or net._device()
        heads = (net.joints.joints, net.pose.pose, net.foot_contact.footcontact, net.velocity.vel)
        self._heads_key = tuple(h.packed_key() for h in heads)
        out = C.c_void_p()
        with torch.cuda.device(self.dev):
            _cabi.check(lib.mp_net_create(C.byref(out), *[k[0] for k in self._heads_key]), 'mp_net_create')
            self.handle = out.value
            _cabi.check(lib.mp_net_set_rec_tile(self.handle, int(rec_tile)), 'mp_net_set_rec_tile')
            self.stream = torch.cuda.Stream(self.dev)
            self.staging = torch.empty(lib.mp_net_host_staging_bytes(B, T), device=self.dev, dtype=torch.uint8)
            self.ws_bytes = lib.mp_net_workspace_bytes(self.handle, B, T)
            self.ws = torch.empty(self.ws_bytes, device=self.dev, dtype=torch.uint8)
        self._physics = None
        self.compact = bool(compact)
        self.pose = (torch.empty(B * T, 16, 6) if compact else torch.empty(B * T, 24, 3, 3)).pin_memory()
        self.joints = torch.empty(B, T, 72).pin_memory()
        self.tran = torch.empty(B, T, 3).pin_memory()
        self.contact = torch.empty(B, T, 2).pin_memory()
        self.lengths = torch.empty(B, dtype=torch.int32).pin_memory()

    def __del__(self):
        try:
            if self.handle is not None:
                self.stream.synchronize()
                _cabi.lib().mp_net_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def _sync_physics(self):
        opt = self.net.dynamics_optimizer
        want = None
        if opt is not None:
            prm = _cabi.PhysicsParams.from_buffer_copy(opt.params)
            prm.vel_scale = amass.vel_scale
            want = bytes(prm)
        if want != self._physics:
            arg = C.byref(_cabi.PhysicsParams.from_buffer_copy(want)) if want is not None else None
            with torch.cuda.device(self.dev):
                _cabi.check(_cabi.lib().mp_net_set_physics(self.handle, arg), 'mp_net_set_physics')
            self._physics = want

    def submit(self, imu_host: torch.Tensor, input_lengths=None):
        """Enqueue one batch; returns immediately.  `imu_host` must stay untouched until wait()."""
        if imu_host.is_cuda or imu_host.dtype != torch.float32 or not imu_host.is_contiguous():
            raise ValueError('HostOffline expects a contiguous float32 host tensor')
        if tuple(imu_host.shape) != (self.B, self.T, 60):
            raise ValueError(f'expected {(self.B, self.T, 60)}, got {tuple(imu_host.shape)}')
        heads = (self.net.joints.joints, self.net.pose.pose, self.net.foot_contact.footcontact, self.net.velocity.vel)
        if tuple(h.packed_key() for h in heads) != self._heads_key:
            raise RuntimeError('the parameters of the net changed after this HostOffline was built; build a new one')
        lens_ptr = None
        if input_lengths is not None:
            lens = self.net._lengths(input_lengths, self.B, self.T)
            self.lengths.copy_(torch.tensor(lens, dtype=torch.int32))
            lens_ptr = self.lengths.data_ptr()
        self._sync_physics()
        entry = _cabi.lib().mp_net_enqueue_offline_host_compact if self.compact else _cabi.lib().mp_net_enqueue_offline_host
        with torch.cuda.device(self.dev):
            _cabi.check(entry(
                self.handle, imu_host.data_ptr(), self.B, self.T, lens_ptr, self.pose.data_ptr(),
                self.joints.data_ptr(), self.tran.data_ptr(), self.contact.data_ptr(), self.staging.data_ptr(),
                self.ws.data_ptr(), self.ws_bytes, self.stream.cuda_stream), 'mp_net_enqueue_offline_host')

    def submit_device(self, imu_dev: torch.Tensor):
        """The same batch step with device-resident input and outputs (no copies): mp_net_forward on this object's
        stream into its own device buffers (`d_pose`, `d_joints`, `d_tran`, `d_contact`, valid after wait())."""
        _require_cuda(imu_dev, 'input batch')
        if tuple(imu_dev.shape) != (self.B, self.T, 60) or imu_dev.dtype != torch.float32 or not imu_dev.is_contiguous():
            raise ValueError(f'expected a contiguous float32 {(self.B, self.T, 60)} device tensor')
        if not hasattr(self, 'd_pose'):
            f32 = dict(device=self.dev, dtype=torch.float32)
            B, T = self.B, self.T
            self.d_pose, self.d_joints = torch.empty(B * T, 24, 3, 3, **f32), torch.empty(B, T, 72, **f32)
            self.d_vel, self.d_contact, self.d_tran = torch.empty(B, T, 72, **f32), torch.empty(B, T, 2, **f32), torch.empty(B, T, 3, **f32)
            self.d_hn, self.d_cn = torch.empty(2, B, 256, **f32), torch.empty(2, B, 256, **f32)
            self.d_imu = torch.empty(B, T, 60, **f32)
            torch.cuda.current_stream(self.dev).synchronize()
        self._sync_physics()
        # the net's CUDA graph is keyed on its pointers: stage the batch in this slot's own buffer (a device-to-device copy of
        # B*T*240 bytes on the slot's stream) so that every call after the second replays the graph, whatever tensor comes in
        self.stream.wait_stream(torch.cuda.current_stream(self.dev))     # whoever produced imu_dev on the caller's stream
        with torch.cuda.stream(self.stream):
            self.d_imu.copy_(imu_dev, non_blocking=True)
        imu_dev.record_stream(self.stream)
        with torch.cuda.device(self.dev):
            _cabi.check(_cabi.lib().mp_net_forward(
                self.handle, self.d_imu.data_ptr(), self.B, self.T, None, None, None, self.d_hn.data_ptr(), self.d_cn.data_ptr(),
                self.d_pose.data_ptr(), self.d_joints.data_ptr(), self.d_vel.data_ptr(), self.d_contact.data_ptr(),
                self.d_tran.data_ptr(), self.ws.data_ptr(), self.ws_bytes, self.stream.cuda_stream), 'mp_net_forward')
        self.last_launches = int(_cabi.lib().mp_launch_count())

    def wait(self):
        """Block until the submitted batch is on the host; -> (pose [B*T,24,3,3], joints, tran, contact) pinned tensors."""
        self.stream.synchronize()
        return self.pose, self.joints, self.tran, self.contact

    def run(self, imu_host: torch.Tensor, input_lengths=None):
        self.submit(imu_host, input_lengths)
        return self.wait()
