"""`PhysicsOptimizer` behind MobilePoserNet's PHYSICS hook (mobileposer/models/net.py:66-69,157-169,211-217).

The reference does `from dynamics import PhysicsOptimizer`, a top-level module that is NOT in its tree (its rbdl
dependency is neither vendored nor pinned), so there is no reference behaviour to match: PARITY
UNPINNED.  This class keeps the hook's interface --

    PhysicsOptimizer(debug=False); .reset_states(); .optimize_frame(pose, jvel, contact, acc) -> (pose, tran)

-- and runs the kinematic-physics optimizer this repository defines (float64 statement in
oracle/physics_port.py) as one sm_100a kernel, one warp per skeleton (csrc/physics.cu, `mp_physics_optimize`).
`optimize_sequences` is the batched form MobilePoserNet.forward_offline uses: B skeletons walk their T frames in one
launch instead of the reference's per-frame Python loop.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi
from .config import FLOOR_Y, amass, datasets
from .modules import _f32c, _require_cuda, current_stream_ptr


class PhysicsOptimizer:
    def __init__(self, debug: bool = False, w_vel: float = 1.0, w_contact: float = 10.0, damping: float = 1.0,
                 damping_abs: float = 1e-2, fps: float = datasets.fps, floor_y: float = FLOOR_Y):
        self.debug = debug
        self.params = _cabi.PhysicsParams(w_vel=w_vel, w_contact=w_contact, damping=damping, damping_abs=damping_abs, fps=fps,
                                          vel_scale=1.0, floor_y=floor_y)
        self.state = None

    # ---- reference hook -------------------------------------------------------------------------------
    def reset_states(self):
        """net.py:69: forget the root position and the previous frame (all skeletons)."""
        if self.state is not None:
            self.state.zero_()

    def _state(self, B, dev):
        if self.state is None or self.state.shape[0] != B or self.state.device != dev:
            self.state = torch.zeros(B, _cabi.PHYSICS_STATE_FLOATS, device=dev, dtype=torch.float32)
        return self.state

    @torch.no_grad()
    def optimize_frame(self, pose, jvel, contact, acc=None):
        """One frame of one skeleton (net.py:166,214): pose [24,3,3] (or [24,9]) local rotations, jvel [24,3] joint
        velocities in m/s (the caller already multiplied by amass.vel_scale), contact [2] logits, acc ignored
        (zeros offline, net.py:159) -> (pose_opt [24,3,3], tran [3])."""
        _require_cuda(pose, 'pose')
        p, t = self._run(pose.reshape(1, 1, 24, 9), jvel.reshape(1, 1, 72), contact.reshape(1, 1, 2), None, 1.0)
        return p.view(24, 3, 3), t.view(3)

    # ---- batched form -------------------------------------------------------------------------------
    @torch.no_grad()
    def optimize_sequences(self, pose, vel, contact, lengths=None, vel_scale: float = amass.vel_scale, out=None):
        """pose [B,T,24,3,3] local rotations, vel [B,T,72] RAW velocity-head output (scaled by vel_scale inside, net.py:162),
        contact [B,T,2] logits, lengths (device int32 [B]) or None -> (pose_opt [B,T,24,3,3], tran [B,T,3]).
        State carries over from the previous call unless `reset_states()` was called."""
        _require_cuda(pose, 'pose')
        B, T = contact.shape[0], contact.shape[1]
        p, t = self._run(pose.reshape(B, T, 24, 9), vel.reshape(B, T, 72), contact, lengths, float(vel_scale), out)
        return p.view(B, T, 24, 3, 3), t

    def _run(self, pose, vel, contact, lengths, vel_scale, out=None, dbg=None, dbg_frame=-1):
        lib = _cabi.lib()
        pose, vel, contact = _f32c(pose), _f32c(vel), _f32c(contact)
        B, T = pose.shape[0], pose.shape[1]
        dev = pose.device
        st = self._state(B, dev)
        pose_out = out if out is not None else torch.empty_like(pose)
        tran = torch.empty(B, T, 3, device=dev, dtype=torch.float32)
        prm = _cabi.PhysicsParams.from_buffer_copy(self.params)
        prm.vel_scale = vel_scale
        lens_ptr = None
        if lengths is not None:
            lengths = lengths.to(device=dev, dtype=torch.int32).contiguous()
            lens_ptr = lengths.data_ptr()
        with torch.cuda.device(dev):
            if dbg is None:
                _cabi.check(lib.mp_physics_optimize(pose.data_ptr(), vel.data_ptr(), contact.data_ptr(), lens_ptr,
                                                    st.data_ptr(), B, T, C.byref(prm), pose_out.data_ptr(), tran.data_ptr(),
                                                    current_stream_ptr(dev)), 'mp_physics_optimize')
            else:
                _cabi.check(lib.mp_physics_optimize_debug(pose.data_ptr(), vel.data_ptr(), contact.data_ptr(), lens_ptr,
                                                          st.data_ptr(), B, T, C.byref(prm), pose_out.data_ptr(),
                                                          tran.data_ptr(), dbg.data_ptr(), dbg_frame,
                                                          current_stream_ptr(dev)), 'mp_physics_optimize_debug')
        return pose_out, tran


@torch.no_grad()
def forward_kinematics(pose):
    """ParametricModel.forward_kinematics(pose) with shape=None, tran=None, calc_mesh=False
    (articulate/model.py:208-232): pose [n,24,3,3] local -> (global rotations [n,24,3,3], joints [n,24,3])."""
    _require_cuda(pose, 'pose')
    pose = _f32c(pose).view(-1, 24, 3, 3)
    n = pose.shape[0]
    glb = torch.empty_like(pose)
    pos = torch.empty(n, 24, 3, device=pose.device, dtype=torch.float32)
    with torch.cuda.device(pose.device):
        _cabi.check(_cabi.lib().mp_physics_fk(pose.data_ptr(), n, glb.data_ptr(), pos.data_ptr(),
                                              current_stream_ptr(pose.device)), 'mp_physics_fk')
    return glb, pos
