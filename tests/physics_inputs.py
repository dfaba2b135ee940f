"""Seeded synthetic inputs of the K8 optimizer tests: smooth local rotations around a bent pose (ignored joints are the
identity, as K5 emits them), velocity-head-like outputs and contact logits."""
import numpy as np

from oracle.physics_port import exp_so3

IGNORED = [0, 7, 8, 10, 11, 20, 21, 22, 23]


def synthetic_motion(B, T, seed=0, amp=0.35):
    rng = np.random.default_rng(seed)
    base = rng.normal(size=(B, 1, 24, 3)) * amp
    freq = rng.uniform(0.02, 0.15, size=(B, 1, 24, 3))
    phase = rng.uniform(0, 6.28, size=(B, 1, 24, 3))
    t = np.arange(T).reshape(1, T, 1, 1)
    w = base + 0.25 * amp * np.sin(freq * t + phase)
    R = exp_so3(w)
    R[:, :, IGNORED[1:]] = np.eye(3)
    vel = rng.normal(size=(B, T, 72)) * 0.15
    contact = 2.0 * np.sin(0.2 * np.arange(T).reshape(1, T, 1) + rng.uniform(0, 6.28, size=(B, 1, 2))) + rng.normal(size=(B, T, 2)) * 0.3
    return R.astype(np.float32), vel.astype(np.float32), contact.astype(np.float32)
