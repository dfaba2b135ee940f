"""Parity metrics of BASELINE.json:north_star: joint angles and root translation within 1e-4 rad / 1e-4 m,
foot-contact logits identical in argmax."""
import torch

ANGLE_TOL = 1e-4   # rad
TRAN_TOL = 1e-4    # m
VALUE_TOL = 1e-4   # head outputs (joint positions in m, velocities, logits)


def geodesic(Ra, Rb):
    """Rotation angle between two batches of rotation matrices [..., 3, 3], in float64.

    2 asin(||Ra - Rb||_F / (2 sqrt 2)) is exact for rotations and well conditioned near zero."""
    d = (Ra.double() - Rb.double()).flatten(-2).norm(dim=-1)
    return 2.0 * torch.asin((d / (2.0 * 2.0 ** 0.5)).clamp(max=1.0))


def max_angle(Ra, Rb):
    return geodesic(Ra.cpu(), Rb.cpu()).max().item()


def max_abs(a, b):
    return (a.detach().cpu().double() - b.detach().cpu().double()).abs().max().item()


def argmax_equal(a, b):
    return torch.equal(a.detach().cpu().argmax(-1), b.detach().cpu().argmax(-1))


def min_margin(contact):
    c = contact.detach().cpu()
    return (c[..., 0] - c[..., 1]).abs().min().item()
