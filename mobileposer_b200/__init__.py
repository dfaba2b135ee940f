"""mobileposer_b200 -- B200-native (sm_100a) implementation of MobilePoser's per-frame hot path.

Drop-in surface (mirrors mobileposer.models / mobileposer.utils.model_utils of the reference):
    MobilePoserNet, Joints, Poser, FootContact, Velocity, RNN, load_model, MODULES
Everything executes in hand-written CUDA behind the C ABI of include/mobileposer_b200.h; there is no
CPU or eager fallback (the package raises if the library or an sm_100 device is missing).
"""
from .config import amass, datasets, joint_set, model_config  # noqa: F401
from .modules import RNN, FootContact, Joints, Poser, Velocity  # noqa: F401
from .net import HostOffline, MobilePoserNet, OnlineStreams, getenv  # noqa: F401
from .model_utils import load_model, reduced_pose_to_full  # noqa: F401

# mobileposer/constants.py:6-11
MODULES = {'poser': Poser, 'joints': Joints, 'foot_contact': FootContact, 'velocity': Velocity}

__all__ = ['MobilePoserNet', 'Joints', 'Poser', 'FootContact', 'Velocity', 'RNN', 'OnlineStreams', 'HostOffline', 'load_model',
           'reduced_pose_to_full', 'MODULES', 'getenv', 'model_config', 'amass', 'datasets', 'joint_set']
