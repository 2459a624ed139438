"""multiposenet.pytorch_b200 -- B200-native hot path of MultiPoseNet behind the reference's API.

    from multiposenet.pytorch_b200 import poseNet, pth_nms, install_dropin
    install_dropin()          # makes `from network.posenet import poseNet` and
                              # `from lib.nms.pth_nms import pth_nms` resolve to this package
"""
from .dropin import install_dropin  # noqa: F401
from .lib.nms.pth_nms import pth_nms  # noqa: F401
from .network.posenet import PoseNet, poseNet  # noqa: F401

__all__ = ["poseNet", "PoseNet", "pth_nms", "install_dropin"]
