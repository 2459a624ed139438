"""Parameter containers with the reference's module tree for the ResNet-50/101 + dual-FPN backbone.

Mirror of /root/reference/network/fpn.py (Bottleneck :9-26, FPN :37-82, FPN50/FPN101 :128-134): the
same child names, shapes and registration order, so `state_dict()` keys match the reference's (and
torchvision-ResNet checkpoints load with strict=False as in training/multipose_keypoint_train.py:74-75).
The modules hold parameters only; the arithmetic runs in libmpn_b200 (see ..engine.Engine), so calling
them directly is an error rather than a silent PyTorch fallback.
"""
import torch.nn as nn

_PLANES = (64, 128, 256, 512)
_STRIDES = (1, 2, 2, 2)


def _no_eager(self, *a, **k):
    raise RuntimeError("%s holds parameters only; run the model through poseNet.forward (libmpn_b200 kernels)" % type(self).__name__)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, in_planes, planes, stride=1):
        super().__init__()
        out_planes = planes * self.expansion
        self.conv1 = nn.Conv2d(in_planes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)  # stride sits on the 3x3
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, out_planes, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(out_planes)
        self.downsample = nn.Sequential()
        if stride != 1 or in_planes != out_planes:
            self.downsample = nn.Sequential(nn.Conv2d(in_planes, out_planes, 1, stride=stride, bias=False),
                                            nn.BatchNorm2d(out_planes))
        self.stride = stride

    forward = _no_eager


class FPN(nn.Module):
    def __init__(self, block, num_blocks):
        super().__init__()
        self.num_blocks = tuple(num_blocks)
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        in_planes = 64
        for i, (planes, n, stride) in enumerate(zip(_PLANES, num_blocks, _STRIDES), start=1):
            blocks = []
            for b in range(n):
                blocks.append(block(in_planes, planes, stride if b == 0 else 1))
                in_planes = planes * block.expansion
            setattr(self, "layer%d" % i, nn.Sequential(*blocks))
        # detection neck (RetinaNet P3..P7)
        self.conv6 = nn.Conv2d(2048, 256, 3, stride=2, padding=1)
        self.conv7 = nn.Conv2d(256, 256, 3, stride=2, padding=1)
        self.latlayer1 = nn.Conv2d(2048, 256, 1)
        self.latlayer2 = nn.Conv2d(1024, 256, 1)
        self.latlayer3 = nn.Conv2d(512, 256, 1)
        self.toplayer0 = nn.Conv2d(256, 256, 3, padding=1)
        self.toplayer1 = nn.Conv2d(256, 256, 3, padding=1)
        self.toplayer2 = nn.Conv2d(256, 256, 3, padding=1)
        # keypoint neck (P2..P5)
        self.toplayer = nn.Conv2d(2048, 256, 1)
        self.flatlayer1 = nn.Conv2d(1024, 256, 1)
        self.flatlayer2 = nn.Conv2d(512, 256, 1)
        self.flatlayer3 = nn.Conv2d(256, 256, 1)
        self.smooth1 = nn.Conv2d(256, 256, 3, padding=1)
        self.smooth2 = nn.Conv2d(256, 256, 3, padding=1)
        self.smooth3 = nn.Conv2d(256, 256, 3, padding=1)

    forward = _no_eager


def FPN50():
    return FPN(Bottleneck, [3, 4, 6, 3])


def FPN101():
    return FPN(Bottleneck, [3, 4, 23, 3])
