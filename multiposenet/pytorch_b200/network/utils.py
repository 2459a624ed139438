"""`BBoxTransform` / `ClipBoxes` with the reference's interface (network/utils.py:6-61).

BBoxTransform.forward runs the decode kernel (mpn_decode_clip without clipping); ClipBoxes.forward
clamps in place exactly as utils.py:56-60 does.  poseNet.forward itself uses the fused decode+clip
launch (engine.entire_forward_device) instead of chaining the two modules.
"""
import torch.nn as nn

from .. import ops as _ops


class BBoxTransform(nn.Module):
    def __init__(self, mean=None, std=None):
        super().__init__()
        if mean is not None or std is not None:
            raise NotImplementedError("only the reference defaults mean=0, std=[.1,.1,.2,.2] are built in")

    def forward(self, boxes, deltas):
        if not deltas.is_cuda:
            raise RuntimeError("BBoxTransform needs CUDA tensors (no CPU path)")
        return _ops.decode_clip(boxes.contiguous(), deltas.contiguous(), 0, 0)


class ClipBoxes(nn.Module):
    def __init__(self, width=None, height=None):
        super().__init__()

    def forward(self, boxes, img):
        _, _, height, width = img.shape
        boxes[:, :, 0].clamp_(min=0)
        boxes[:, :, 1].clamp_(min=0)
        boxes[:, :, 2].clamp_(max=width)
        boxes[:, :, 3].clamp_(max=height)
        return boxes
