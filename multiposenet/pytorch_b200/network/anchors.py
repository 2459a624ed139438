"""`Anchors` with the reference's interface (network/anchors.py:6-37): forward(image) -> [1, A, 4] fp32.

The float64 arithmetic of generate_anchors/shift (anchors.py:39-70,106-126) is done once per image
size by libmpn_b200 (mpn_generate_anchors, bit-identical) and cached on the device, instead of numpy
on the CPU plus a host->device copy on every forward.
"""
import torch.nn as nn

from .. import ops as _ops


class Anchors(nn.Module):
    def __init__(self, pyramid_levels=None, strides=None, sizes=None, ratios=None, scales=None):
        super().__init__()
        if any(a is not None for a in (pyramid_levels, strides, sizes, ratios, scales)):
            raise NotImplementedError("only the reference defaults (levels 3-7, 3 ratios x 3 scales) are built in")
        self.pyramid_levels = [3, 4, 5, 6, 7]
        self.strides = [2 ** x for x in self.pyramid_levels]
        self.sizes = [2 ** (x + 2) for x in self.pyramid_levels]

    def forward(self, image):
        if not image.is_cuda:
            raise RuntimeError("Anchors.forward needs a CUDA image batch (no CPU path)")
        return _ops.anchors_for(image.shape[2], image.shape[3], image.device)
