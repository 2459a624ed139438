"""Experiment: time mpn_conv2d_fwd_multi (a RetinaNet tower layer, 256 -> 256, 3x3) on subsets of the pyramid levels with the stored
(derive=False) and the derived (derive=True) e5m2 copy plane."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from multiposenet.pytorch_b200 import ops

B = 32
LEVELS = [(60, 80), (30, 40), (15, 20), (8, 10), (4, 5)]
g = torch.Generator().manual_seed(0)
w = (torch.randn(256, 256, 3, 3, generator=g) / 48.0).cuda()
pc = ops.pack_conv(w, torch.randn(256, generator=g).cuda(), None, 3)
acts = [ops.act_from_nchw(torch.randn(B, 256, h, w_, generator=g).cuda(), 3) for h, w_ in LEVELS]
for sub in ([0], [1], [2], [3], [4], [0, 1], [0, 1, 2], [0, 1, 2, 3], [0, 1, 2, 3, 4], [2, 3, 4], [3, 4]):
    xs = [acts[i] for i in sub]
    res = []
    for derive in (False, True):
        for _ in range(3):
            ops.conv2d_multi(xs, pc, pad=1, relu=True, derive=derive, want_h8=not derive)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.conv2d_multi(xs, pc, pad=1, relu=True, derive=derive, want_h8=not derive)
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 20 * 1e3)
    print("levels %-16s stored %8.1f us   derived %8.1f us" % (sub, res[0], res[1]))
