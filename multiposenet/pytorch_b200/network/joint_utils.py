"""Device implementation of the reference's heat-map peak extraction (network/joint_utils.py:19-32, 61-152).

Only the hot part is mirrored -- `NMS` / `get_joint_list`, called once per image by evaluate/tester.py:215-221 on a heat
map that the reference first copies to the host; the plotting helpers of the reference module are not part of the path.
"""
import numpy as np
import torch

from .. import ops

NUM_JOINTS = 18


def joint_lists(heat, thre1=0.1, factor=4, scales=None, max_peaks=1024):
    """heat: CUDA fp32 [B, C>=18, H, W].  Returns one float64 [P_b, 5] array per image with the reference's rows
    (x, y, score, id, joint_type), x / y multiplied by scales[b] (joint_utils.py:146-147)."""
    rows, count = ops.heatmap_peaks(heat, thre1=thre1, factor=factor, max_peaks=max_peaks, channels=NUM_JOINTS)
    cnt = count.cpu().tolist()
    if max(cnt) > max_peaks:  # rare: more peaks than the fast-path capacity -> redo with room (no truncation)
        rows, count = ops.heatmap_peaks(heat, thre1=thre1, factor=factor, max_peaks=max(cnt), channels=NUM_JOINTS)
    rows = rows.cpu().numpy().astype(np.float64)
    out = []
    for b, n in enumerate(cnt):
        r = rows[b, :n].copy()
        if scales is not None:
            r[:, :2] *= float(scales[b])
        out.append(r)
    return out


def get_joint_list(img_orig, param, heatmaps, scale):
    """Same signature and result as the reference's get_joint_list (joint_utils.py:141-152).

    heatmaps: the numpy [H, W, 18] array the reference passes (uploaded again), or -- to skip the round trip -- the CUDA
    tensor [18(+), H, W] / [1, 18(+), H, W] straight from poseNet.forward."""
    if isinstance(heatmaps, np.ndarray):
        h = torch.from_numpy(np.ascontiguousarray(heatmaps.transpose(2, 0, 1), dtype=np.float32)).cuda()[None]
    else:
        h = heatmaps if heatmaps.dim() == 4 else heatmaps[None]
        if h.shape[0] != 1:
            raise ValueError("get_joint_list takes one image; use joint_lists for a batch")
    factor = img_orig.shape[0] / float(h.shape[2])  # joint_utils.py:143-144
    if factor != int(factor) or not 1 <= int(factor) <= 8:
        raise NotImplementedError("heat-map peak refinement is built for integer upsampling factors 1..8 (got %r)" % factor)
    return joint_lists(h.float(), thre1=param["thre1"], factor=int(factor), scales=[scale])[0]
