"""`pth_nms(dets, thresh)` with the reference's contract (lib/nms/pth_nms.py:5-44).

dets: float32 [N,5] tensor (x1,y1,x2,y2,score).  Returns a LongTensor of kept indices into dets in
descending-score order.  CUDA dets follow the reference's GPU branch (suppress at IoU > thresh,
nms_kernel.cu:63); CPU dets follow its CPU branch (ovr >= thresh, nms.c:59) but are still computed by the
CUDA kernels (copied to the current device and back) -- there is no host implementation.
The reference entry points return 1 on success (nms.c:68); the C ABI returns 0 and raises here otherwise.
"""
import torch

from ... import ops as _ops


def pth_nms(dets, thresh):
    if dets.dim() != 2 or dets.size(1) != 5:
        raise RuntimeError("dets must be [N,5]")
    if dets.is_cuda:
        with torch.cuda.device(dets.device):
            return _ops.nms(dets.float(), float(thresh), ge=False)
    if not torch.cuda.is_available():
        raise RuntimeError("pth_nms: no CUDA device (libmpn_b200 has no CPU path)")
    keep = _ops.nms(dets.float().cuda(), float(thresh), ge=True)
    return keep.cpu()
