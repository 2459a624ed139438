"""Drop-in for the reference's network/losses.py (RetinaNet focal + smooth-L1 loss of the detection subnet).

Mirror of /root/reference/network/losses.py: `calc_iou(a, b)` :5-24 and `FocalLoss.forward(classifications, regressions,
anchors, annotations)` :27-137, same arguments and return values ((classification_loss [1], regression_loss [1]), the means of
the per-image losses).  The reference walks the batch in a Python loop of ~40 small torch kernels per image; here one
`mpn_focal_loss` call (csrc/mpn_focal.cu) does the anchor <-> annotation assignment, both losses and their gradients for the
whole batch, and `loss.backward()` of the reference training loop (training/trainer.py:245-259) reaches it through a
torch.autograd.Function.  CUDA tensors only (no CPU path).
"""
import torch
import torch.nn as nn

from .. import ops


def calc_iou(a, b):
    """losses.py:5-24: IoU of every box of a [N,4] with every box of b [M,4] (no +1 convention, union clamped at 1e-8).
    Elementwise torch expression kept for API compatibility; FocalLoss evaluates the same arithmetic inside its kernel."""
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    iw = torch.min(torch.unsqueeze(a[:, 2], dim=1), b[:, 2]) - torch.max(torch.unsqueeze(a[:, 0], 1), b[:, 0])
    ih = torch.min(torch.unsqueeze(a[:, 3], dim=1), b[:, 3]) - torch.max(torch.unsqueeze(a[:, 1], 1), b[:, 1])
    iw = torch.clamp(iw, min=0)
    ih = torch.clamp(ih, min=0)
    ua = torch.unsqueeze((a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]), dim=1) + area - iw * ih
    ua = torch.clamp(ua, min=1e-8)
    return iw * ih / ua


class _FocalLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, classifications, regressions, anchors, annotations):
        need = classifications.requires_grad or regressions.requires_grad
        cl, rl, dcls, dreg = ops.focal_loss(classifications.detach(), regressions.detach(), anchors.detach(), annotations.detach(),
                                            want_grads=need)
        ctx.save_for_backward(dcls, dreg) if need else None
        ctx.need = need
        return cl.mean(dim=0, keepdim=True), rl.mean(dim=0, keepdim=True)   # losses.py:137

    @staticmethod
    def backward(ctx, g_cls, g_reg):
        if not ctx.need:
            return None, None, None, None
        dcls, dreg = ctx.saved_tensors
        # the kernel's gradients are those of mean(cls_loss) and mean(reg_loss) with unit upstream gradients
        return dcls * g_cls.reshape(()), dreg * g_reg.reshape(()), None, None


class FocalLoss(nn.Module):
    def forward(self, classifications, regressions, anchors, annotations):
        if not (classifications.is_cuda and regressions.is_cuda):
            raise RuntimeError("FocalLoss runs on the device (libmpn_b200): CUDA tensors expected, there is no CPU path")
        annotations = annotations.to(device=classifications.device, dtype=torch.float32)
        anchors = anchors.to(device=classifications.device, dtype=torch.float32)
        return _FocalLossFn.apply(classifications.float(), regressions.float(), anchors, annotations)
