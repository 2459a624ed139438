"""torch restatement of the reference's detection loss (TEST INFRASTRUCTURE): network/losses.py:5-24 (calc_iou) and :27-137
(FocalLoss.forward), with the reference's operation order, differentiable through torch autograd.  Pinned by
tests/golden/focal_loss.npz = the values and gradients of the reference's own FocalLoss run in this container
(oracle/make_goldens.py focal_golden)."""
import numpy as np
import torch


def calc_iou(a, b):
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    iw = torch.min(torch.unsqueeze(a[:, 2], dim=1), b[:, 2]) - torch.max(torch.unsqueeze(a[:, 0], 1), b[:, 0])
    ih = torch.min(torch.unsqueeze(a[:, 3], dim=1), b[:, 3]) - torch.max(torch.unsqueeze(a[:, 1], 1), b[:, 1])
    iw = torch.clamp(iw, min=0)
    ih = torch.clamp(ih, min=0)
    ua = torch.unsqueeze((a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]), dim=1) + area - iw * ih
    ua = torch.clamp(ua, min=1e-8)
    return iw * ih / ua


def focal_loss(classifications, regressions, anchors, annotations, alpha=0.25, gamma=2.0):
    """Returns (classification_loss [1], regression_loss [1], per-image lists)."""
    dev = classifications.device
    cls_losses, reg_losses = [], []
    anchor = anchors[0]
    aw, ah = anchor[:, 2] - anchor[:, 0], anchor[:, 3] - anchor[:, 1]
    acx, acy = anchor[:, 0] + 0.5 * aw, anchor[:, 1] + 0.5 * ah
    for j in range(classifications.shape[0]):
        classification, regression = classifications[j], regressions[j]
        ann = annotations[j]
        ann = ann[ann[:, 4] != -1]
        if ann.shape[0] == 0:                                                   # :53-57
            reg_losses.append(torch.zeros((), device=dev))
            cls_losses.append(torch.zeros((), device=dev))
            continue
        classification = torch.clamp(classification, 1e-4, 1.0 - 1e-4)           # :59
        iou = calc_iou(anchor, ann[:, :4])
        iou_max, iou_arg = torch.max(iou, dim=1)
        targets = torch.ones_like(classification) * -1
        targets[iou_max < 0.4, :] = 0
        pos = iou_max >= 0.5
        npos = pos.sum()
        assigned = ann[iou_arg]
        targets[pos, :] = 0
        targets[pos, assigned[pos, 4].long()] = 1
        alpha_factor = torch.where(targets == 1., torch.full_like(targets, alpha), torch.full_like(targets, 1. - alpha))
        focal_weight = torch.where(targets == 1., 1. - classification, classification)
        focal_weight = alpha_factor * torch.pow(focal_weight, gamma)
        bce = -(targets * torch.log(classification) + (1.0 - targets) * torch.log(1.0 - classification))
        cls_loss = focal_weight * bce
        cls_loss = torch.where(targets != -1.0, cls_loss, torch.zeros_like(cls_loss))
        cls_losses.append(cls_loss.sum() / torch.clamp(npos.float(), min=1.0))   # :94
        if npos > 0:
            asg = assigned[pos]
            gw, gh = asg[:, 2] - asg[:, 0], asg[:, 3] - asg[:, 1]
            gcx, gcy = asg[:, 0] + 0.5 * gw, asg[:, 1] + 0.5 * gh
            gw, gh = torch.clamp(gw, min=1), torch.clamp(gh, min=1)
            t = torch.stack(((gcx - acx[pos]) / aw[pos], (gcy - acy[pos]) / ah[pos], torch.log(gw / aw[pos]), torch.log(gh / ah[pos]))).t()
            t = t / torch.tensor([[0.1, 0.1, 0.2, 0.2]], device=dev)
            diff = torch.abs(t - regression[pos])
            rl = torch.where(diff <= 1.0 / 9.0, 0.5 * 9.0 * torch.pow(diff, 2), diff - 0.5 / 9.0)
            reg_losses.append(rl.mean())
        else:
            reg_losses.append(torch.zeros((), device=dev))
    return (torch.stack(cls_losses).mean(dim=0, keepdim=True), torch.stack(reg_losses).mean(dim=0, keepdim=True),
            cls_losses, reg_losses)


def focal_case(seed=0, hw=(96, 128), batch=3, max_ann=6, with_empty=True):
    """Seeded inputs: anchors of an hw image, sigmoid-like scores, regressions, annotations (rows padded with -1; with_empty:
    image 1 has no annotation at all -- losses.py:53-57, a branch that only runs on torch 0.4: `torch.tensor(0,
    requires_grad=True)` is an error on torch >= 1.0, so the golden made from the live reference uses with_empty=False)."""
    from .anchors_oracle import anchors_for_image
    rng = np.random.Generator(np.random.PCG64(seed))
    anchors = anchors_for_image(hw[0], hw[1])[None].astype(np.float32)
    A = anchors.shape[1]
    cls = (1.0 / (1.0 + np.exp(-rng.normal(-2.0, 2.5, (batch, A, 1))))).astype(np.float32)
    cls[0, :5, 0] = [0.0, 1.0, 5e-5, 1.0 - 5e-5, 0.5]                       # outside / at the clamp range
    reg = rng.normal(0, 1.0, (batch, A, 4)).astype(np.float32)
    ann = -np.ones((batch, max_ann, 5), dtype=np.float32)
    for b in range(batch):
        n = 0 if (b == 1 and with_empty) else int(rng.integers(2, max_ann + 1))
        for i in range(n):
            w, h = rng.uniform(12, 70), rng.uniform(16, 90)
            x, y = rng.uniform(0, hw[1] - w), rng.uniform(0, hw[0] - h)
            ann[b, i] = [x, y, x + w, y + h, 0]
    ann[batch - 1, 1, 2:4] = ann[batch - 1, 1, 0:2] + np.float32(0.4)                        # a sub-pixel box: gt width / height clamp at 1
    return cls, reg, anchors, ann
