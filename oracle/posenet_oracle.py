"""Functional fp32 restatement of the reference forward (TEST INFRASTRUCTURE).

Every function takes a flat state dict `sd` (name -> torch tensor, the reference's
state_dict keys) and mirrors one reference routine; citations are into /root/reference.
torch is used here only as the fp32 array library the reference itself calls
(conv2d / batch_norm / max_pool2d are the third-party arithmetic, see oracle/__init__).

Pinned against the real reference modules by oracle/make_goldens.py (container-side) and
tests/test_oracle.py; tests/golden/*.npz hold the reference's outputs.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .anchors_oracle import anchors_for_image
from .nms_oracle import nms_gpu_semantics

BLOCKS = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}


def _conv(sd, name, x, stride=1, pad=0):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=pad)


RELU_MASKS = None  # iterator of {0,1} tensors: _relu(x) = x * mask (lets a test impose ANOTHER implementation's ReLU
                   # pattern so that a handful of sign flips at |x| ~ 1e-4 do not dominate a max-norm gradient comparison)
POOL_INDICES = None  # same idea for the stem max-pool: impose another implementation's arg-max selection
TRACE = None       # set to a dict to record block outputs (tests / debugging only)
_BN_TRAIN = False  # forward_train_keypoint() flips this: batch statistics, as under model.train() (trainer.py:170-174)


def _bn(sd, name, x):
    # BatchNorm2d, eps 1e-5 (fpn.py:15); eval mode (poseNet.freeze_bn posenet.py:220-224) unless _BN_TRAIN
    if _BN_TRAIN:
        return F.batch_norm(x, None, None, sd[name + ".weight"], sd[name + ".bias"], True, 0.0, 1e-5)
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"],
                        sd[name + ".weight"], sd[name + ".bias"], False, 0.0, 1e-5)


def _relu(x):
    if RELU_MASKS is None:
        return F.relu(x)
    m = next(RELU_MASKS)
    assert m.shape == x.shape, (m.shape, x.shape)
    return x * m.to(x.dtype)


def bottleneck(sd, p, x, stride):
    """fpn.py:28-34."""
    out = _relu(_bn(sd, p + "bn1", _conv(sd, p + "conv1", x)))
    out = _relu(_bn(sd, p + "bn2", _conv(sd, p + "conv2", out, stride=stride, pad=1)))
    out = _bn(sd, p + "bn3", _conv(sd, p + "conv3", out))
    if (p + "downsample.0.weight") in sd:
        sc = _bn(sd, p + "downsample.1", _conv(sd, p + "downsample.0", x, stride=stride))
    else:
        sc = x
    return _relu(out + sc)


def upsample_add(x, y):
    """fpn.py:84-95: nearest upsample of x to y's size, plus y."""
    return F.interpolate(x, size=y.shape[2:], mode="nearest") + y


def backbone(sd, layers, x, pre="fpn."):
    """fpn.py:99-105."""
    c1 = _relu(_bn(sd, pre + "bn1", _conv(sd, pre + "conv1", x, stride=2, pad=3)))
    if POOL_INDICES is None:
        c = F.max_pool2d(c1, kernel_size=3, stride=2, padding=1)
    else:
        idx = POOL_INDICES
        c = c1.flatten(2).gather(2, idx.flatten(2)).view(idx.shape)
    feats = []
    for li, (nblk, stride) in enumerate(zip(BLOCKS[layers], (1, 2, 2, 2)), start=1):
        for b in range(nblk):
            c = bottleneck(sd, "%slayer%d.%d." % (pre, li, b), c, stride if b == 0 else 1)
            if TRACE is not None and c.requires_grad:  # diagnostics: keep each block output and its gradient
                c.retain_grad()
                TRACE["%slayer%d.%d" % (pre, li, b)] = c
        feats.append(c)
    return feats  # c2, c3, c4, c5


def detection_neck(sd, c3, c4, c5, pre="fpn."):
    """fpn.py:107-114."""
    p6 = _conv(sd, pre + "conv6", c5, stride=2, pad=1)
    p7 = _conv(sd, pre + "conv7", F.relu(p6), stride=2, pad=1)
    p5 = _conv(sd, pre + "latlayer1", c5)
    p4 = upsample_add(p5, _conv(sd, pre + "latlayer2", c4))
    p3 = upsample_add(p4, _conv(sd, pre + "latlayer3", c3))
    p5 = _conv(sd, pre + "toplayer0", p5, pad=1)
    p4 = _conv(sd, pre + "toplayer1", p4, pad=1)
    p3 = _conv(sd, pre + "toplayer2", p3, pad=1)
    return [p3, p4, p5, p6, p7]


def keypoint_neck(sd, c2, c3, c4, c5, pre="fpn."):
    """fpn.py:117-124."""
    fp5 = _conv(sd, pre + "toplayer", c5)
    fp4 = upsample_add(fp5, _conv(sd, pre + "flatlayer1", c4))
    fp3 = upsample_add(fp4, _conv(sd, pre + "flatlayer2", c3))
    fp2 = upsample_add(fp3, _conv(sd, pre + "flatlayer3", c2))
    fp4 = _conv(sd, pre + "smooth1", fp4, pad=1)
    fp3 = _conv(sd, pre + "smooth2", fp3, pad=1)
    fp2 = _conv(sd, pre + "smooth3", fp2, pad=1)
    return [fp2, fp3, fp4, fp5]


def keypoint_head(sd, p2, p3, p4, p5):
    """posenet.py:243-257 / 302-315."""
    q5 = _conv(sd, "convs1", _conv(sd, "convt1", p5, pad=1), pad=1)
    q4 = _conv(sd, "convs2", _conv(sd, "convt2", p4, pad=1), pad=1)
    q3 = _conv(sd, "convs3", _conv(sd, "convt3", p3, pad=1), pad=1)
    q2 = _conv(sd, "convs4", _conv(sd, "convt4", p2, pad=1), pad=1)
    q5 = F.interpolate(q5, scale_factor=8, mode="nearest")
    q4 = F.interpolate(q4, scale_factor=4, mode="nearest")
    q3 = F.interpolate(q3, scale_factor=2, mode="nearest")
    cat = torch.cat((q5, q4, q3, q2), 1)
    return _conv(sd, "convfin", _relu(_conv(sd, "conv2", cat, pad=1)))


def intermediate_heads(sd, p2, p3, p4, p5):
    """posenet.py:296-299."""
    return [
        _conv(sd, "convfin_k2", p2),
        F.interpolate(_conv(sd, "convfin_k3", p3), scale_factor=2, mode="nearest"),
        F.interpolate(_conv(sd, "convfin_k4", p4), scale_factor=4, mode="nearest"),
        F.interpolate(_conv(sd, "convfin_k5", p5), scale_factor=8, mode="nearest"),
    ]


def retina_head(sd, head, feat, sigmoid):
    """posenet.py:51-69 (RegressionModel.forward) / :96-117 (ClassificationModel.forward)."""
    out = feat
    for n in ("conv1", "conv2", "conv3", "conv4"):
        out = F.relu(_conv(sd, "%s.%s" % (head, n), out, pad=1))
    out = _conv(sd, head + ".output", out, pad=1)
    if sigmoid:
        out = torch.sigmoid(out)
    out = out.permute(0, 2, 3, 1).contiguous()
    return out.view(out.shape[0], -1, 1 if sigmoid else 4)


def detection_heads(sd, feats):
    """posenet.py:262-263."""
    reg = torch.cat([retina_head(sd, "regressionModel", f, False) for f in feats], dim=1)
    cls = torch.cat([retina_head(sd, "classificationModel", f, True) for f in feats], dim=1)
    return cls, reg


def decode_boxes(anchors, deltas):
    """network/utils.py:19-43 (BBoxTransform.forward), mean 0, std [.1,.1,.2,.2]."""
    std = torch.tensor([0.1, 0.1, 0.2, 0.2], dtype=torch.float32, device=deltas.device)
    widths = anchors[:, :, 2] - anchors[:, :, 0]
    heights = anchors[:, :, 3] - anchors[:, :, 1]
    ctr_x = anchors[:, :, 0] + 0.5 * widths
    ctr_y = anchors[:, :, 1] + 0.5 * heights
    dx = deltas[:, :, 0] * std[0] + 0
    dy = deltas[:, :, 1] * std[1] + 0
    dw = deltas[:, :, 2] * std[2] + 0
    dh = deltas[:, :, 3] * std[3] + 0
    pcx = ctr_x + dx * widths
    pcy = ctr_y + dy * heights
    pw = torch.exp(dw) * widths
    ph = torch.exp(dh) * heights
    return torch.stack([pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph], dim=2)


def clip_boxes(boxes, height, width):
    """network/utils.py:51-61 (only these four clamps)."""
    boxes = boxes.clone()
    boxes[:, :, 0] = torch.clamp(boxes[:, :, 0], min=0)
    boxes[:, :, 1] = torch.clamp(boxes[:, :, 1], min=0)
    boxes[:, :, 2] = torch.clamp(boxes[:, :, 2], max=width)
    boxes[:, :, 3] = torch.clamp(boxes[:, :, 3], max=height)
    return boxes


def postprocess_image(cls_i, boxes_i, score_thresh=0.05, iou_thresh=0.5, ge=False):
    """posenet.py:269-285 for ONE image: filter > 0.05, NMS, gather.

    cls_i [A,1], boxes_i [A,4] (decoded + clipped).  Returns (scores[K], classes[K] int64,
    boxes[K,4], keep_idx[K] int64 = indices into the filtered set, descending score)."""
    scores = cls_i.max(dim=1)[0]
    m = scores > score_thresh
    if int(m.sum()) == 0:
        z = torch.zeros(0)
        return z, torch.zeros(0, dtype=torch.int64), torch.zeros(0, 4), torch.zeros(0, dtype=torch.int64)
    fcls, fbox, fsc = cls_i[m], boxes_i[m], scores[m]
    dets = torch.cat([fbox, fsc[:, None]], dim=1).cpu().numpy().astype(np.float32)
    keep = torch.from_numpy(nms_gpu_semantics(dets, iou_thresh, ge=ge))
    ksc, kcl = fcls[keep].max(dim=1)
    return ksc, kcl, fbox[keep], keep


def forward(sd, layers, img, subnet_name):
    """posenet.py:226-285 dispatch.  img fp32 NCHW."""
    H, W = img.shape[2:]
    if subnet_name == "prn_subnet":
        return prn_forward(sd, img)
    c2, c3, c4, c5 = backbone(sd, layers, img)
    if subnet_name == "keypoint_subnet":
        p2, p3, p4, p5 = keypoint_neck(sd, c2, c3, c4, c5)
        saved = intermediate_heads(sd, p2, p3, p4, p5)
        heat = keypoint_head(sd, p2, p3, p4, p5)
        saved.append(heat)
        return heat, saved
    if subnet_name == "detection_subnet":
        feats = detection_neck(sd, c3, c4, c5)
        cls, reg = detection_heads(sd, feats)
        anchors = torch.from_numpy(anchors_for_image(H, W))[None].to(img.device)
        return [], [cls, reg, anchors]
    # entire_net
    p2, p3, p4, p5 = keypoint_neck(sd, c2, c3, c4, c5)
    heat = keypoint_head(sd, p2, p3, p4, p5)
    feats = detection_neck(sd, c3, c4, c5)
    cls, reg = detection_heads(sd, feats)
    anchors = torch.from_numpy(anchors_for_image(H, W))[None].to(img.device)
    boxes = clip_boxes(decode_boxes(anchors, reg), H, W)
    ksc, kcl, kbox, _ = postprocess_image(cls[0], boxes[0])
    return heat, [ksc, kcl, kbox], dict(cls=cls, reg=reg, boxes=boxes)


def forward_train_keypoint(sd, layers, img):
    """keypoint_subnet forward with BatchNorm in TRAIN mode (the reference's training configuration,
    training/trainer.py:170-174).  Returns the 5 supervised maps [k2, k3, k4, k5, heat]."""
    global _BN_TRAIN
    _BN_TRAIN = True
    try:
        heat, saved = forward(sd, layers, img, "keypoint_subnet")
    finally:
        _BN_TRAIN = False
    return saved


def keypoint_loss(saved, heat_gt, heat_weight):
    """posenet.py:367-403: sum over the 5 maps of MSELoss(mean)(pred[:, :18] * w, w * gt)."""
    total = 0
    for s in saved:
        total = total + F.mse_loss(s[:, :18] * heat_weight, heat_weight * heat_gt)
    return total


def prn_forward(sd, x):
    """posenet.py:337-350 (eval mode: dropout is identity)."""
    res = x.reshape(x.shape[0], -1)
    out = F.relu(F.linear(res, sd["prn.dens1.weight"], sd["prn.dens1.bias"]))
    out = F.relu(F.linear(out, sd["prn.bneck.weight"], sd["prn.bneck.bias"]))
    out = F.relu(F.linear(out, sd["prn.dens2.weight"], sd["prn.dens2.bias"]))
    out = torch.softmax(out + res, dim=1)
    out = out.view(x.shape[0], x.shape[1], x.shape[2], 17)
    return out, [out]


def conv_flops_entire(layers, H=480, W=640):
    from multiposenet.pytorch_b200.synthetic import conv_flops_entire as f
    return f(layers, H, W)
