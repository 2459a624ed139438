"""Training step of the detection subnet on the C ABI.

Reference: training/multipose_detection_train.py freezes the ResNet trunk, the keypoint neck / head and the PRN (:64-79) and trains
the RetinaNet neck (conv6, conv7, latlayer1-3, toplayer0-2) and the two towers with Adam; Trainer puts the model in train() mode
and then freezes every BatchNorm (training/trainer.py:170-174: `freeze_bn()` for every subnet but the keypoint one), so the trunk
is the eval-mode forward and no batch statistics exist anywhere on this path.  The step is
    out, saved = model([img, 'detection_subnet'])                                  posenet.py:320-335
    loss, log = model.module.build_loss(saved, 'detection_subnet', annotations)    posenet.py:405-424, network/losses.py:27-137
    loss.backward(); optimizer.step()                                              trainer.py:245-259

Here: the trunk runs on the inference engine (folded BatchNorm, bf16x3 planes: the weight-gradient kernel reads bf16), the
detection neck and towers run with saved activations, and the backward is an explicit reverse schedule of libmpn_b200 launches
(data gradients = the forward tcgen05 kernel on flipped filters, weight gradients = wgrad_tc_kernel, bias gradients = channel
sums, ReLU masks, block sums for the nearest-upsample adds).  `DetectionTrainFunction` exposes it to autograd so the reference
loop above runs unchanged; the loss is `network.losses.FocalLoss` (one kernel for the batch).
"""
import torch

from . import ops
from . import train_ops as T
from ._lib import OUT_F32_NHWC

NECK = ("conv6", "conv7", "latlayer1", "latlayer2", "latlayer3", "toplayer0", "toplayer1", "toplayer2")
TOWER = ("conv1", "conv2", "conv3", "conv4", "output")


class _Saved(object):
    pass


class DetectionTrainEngine(object):
    def __init__(self, model, precision="bf16x3"):
        if precision not in ("bf16x3", "bf16"):
            raise ValueError("training runs on the tcgen05 path: precision must be bf16x3 or bf16")
        self.model, self.precision, self.fmt = model, precision, ops.PRECISIONS[precision]
        self.last_saved = None

    # ------------------------------------------------------------------ parameters
    def trainable_parameters(self):
        """(name, parameter) of the detection neck and towers that require a gradient, in a fixed order."""
        m = self.model
        out = []
        for n in NECK:
            for pn, p in getattr(m.fpn, n).named_parameters():
                out.append(("fpn.%s.%s" % (n, pn), p))
        for h in ("regressionModel", "classificationModel"):
            for n in TOWER:
                for pn, p in getattr(getattr(m, h), n).named_parameters():
                    out.append(("%s.%s.%s" % (h, n, pn), p))
        return [(n, p) for n, p in out if p.requires_grad]

    def check_frozen(self):
        m = self.model
        det = {id(p) for _, p in self.trainable_parameters()}
        loose = [n for n, p in m.named_parameters() if p.requires_grad and id(p) not in det]
        if loose:
            raise NotImplementedError("detection-subnet training is built for the reference's configuration (trunk, keypoint subnet "
                                      "and PRN frozen, training/multipose_detection_train.py:64-79); these parameters still require "
                                      "gradients: %s ..." % ", ".join(loose[:4]))
        if any(b.training for b in m.modules() if isinstance(b, torch.nn.BatchNorm2d)):
            raise NotImplementedError("detection-subnet training runs with frozen BatchNorm (training/trainer.py:173-174 calls "
                                      "freeze_bn() for this subnet); call model.freeze_bn() after model.train()")

    def _conv(self, x, conv, **kw):
        return ops.conv2d(x, ops.pack_conv(conv.weight, conv.bias, None, self.fmt), stride=conv.stride[0], pad=conv.padding[0], **kw)

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, img):
        m, f = self.model, self.model.fpn
        S = _Saved()
        eng = m.engine(self.precision)
        eng._ensure_packed()
        _, S.c3, S.c4, S.c5 = eng.backbone(img)                                  # frozen trunk, eval-mode BatchNorm folded
        # detection neck (fpn.py:107-114)
        S.p6 = self._conv(S.c5, f.conv6)
        S.r6 = ops.relu(S.p6)
        S.p7 = self._conv(S.r6, f.conv7)
        S.p5 = self._conv(S.c5, f.latlayer1)
        S.p4 = self._conv(S.c4, f.latlayer2, up=S.p5)
        S.p3 = self._conv(S.c3, f.latlayer3, up=S.p4)
        S.s5 = self._conv(S.p5, f.toplayer0)
        S.s4 = self._conv(S.p4, f.toplayer1)
        S.s3 = self._conv(S.p3, f.toplayer2)
        for a, b in ((S.p4, S.p5), (S.p3, S.p4)):
            if a.H != 2 * b.H or a.W != 2 * b.W:
                raise NotImplementedError("detection training needs image sizes whose pyramid levels halve exactly (multiples of 32: "
                                          "the reference trains at 608x608); got levels %dx%d <- %dx%d" % (a.H, a.W, b.H, b.W))
        S.feats = [S.s3, S.s4, S.s5, S.p6, S.p7]
        # towers (posenet.py:33-117, 262-263); outputs land in the concatenated [B, A, 1 | 4] tensors
        B, dev = img.shape[0], img.device
        cells = [t.H * t.W for t in S.feats]
        A = 9 * sum(cells)
        S.cls = torch.empty((B, A, 1), dtype=torch.float32, device=dev)
        S.reg = torch.empty((B, A, 4), dtype=torch.float32, device=dev)
        S.h = {}
        for head, hname, out, per, sig in ((m.regressionModel, "reg", S.reg, 4, False), (m.classificationModel, "cls", S.cls, 1, True)):
            pcs = [ops.pack_conv(getattr(head, n).weight, getattr(head, n).bias, None, self.fmt) for n in TOWER]
            off = 0
            for li, t in enumerate(S.feats):
                hs, o = [], t
                for i in range(4):
                    o = ops.conv2d(o, pcs[i], pad=1, relu=True)
                    hs.append(o)
                ops.conv2d(o, pcs[4], pad=1, sigmoid=sig, out_mode=OUT_F32_NHWC, out_tensor=out, out_elem_offset=off * per,
                           out_cstride=9 * per, out_nstride=A * per)
                S.h[(hname, li)] = hs
                off += cells[li] * 9
        S.cells, S.A = cells, A
        H, W = img.shape[2], img.shape[3]
        S.anchors = ops.anchors_for(H, W, dev)
        self.last_saved = S
        return (S.cls, S.reg, S.anchors), S

    # ------------------------------------------------------------------ backward
    def _out_grad_act(self, g, off, cells, hw, per):
        """Slice of d loss / d (tower output) of one pyramid level, [B, A, per] fp32 -> NHWC Act with the 9*per channels padded to 64."""
        B = g.shape[0]
        c = 9 * per
        lv = g.reshape(B, -1)[:, off * per:(off + cells * 9) * per].reshape(B, hw[0], hw[1], c).permute(0, 3, 1, 2)
        return ops.act_from_nchw(torch.nn.functional.pad(lv, (0, 0, 0, 0, 0, 64 - c)).contiguous(), self.fmt)

    @torch.no_grad()
    def backward(self, S, dcls, dreg):
        """dcls [B,A,1] = d loss / d (sigmoid scores), dreg [B,A,4].  Returns {state_dict name: gradient}."""
        m, f = self.model, self.model.fpn
        grads = {}

        def acc(name, g):
            grads[name] = g if name not in grads else grads[name].add_(g)   # shared tower weights: sum over the 5 levels

        dlogit = dcls.float() * S.cls * (1.0 - S.cls)                        # through the fused sigmoid (posenet.py:109)
        d_feat = [None] * 5
        for head, hname, hkey, g, per in ((m.regressionModel, "regressionModel", "reg", dreg.float(), 4),
                                          (m.classificationModel, "classificationModel", "cls", dlogit, 1)):
            convs = [getattr(head, n) for n in TOWER]
            dpc = [T.pack_dgrad_filter(c.weight, self.fmt, cout_pad=64 if i == 4 else None) for i, c in enumerate(convs)]
            off = 0
            for li, t in enumerate(S.feats):
                hs = S.h[(hkey, li)]
                d = self._out_grad_act(g, off, S.cells[li], (t.H, t.W), per)
                acc("%s.output.weight" % hname, T.unpack_filter_grad(T.conv_wgrad(hs[3], d, 9 * per, 3, 3, 1, 1)))
                acc("%s.output.bias" % hname, T.channel_sum(d, C=9 * per))
                d = T.conv_dgrad(d, convs[4].weight, 1, 1, (t.H, t.W), self.fmt, pc=dpc[4])
                for i in (3, 2, 1, 0):
                    d = T.relu_backward(d, hs[i])
                    x_in = hs[i - 1] if i > 0 else t
                    acc("%s.%s.weight" % (hname, TOWER[i]), T.unpack_filter_grad(T.conv_wgrad(x_in, d, 256, 3, 3, 1, 1)))
                    acc("%s.%s.bias" % (hname, TOWER[i]), T.channel_sum(d))
                    # the second tower's contribution to a level's features rides on the residual input of the last data gradient
                    d = T.conv_dgrad(d, convs[i].weight, 1, 1, (t.H, t.W), self.fmt, pc=dpc[i], residual=d_feat[li] if i == 0 else None)
                d_feat[li] = d
                off += S.cells[li] * 9
        d_s3, d_s4, d_s5, d_p6, d_p7 = d_feat

        def conv_grads(name, conv, x, dy, need_dx=True, dx_hw=None):
            Cout, _, R, Sx = conv.weight.shape
            grads["fpn.%s.weight" % name] = T.unpack_filter_grad(T.conv_wgrad(x, dy, Cout, R, Sx, conv.stride[0], conv.padding[0]))
            grads["fpn.%s.bias" % name] = T.channel_sum(dy, C=Cout)
            if not need_dx:
                return None
            return T.conv_dgrad(dy, conv.weight, conv.stride[0], conv.padding[0], dx_hw or (x.H, x.W), self.fmt)
        # p7 = conv7(relu(p6)); p6 = conv6(c5)  (fpn.py:108-109); the trunk is frozen: no data gradient into c3 / c4 / c5
        d_r6 = conv_grads("conv7", f.conv7, S.r6, d_p7)
        d_p6 = T.add(d_p6, T.relu_backward(d_r6, S.r6))
        conv_grads("conv6", f.conv6, S.c5, d_p6, need_dx=False)
        # smooth convs, then the top-down pathway (fpn.py:110-114)
        d_p3 = conv_grads("toplayer2", f.toplayer2, S.p3, d_s3)
        d_p4 = conv_grads("toplayer1", f.toplayer1, S.p4, d_s4)
        d_p5 = conv_grads("toplayer0", f.toplayer0, S.p5, d_s5)
        conv_grads("latlayer3", f.latlayer3, S.c3, d_p3, need_dx=False)
        d_p4 = T.add(d_p4, T.block_sum(d_p3, 2))
        conv_grads("latlayer2", f.latlayer2, S.c4, d_p4, need_dx=False)
        d_p5 = T.add(d_p5, T.block_sum(d_p4, 2))
        conv_grads("latlayer1", f.latlayer1, S.c5, d_p5, need_dx=False)
        return grads


class DetectionTrainFunction(torch.autograd.Function):
    """model([img, 'detection_subnet']) under torch.enable_grad() in train mode with frozen BatchNorm: class scores and box
    regressions come out as autograd-tracked tensors, so `build_loss(...).backward()` of the reference loop reaches the kernels."""

    @staticmethod
    def forward(ctx, engine, img, *params):
        (cls, reg, anchors), S = engine.forward(img)
        ctx.engine, ctx.saved, ctx.names = engine, S, [n for n, _ in engine.trainable_parameters()]
        ctx.mark_non_differentiable(anchors)
        return cls, reg, anchors

    @staticmethod
    def backward(ctx, dcls, dreg, _danchors):
        S = ctx.saved
        dev = S.cls.device
        dcls = torch.zeros_like(S.cls) if dcls is None else dcls.contiguous()
        dreg = torch.zeros_like(S.reg) if dreg is None else dreg.contiguous()
        with torch.cuda.device(dev):
            grads = ctx.engine.backward(S, dcls, dreg)
        return (None, None) + tuple(grads.get(n) for n in ctx.names)
