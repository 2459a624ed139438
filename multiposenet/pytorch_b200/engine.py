"""Layer schedule of the hot path on top of the C ABI (the host half of the B200 design).

Mirrors the reference call graph -- FPN.forward (network/fpn.py:97-126), the keypoint / detection /
entire_net branches of poseNet.forward (network/posenet.py:226-335) -- as a flat sequence of
libmpn_b200 launches on NHWC activations:

  stem     : NCHW->NHWC, 7x7/2 conv + BN + ReLU (CUDA-core kernel, fp32 in), 3x3/2 max-pool
  backbone : per Bottleneck 3 launches (4 with a projection shortcut); BN folded to scale/bias,
             residual add + ReLU fused in conv3's epilogue
  necks    : lateral 1x1 with the nearest-upsample-add fused in its epilogue, then 3x3 smooth
  kp head  : convt/convs per level; the x8 / x4 quarters of conv2 are evaluated at low resolution as phase-class
             convolutions and gathered in conv2's epilogue, the x2 and native quarters go through a 256-channel
             concat buffer (replication fused in the store); conv2+ReLU; convfin writes fp32 NCHW directly
  det head : shared tower on 5 levels; outputs land in the concatenated [B, A, 1|4] tensors
  post     : decode+clip, filter, radix sort, bit-mask NMS, gather -- all on the device

Precision modes (env MPN_PRECISION or Engine(precision=...)):
  "f16f8" (default)  tcgen05, fp16 plane + two fp8 cross-term planes, 8 MMA slots / K block: fp32-grade parity (<1e-3);
                     activations saturate at +-65504 (Engine.check_range / MPN_RANGE_CHECK=1 reports a clamped layer)
  "bf16x3"           tcgen05, hi/lo bf16 split, 3 MMAs / K-step: fp32-grade parity (<1e-3), full fp32 range
  "bf16"             tcgen05, single bf16 plane: fastest, ~1e-2 relative error after 100 layers
  "fp32"             CUDA-core fp32 FMA kernel: exact-mode reference on the device
"""
import os

import torch

from . import ops
from ._lib import FMT_BF16, FMT_BF16X2, FMT_F32, OUT_ACT, OUT_F32_NCHW, OUT_F32_NHWC

# Inference default: f16f8 (fp16 hi plane + two fp8 cross-term planes, 8 MMA slots per K block) -- within 3.3e-4 of the fp32
# reference on every output (profiles/r01s_parity_margin.txt; the bar is 1e-3) and ~15 % faster than bf16x3, which has the
# same error.  Training (train_engine) runs on bf16x3 / bf16 planes.
DEFAULT_PRECISION = os.environ.get("MPN_PRECISION", "f16f8")
# The public forward() replays a captured CUDA graph from the SECOND call with the same (mode, shape, dtype, device) on
# (the first call runs eagerly: a caller that never repeats a shape -- multi-scale TTA -- never pays a capture); at most
# MAX_GRAPHS captured graphs are kept per engine (least recently used evicted: each holds its own activation pool).
USE_GRAPHS = os.environ.get("MPN_CUDA_GRAPH", "1") == "1"
MAX_GRAPHS = int(os.environ.get("MPN_MAX_GRAPHS", "4"))
TC_STEM = os.environ.get("MPN_TC_STEM", "1") == "1"
USE_STREAMS = os.environ.get("MPN_STREAMS", "0") == "1"  # measured: no gain on B200 (r01c), kept as an option
LEVEL_STREAMS = os.environ.get("MPN_LEVEL_STREAMS", "1") == "1"  # small pyramid levels of the RetinaNet towers side by side
# one launch per tower layer over the five pyramid levels (mpn_conv2d_fwd_multi): 10 launches per step instead of 50
TOWER_MULTI = os.environ.get("MPN_TOWER_MULTI", "1") == "1"
# keypoint head: the x8 / x4 quarters of conv2 evaluated at low resolution as phase-class convolutions (no replicated concat)
CONV2_GATHER = os.environ.get("MPN_CONV2_GATHER", "1") == "1"
SMALL_LEVEL_PIXELS = 120 * 80  # batch x H x W of a level that cannot fill the GPU (<= 40 CTA pairs of 2 x 120 pixels)


def _bn_tuple(bn):
    return (bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)


class Engine(object):
    def __init__(self, model, precision=None):
        self.model = model
        self.precision = precision or DEFAULT_PRECISION
        if self.precision not in ops.PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(ops.PRECISIONS))
        self.fmt = ops.PRECISIONS[self.precision]
        self._packed = {}
        self._sig = None
        self._graphs = {}
        self._graph_sig = None
        self._streams = {}
        self.last_detections = None

    def _side(self, i, device):
        """Side streams for the independent branches of the graph (keypoint head | detection neck+class tower |
        box tower): the persistent conv kernels own whole SMs, so a second stream fills the tail of each launch
        and lets the small pyramid levels of the two towers run side by side."""
        key = (i, str(device))
        st = self._streams.get(key)
        if st is None:
            st = torch.cuda.Stream(device=device)
            self._streams[key] = st
        return st

    # ------------------------------------------------------------------ weights
    def _collect_slots(self):
        """One walk over the module tree (~4 ms for R101): the (owner dict, name) slot of every parameter and buffer, the
        (parent._modules, name, child) link of every submodule and the BatchNorm modules (their eps is part of the fold)."""
        slots, links, bns = [], [], []
        for mod in self.model.modules():
            slots.extend((mod._parameters, n) for n, t in mod._parameters.items() if t is not None)
            slots.extend((mod._buffers, n) for n, t in mod._buffers.items() if t is not None)
            links.extend((mod._modules, n, c) for n, c in mod._modules.items() if c is not None)
            if isinstance(mod, torch.nn.modules.batchnorm._BatchNorm):
                bns.append(mod)
        self._slots, self._links, self._bns = slots, links, bns

    def invalidate(self):
        """Forget the packed filters, the captured graphs and the cached module walk.  Needed after edits the signature
        cannot see: in-place writes through `.data` / `.detach()` views (`p.data.mul_()`, `p.data.fill_()` -- those do not
        bump the version counter of `p`).  poseNet.load_state_dict / _apply (.to, .cuda, .half ...) call it themselves."""
        self._packed, self._sig = {}, None
        self._graphs, self._graph_sig = {}, None
        self._slots = None

    def _signature(self):
        """(storage address, version counter) of every parameter and buffer, the identity of every submodule link and every
        BatchNorm eps: in-place updates (optimizer.step, load_state_dict), .to()/.cuda(), replaced Parameter objects,
        replaced submodules (`m.convfin = nn.Conv2d(...)`) and a changed `bn.eps` all change it.  The slots are collected
        once -- walking the module tree costs ~4 ms for R101, which a synchronous caller pays as GPU idle time in front of
        every graph replay; reading them costs ~0.3 ms.  NOT seen: writes through `.data` views -> call invalidate()."""
        if self.__dict__.get("_slots") is None:
            self._collect_slots()
        try:
            for d, n, c in self._links:
                if d[n] is not c:       # a submodule was replaced: the cached walk is stale
                    raise KeyError(n)
            ts = [d[n] for d, n in self._slots]
            return (tuple(map(torch.Tensor.data_ptr, ts)), tuple([t._version for t in ts]), tuple([b.eps for b in self._bns]))
        except (KeyError, TypeError, AttributeError):  # something was removed, replaced or set to None: walk again next time
            self._slots = None
            return (object(),)

    def _ensure_packed(self):
        sig = self._signature()
        if sig != self._sig:
            self._packed = {}
            self._sig = sig

    def _pc(self, name, conv, bn=None, fmt=None, no_h8=False):
        """no_h8 (f16f8 only): pack for an input stored without its e5m2 copy plane (ops.pack_conv in_no_h8)."""
        fmt = self.fmt if fmt is None else fmt
        no_h8 = bool(no_h8) and fmt == ops.FMT_F16F8 and ops.NO_H8 and not ops.DERIVE_H8   # derived copy plane: ordinary packing
        key = (name, fmt, no_h8)
        pc = self._packed.get(key)
        if pc is None:
            if not conv.weight.is_cuda:
                raise RuntimeError("poseNet must be on a CUDA device (no CPU path); call .cuda() first")
            pc = ops.pack_conv(conv.weight, conv.bias, _bn_tuple(bn) if bn is not None else None, fmt, in_no_h8=no_h8)
            self._packed[key] = pc
        return pc

    @property
    def _slim(self):
        """f16f8: tensors that only 1x1 convolutions (and shortcut adds) read are stored without their e5m2 copy plane."""
        return self.fmt == ops.FMT_F16F8 and (ops.NO_H8 or ops.DERIVE_H8)

    @property
    def _derive(self):
        """f16f8: the kernels derive the e5m2 copy in shared memory (ops.DERIVE_H8): no tensor stores it."""
        return self.fmt == ops.FMT_F16F8 and ops.DERIVE_H8

    # ------------------------------------------------------------------ backbone
    def backbone(self, img):
        """fpn.py:99-105 -> c2, c3, c4, c5 (Act)."""
        fpn = self.model.fpn
        raw_u8 = img.is_cuda and img.dtype == torch.uint8 and img.dim() == 4 and img.shape[3] == 3
        if raw_u8 and not (self.fmt != FMT_F32 and TC_STEM):
            img = ops.resnet_preprocess_u8(img)  # extension: raw cv2 images, resnet_preprocess on the device
            raw_u8 = False
        if not raw_u8 and not (img.is_cuda and img.dtype == torch.float32 and img.dim() == 4 and img.shape[1] == 3):
            raise RuntimeError("expected a CUDA fp32 [B,3,H,W] image batch (or uint8 [B,H,W,3] BGR raw images)")
        if self.fmt != FMT_F32 and TC_STEM:
            # tensor-core stem: 7x7/2 on the image == 4x4/1 on the zero-padded space-to-depth tensor
            # with 64 output channels the stem is bound by its A-operand traffic (L2 -> shared memory), not by the tensor pipe:
            # the no-h8 operand variant moves 192 instead of 256 bytes per pixel and K block
            key = ("fpn.conv1", "tcstem", self.fmt, self._slim)
            pc = self._packed.get(key)
            if pc is None:
                pc = ops.pack_stem_filter(fpn.conv1.weight, _bn_tuple(fpn.bn1), self.fmt, in_no_h8=self._slim)
                self._packed[key] = pc
            xs = (ops.stem_pack_input_u8 if raw_u8 else ops.stem_pack_input)(img, self.fmt, want_h8=not self._slim)
            c1 = ops.conv2d(xs, pc, relu=True, want_h8=not self._slim)   # read by the max-pool only
        else:
            x = ops.act_from_nchw(img, FMT_F32)
            # fp32-packed stem filter on the CUDA-core kernel; the epilogue emits the engine's activation format
            c1 = ops.conv2d(x, self._stem_pc(), stride=2, pad=3, relu=True, f32_input=True)
        c = ops.maxpool3x3s2(c1, want_h8=not self._slim)
        feats = []
        for li in range(1, 5):
            layer = getattr(fpn, "layer%d" % li)
            for bi, blk in enumerate(layer):
                # c5 (the last block of layer4) also feeds the 3x3 conv6 (fpn.py:108): it keeps its e5m2 copy plane
                c = self._bottleneck("fpn.layer%d.%d" % (li, bi), blk, c, out_h8=(li == 4 and bi == len(layer) - 1))
            feats.append(c)
        return feats

    def _stem_pc(self):
        key = ("fpn.conv1", "stem", self.fmt)
        pc = self._packed.get(key)
        if pc is None:
            fpn = self.model.fpn
            if not fpn.conv1.weight.is_cuda:
                raise RuntimeError("poseNet must be on a CUDA device (no CPU path); call .cuda() first")
            pc = ops.pack_conv(fpn.conv1.weight, None, _bn_tuple(fpn.bn1), FMT_F32)
            pc.fmt = self.fmt  # output / epilogue format
            self._packed[key] = pc
        return pc

    def _bottleneck(self, name, blk, x, out_h8=False):
        """fpn.py:28-34.  f16f8: the block input / output and the 3x3's output are read by 1x1 convolutions and shortcut adds
        only -> stored without the e5m2 copy plane (3 B per element), their consumers packed with no_h8; the 1x1 -> 3x3 tensor
        keeps it."""
        stride = blk.conv2.stride[0]
        slim = self._slim
        nh = slim and not x.has_h8
        # a 3x3 with <= 64 output channels (layer1) is bound by its A-operand traffic, not by the tensor pipe: it takes the no-h8
        # operand variant too, and its input is then stored without the copy plane
        thin = slim and (blk.conv2.out_channels <= 64 or self._derive)
        o = ops.conv2d(x, self._pc(name + ".conv1", blk.conv1, blk.bn1, no_h8=nh), relu=True, want_h8=not thin)
        o = ops.conv2d(o, self._pc(name + ".conv2", blk.conv2, blk.bn2, no_h8=thin), stride=stride, pad=1, relu=True, want_h8=not slim)
        if len(blk.downsample) > 0:
            sc = ops.conv2d(x, self._pc(name + ".downsample", blk.downsample[0], blk.downsample[1], no_h8=nh), stride=stride,
                            want_h8=not slim)
        else:
            sc = x
        return ops.conv2d(o, self._pc(name + ".conv3", blk.conv3, blk.bn3, no_h8=slim), relu=True, residual=sc,
                          want_h8=((not slim) or out_h8) and not self._derive)

    # ------------------------------------------------------------------ necks
    def detection_neck(self, c3, c4, c5):
        """fpn.py:107-114 -> [p3, p4, p5, p6, p7]."""
        f = self.model.fpn
        p6 = ops.conv2d(c5, self._pc("fpn.conv6", f.conv6), stride=2, pad=1)
        p7 = ops.conv2d(ops.relu(p6), self._pc("fpn.conv7", f.conv7), stride=2, pad=1)
        nh = lambda t: not t.has_h8   # stage outputs stored without the e5m2 copy plane -> laterals packed accordingly
        p5 = ops.conv2d(c5, self._pc("fpn.latlayer1", f.latlayer1, no_h8=nh(c5)))
        p4 = ops.conv2d(c4, self._pc("fpn.latlayer2", f.latlayer2, no_h8=nh(c4)), up=p5)
        p3 = ops.conv2d(c3, self._pc("fpn.latlayer3", f.latlayer3, no_h8=nh(c3)), up=p4)
        p5 = ops.conv2d(p5, self._pc("fpn.toplayer0", f.toplayer0), pad=1)
        p4 = ops.conv2d(p4, self._pc("fpn.toplayer1", f.toplayer1), pad=1)
        p3 = ops.conv2d(p3, self._pc("fpn.toplayer2", f.toplayer2), pad=1)
        return [p3, p4, p5, p6, p7]

    def keypoint_neck(self, c2, c3, c4, c5):
        """fpn.py:117-124 -> [fp2, fp3, fp4, fp5]."""
        f = self.model.fpn
        nh = lambda t: not t.has_h8
        fp5 = ops.conv2d(c5, self._pc("fpn.toplayer", f.toplayer, no_h8=nh(c5)))
        fp4 = ops.conv2d(c4, self._pc("fpn.flatlayer1", f.flatlayer1, no_h8=nh(c4)), up=fp5)
        fp3 = ops.conv2d(c3, self._pc("fpn.flatlayer2", f.flatlayer2, no_h8=nh(c3)), up=fp4)
        fp2 = ops.conv2d(c2, self._pc("fpn.flatlayer3", f.flatlayer3, no_h8=nh(c2)), up=fp3)
        fp4 = ops.conv2d(fp4, self._pc("fpn.smooth1", f.smooth1), pad=1)
        fp3 = ops.conv2d(fp3, self._pc("fpn.smooth2", f.smooth2), pad=1)
        fp2 = ops.conv2d(fp2, self._pc("fpn.smooth3", f.smooth3), pad=1)
        return [fp2, fp3, fp4, fp5]

    # ------------------------------------------------------------------ heads
    def keypoint_head(self, p2, p3, p4, p5):
        """posenet.py:243-257: returns heat [B,18,H/4,W/4] fp32 NCHW."""
        m = self.model
        if not (p3.H * 2 == p2.H and p4.H * 4 == p2.H and p5.H * 8 == p2.H and p3.W * 2 == p2.W and p4.W * 4 == p2.W
                and p5.W * 8 == p2.W):
            raise RuntimeError("keypoint head needs H and W to be multiples of 32 (evaluate/tester.py:285 pads to 32)")
        if CONV2_GATHER and self.fmt != FMT_F32:
            return self._keypoint_head_classes(p2, p3, p4, p5)
        cat = ops.Act(self.fmt, p2.N, p2.H, p2.W, 512, p2.hi.device, has_h8=not self._derive)
        for src, t, s, rep, off in ((p5, "convt1", "convs1", 8, 0), (p4, "convt2", "convs2", 4, 128),
                                    (p3, "convt3", "convs3", 2, 256), (p2, "convt4", "convs4", 1, 384)):
            q = ops.conv2d(src, self._pc(t, getattr(m, t)), pad=1)
            ops.conv2d(q, self._pc(s, getattr(m, s)), pad=1, out=cat, out_coffset=off, out_rep=rep)
        h = ops.conv2d(cat, self._pc("conv2", m.conv2), pad=1, relu=True, want_h8=not self._slim)   # read by the 1x1 convfin only
        return ops.conv2d(h, self._pc("convfin", m.convfin, no_h8=self._slim), out_mode=OUT_F32_NCHW)

    def _keypoint_head_classes(self, p2, p3, p4, p5):
        """The same head without materialising the x8 / x4 upsampled quarters of the concat (posenet.py:250-254): conv2 is
        linear in its input channels, and a 3x3 convolution of a nearest-upsampled map is, per output phase, one of nine 3x3
        convolutions of the low-resolution map (ops.phase_class_filter).  The q5 / q4 quarters of conv2 are evaluated at 1/32 and
        1/16 resolution (9 * 256 output channels: 0.14x and 0.56x the tensor work of their quarter) and conv2 -- now over the
        256 channels of q3 (x2, still replicated into the concat) and q2 -- adds them per pixel in its epilogue."""
        m = self.model
        slim = self._slim
        key = ("conv2.classes", self.fmt, slim)
        pcs = self._packed.get(key)
        if pcs is None:
            w = m.conv2.weight.detach()
            if not w.is_cuda:
                raise RuntimeError("poseNet must be on a CUDA device (no CPU path); call .cuda() first")
            pcs = (ops.pack_conv(ops.phase_class_filter(w[:, 0:128]), None, None, self.fmt),
                   ops.pack_conv(ops.phase_class_filter(w[:, 128:256]), None, None, self.fmt),
                   ops.pack_conv(w[:, 256:512].contiguous(), m.conv2.bias, None, self.fmt))
            self._packed[key] = pcs
        z = []
        for src, t, s, pcz in ((p5, "convt1", "convs1", pcs[0]), (p4, "convt2", "convs2", pcs[1])):
            q = ops.conv2d(src, self._pc(t, getattr(m, t)), pad=1)
            q = ops.conv2d(q, self._pc(s, getattr(m, s)), pad=1)
            z.append(ops.conv2d(q, pcz, pad=1, want_h8=False))            # read by conv2's epilogue only
        cat = ops.Act(self.fmt, p2.N, p2.H, p2.W, 256, p2.hi.device, has_h8=not self._derive)
        for src, t, s, rep, off in ((p3, "convt3", "convs3", 2, 0), (p2, "convt4", "convs4", 1, 128)):
            q = ops.conv2d(src, self._pc(t, getattr(m, t)), pad=1)
            ops.conv2d(q, self._pc(s, getattr(m, s)), pad=1, out=cat, out_coffset=off, out_rep=rep)
        h = ops.conv2d(cat, pcs[2], pad=1, relu=True, want_h8=not slim, gather=[(z[0], 3), (z[1], 2)])
        return ops.conv2d(h, self._pc("convfin", m.convfin, no_h8=slim), out_mode=OUT_F32_NCHW)

    def intermediate_heads(self, p2, p3, p4, p5):
        """posenet.py:296-299."""
        m = self.model
        outs = []
        for src, name, rep in ((p2, "convfin_k2", 1), (p3, "convfin_k3", 2), (p4, "convfin_k4", 4), (p5, "convfin_k5", 8)):
            outs.append(ops.conv2d(src, self._pc(name, getattr(m, name)), out_mode=OUT_F32_NCHW, out_rep=rep))
        return outs

    def _tower_level(self, head, hname, f, off, A, out, per, sigmoid):
        o = f
        for n in ("conv1", "conv2", "conv3", "conv4"):
            o = ops.conv2d(o, self._pc("%s.%s" % (hname, n), getattr(head, n)), pad=1, relu=True)
        ops.conv2d(o, self._pc(hname + ".output", head.output), pad=1, sigmoid=sigmoid, out_mode=OUT_F32_NHWC,
                   out_tensor=out, out_elem_offset=off * per, out_cstride=9 * per, out_nstride=A * per)

    def _tower(self, head, hname, feats, out, per, sigmoid, levels=None):
        cells = [f.H * f.W for f in feats]
        A = 9 * sum(cells)
        off = 0
        for i, (f, ncell) in enumerate(zip(feats, cells)):
            if levels is None or i in levels:
                self._tower_level(head, hname, f, off, A, out, per, sigmoid)
            off += ncell * 9

    def detection_heads(self, feats):
        """posenet.py:262-263 -> cls [B,A,1], reg [B,A,4] fp32 (levels concatenated in place)."""
        m = self.model
        B = feats[0].N
        A = 9 * sum(f.H * f.W for f in feats)
        dev = feats[0].hi.device
        cls = torch.empty((B, A, 1), dtype=torch.float32, device=dev)
        reg = torch.empty((B, A, 4), dtype=torch.float32, device=dev)
        small = [i for i, f in enumerate(feats) if f.N * f.H * f.W <= SMALL_LEVEL_PIXELS]
        if TOWER_MULTI and self.fmt != FMT_F32 and len(feats) <= 5:
            # posenet.py:262-263: the tower's weights are shared by the levels -> one persistent launch per layer whose tile
            # list is the concatenation of the five levels' tiles (the P5-P7 levels alone are latency-bound launches that
            # leave most SMs idle); the two towers alternate so that consecutive launches are independent
            cells = [f.H * f.W for f in feats]
            offs = [9 * sum(cells[:i]) for i in range(len(feats))]
            hr, hc = list(feats), list(feats)
            slim = self._slim   # the 36- / 9-channel output convs are A-traffic bound: no-h8 operand variant, conv4 stores no h8
            for n in ("conv1", "conv2", "conv3", "conv4"):
                h8 = not (slim and n == "conv4") and not self._derive
                hr = ops.conv2d_multi(hr, self._pc("regressionModel." + n, getattr(m.regressionModel, n)), pad=1, relu=True, want_h8=h8)
                hc = ops.conv2d_multi(hc, self._pc("classificationModel." + n, getattr(m.classificationModel, n)), pad=1, relu=True, want_h8=h8)
            ops.conv2d_multi(hr, self._pc("regressionModel.output", m.regressionModel.output, no_h8=slim), pad=1, out_mode=OUT_F32_NHWC,
                             out_tensor=reg, out_elem_offsets=[o * 4 for o in offs], out_cstride=36, out_nstride=A * 4)
            ops.conv2d_multi(hc, self._pc("classificationModel.output", m.classificationModel.output, no_h8=slim), pad=1, sigmoid=True,
                             out_mode=OUT_F32_NHWC, out_tensor=cls, out_elem_offsets=offs, out_cstride=9, out_nstride=A)
        elif LEVEL_STREAMS and small and len(small) < len(feats):
            # The five convs of a small pyramid level (P5-P7: at most 40 CTA pairs, a serial K loop per CTA) are latency
            # bound and leave most SMs idle.  The large levels run first on the current stream (their persistent kernels
            # want every SM); then the small-level chains of both towers run side by side, one stream per (tower, level).
            big = [i for i in range(len(feats)) if i not in small]
            self._tower(m.regressionModel, "regressionModel", feats, reg, 4, False, levels=big)
            self._tower(m.classificationModel, "classificationModel", feats, cls, 1, True, levels=big)
            cur = torch.cuda.current_stream()
            sides = []
            for ti, (head, hname, out, per, sig) in enumerate(((m.regressionModel, "regressionModel", reg, 4, False),
                                                               (m.classificationModel, "classificationModel", cls, 1, True))):
                for li in small:
                    if ti == 0 and li == small[0]:
                        continue  # the first chain stays on the current stream (below)
                    side = self._side(10 + ti * 8 + li, dev)
                    side.wait_stream(cur)
                    with torch.cuda.stream(side):
                        self._tower(head, hname, feats, out, per, sig, levels=[li])
                    sides.append(side)
            self._tower(m.regressionModel, "regressionModel", feats, reg, 4, False, levels=[small[0]])
            for side in sides:
                cur.wait_stream(side)
        elif USE_STREAMS:
            cur, side = torch.cuda.current_stream(), self._side(1, dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                self._tower(m.regressionModel, "regressionModel", feats, reg, 4, False)
            self._tower(m.classificationModel, "classificationModel", feats, cls, 1, True)
            cur.wait_stream(side)
        else:
            self._tower(m.regressionModel, "regressionModel", feats, reg, 4, False)
            self._tower(m.classificationModel, "classificationModel", feats, cls, 1, True)
        return cls, reg

    # ------------------------------------------------------------------ subnet entry points
    @torch.no_grad()
    def keypoint_forward(self, img):
        self._ensure_packed()
        c2, c3, c4, c5 = self.backbone(img)
        p2, p3, p4, p5 = self.keypoint_neck(c2, c3, c4, c5)  # the unused detection neck is skipped (fpn.py:126)
        saved = self.intermediate_heads(p2, p3, p4, p5)
        heat = self.keypoint_head(p2, p3, p4, p5)
        saved.append(heat)
        return heat, saved

    @torch.no_grad()
    def detection_forward(self, img):
        self._ensure_packed()
        _, c3, c4, c5 = self.backbone(img)
        cls, reg = self.detection_heads(self.detection_neck(c3, c4, c5))
        H, W = (img.shape[1], img.shape[2]) if img.dtype == torch.uint8 else (img.shape[2], img.shape[3])
        anchors = ops.anchors_for(H, W, img.device)
        return [], [cls, reg, anchors]

    @torch.no_grad()
    def fpn_forward(self, img):
        """FPN.forward (fpn.py:97-126) with the reference's return structure, as fp32 NCHW tensors."""
        self._ensure_packed()
        c2, c3, c4, c5 = self.backbone(img)
        kp = self.keypoint_neck(c2, c3, c4, c5)
        det = self.detection_neck(c3, c4, c5)
        return [[a.to_nchw() for a in kp], [a.to_nchw() for a in det]]

    @torch.no_grad()
    def entire_forward_device(self, img, score_thresh=0.05, iou_thresh=0.5, max_cand=4096, ge=False):
        """Everything on the device, no host sync: returns (heat, cls, reg, boxes, Detections)."""
        self._ensure_packed()
        H, W = (img.shape[1], img.shape[2]) if img.dtype == torch.uint8 else (img.shape[2], img.shape[3])
        c2, c3, c4, c5 = self.backbone(img)
        anchors = ops.anchors_for(H, W, img.device)

        def detect_branch():
            cls, reg = self.detection_heads(self.detection_neck(c3, c4, c5))
            boxes = ops.decode_clip(anchors, reg, H, W)
            return cls, reg, boxes, ops.filter_sort_nms(cls, boxes, score_thresh, iou_thresh, ge=ge, max_cand=max_cand)

        if USE_STREAMS:
            cur, side = torch.cuda.current_stream(), self._side(0, img.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                cls, reg, boxes, det = detect_branch()
            heat = self.keypoint_head(*self.keypoint_neck(c2, c3, c4, c5))
            cur.wait_stream(side)
        else:
            heat = self.keypoint_head(*self.keypoint_neck(c2, c3, c4, c5))
            cls, reg, boxes, det = detect_branch()
        return heat, cls, reg, boxes, det

    # ------------------------------------------------------------------ pose residual network
    @torch.no_grad()
    def prn_forward(self, x):
        """posenet.py:337-350 (eval: dropout = identity) for a whole batch of persons at once:
        flatten -> FC+ReLU -> FC+ReLU -> FC+ReLU -> +input -> softmax, the three FCs as 1x1 convs on the
        tcgen05 kernel (a person = a pixel), K padded to a multiple of 64."""
        self._ensure_packed()
        prn = self.model.prn
        if not (x.is_cuda and x.dtype == torch.float32):
            raise RuntimeError("expected a CUDA fp32 [P, H, W, 17] batch")
        P = x.shape[0]
        res = x.reshape(P, -1).contiguous()
        D = res.shape[1]
        if self.fmt == FMT_F32:
            raise NotImplementedError("PRN runs on the tcgen05 path (precision f16f8, bf16x3 or bf16), not in the fp32 CUDA-core mode")
        Dp = (D + 63) // 64 * 64

        def pc(name, lin, cin_pad, cout_pad=None):
            key = (name, "prn", self.fmt)
            p_ = self._packed.get(key)
            if p_ is None:
                w, b = lin.weight.detach(), lin.bias.detach()
                cout_pad = cout_pad or w.shape[0]
                if cin_pad != w.shape[1] or cout_pad != w.shape[0]:  # zero rows / columns: exact
                    w = torch.nn.functional.pad(w, (0, cin_pad - w.shape[1], 0, cout_pad - w.shape[0]))
                    b = torch.nn.functional.pad(b, (0, cout_pad - b.shape[0]))
                p_ = ops.pack_conv(w.reshape(w.shape[0], w.shape[1], 1, 1).contiguous(), b, None, self.fmt)
                self._packed[key] = p_
            return p_

        # a person = one 1x1 "image" with D channels, zero-padded to Dp (NCHW [P,Dp,1,1] == NHWC [P,1,1,Dp])
        a = ops.act_from_nchw(torch.nn.functional.pad(res, (0, Dp - D)).view(P, Dp, 1, 1), self.fmt)
        h = ops.conv2d(a, pc("prn.dens1", prn.dens1, Dp), relu=True)
        h = ops.conv2d(h, pc("prn.bneck", prn.bneck, prn.bneck.weight.shape[1]), relu=True)
        o = ops.conv2d(h, pc("prn.dens2", prn.dens2, prn.dens2.weight.shape[1], Dp), relu=True)  # [P, Dp], first D used
        out = ops.add_softmax_rows(o, res)
        return out.view(P, prn.height, prn.width, 17)

    # ------------------------------------------------------------------ CUDA-graph replay of a whole step
    def graphed(self, kind, img, **kw):
        """Run `kind` ('entire', 'keypoint', 'detection') through a captured CUDA graph of its launch
        sequence (captured on first use per input shape; ~250 launches replayed with one driver call).
        The returned tensors are the graph's static outputs: they are overwritten by the next replay."""
        key = self._graph_key(kind, img, kw)
        g = self._graphs.get(key) if self._sig is not None and self._graph_sig == self._sig else None
        if g is not None:
            self._graphs[key] = self._graphs.pop(key)  # most recently used last
            # Fast path: replay first, then validate the weights while the GPU runs -- the signature walk (~0.4 ms) would
            # otherwise be GPU idle time in front of every replay of a synchronous caller.  If the weights did change
            # since the capture, this replay's outputs are discarded and the slow path below repacks, recaptures and reruns.
            graph, static_in, out = g
            static_in.copy_(img, non_blocking=True)
            graph.replay()
            if self._signature() == self._sig:
                return out
        self._ensure_packed()
        if self._graph_sig != self._sig:
            self._graphs, self._graph_sig = {}, self._sig  # weights changed -> recapture
        g = self._graphs.get(key)
        fn = {"entire": self.entire_forward_device, "keypoint": self.keypoint_forward, "detection": self.detection_forward}[kind]
        if g is None:
            static_in = torch.empty_like(img)
            static_in.copy_(img)
            side = torch.cuda.Stream(device=img.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # warm-up off the capture: packs filters, caches anchors, sizes func attributes
                for _ in range(2):
                    fn(static_in, **kw)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = fn(static_in, **kw)
            g = (graph, static_in, out)
            while len(self._graphs) >= max(1, MAX_GRAPHS):
                self._graphs.pop(next(iter(self._graphs)))  # least recently used
            self._graphs[key] = g
        graph, static_in, out = g
        static_in.copy_(img, non_blocking=True)
        graph.replay()
        return out

    @staticmethod
    def _graph_key(kind, img, kw):
        return (kind, tuple(img.shape), str(img.device), str(img.dtype), tuple(sorted(kw.items())))

    def _wants_graph(self, kind, img, **kw):
        """Graph replay for the public calls: on from the second call with the same key (see USE_GRAPHS)."""
        if not USE_GRAPHS or (img.is_cuda and torch.cuda.is_current_stream_capturing()):
            return False
        key = self._graph_key(kind, img, kw)
        if key in self._graphs:
            return True
        seen = self.__dict__.setdefault("_seen", {})
        if len(seen) > 256:
            seen.clear()
        seen[key] = seen.get(key, 0) + 1
        return seen[key] >= 2

    def check_range(self, img, kind="entire"):
        """Overflow guard of the f16f8 format (fp16 planes saturate at +-65504): runs one eager forward that reads back the
        largest magnitude of every activation tensor and raises MpnError naming the first clamped layer.  Debug aid for a
        new checkpoint; MPN_RANGE_CHECK=1 does the same on every eager conv call."""
        fn = {"entire": self.entire_forward_device, "keypoint": self.keypoint_forward, "detection": self.detection_forward}[kind]
        prev, ops.stats["range_check"] = ops.stats.get("range_check"), True
        try:
            fn(img)
        finally:
            ops.stats["range_check"] = prev
        return True

    @torch.no_grad()
    def _counts_to_host(self, det):
        """Candidate and keep counts of the batch with ONE stream synchronisation (two async copies into a pinned buffer): the
        GPU idles while the host reads them, every separate `.cpu()` / `.item()` is another round trip."""
        B = det.cand_cnt.numel()
        key = (B, str(det.cand_cnt.device))
        pin = self.__dict__.setdefault("_cnt_pin", {}).get(key)
        if pin is None:
            pin = self._cnt_pin[key] = torch.empty((2, B), dtype=torch.int32).pin_memory()
        pin[0].copy_(det.cand_cnt, non_blocking=True)
        pin[1].copy_(det.keep_cnt, non_blocking=True)
        torch.cuda.current_stream(det.cand_cnt.device).synchronize()
        return pin[0].clone(), pin[1].clone()

    def entire_forward(self, img, max_cand=4096):
        """posenet.py:236-285: (heat, [nms_scores, nms_class, boxes]) for image 0, like the reference;
        the per-image results of the whole batch stay in self.last_detections."""
        if self._wants_graph("entire", img, max_cand=max_cand):
            heat, cls, reg, boxes, det = self.graphed("entire", img, max_cand=max_cand)
            heat = heat.clone()  # static graph output -> caller-owned
        else:
            heat, cls, reg, boxes, det = self.entire_forward_device(img, max_cand=max_cand)
        cnt, kcnt = self._counts_to_host(det)
        if int(cnt.max()) > det.max_cand:  # rare: more survivors than the fast-path capacity -> redo with room
            det = ops.filter_sort_nms(cls, boxes, 0.05, 0.5, max_cand=int(cnt.max()))
            cnt, kcnt = self._counts_to_host(det)
        self.last_detections = det
        if int(cnt[0]) == 0:  # posenet.py:273-275 (CPU tensors, as in the reference)
            return heat, [torch.zeros(0), torch.zeros(0), torch.zeros(0, 4)]
        k = int(kcnt[0])
        scores = det.scores[0, :k].clone()
        classes = torch.zeros((k,), dtype=torch.int64, device=img.device)  # single class: argmax over 1 column
        return heat, [scores, classes, det.boxes[0, :k].clone()]
