"""Parameter inventory, seeded synthetic weight sets and FLOP accounting of the path (data only, no arithmetic of the path).

Shared by bench.py (synthetic weights of the benchmarked architecture, algorithmic FLOPs) and, re-exported through
oracle/weights.py, by the tests.

`param_spec(layers)` lists every state_dict entry of the reference's
`poseNet(layers)` (network/posenet.py:154-211, network/fpn.py:9-82) in registration
order.  It is written from the constructor code, not imported, so that it is
available on the GPU box where /root/reference does not exist; the not-gpu test
`tests/test_oracle.py` checks it key-by-key against the real module.

Two weight sets (last row / section 8(d)):
  * "refinit"     -- what the reference constructor produces: normal(std=0.01) convs,
                     zero conv bias, cls output weight 0 / bias -log(99), reg output 0,
                     default BN (posenet.py:205-218).  Drawn here from numpy PCG64 so
                     it is reproducible without torch's RNG.
  * "conditioned" -- He-normal convs, randomised BN gamma/beta/mean/var, non-zero head
                     outputs with the cls bias placed so ~1-5 % of anchors pass 0.05.
All arrays are float32 numpy (int64 for num_batches_tracked).
"""
import math
from collections import OrderedDict

import numpy as np

BLOCKS = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}
CLS_OUT_GAIN = 0.65
REG_OUT_GAIN = 0.7


def _conv(spec, name, cout, cin, k, bias):
    spec[name + ".weight"] = (cout, cin, k, k)
    if bias:
        spec[name + ".bias"] = (cout,)


def _bn(spec, name, c):
    spec[name + ".weight"] = (c,)
    spec[name + ".bias"] = (c,)
    spec[name + ".running_mean"] = (c,)
    spec[name + ".running_var"] = (c,)
    spec[name + ".num_batches_tracked"] = ()


def fpn_spec(layers, prefix="fpn."):
    """network/fpn.py:37-74 (FPN.__init__) and :9-26 (Bottleneck.__init__)."""
    s = OrderedDict()
    _conv(s, prefix + "conv1", 64, 3, 7, False)
    _bn(s, prefix + "bn1", 64)
    in_planes = 64
    for li, (planes, nblk, stride) in enumerate(zip((64, 128, 256, 512), BLOCKS[layers], (1, 2, 2, 2)), start=1):
        for b in range(nblk):
            st = stride if b == 0 else 1
            p = "%slayer%d.%d." % (prefix, li, b)
            _conv(s, p + "conv1", planes, in_planes, 1, False)
            _bn(s, p + "bn1", planes)
            _conv(s, p + "conv2", planes, planes, 3, False)
            _bn(s, p + "bn2", planes)
            _conv(s, p + "conv3", planes * 4, planes, 1, False)
            _bn(s, p + "bn3", planes * 4)
            if st != 1 or in_planes != planes * 4:
                _conv(s, p + "downsample.0", planes * 4, in_planes, 1, False)
                _bn(s, p + "downsample.1", planes * 4)
            in_planes = planes * 4
    _conv(s, prefix + "conv6", 256, 2048, 3, True)
    _conv(s, prefix + "conv7", 256, 256, 3, True)
    _conv(s, prefix + "latlayer1", 256, 2048, 1, True)
    _conv(s, prefix + "latlayer2", 256, 1024, 1, True)
    _conv(s, prefix + "latlayer3", 256, 512, 1, True)
    for n in ("toplayer0", "toplayer1", "toplayer2"):
        _conv(s, prefix + n, 256, 256, 3, True)
    _conv(s, prefix + "toplayer", 256, 2048, 1, True)
    _conv(s, prefix + "flatlayer1", 256, 1024, 1, True)
    _conv(s, prefix + "flatlayer2", 256, 512, 1, True)
    _conv(s, prefix + "flatlayer3", 256, 256, 1, True)
    for n in ("smooth1", "smooth2", "smooth3"):
        _conv(s, prefix + n, 256, 256, 3, True)
    return s


def param_spec(layers, prn_node_count=1024, prn_coeff=2):
    """network/posenet.py:155-199 (poseNet.__init__), registration order."""
    s = fpn_spec(layers)
    for n in ("convfin_k2", "convfin_k3", "convfin_k4", "convfin_k5"):
        _conv(s, n, 19, 256, 1, True)
    for n in ("convt1", "convt2", "convt3", "convt4"):
        _conv(s, n, 128, 256, 3, True)
    for n in ("convs1", "convs2", "convs3", "convs4"):
        _conv(s, n, 128, 128, 3, True)
    _conv(s, "conv2", 256, 512, 3, True)
    _conv(s, "convfin", 18, 256, 1, True)
    for head, cout in (("regressionModel", 36), ("classificationModel", 9)):
        for n in ("conv1", "conv2", "conv3", "conv4"):
            _conv(s, "%s.%s" % (head, n), 256, 256, 3, True)
        _conv(s, head + ".output", cout, 256, 3, True)
    hw17 = (prn_coeff * 28) * (prn_coeff * 18) * 17
    s["prn.dens1.weight"] = (prn_node_count, hw17)
    s["prn.dens1.bias"] = (prn_node_count,)
    s["prn.bneck.weight"] = (prn_node_count, prn_node_count)
    s["prn.bneck.bias"] = (prn_node_count,)
    s["prn.dens2.weight"] = (hw17, prn_node_count)
    s["prn.dens2.bias"] = (hw17,)
    return s


def _gain(name):
    """He gain sqrt(2) for convs whose output feeds a ReLU, 1 for the linear ones (FPN necks,
    convt/convs, convfin*, head outputs, downsample, bottleneck conv3)."""
    if ".layer" in name:
        return math.sqrt(2.0) if name.endswith(("conv1.weight", "conv2.weight")) else 1.0
    if name == "fpn.conv1.weight" or name == "conv2.weight":
        return math.sqrt(2.0)
    if name.startswith(("regressionModel", "classificationModel")) and not name.endswith("output.weight"):
        return math.sqrt(2.0)
    if name.startswith(("fpn.latlayer", "fpn.flatlayer", "fpn.toplayer.", "fpn.conv6")):
        return 0.35  # laterals see post-ReLU trunk features with E[x^2] >> 1
    return 1.0


def _is_conv_weight(name, shape):
    return name.endswith(".weight") and len(shape) == 4


def make_weights(layers, kind="conditioned", seed=0, include_prn=False):
    """Return OrderedDict name -> numpy array for every entry of param_spec.

    PRN matrices (285 MB) are skipped unless include_prn (they are off the conv path).
    """
    assert kind in ("refinit", "conditioned")
    rng = np.random.Generator(np.random.PCG64(seed))
    spec = param_spec(layers)
    out = OrderedDict()
    for name, shape in spec.items():
        if name.startswith("prn.") and not include_prn:
            continue
        if name.endswith("num_batches_tracked"):
            out[name] = np.zeros((), dtype=np.int64)
            continue
        is_bn = (".bn" in name or "downsample.1" in name or name == "fpn.bn1.weight")
        if name.startswith("prn."):
            fan_in = shape[-1] if len(shape) == 2 else 1
            if len(shape) == 2:
                out[name] = (rng.standard_normal(shape, dtype=np.float32) / np.float32(math.sqrt(fan_in)))
            else:
                out[name] = np.zeros(shape, dtype=np.float32)
            continue
        if kind == "refinit":
            if _is_conv_weight(name, shape):
                out[name] = rng.standard_normal(shape, dtype=np.float32) * np.float32(0.01)
            elif is_bn and name.endswith((".weight", ".running_var")):
                out[name] = np.ones(shape, dtype=np.float32)
            else:  # conv bias, bn bias, running_mean
                out[name] = np.zeros(shape, dtype=np.float32)
        else:
            if _is_conv_weight(name, shape):
                fan_in = shape[1] * shape[2] * shape[3]
                out[name] = rng.standard_normal(shape, dtype=np.float32) * np.float32(_gain(name) / math.sqrt(fan_in))
            elif is_bn and name.endswith("bn3.weight"):
                # small residual-branch gain keeps the trunk O(1) through 33 blocks
                out[name] = rng.uniform(0.2, 0.4, shape).astype(np.float32)
            elif is_bn and name.endswith(".weight"):
                out[name] = rng.uniform(0.5, 1.5, shape).astype(np.float32)
            elif is_bn and name.endswith(".running_var"):
                out[name] = rng.uniform(0.5, 1.5, shape).astype(np.float32)
            elif is_bn and name.endswith(".running_mean"):
                out[name] = (rng.standard_normal(shape) * 0.1).astype(np.float32)
            else:  # biases (conv and bn)
                out[name] = (rng.standard_normal(shape) * 0.1).astype(np.float32)
    if kind == "refinit":
        # posenet.py:205-209
        out["classificationModel.output.weight"][...] = 0
        out["classificationModel.output.bias"][...] = -math.log((1.0 - 0.01) / 0.01)
        out["regressionModel.output.weight"][...] = 0
        out["regressionModel.output.bias"][...] = 0
    else:
        # Score head: logits ~ N(-5.0, ~1.3-1.5) -> a few % of anchors above logit(0.05) = -2.94;
        # box head: deltas with std ~1 (x std .1/.2 in the decode -> mild box motion).
        out["classificationModel.output.weight"] *= np.float32(CLS_OUT_GAIN)
        out["classificationModel.output.bias"][...] = -5.0
        out["regressionModel.output.weight"] *= np.float32(REG_OUT_GAIN)
    return out


def to_torch_state_dict(weights):
    import torch

    return OrderedDict((k, torch.from_numpy(np.ascontiguousarray(v))) for k, v in weights.items())


def conv_flops_entire(layers, H=480, W=640):
    """Algorithmic conv FLOPs/img of the entire_net graph (2*MAC), SURVEY.md 8(d)."""
    spec = param_spec(layers)
    total = 0.0
    # spatial size of each conv's OUTPUT, derived from the graph
    h4, w4 = H // 4, W // 4
    def out_hw(name):
        if name == "fpn.conv1":
            return H // 2, W // 2
        if name.startswith("fpn.layer"):
            li = int(name[len("fpn.layer")])
            b = int(name.split(".")[2])
            s = (1, 2, 4, 8)[li - 1]
            hh, ww = h4 // s, w4 // s
            if li > 1 and b == 0 and name.endswith("conv1"):
                return hh * 2, ww * 2  # stride sits on conv2 (fpn.py:16)
            return hh, ww
        m = {"fpn.conv6": 64, "fpn.conv7": 128, "fpn.latlayer1": 32, "fpn.latlayer2": 16, "fpn.latlayer3": 8,
             "fpn.toplayer0": 32, "fpn.toplayer1": 16, "fpn.toplayer2": 8, "fpn.toplayer": 32,
             "fpn.flatlayer1": 16, "fpn.flatlayer2": 8, "fpn.flatlayer3": 4, "fpn.smooth1": 16,
             "fpn.smooth2": 8, "fpn.smooth3": 4, "convt1": 32, "convs1": 32, "convt2": 16, "convs2": 16,
             "convt3": 8, "convs3": 8, "convt4": 4, "convs4": 4, "conv2": 4, "convfin": 4}
        if name in m:
            d = m[name]
            return -(-H // d), -(-W // d)
        return None
    for k, shp in spec.items():
        if not (k.endswith(".weight") and len(shp) == 4):
            continue
        name = k[:-len(".weight")]
        if name.startswith("convfin_k"):
            continue  # keypoint_subnet mode only
        mac = shp[0] * shp[1] * shp[2] * shp[3]
        if name.startswith(("regressionModel", "classificationModel")):
            cells = sum((-(-H // d)) * (-(-W // d)) for d in (8, 16, 32, 64, 128))
            total += 2.0 * mac * cells
            continue
        hw = out_hw(name)
        assert hw is not None, name
        total += 2.0 * mac * hw[0] * hw[1]
    return total


def calibrate_output_bias(model, x, what, per_image, threshold):
    """Synthetic-weight construction for random-weight benchmarks / tests: shift ONE bias so that about `per_image`
    outputs per image exceed `threshold` on the probe batch x (CUDA fp32 [B,3,H,W]).
      what == 'cls'  : class-head output bias, sigmoid scores (posenet.py:262-279 score filter / tester.py:236 box filter)
      what == 'heat' : convfin bias, the 18 supervised heat-map channels (joint_utils.py:19-32 thre1)
    Returns the shift that was added."""
    import math
    import torch
    with torch.no_grad():
        if what == "cls":
            _, (cls, _, _) = model((x, "detection_subnet"))
            p = cls[:, :, 0].double().clamp(1e-12, 1 - 1e-12)
            v = torch.log(p / (1 - p))
            thr = math.log(threshold / (1.0 - threshold))
            bias = model.classificationModel.output.bias
        elif what == "heat":
            heat, _ = model((x, "keypoint_subnet"))
            v = heat[:, :18].double().flatten(1)
            thr = float(threshold)
            bias = model.convfin.bias
        else:
            raise ValueError(what)
        flat = v.flatten()
        k = max(1, min(flat.numel() - 1, int(round(per_image * v.shape[0]))))
        cut = float(torch.topk(flat, k + 1).values[-1])          # the (k+1)-th largest value lands exactly on the threshold
        shift = thr - cut
        bias += shift
    return shift


def cfg3_detections(batch, seed=0, H=480, W=640, persons=100, per_person=41, anchors=57600):
    """SURVEY 8(d) cfg3 detection load, fed to the filter/sort/NMS stage directly: per image `persons` ground-truth boxes
    (w ~ U(20,120), h ~ U(60,300), centres uniform in the image) x `per_person` jittered candidates (+-10 % centre / size) =
    ~4100 candidates with DISTINCT scores in (0.05, 1) scattered over the `anchors` slots; every other slot scores 0.01.
    Returns (cls [batch, anchors, 1] fp32, boxes [batch, anchors, 4] fp32) numpy arrays; expected kept boxes ~ persons."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = persons * per_person
    cls = np.full((batch, anchors, 1), 0.01, dtype=np.float32)
    boxes = np.zeros((batch, anchors, 4), dtype=np.float32)
    boxes[:, :, 2:] = 1.0
    for b in range(batch):
        w = rng.uniform(20, 120, persons); h = rng.uniform(60, 300, persons)
        cx = rng.uniform(0, W, persons); cy = rng.uniform(0, H, persons)
        jw = (w[:, None] * rng.uniform(0.9, 1.1, (persons, per_person))).reshape(-1)
        jh = (h[:, None] * rng.uniform(0.9, 1.1, (persons, per_person))).reshape(-1)
        jx = (cx[:, None] + w[:, None] * rng.uniform(-0.1, 0.1, (persons, per_person))).reshape(-1)
        jy = (cy[:, None] + h[:, None] * rng.uniform(-0.1, 0.1, (persons, per_person))).reshape(-1)
        bx = np.stack([jx - jw / 2, jy - jh / 2, jx + jw / 2, jy + jh / 2], 1)
        bx[:, 0::2] = np.clip(bx[:, 0::2], 0, W); bx[:, 1::2] = np.clip(bx[:, 1::2], 0, H)
        slots = np.sort(rng.choice(anchors, n, replace=False))
        boxes[b, slots] = bx.astype(np.float32)
        cls[b, slots, 0] = rng.permutation(np.linspace(0.0501, 0.9999, n)).astype(np.float32)
    return cls, boxes
