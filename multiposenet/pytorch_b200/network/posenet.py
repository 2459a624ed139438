"""Drop-in `poseNet` with the reference's module tree, state_dict and forward() contract.

Mirror of /root/reference/network/posenet.py:
  poseNet(layers, prn_node_count=1024, prn_coeff=2)            :154-211
  forward([img_batch, subnet_name])                            :226-285
      'keypoint_subnet'  -> (heat[B,18,H/4,W/4], [k2,k3,k4,k5 each [B,19,H/4,W/4], heat])   :288-318
      'detection_subnet' -> ([], [cls[B,A,1], reg[B,A,4], anchors[1,A,4]])                  :320-335
      'prn_subnet'       -> (out[B,56,36,17], [out])                                        :337-350
      anything else      -> (heat, [nms_scores[K], nms_class[K], boxes[K,4]]) for image 0   :236-285
  build_loss / freeze_bn                                        :220-224, :352-364
Inference (no grad) runs on libmpn_b200's sm_100a kernels through ..engine.Engine; there is no eager
PyTorch fallback: CPU tensors or a missing library raise.
"""
import math
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import init

from .. import engine as _engine
from .. import ops as _ops
from .fpn import FPN50, FPN101


def nms(dets, thresh):
    """posenet.py:19-22: dispatch to pth_nms."""
    from ..lib.nms.pth_nms import pth_nms
    return pth_nms(dets, thresh)


class _ParamsOnly(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("%s holds parameters only; use poseNet.forward" % type(self).__name__)


class RegressionModel(_ParamsOnly):
    def __init__(self, num_features_in, num_anchors=9, feature_size=256):
        super().__init__()
        self.conv1 = nn.Conv2d(num_features_in, feature_size, 3, padding=1)
        self.act1 = nn.ReLU()
        self.conv2 = nn.Conv2d(feature_size, feature_size, 3, padding=1)
        self.act2 = nn.ReLU()
        self.conv3 = nn.Conv2d(feature_size, feature_size, 3, padding=1)
        self.act3 = nn.ReLU()
        self.conv4 = nn.Conv2d(feature_size, feature_size, 3, padding=1)
        self.act4 = nn.ReLU()
        self.output = nn.Conv2d(feature_size, num_anchors * 4, 3, padding=1)


class ClassificationModel(_ParamsOnly):
    def __init__(self, num_features_in, num_anchors=9, num_classes=80, prior=0.01, feature_size=256):
        super().__init__()
        self.num_classes, self.num_anchors = num_classes, num_anchors
        self.conv1 = nn.Conv2d(num_features_in, feature_size, 3, padding=1)
        self.act1 = nn.ReLU()
        self.conv2 = nn.Conv2d(feature_size, feature_size, 3, padding=1)
        self.act2 = nn.ReLU()
        self.conv3 = nn.Conv2d(feature_size, feature_size, 3, padding=1)
        self.act3 = nn.ReLU()
        self.conv4 = nn.Conv2d(feature_size, feature_size, 3, padding=1)
        self.act4 = nn.ReLU()
        self.output = nn.Conv2d(feature_size, num_anchors * num_classes, 3, padding=1)
        self.output_act = nn.Sigmoid()


class Flatten(nn.Module):
    def forward(self, x):
        return x.view(x.size(0), -1)


class Add(nn.Module):
    def forward(self, a, b):
        return torch.add(a, b)


class Concat(nn.Module):
    def forward(self, a, b, c, d):
        return torch.cat((a, b, c, d), 1)


class PRN(nn.Module):
    """posenet.py:130-152.  Off the conv hot path: three nn.Linear library calls."""

    def __init__(self, node_count, coeff):
        super().__init__()
        self.flatten = Flatten()
        self.height, self.width = coeff * 28, coeff * 18
        n = self.height * self.width * 17
        self.dens1 = nn.Linear(n, node_count)
        self.bneck = nn.Linear(node_count, node_count)
        self.dens2 = nn.Linear(node_count, n)
        self.drop = nn.Dropout()
        self.add = Add()
        self.softmax = nn.Softmax(dim=1)

    def forward(self, x):
        res = self.flatten(x)
        out = self.drop(F.relu(self.dens1(res)))
        out = self.drop(F.relu(self.bneck(out)))
        out = F.relu(self.dens2(out))
        out = self.softmax(self.add(out, res))
        return out.view(out.size(0), self.height, self.width, 17)


class poseNet(nn.Module):
    def __init__(self, layers, prn_node_count=1024, prn_coeff=2, precision=None):
        super().__init__()
        if layers == 101:
            self.fpn = FPN101()
        elif layers == 50:
            self.fpn = FPN50()
        else:
            raise ValueError("layers must be 50 or 101")
        self.layers = layers
        # keypoint subnet: intermediate supervision heads, per-level towers, fusion
        for k in (2, 3, 4, 5):
            setattr(self, "convfin_k%d" % k, nn.Conv2d(256, 19, 1))
        for i in (1, 2, 3, 4):
            setattr(self, "convt%d" % i, nn.Conv2d(256, 128, 3, padding=1))
        for i in (1, 2, 3, 4):
            setattr(self, "convs%d" % i, nn.Conv2d(128, 128, 3, padding=1))
        self.upsample1 = nn.Upsample(scale_factor=8, mode="nearest")
        self.upsample2 = nn.Upsample(scale_factor=4, mode="nearest")
        self.upsample3 = nn.Upsample(scale_factor=2, mode="nearest")
        self.concat = Concat()
        self.conv2 = nn.Conv2d(512, 256, 3, padding=1)
        self.convfin = nn.Conv2d(256, 18, 1)
        # detection subnet
        self.regressionModel = RegressionModel(256)
        self.classificationModel = ClassificationModel(256, num_classes=1)
        from .anchors import Anchors
        from .utils import BBoxTransform, ClipBoxes
        self.anchors = Anchors()
        self.regressBoxes = BBoxTransform()
        self.clipBoxes = ClipBoxes()
        # pose residual network
        self.prn = PRN(prn_node_count, prn_coeff)
        self._initialize_weights_norm()
        prior = 0.01
        with torch.no_grad():
            self.classificationModel.output.weight.fill_(0)
            self.classificationModel.output.bias.fill_(-math.log((1.0 - prior) / prior))
            self.regressionModel.output.weight.fill_(0)
            self.regressionModel.output.bias.fill_(0)
        self.freeze_bn()
        object.__setattr__(self, "_engines", {})
        self._precision = precision

    # -- reference API ---------------------------------------------------------------------------
    def _initialize_weights_norm(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                init.normal_(m.weight, std=0.01)
                if m.bias is not None:
                    init.constant_(m.bias, 0.0)

    def freeze_bn(self):
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()

    # -- engine cache: packed filters, streams and CUDA graphs live OUTSIDE the module state ------------------------
    def invalidate_engines(self):
        """Drop packed filters / captured graphs of every engine of this module.  Called by load_state_dict and _apply;
        call it yourself after in-place edits through `.data` views (`p.data.mul_()`), which no version counter records."""
        for e in self.__dict__.get("_engines", {}).values():
            inv = getattr(e, "invalidate", None)
            if inv is not None:
                inv()

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_engines()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate_engines()
        return out

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_engines"] = {}  # torch.save(model) / pickling: engines hold streams, graphs and packed device buffers
        return state

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = {} if k == "_engines" else copy.deepcopy(v, memo)
        return new

    def engine(self, precision=None):
        precision = precision or self._precision or _engine.DEFAULT_PRECISION
        # nn.DataParallel replicas share this dict (shallow __dict__ copy): key by module identity
        engines = self.__dict__.setdefault("_engines", {})
        key = (precision, id(self))
        e = engines.get(key)
        if e is None or e.model is not self:
            if len(engines) > 32:  # nn.DataParallel makes fresh replicas every call: drop the engines of dead replicas only
                for k in [k for k, v in engines.items() if v.model is not self and k[0] != "train"][:16]:
                    del engines[k]
            e = _engine.Engine(self, precision)
            engines[key] = e
        return e

    def train_engine(self, precision=None):
        from ..train_engine import TrainEngine
        precision = precision or self._precision or _engine.DEFAULT_PRECISION
        if precision == "f16f8":  # inference-only format: the weight-gradient kernel reads bf16 planes
            precision = "bf16x3"
        engines = self.__dict__.setdefault("_engines", {})
        key = ("train", precision, id(self))
        e = engines.get(key)
        if e is None or e.model is not self:
            e = TrainEngine(self, precision)
            engines[key] = e
        return e

    def detection_train_engine(self, precision=None):
        from ..detect_train import DetectionTrainEngine
        precision = precision or self._precision or _engine.DEFAULT_PRECISION
        if precision in ("f16f8", "fp32"):  # the weight-gradient kernel reads bf16 planes
            precision = "bf16x3"
        engines = self.__dict__.setdefault("_engines", {})
        key = ("train-det", precision, id(self))
        e = engines.get(key)
        if e is None or e.model is not self:
            e = DetectionTrainEngine(self, precision)
            engines[key] = e
        return e

    def forward(self, x):
        img_batch, subnet_name = x
        if subnet_name == "prn_subnet":
            return self.prn_forward(img_batch)
        bn_batch_stats = self.training and self.fpn.bn1.training  # model.train() without freeze_bn (trainer.py:170-174)
        wants_grad = torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters())
        if wants_grad and subnet_name == "detection_subnet":
            # detection-subnet training (training/multipose_detection_train.py): frozen trunk and BatchNorm, RetinaNet neck + towers
            if not (torch.is_tensor(img_batch) and img_batch.is_cuda):
                raise RuntimeError("poseNet.forward needs a CUDA image batch (no CPU / eager fallback)")
            from ..detect_train import DetectionTrainFunction
            deng = self.detection_train_engine()
            deng.check_frozen()
            with torch.cuda.device(img_batch.device):
                params = [p for _, p in deng.trainable_parameters()]
                cls, reg, anchors = DetectionTrainFunction.apply(deng, img_batch, *params)
            return [], [cls, reg, anchors]
        if wants_grad or (bn_batch_stats and subnet_name == "keypoint_subnet"):
            if subnet_name != "keypoint_subnet":
                raise NotImplementedError(
                    "the keypoint-subnet (BASELINE config 4) and detection-subnet training steps have backward kernels; "
                    "training through the entire_net branch is not a reference configuration and there is no eager fallback")
            if not bn_batch_stats:
                raise NotImplementedError("keypoint-subnet training with frozen BatchNorm (freeze_bn() in train mode) is not "
                                          "built: the reference trains this subnet with batch statistics (trainer.py:170-174)")
            if not (torch.is_tensor(img_batch) and img_batch.is_cuda):
                raise RuntimeError("poseNet.forward needs a CUDA image batch (no CPU / eager fallback)")
            from ..train_engine import KeypointTrainFunction
            teng = self.train_engine()
            with torch.cuda.device(img_batch.device):
                if wants_grad:
                    params = [p for _, p in teng.trainable_parameters()]
                    outs = KeypointTrainFunction.apply(teng, img_batch, *params)
                else:  # train mode under no_grad: batch statistics (and running-stat updates), nothing saved
                    with torch.no_grad():
                        outs, _ = teng.forward(img_batch)
            return outs[4], [outs[0], outs[1], outs[2], outs[3], outs[4]]
        if not (torch.is_tensor(img_batch) and img_batch.is_cuda):
            raise RuntimeError("poseNet.forward needs a CUDA image batch: the path runs on libmpn_b200 (sm_100a) only, "
                               "there is no CPU / eager fallback")
        eng = self.engine()
        with torch.cuda.device(img_batch.device):
            if subnet_name == "keypoint_subnet":
                return eng.keypoint_forward(img_batch)
            if subnet_name == "detection_subnet":
                return eng.detection_forward(img_batch)
            return eng.entire_forward(img_batch)

    def keypoint_forward(self, img_batch):
        return self.forward((img_batch, "keypoint_subnet"))

    def detection_forward(self, img_batch):
        return self.forward((img_batch, "detection_subnet"))

    def prn_forward(self, img_batch):
        if img_batch.is_cuda and not self.training:
            with torch.cuda.device(img_batch.device), torch.no_grad():
                out = self.engine().prn_forward(img_batch)  # batched, tensor cores (eval: dropout is the identity)
        else:
            # model.train(): nn.Dropout is active in the reference (posenet.py:341-342) whether or not grad is enabled, and
            # PRN training (multipose_prn_train.py) stays three nn.Linear library calls, SURVEY 8(f)
            out = self.prn(img_batch)
        return out, [out]

    @staticmethod
    def build_loss(saved_for_loss, *args):
        subnet_name = args[0]
        if subnet_name == "keypoint_subnet":
            return build_keypoint_loss(saved_for_loss, args[1], args[2])
        if subnet_name == "detection_subnet":
            return build_detection_loss(saved_for_loss, args[1])
        if subnet_name == "prn_subnet":
            return build_prn_loss(saved_for_loss, args[1])
        return 0


PoseNet = poseNet  # alias (BASELINE.json spells it with a capital P)


def build_names():
    names = []
    for j in range(2, 6):
        names += ["heatmap_loss_k%d" % j, "seg_loss_k%d" % j]
    return names + ["heatmap_loss", "seg_loss"]


def build_keypoint_loss(saved_for_loss, heat_temp, heat_weight):
    """posenet.py:367-403: sum over the 5 maps of MSE(pred[:, :18] * w, w * gt), mean reduction."""
    names = build_names()
    log = OrderedDict()
    total = 0
    target = heat_weight * heat_temp
    for j in range(5):
        loss = F.mse_loss(saved_for_loss[j][:, :18] * heat_weight, target)
        total = total + loss
        log[names[2 * j]] = loss.item()
    last = saved_for_loss[-1].detach()[:, :18]
    log["max_ht"] = last.max().item()
    log["min_ht"] = last.min().item()
    return total, log


def build_detection_loss(saved_for_loss, anno):
    """posenet.py:405-424: saved_for_loss = [classifications, regressions, anchors]; mean focal + smooth-L1 loss."""
    from .losses import FocalLoss
    log = OrderedDict()
    classification_loss, regression_loss = FocalLoss()(*saved_for_loss, anno)
    classification_loss = classification_loss.mean()
    regression_loss = regression_loss.mean()
    total = classification_loss + regression_loss
    log["total_loss"] = total.item()
    log["classification_loss"] = classification_loss.item()
    log["regression_loss"] = regression_loss.item()
    return total, log


def build_prn_loss(saved_for_loss, label):
    """posenet.py:426-445."""
    loss = F.binary_cross_entropy(saved_for_loss[0], label)
    return loss, OrderedDict([("PRN loss", loss.item())])
