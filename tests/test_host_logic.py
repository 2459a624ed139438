"""CPU tests of the host-side logic around the kernels (no GPU, no compute calls)."""
import numpy as np
import torch


def test_weight_signature_sees_every_kind_of_update():
    """engine.Engine._signature guards the packed-filter cache and the captured graphs: it must change for in-place
    updates, load_state_dict, dtype/device moves, .data swaps and removed parameters, and stay equal otherwise."""
    from multiposenet.pytorch_b200 import engine as E, poseNet
    m = poseNet(50)
    e = E.Engine(m, "bf16x3")
    s0 = e._signature()
    assert e._signature() == s0
    with torch.no_grad():
        m.conv2.weight.add_(1.0)
    s1 = e._signature()
    assert s1 != s0
    m.load_state_dict(m.state_dict())
    s2 = e._signature()
    assert s2 != s1
    m.convfin.weight.data = torch.zeros_like(m.convfin.weight)   # no version bump: caught by the storage address
    s3 = e._signature()
    assert s3 != s2
    m.double()
    s4 = e._signature()
    assert s4 != s3
    m.conv2.bias = None                                           # slot disappears: signature changes, then settles
    s5 = e._signature()
    assert s5 != s4
    s6, s7 = e._signature(), e._signature()
    assert s6 == s7 and s6 != s4
    with torch.no_grad():                                         # batch-norm buffers count too (running stats)
        m.fpn.bn1.running_mean.add_(1.0)
    assert e._signature() != s7


def test_boxes_for_prn_matches_the_reference_loop():
    """tester.py:232-240 multiplies each kept box by the scale on its own; the vectorised version must give the same
    python floats, keep the descending-score order and honour the cap."""
    from multiposenet.pytorch_b200.evaluate.pipeline import boxes_for_prn
    rng = np.random.default_rng(3)
    boxes = (rng.random((40, 4)) * 600).astype(np.float32)
    scores = np.sort(rng.random(40).astype(np.float32))[::-1].copy()
    for scale in (1.0, 1.37, 0.5):
        want = [(np.asarray(boxes[i]) * scale).tolist() for i in np.where(scores > 0.5)[0]]
        got = boxes_for_prn(scores, boxes, scale)
        assert got == want and all(isinstance(v, float) for b in got for v in b)
        assert boxes_for_prn(scores, boxes, scale, limit=3) == want[:3]
    assert boxes_for_prn(scores, boxes, 1.0, score_thresh=2.0) == []
    assert boxes_for_prn(np.zeros(0, np.float32), np.zeros((0, 4), np.float32), 1.0) == []


def test_precision_defaults():
    """Inference engines default to f16f8; an f16f8 model trains on bf16x3 planes; explicit choices are kept."""
    import os
    from multiposenet.pytorch_b200 import engine as E, poseNet
    if not os.environ.get("MPN_PRECISION"):
        assert E.DEFAULT_PRECISION == "f16f8"
        assert poseNet(50).engine().precision == "f16f8"
    m = poseNet(50, precision="f16f8")
    assert m.engine().precision == "f16f8"
    assert m.train_engine().precision == "bf16x3"
    assert poseNet(50, precision="bf16").train_engine().precision == "bf16"
    assert poseNet(50, precision="bf16x3").engine().precision == "bf16x3"


def test_joints_for_prn_and_regroup():
    """tester.py:222-229 drops the neck rows (type 1) and shifts later joint types down by one; prn_process regroups the
    rows by joint type in their original order (tester.py:337-350)."""
    from multiposenet.pytorch_b200.evaluate.pipeline import joints_for_prn
    from multiposenet.pytorch_b200.evaluate.prn_assign import _regroup
    rows = np.array([[10, 11, 0.9, 0, 0], [20, 21, 0.8, 1, 1], [30, 31, 0.7, 2, 2], [40, 41, 0.6, 3, 17], [50, 51, 0.5, 4, 2],
                     [60, 61, 0.4, 5, 0]], dtype=np.float64)
    out = joints_for_prn(rows)
    assert out[:, 4].tolist() == [0, 1, 16, 1, 0]                 # neck gone, types 2.. shifted down, type 0 kept
    assert out[:, 0].tolist() == [10, 30, 40, 50, 60] and out.dtype == np.float64
    assert joints_for_prn(np.zeros((0, 5))).shape == (0, 5)
    xy, ty = _regroup(out)
    assert ty.tolist() == [0, 0, 1, 1, 16] and ty.dtype == np.int32
    assert xy[:, 0].tolist() == [10, 60, 30, 50, 40]              # stable within a joint type
    xy0, ty0 = _regroup([])
    assert xy0.shape == (0, 2) and ty0.shape == (0,)
    xy1, ty1 = _regroup(np.array([[1, 2, 0.5, 0, 17], [3, 4, 0.5, 1, 3]], dtype=np.float64))   # types outside 0..16 are ignored
    assert ty1.tolist() == [3] and xy1.tolist() == [[3.0, 4.0]]


def test_signature_sees_module_replacement_and_bn_eps_and_invalidate():
    """a replaced submodule (`m.convfin = nn.Conv2d(...)`) and a changed bn.eps must change the signature; writes
    through `.data` views bump no version counter, so Engine.invalidate() / poseNet.invalidate_engines() is the contract."""
    import copy
    import pickle
    from multiposenet.pytorch_b200 import poseNet
    m = poseNet(50, precision="bf16x3")
    e = m.engine()
    s0 = e._signature()
    m.convfin = torch.nn.Conv2d(256, 18, 1)
    s1 = e._signature()                       # stale walk detected ...
    s2, s3 = e._signature(), e._signature()   # ... re-collected, then stable
    assert s1 != s0 and s2 == s3 and s2 != s0
    m.fpn.layer1[0].bn1.eps = 1e-3
    s4 = e._signature()
    assert s4 != s3
    m.fpn.layer2[1] = copy.deepcopy(m.fpn.layer2[1])   # a replaced block deep in the tree
    assert e._signature() != s4
    s5 = e._signature()
    assert e._signature() == s5
    # .data in-place edits are invisible to the signature (documented) -> explicit invalidation
    e._sig, e._packed = s5, {"x": 1}
    m.conv2.weight.data.mul_(2.0)
    assert e._signature() == s5
    m.invalidate_engines()
    assert e._packed == {} and e._sig is None and e._graphs == {}
    # load_state_dict and _apply invalidate by themselves
    e._packed = {"x": 1}
    m.load_state_dict(m.state_dict())
    assert e._packed == {}
    e._packed = {"x": 1}
    m.float()
    assert e._packed == {}
    # engines stay out of copies and pickles (they hold streams, graphs and packed device buffers)
    m2 = copy.deepcopy(m)
    assert m2.__dict__["_engines"] == {} and m.__dict__["_engines"]
    assert m2.engine() is not e and m2.engine().model is m2
    m3 = pickle.loads(pickle.dumps(m))
    assert m3.__dict__["_engines"] == {}
    assert torch.equal(m3.conv2.weight, m.conv2.weight)


def test_graph_policy_second_call_and_lru():
    """The public forward replays a CUDA graph from the second call with the same key; captured graphs are LRU-capped."""
    from multiposenet.pytorch_b200 import engine as E, poseNet
    e = E.Engine(poseNet(50), "f16f8")
    x = torch.zeros(2, 3, 64, 96)
    if not E.USE_GRAPHS:
        return
    assert e._wants_graph("entire", x, max_cand=4096) is False     # first sight of this shape: eager
    assert e._wants_graph("entire", x, max_cand=4096) is True      # second call: capture + replay
    assert e._wants_graph("entire", torch.zeros(2, 3, 96, 96), max_cand=4096) is False
    assert e._graph_key("entire", x, {"max_cand": 4096}) != e._graph_key("keypoint", x, {})


def test_prn_degenerate_boxes_follow_the_reference():
    """tester.py divides by ceil(w), ceil(h) only for boxes that contain a peak (impossible for w <= 0) or in the
    fallback branch (some joint type without a peak inside any box), where ceil == 0 raises ZeroDivisionError."""
    from multiposenet.pytorch_b200.evaluate.prn_assign import _reference_would_divide_by_zero as zdiv
    xy = np.array([[50.0, 60.0]] * 17)
    ty = np.arange(17, dtype=np.int32)
    good = np.array([[10.0, 10.0, 100.0, 200.0]])
    flat = np.array([[10.0, 10.0, -0.5, 50.0]])       # ceil(w) == 0
    neg = np.array([[10.0, 10.0, -30.0, 50.0]])       # ceil(w) == -30: the reference divides happily
    assert not zdiv(xy, ty, good, 0.21)
    assert not zdiv(xy, ty, np.concatenate([good, flat]), 0.21)        # every type has a peak in `good`: fallback not reached
    assert zdiv(xy[:16], ty[:16], np.concatenate([good, flat]), 0.21)  # type 16 has no peak -> fallback over ALL boxes -> 1/0
    assert not zdiv(xy[:16], ty[:16], np.concatenate([good, neg]), 0.21)
    assert zdiv(np.zeros((0, 2)), np.zeros((0,), np.int32), flat, 0.21)
    # cross-check against the python restatement of the reference method
    from oracle import prn_oracle
    import pytest
    kps = [[50.0, 60.0, 0.9, i, i] for i in range(16)]
    with pytest.raises(ZeroDivisionError):
        prn_oracle.prn_process(kps, [[10, 10, 110, 210], [10, 10, 9.5, 60]], prn_oracle.synthetic_prn(0))


def test_compat_batch_processor_contract():
    """training/batch_processor.py:10-60: (inputs, gts, saved_for_eval) for the three subnets; the mirror must import on this
    Python (the reference file does not: `async=` keyword) and be registered by install_dropin()."""
    import sys
    import types
    from multiposenet.pytorch_b200 import install_dropin
    from multiposenet.pytorch_b200.training import batch_processor
    install_dropin()
    assert sys.modules["training.batch_processor"].batch_processor is batch_processor
    assert hasattr(sys.modules["network.losses"], "FocalLoss")
    moved = []

    class T(object):   # stands in for a CPU tensor: records the device it is sent to
        def to(self, dev, non_blocking=False):
            moved.append(str(dev))
            return self

        def float(self):
            return self
    state = types.SimpleNamespace(params=types.SimpleNamespace(gpus=[0], subnet_name="keypoint_subnet"),
                                  model=types.SimpleNamespace(training=True))
    a, b, c = T(), T(), T()
    inputs, gts, saved = batch_processor(state, (a, b, c))
    assert inputs == [[a, "keypoint_subnet"]] and gts == ["keypoint_subnet", b, c] and saved == []
    assert moved == ["cuda:0"] * 3
    state.params.subnet_name = "detection_subnet"
    state.model.training = False
    inputs, gts, _ = batch_processor(state, (a, b))
    assert inputs == [[a, "detection_subnet"]] and gts == ["detection_subnet", b]
    state.params.subnet_name = "prn_subnet"
    inputs, gts, _ = batch_processor(state, (a, b))
    assert inputs[0][1] == "prn_subnet" and gts[0] == "prn_subnet"


def test_phase_class_filter_identity():
    """ops.phase_class_filter: conv3x3(nearest_upsample(x, f), W) == per output phase one of nine 3x3 convolutions of x
    (the identity behind mpn_conv_desc.gat_*), for f = 2, 4, 8 including the zero-padded borders."""
    import torch
    import torch.nn.functional as F
    from multiposenet.pytorch_b200.ops import phase_class_filter
    g = torch.Generator().manual_seed(0)
    for sh in (1, 2, 3):
        f = 1 << sh
        x = torch.randn(2, 5, 3, 4, generator=g, dtype=torch.float64)
        w = torch.randn(7, 5, 3, 3, generator=g)
        ref = F.conv2d(F.interpolate(x, scale_factor=f, mode="nearest"), w.double(), padding=1)
        z = F.conv2d(x, phase_class_filter(w).double(), padding=1).reshape(2, 9, 7, 3, 4)
        oh, ow = torch.arange(3 * f), torch.arange(4 * f)
        rc = torch.where(oh % f == 0, 0, torch.where(oh % f == f - 1, 2, 1))
        cc = torch.where(ow % f == 0, 0, torch.where(ow % f == f - 1, 2, 1))
        out = z[:, rc[:, None] * 3 + cc[None, :], :, (oh // f)[:, None], (ow // f)[None, :]].permute(2, 3, 0, 1)
        assert float((out - ref).abs().max()) < 1e-5
