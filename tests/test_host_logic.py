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
