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
