"""CPU: the C-ABI library loads and exports every symbol include/mpn_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    h = open(os.path.join(ROOT, "include", "mpn_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(mpn_[a-z0-9_]+)\s*\(", h)))


def test_header_symbols_exported():
    from multiposenet.pytorch_b200 import _lib
    from multiposenet.pytorch_b200.csrc import build
    build.build()
    names = _declared()
    assert len(names) >= 20
    assert set(names) == set(_lib.EXPORTS), set(names) ^ set(_lib.EXPORTS)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), n
    L = _lib.lib()
    assert L.mpn_version() >= 100
    assert L.mpn_num_anchors(480, 640) == 57600
    # the two by-pointer structs have the layout the library was compiled with (also checked at every load)
    assert ctypes.sizeof(_lib.ConvDesc) == L.mpn_sizeof_conv_desc()
    assert ctypes.sizeof(_lib.ConvPtrs) == L.mpn_sizeof_conv_ptrs()


def test_host_side_entry_points_and_errors():
    from multiposenet.pytorch_b200 import _lib
    L = _lib.lib()
    g = np.load(os.path.join(ROOT, "tests", "golden", "anchors.npz"))
    for key in g.files:
        h, w = map(int, key.split("x"))
        out = np.zeros((L.mpn_num_anchors(h, w), 4), np.float32)
        _lib.check(L.mpn_generate_anchors(h, w, out.ctypes.data_as(ctypes.c_void_p)))
        assert np.array_equal(out.view(np.uint32), g[key].view(np.uint32)), key
    # argument errors are reported, not crashed on (reference: THArgCheck -> RuntimeError, nms.c:6-9)
    rc = L.mpn_generate_anchors(0, 10, None)
    assert rc != 0
    with pytest.raises(_lib.MpnError):
        _lib.check(rc, "mpn_generate_anchors")
    d = _lib.ConvDesc()
    p = _lib.ConvPtrs()
    assert L.mpn_conv2d_fwd(ctypes.byref(d), ctypes.byref(p), None) != 0
    assert b"conv" in L.mpn_last_error()


def test_state_dict_matches_reference_inventory():
    from multiposenet.pytorch_b200 import poseNet
    from oracle import weights
    for layers, n in ((50, 402), (101, 708)):
        spec = weights.param_spec(layers)
        sd = poseNet(layers).state_dict()
        assert len(sd) == n and list(sd.keys()) == list(spec.keys())
        assert all(tuple(sd[k].shape) == tuple(spec[k]) for k in sd)


def test_reference_init_statistics():
    import math
    from multiposenet.pytorch_b200 import poseNet
    m = poseNet(50)
    assert float(m.classificationModel.output.weight.abs().max()) == 0.0
    assert abs(float(m.classificationModel.output.bias[0]) + math.log(99.0)) < 1e-6
    assert float(m.regressionModel.output.weight.abs().max()) == 0.0
    assert abs(float(m.conv2.weight.std()) - 0.01) < 1e-3 and float(m.conv2.bias.abs().max()) == 0.0
    assert not m.fpn.bn1.training  # freeze_bn (posenet.py:220-224)
    m.train()
    m.freeze_bn()
    assert not m.fpn.layer1[0].bn1.training


def test_no_cpu_fallback():
    import torch
    from multiposenet.pytorch_b200 import poseNet, pth_nms
    m = poseNet(50).eval()
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            m([torch.zeros(1, 3, 64, 64), "keypoint_subnet"])
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            pth_nms(torch.zeros(4, 5), 0.5)
