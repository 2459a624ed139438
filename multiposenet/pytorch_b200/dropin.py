"""Register this package under the reference's import names.

The reference's harness does `from network.posenet import poseNet` (evaluate/multipose_test.py:6,
training/multipose_keypoint_train.py:11) and `from lib.nms.pth_nms import pth_nms`
(network/posenet.py:16).  install_dropin() aliases those module names to this package's mirrors, so
evaluate/ and training/ scripts import the B200 implementation without being edited.  Other
`network.*` / `lib.*` submodules of a reference checkout on sys.path (joint_utils, net_utils, lib.utils)
keep resolving to the checkout: only the hot-path modules are replaced, plus `network.losses` (FocalLoss on the device) and
`training.batch_processor` (the reference's file is a SyntaxError on Python >= 3.7).
"""
import importlib
import sys
import types


def install_dropin(reference_root=None):
    from .lib.nms import pth_nms as _pth
    from .network import anchors as _anchors
    from .network import fpn as _fpn
    from .network import posenet as _posenet
    from .network import utils as _utils

    if reference_root and reference_root not in sys.path:
        sys.path.insert(0, reference_root)

    def _pkg(name):
        m = sys.modules.get(name)
        if m is None:
            try:
                m = importlib.import_module(name)  # a real package on sys.path (reference checkout)
            except Exception:
                m = types.ModuleType(name)
                m.__path__ = []
            sys.modules[name] = m
        return m

    from .network import losses as _losses
    _bp = importlib.import_module(__package__ + ".training.batch_processor")   # the module (the package re-exports the function)
    for pkg in ("network", "lib", "lib.nms", "training"):
        _pkg(pkg)
    # network.losses: FocalLoss on the device (the reference's own file needs two torch-0.4 idioms that fail on torch >= 1.2);
    # training.batch_processor: the reference file is a SyntaxError on Python >= 3.7 (`async=` keyword)
    for name, mod in (("network.posenet", _posenet), ("network.fpn", _fpn), ("network.anchors", _anchors),
                      ("network.utils", _utils), ("network.losses", _losses), ("lib.nms.pth_nms", _pth),
                      ("training.batch_processor", _bp)):
        sys.modules[name] = mod
        parent, _, leaf = name.rpartition(".")
        setattr(sys.modules[parent], leaf, mod)
    return _posenet.poseNet
