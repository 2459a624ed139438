"""Import the UNMODIFIED reference network modules from /root/reference (TEST INFRASTRUCTURE).

Only usable where /root/reference exists (the build container).  Two shims, SURVEY.md 8(c):
  1. lib.nms.pth_nms is pre-seeded with a stub (its cffi binary needs torch.utils.ffi, removed
     in torch >= 1.0); the stub routes to the C restatement oracle/nms_oracle.c.
  2. on a CPU-only box torch.Tensor.cuda is made the identity (utils.py:11,15, anchors.py:37).
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

REF_ROOT = os.environ.get("MPN_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "network"))


def import_reference():
    """Returns the reference's network.posenet module."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    from . import nms_oracle

    def pth_nms(dets, thresh):
        d = dets.detach().cpu().numpy().astype(np.float32)
        if dets.is_cuda:
            keep = nms_oracle.nms_gpu_semantics(d, float(thresh))
        else:
            # evaluate/ feeds CUDA dets -> GPU-branch semantics are the contract
            keep = nms_oracle.nms_gpu_semantics(d, float(thresh))
        return torch.from_numpy(keep).to(dets.device)

    for name in ("lib", "lib.nms"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    stub = types.ModuleType("lib.nms.pth_nms")
    stub.pth_nms = pth_nms
    sys.modules["lib.nms.pth_nms"] = stub
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # our drop-in may have registered itself under the same module names: evict it
    for k in [k for k in sys.modules if k == "network" or k.startswith("network.")]:
        mod = sys.modules[k]
        f = getattr(mod, "__file__", "") or ""
        if not f.startswith(REF_ROOT):
            del sys.modules[k]
    warnings.filterwarnings("ignore", category=UserWarning)
    import network.posenet as ref_posenet  # noqa: E402

    assert ref_posenet.__file__.startswith(REF_ROOT), ref_posenet.__file__
    return ref_posenet


def build_reference_model(layers, weights):
    """poseNet(layers) from the reference with `weights` (numpy dict from oracle.weights) loaded."""
    ref = import_reference()
    torch.manual_seed(0)
    model = ref.poseNet(layers)
    sd = model.state_dict()
    new = {}
    for k, v in sd.items():
        if k in weights:
            assert tuple(v.shape) == tuple(weights[k].shape), (k, v.shape, weights[k].shape)
            new[k] = torch.from_numpy(np.ascontiguousarray(weights[k]))
        else:
            new[k] = v  # PRN matrices when not generated
    model.load_state_dict(new)
    model.eval()
    return model
