"""Multi-scale / flip TTA on the device vs the host restatement of evaluate/tester.py:256-331 (cv2)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_resize_cubic_vs_cv2():
    cv2 = pytest.importorskip("cv2")
    from multiposenet.pytorch_b200 import ops
    rng = np.random.Generator(np.random.PCG64(0))
    for (sh, sw, dh, dw) in ((12, 16, 48, 64), (30, 41, 97, 131), (45, 33, 20, 17), (7, 5, 7, 5)):
        src = rng.standard_normal((6, sh + 3, sw + 2)).astype(np.float32)        # valid region smaller than the pitch
        want = np.stack([cv2.resize(np.ascontiguousarray(p[:sh, :sw]), (dw, dh), interpolation=cv2.INTER_CUBIC) for p in src])
        got = ops.resize_cubic(torch.from_numpy(src).cuda(), sh, sw, dh, dw, 1.0 / (float(dw) / sw), 1.0 / (float(dh) / sh)).cpu().numpy()
        # measured on B200 vs the wheel's cv2 4.13: 3.9e-6 of the plane maximum at non-integer scales (typical element: 1-3 ulp);
        # OpenCV's own dispatch (IPP / AVX vector body / scalar tail) moves results by the same order between calls
        assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max(), (sh, sw, dh, dw)
    src = rng.standard_normal((18, 10, 12)).astype(np.float32)
    want = np.stack([cv2.resize(p, None, fx=4, fy=4, interpolation=cv2.INTER_CUBIC) for p in src])
    got = ops.resize_cubic(torch.from_numpy(src).cuda(), 10, 12, 40, 48, 0.25, 0.25).cpu().numpy()
    assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()


def test_tta_device_path_vs_reference_golden(golden_dir):
    """The stub model of the golden behind the device-side resize / accumulate / mirror / swap / average."""
    pytest.importorskip("cv2")
    from multiposenet.pytorch_b200.evaluate import tta
    from oracle import tta_oracle as to
    g = np.load(os.path.join(golden_dir, "tta.npz"))
    meta = json.loads(str(g["meta"]))
    img = to.test_image()

    class Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1))
            self.last = None

        def forward(self, x):
            im = x[0].cpu().numpy()
            heats, outs = [], None
            for b in range(im.shape[0]):
                h, s, c, bx = to.stub_model(im[b:b + 1])
                heats.append(h)
                outs = outs or (s, c, bx)
            return torch.from_numpy(np.concatenate(heats)).cuda(), [torch.from_numpy(a).cuda() for a in outs]

        def engine(self):
            class E(object):
                last_detections = None
            return E()
    m = Stub().cuda()
    mult = tta.get_multiplier(img, meta["inp_size"])
    assert np.array_equal(np.array(mult), g["multiplier"])
    hn, bn = tta.get_outputs(m, mult, img)
    scale = np.abs(g["heat_normal"]).max()
    assert hn.shape == g["heat_normal"].shape and hn.dtype == np.float64
    assert np.abs(hn - g["heat_normal"]).max() <= 1e-5 * scale
    assert bn == json.loads(str(g["bbox_normal"]))
    h64, h32, b_n, b_f = tta.multi_scale_flip(m, img, inp_size=meta["inp_size"])
    avg = h64.permute(1, 2, 0).cpu().numpy()
    assert np.abs(avg - g["heat_avg"]).max() <= 1e-5 * scale
    assert h32.shape == (1, 18, img.shape[0], img.shape[1]) and h32.dtype == torch.float32
    assert np.abs(tta.handle_heat(g["heat_normal"].astype(np.float64), g["heat_flipped"].astype(np.float64)) - g["heat_avg"]).max() <= 1e-6 * scale


def test_tta_real_model_batch2_equals_two_batch1_passes():
    """The original + mirrored image of a scale as ONE batch-2 forward gives the heat maps of two batch-1 forwards, and the
    whole device pipeline equals the host restatement fed with the same network."""
    pytest.importorskip("cv2")
    from gpu_util import load_model
    from multiposenet.pytorch_b200.evaluate import tta
    from multiposenet.pytorch_b200.network import joint_utils
    from oracle import tta_oracle as to
    m, _ = load_model(50, "conditioned", "f16f8")
    img = to.test_image(seed=5, hw=(70, 94))

    def model_fn(im_data):
        with torch.no_grad():
            heat, (s, c, b) = m([torch.from_numpy(im_data).cuda(), "both"])
        return heat.cpu().numpy(), s.cpu().numpy(), c.cpu().numpy(), b.cpu().numpy()
    mult = to.get_multiplier(img, 64)
    want_n, bb_n = to.get_outputs(model_fn, mult, img)
    want_f, _ = to.get_outputs(model_fn, mult, img[:, ::-1, :])
    want = to.handle_heat(want_n, want_f)
    h64, h32, b_n, b_f = tta.multi_scale_flip(m, img, inp_size=64)
    got = h64.permute(1, 2, 0).cpu().numpy()
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
    assert b_n == bb_n
    rows = joint_utils.get_joint_list(img, {"thre1": 0.1}, h32, 1)   # tester.py:158 on the device tensor
    assert rows.shape[1] == 5
