"""GPU parity of the full drop-in poseNet vs the reference goldens (tests/golden, made from /root/reference)
and vs the oracle restatement run live in fp32."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# max|a-b|/max|b| per output tensor (north_star: 1e-3 vs the fp32 reference)
TOL = {"fp32": 1e-4, "bf16x3": 1e-3, "f16f8": 1e-3}


def _run_all(m, x):
    with torch.no_grad():
        heat, saved = m([x, "keypoint_subnet"])
        _, (cls, reg, anc) = m([x, "detection_subnet"])
        heat2, (sc, cl, bx) = m((x, "both"))
    return heat, saved, cls, reg, anc, heat2, sc, cl, bx


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "f16f8"])
@pytest.mark.parametrize("name", ["r50_cond_64x96_b2", "r50_refinit_64x96_b1", "r101_cond_64x96_b1"])
def test_network_vs_reference_golden(golden_dir, name, precision):
    from gpu_util import image, load_model, nerr
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    meta = json.loads(str(g["meta"]))
    m, _ = load_model(meta["layers"], meta["kind"], precision)
    x = image(meta["img_seed"], (meta["batch"], 3) + tuple(meta["hw"]))
    heat, saved, cls, reg, anc, heat2, sc, cl, bx = _run_all(m, x)
    tol = TOL[precision]
    T = lambda k: torch.from_numpy(g[k]).cuda()
    assert heat.shape == T("kp_heat").shape
    assert nerr(heat, T("kp_heat")) <= tol
    for i in range(4):
        assert saved[i].shape == T("kp_saved%d" % i).shape
        assert nerr(saved[i], T("kp_saved%d" % i)) <= tol, i
    assert saved[4] is heat
    assert nerr(cls, T("det_cls")) <= tol and nerr(reg, T("det_reg")) <= tol
    assert np.array_equal(anc.cpu().numpy(), g["det_anchors"])
    assert nerr(heat2, T("both_heat")) <= tol
    # detections: same K and the same boxes unless a score/IoU sits within rounding of a threshold
    assert len(sc) == len(g["both_scores"])
    if len(g["both_scores"]):
        assert cl.dtype == torch.int64 and np.array_equal(cl.cpu().numpy(), g["both_classes"])
        np.testing.assert_allclose(sc.cpu().numpy(), g["both_scores"], rtol=5e-3)
        np.testing.assert_allclose(bx.cpu().numpy(), g["both_boxes"], rtol=5e-3, atol=0.05)
    else:
        assert not sc.is_cuda and bx.shape == (0, 4)  # reference returns CPU zeros (posenet.py:275)


def test_network_480x640_vs_golden_and_oracle(golden_dir):
    from gpu_util import image, load_model, nerr, no_tf32
    from oracle import posenet_oracle as po, weights
    no_tf32()
    g = np.load(os.path.join(golden_dir, "r50_cond_480x640_b1.npz"))
    meta = json.loads(str(g["meta"]))
    m, w = load_model(50, "conditioned", "bf16x3")
    x = image(meta["img_seed"], (1, 3, 480, 640))
    heat, saved, cls, reg, anc, heat2, sc, cl, bx = _run_all(m, x)
    assert heat.shape == (1, 18, 120, 160) and cls.shape == (1, 57600, 1) and reg.shape == (1, 57600, 4)
    T = lambda k: torch.from_numpy(g[k]).cuda()
    assert nerr(heat[:, :, ::4, ::4], T("kp_heat")) <= 1e-3
    assert nerr(cls, T("det_cls")) <= 1e-3
    assert nerr(reg[:, ::5], T("det_reg")) <= 1e-3
    # live oracle on the GPU in fp32 (cuDNN, TF32 off): full-resolution comparison
    sd = {k: v.cuda() for k, v in weights.to_torch_state_dict(w).items()}
    with torch.no_grad():
        oheat, osaved = po.forward(sd, 50, x, "keypoint_subnet")
    assert nerr(heat, oheat) <= 1e-3
    for a, b in zip(saved[:4], osaved[:4]):
        assert nerr(a, b) <= 1e-3
    # NMS stage fed with OUR cls/boxes through the oracle: kept indices bit-exact
    from multiposenet.pytorch_b200 import ops
    from oracle import nms_oracle
    boxes = ops.decode_clip(anc, reg, 480, 640)
    mask = (cls[0, :, 0] > 0.05)
    d = torch.cat([boxes[0][mask], cls[0][mask]], 1).cpu().numpy()
    want = nms_oracle.nms_gpu_semantics(d, 0.5)
    assert len(sc) == len(want)
    assert np.array_equal(sc.cpu().numpy(), d[want, 4]) and np.array_equal(bx.cpu().numpy(), d[want, :4])
    det = m.engine().last_detections
    assert np.array_equal(det.keep_idx[0, :len(want)].cpu().numpy(), want)


def test_bf16_fast_mode_error_is_bounded(golden_dir):
    from gpu_util import image, load_model, nerr
    g = np.load(os.path.join(golden_dir, "r50_cond_64x96_b2.npz"))
    m, _ = load_model(50, "conditioned", "bf16")
    x = image(11, (2, 3, 64, 96))
    with torch.no_grad():
        heat, _ = m([x, "keypoint_subnet"])
    e = nerr(heat, torch.from_numpy(g["kp_heat"]).cuda())
    print("bf16 single-pass heat error: %.3g" % e)
    assert e < 0.1


def test_dropin_surface():
    import sys
    from multiposenet.pytorch_b200 import install_dropin
    install_dropin()
    from network.posenet import poseNet  # noqa: the reference's import line (evaluate/multipose_test.py:6)
    from lib.nms.pth_nms import pth_nms  # noqa
    m = poseNet(50)
    names = [n for n, _ in m.named_children()]
    for want in ["fpn", "convfin_k2", "convt1", "convs4", "upsample1", "conv2", "convfin", "regressionModel",
                 "classificationModel", "prn"]:
        assert want in names
    assert hasattr(m, "freeze_bn") and hasattr(m, "build_loss")
    with pytest.raises(RuntimeError):
        m([torch.zeros(1, 3, 64, 64), "keypoint_subnet"])  # CPU tensors: no fallback
    for k in [k for k in sys.modules if k.startswith(("network", "lib.nms"))]:
        pass


def test_cuda_graph_replay_matches_eager():
    from gpu_util import image, load_model
    m, _ = load_model(50, "conditioned", "bf16x3")
    eng = m.engine()
    x1, x2 = image(21, (2, 3, 96, 128)), image(22, (2, 3, 96, 128))
    for x in (x1, x2, x1):
        heat_e, cls_e, reg_e, boxes_e, det_e = eng.entire_forward_device(x, max_cand=4096)
        keep_e = det_e.keep_idx.clone(); cnt_e = det_e.keep_cnt.clone(); heat_e = heat_e.clone()
        heat_g, cls_g, reg_g, boxes_g, det_g = eng.graphed("entire", x, max_cand=4096)
        torch.cuda.synchronize()
        assert torch.equal(heat_g, heat_e)
        assert torch.equal(det_g.keep_cnt, cnt_e)
        k = int(cnt_e[0])
        assert torch.equal(det_g.keep_idx[0, :k], keep_e[0, :k])


def test_cuda_graph_follows_weight_updates():
    """graphed() replays first and validates the weights while the GPU runs: after an in-place update, a load_state_dict
    or a dtype round trip the stale replay must be discarded and the result must be the eager result of the NEW weights."""
    from gpu_util import image, load_model
    m, _ = load_model(50, "conditioned", "f16f8")
    eng = m.engine()
    x = image(23, (2, 3, 96, 128))
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    h0 = eng.graphed("keypoint", x)[0].clone()
    assert torch.equal(eng.graphed("keypoint", x)[0], h0)          # fast path, unchanged weights
    with torch.no_grad():
        m.convfin.bias.add_(0.25)
    h1 = eng.graphed("keypoint", x)[0].clone()
    assert torch.equal(h1, eng.keypoint_forward(x)[0])
    assert float((h1 - h0).abs().min()) > 0.2                       # every heat-map value moved with the bias
    m.load_state_dict(sd0)                                          # the original weights, bit for bit
    h2 = eng.graphed("keypoint", x)[0].clone()
    assert torch.equal(h2, h0)
    m.double().float()                                              # new storages, same values
    assert torch.equal(eng.graphed("keypoint", x)[0], h0)
    assert torch.equal(eng.graphed("keypoint", x)[0], h0)


@pytest.mark.parametrize("hw", [(100, 130), (75, 50)])
def test_detection_subnet_ragged_sizes_vs_oracle(hw):
    """Odd image sizes: ceil-shaped pyramid levels, non-2x nearest upsample-add (fpn.py:84-95), 1x1 P7."""
    from gpu_util import image, load_model, nerr, no_tf32
    from oracle import posenet_oracle as po, weights
    no_tf32()
    m, w = load_model(50, "conditioned", "bf16x3")
    x = image(31, (2, 3) + hw)
    with torch.no_grad():
        _, (cls, reg, anc) = m([x, "detection_subnet"])
        sd = {k: v.cuda() for k, v in weights.to_torch_state_dict(w).items()}
        _, (ocls, oreg, oanc) = po.forward(sd, 50, x, "detection_subnet")
    assert cls.shape == ocls.shape and reg.shape == oreg.shape
    assert torch.equal(anc, oanc)
    assert nerr(cls, ocls) <= 1e-3 and nerr(reg, oreg) <= 1e-3


@pytest.mark.parametrize("precision,tol", [("bf16x3", 1e-3), ("bf16", 5e-2)])
def test_prn_batched_vs_oracle(precision, tol):
    """posenet.py:337-350 on the tensor-core path (small PRN: coeff 1, 256 nodes; K = 8568 is not a multiple of 64)."""
    from gpu_util import nerr, no_tf32
    from multiposenet.pytorch_b200 import poseNet
    from oracle import posenet_oracle as po
    no_tf32()
    torch.manual_seed(0)
    m = poseNet(50, prn_node_count=256, prn_coeff=1, precision=precision).cuda().eval()
    with torch.no_grad():
        for lin in (m.prn.dens1, m.prn.bneck, m.prn.dens2):
            lin.weight.normal_(0, 1.0 / lin.weight.shape[1] ** 0.5)
            lin.bias.normal_(0, 0.1)
    x = torch.rand(37, 28, 18, 17, device="cuda") * 3
    with torch.no_grad():
        out, saved = m([x, "prn_subnet"])
        want, _ = po.prn_forward({k: v for k, v in m.state_dict().items()}, x)
    assert out.shape == (37, 28, 18, 17) and saved[0] is out
    assert nerr(out, want) <= tol
    assert torch.allclose(out.reshape(37, -1).sum(1), torch.ones(37, device="cuda"), atol=1e-4)


def test_uint8_input_with_fused_resnet_preprocess(golden_dir):
    """raw cv2-style images; the normalisation is bit-identical to the reference's numpy code."""
    from gpu_util import image, load_model, nerr
    from multiposenet.pytorch_b200 import ops
    from oracle.preprocess_oracle import resnet_preprocess
    g = np.load(os.path.join(golden_dir, "preprocess.npz"))
    for a, b in (("img", "out"), ("full", "out_full")):
        got = ops.resnet_preprocess_u8(torch.from_numpy(g[a])[None].cuda())[0].cpu().numpy()
        assert np.array_equal(got.view(np.uint32), g[b].view(np.uint32))
    rng = np.random.Generator(np.random.PCG64(5))
    u8 = rng.integers(0, 256, (2, 64, 96, 3), dtype=np.uint8)
    x = torch.from_numpy(np.stack([resnet_preprocess(im) for im in u8])).cuda()
    m, _ = load_model(50, "conditioned", "bf16x3")
    with torch.no_grad():
        h_ref, _ = m([x, "keypoint_subnet"])
        h_u8, _ = m([torch.from_numpy(u8).cuda(), "keypoint_subnet"])
        _, (cls_ref, _, _) = m([x, "detection_subnet"])
        _, (cls_u8, _, anc) = m([torch.from_numpy(u8).cuda(), "detection_subnet"])
    assert torch.equal(h_ref, h_u8) and torch.equal(cls_ref, cls_u8)   # same bits in, same bits out
    assert anc.shape[1] == ops.anchors_for(64, 96, anc.device).shape[1]


@pytest.mark.parametrize("precision", ["f16f8", "bf16x3"])
def test_keypoint_head_with_and_without_the_replicated_concat(precision, monkeypatch):
    """engine.CONV2_GATHER: the x8 / x4 quarters of conv2 as low-resolution phase-class convolutions gathered in its epilogue
    (default) against the replicated 512-channel concat of round 1 (MPN_CONV2_GATHER=0): same heat maps to the rounding of the
    regrouped sums."""
    from gpu_util import image, load_model, nerr
    from multiposenet.pytorch_b200 import engine as eng_mod
    m, _ = load_model(50, "conditioned", precision)
    x = image(5, (2, 3, 128, 192))
    eng = m.engine()
    monkeypatch.setattr(eng_mod, "CONV2_GATHER", True)
    h1 = eng.keypoint_forward(x)[0].clone()
    monkeypatch.setattr(eng_mod, "CONV2_GATHER", False)
    h0 = eng.keypoint_forward(x)[0].clone()
    torch.cuda.synchronize()
    assert h1.shape == h0.shape and torch.isfinite(h1).all()
    assert nerr(h1, h0) <= 2e-4
