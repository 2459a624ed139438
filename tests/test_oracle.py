"""CPU tests: the oracle restatement against the committed golden vectors (made from the reference)
and, where /root/reference is present, against the live reference modules."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import anchors_oracle, nms_oracle, posenet_oracle as po, refshim, weights



@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_anchors_bit_exact(golden_dir):
    g = _load(golden_dir, "anchors.npz")
    for key in g.files:
        h, w = map(int, key.split("x"))
        a = anchors_oracle.anchors_for_image(h, w)
        assert a.dtype == np.float32 and a.shape == g[key].shape
        assert np.array_equal(a.view(np.uint32), g[key].view(np.uint32)), key
    assert anchors_oracle.anchors_for_image(480, 640).shape == (57600, 4)


def test_decode_clip(golden_dir):
    g = _load(golden_dir, "decode.npz")
    b = po.clip_boxes(po.decode_boxes(torch.from_numpy(g["anchors"]), torch.from_numpy(g["deltas"])), 64, 96)
    assert np.array_equal(b.numpy(), g["boxes"])


def test_nms_goldens_and_bruteforce(golden_dir):
    g = _load(golden_dir, "nms.npz")
    names = sorted({k[: -len("_dets")] for k in g.files if k.endswith("_dets")})
    assert names
    for n in names:
        dets = g[n + "_dets"]
        for thr in (0.5, 0.3):
            kg = nms_oracle.nms_gpu_semantics(dets, thr)
            kc = nms_oracle.nms_cpu_semantics(dets, thr)
            assert np.array_equal(kg, g["%s_keep_gt_%g" % (n, thr)])
            assert np.array_equal(kc, g["%s_keep_ge_%g" % (n, thr)])
            assert np.array_equal(kg, nms_oracle.nms_numpy_bruteforce(dets, thr))


def test_nms_known_answers():
    # IoU with the +1 convention: boxes [0,0,9,9] and [0,0,9,19]: inter 100, union 200 -> exactly 0.5
    d = np.array([[0, 0, 9, 9, 0.9], [0, 0, 9, 19, 0.8], [100, 100, 120, 120, 0.7]], np.float32)
    assert nms_oracle.nms_gpu_semantics(d, 0.5).tolist() == [0, 1, 2]      # '>' keeps the tie
    assert nms_oracle.nms_cpu_semantics(d, 0.5).tolist() == [0, 2]         # '>=' suppresses it
    assert nms_oracle.nms_gpu_semantics(d, 0.5, ge=True).tolist() == [0, 2]
    assert nms_oracle.nms_gpu_semantics(np.zeros((0, 5), np.float32), 0.5).shape == (0,)
    # order: descending score, indices refer to the input rows
    d2 = d[[2, 0, 1]]
    assert nms_oracle.nms_gpu_semantics(d2, 0.5).tolist() == [1, 2, 0]
    # mask/reduce decomposition equals the fused entry point
    rng = np.random.default_rng(0)
    xy = rng.uniform(0, 200, (300, 2)); wh = rng.uniform(5, 80, (300, 2))
    dets = np.concatenate([xy, xy + wh, rng.permutation(300)[:, None] / 300.0], 1).astype(np.float32)
    order = nms_oracle.stable_desc_order(dets[:, 4])
    m = nms_oracle.nms_mask(dets[order], 0.5)
    assert np.array_equal(order[nms_oracle.reduce_mask(m, 300)], nms_oracle.nms_gpu_semantics(dets, 0.5))


@pytest.mark.parametrize("name", ["r50_cond_64x96_b2", "r50_refinit_64x96_b1", "r101_cond_64x96_b1"])
def test_network_restatement_vs_golden(golden_dir, name):
    g = _load(golden_dir, name + ".npz")
    meta = json.loads(str(g["meta"]))
    sd = weights.to_torch_state_dict(weights.make_weights(meta["layers"], meta["kind"], seed=0))
    x = torch.from_numpy(np.random.Generator(np.random.PCG64(meta["img_seed"])).standard_normal(
        (meta["batch"], 3) + tuple(meta["hw"]), dtype=np.float32))
    heat, saved = po.forward(sd, meta["layers"], x, "keypoint_subnet")
    # same torch build, same op sequence -> identical; a different oneDNN path may reorder sums
    tol = dict(rtol=1e-4, atol=1e-5 * float(np.abs(g["kp_heat"]).max()))
    np.testing.assert_allclose(heat.numpy(), g["kp_heat"], **tol)
    for i in range(4):
        np.testing.assert_allclose(saved[i].numpy(), g["kp_saved%d" % i], rtol=1e-4,
                                   atol=1e-5 * float(np.abs(g["kp_saved%d" % i]).max()))
    _, (cls, reg, anc) = po.forward(sd, meta["layers"], x, "detection_subnet")
    np.testing.assert_allclose(cls.numpy(), g["det_cls"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(reg.numpy(), g["det_reg"], rtol=1e-4, atol=1e-5 * float(np.abs(g["det_reg"]).max()))
    assert np.array_equal(anc.numpy(), g["det_anchors"])
    heat2, (sc, cl, bx), _ = po.forward(sd, meta["layers"], x, "both")
    np.testing.assert_allclose(heat2.numpy(), g["both_heat"], **tol)
    assert sc.shape == g["both_scores"].shape
    if len(sc):
        np.testing.assert_allclose(sc.numpy(), g["both_scores"], rtol=1e-4)
        np.testing.assert_allclose(bx.numpy(), g["both_boxes"], rtol=1e-4, atol=1e-3)
        assert np.array_equal(cl.numpy(), g["both_classes"])


def test_param_spec_counts():
    assert len(weights.param_spec(50)) == 402 and len(weights.param_spec(101)) == 708
    assert abs(po.conv_flops_entire(101) / 1e9 - 270.12) < 0.01
    assert abs(po.conv_flops_entire(50) / 1e9 - 224.67) < 0.01


@pytest.mark.skipif(not refshim.available(), reason="/root/reference not present")
def test_live_reference_keys_and_forward():
    w = weights.make_weights(50, "conditioned", seed=0)
    m = refshim.build_reference_model(50, w)
    spec = weights.param_spec(50)
    assert list(m.state_dict().keys()) == list(spec.keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(spec[k]), k
    x = torch.from_numpy(np.random.Generator(np.random.PCG64(5)).standard_normal((1, 3, 64, 64), dtype=np.float32))
    sd = weights.to_torch_state_dict(w)
    heat, saved = m([x, "keypoint_subnet"])
    oh, osv = po.forward(sd, 50, x, "keypoint_subnet")
    assert torch.equal(heat, oh) and all(torch.equal(a, b) for a, b in zip(saved, osv))
    h2, (s, c, b) = m((x, "both"))
    oh2, (os_, oc, ob), _ = po.forward(sd, 50, x, "both")
    assert torch.equal(h2, oh2) and torch.equal(s, os_) and torch.equal(b, ob)


def test_resnet_preprocess_bit_exact(golden_dir):
    from oracle.preprocess_oracle import resnet_preprocess
    g = _load(golden_dir, "preprocess.npz")
    for a, b in (("img", "out"), ("full", "out_full")):
        assert np.array_equal(resnet_preprocess(g[a]).view(np.uint32), g[b].view(np.uint32))


# ---------------------------------------------------------------------------------------------
# heat-map peak extraction (joint_utils.py:19-32, 61-152)
def _check_rows(mine, gold, score_tol=1e-6):
    assert mine.shape == gold.shape
    assert np.array_equal(mine[:, [0, 1, 3, 4]], gold[:, [0, 1, 3, 4]])   # x, y, id, joint type: exact
    assert np.abs(mine[:, 2] - gold[:, 2]).max() <= score_tol            # cv2's SIMD build vs unfused float32


def test_peaks_oracle_vs_reference_goldens(golden_dir):
    from oracle import peaks_oracle
    g = _load(golden_dir, "peaks.npz")
    _check_rows(peaks_oracle.joint_list(g["small_heat"], 0.1, 4), g["small_rows"])
    _check_rows(peaks_oracle.joint_list(g["small_heat"], 0.1, 2), g["small_rows_f2"])
    hm = peaks_oracle.synthetic_heatmaps(1)
    assert abs(float(hm.astype(np.float64).sum()) - float(g["seed1_sum"])) < 1e-3   # the seeded maps are the golden's maps
    _check_rows(peaks_oracle.joint_list(hm, 0.1, 4), g["seed1_rows"])


def test_peaks_oracle_edge_cases():
    from oracle import peaks_oracle
    H, W = 24, 32
    hm = peaks_oracle.synthetic_heatmaps(5, C=6, H=H, W=W, persons=1)
    as_set = lambda a: {tuple(int(v) for v in r) for r in a}
    assert (0, 0) in as_set(peaks_oracle.find_peaks(hm[0], 0.1))                                 # corner
    assert (7, H - 1) in as_set(peaks_oracle.find_peaks(hm[1], 0.1))                             # bottom edge
    assert {(W // 3, H // 3), (W // 3 + 1, H // 3)} <= as_set(peaks_oracle.find_peaks(hm[2], 0.1))  # both plateau pixels
    assert (W - 1, H // 2) in as_set(peaks_oracle.find_peaks(hm[3], 0.1))                        # right edge
    assert (10, 10) not in as_set(peaks_oracle.find_peaks(hm[4], 0.1))                           # == thre1 is not a peak
    rows = peaks_oracle.joint_list(hm, 0.1, 4)
    assert np.array_equal(rows[:, 3], np.arange(len(rows)))                                      # ids count up over joint types
    assert (np.diff(rows[:, 4]) >= 0).all()                                                      # grouped by joint type
    assert len(peaks_oracle.joint_list(np.zeros((3, 8, 8), np.float32), 0.1, 4)) == 0            # empty maps


def test_peaks_oracle_vs_live_reference():
    pytest.importorskip("cv2")
    from oracle import peaks_oracle, refshim
    if not os.path.isdir(refshim.REF_ROOT):
        pytest.skip("reference checkout not present (GPU box)")
    refshim.import_reference()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from network.joint_utils import get_joint_list
    hm = peaks_oracle.synthetic_heatmaps(9, C=18, H=40, W=56, persons=3)
    gold = get_joint_list(np.zeros((160, 224, 3), np.float32), {"thre1": 0.1}, np.ascontiguousarray(hm.transpose(1, 2, 0)), 1.0)
    _check_rows(peaks_oracle.joint_list(hm, 0.1, 4), gold)


# ---------------------------------------------------------------------------------------------
# train-mode forward + loss + autograd backward and the PRN MLP (a16): the restatement against vectors made
# by the live reference (oracle/make_goldens.py train_golden / prn_forward_golden)
def test_train_mode_oracle_vs_reference_golden(golden_dir):
    from oracle import make_goldens as mg
    g = _load(golden_dir, "train_step.npz")
    meta = json.loads(str(g["meta"]))
    sd = weights.to_torch_state_dict(weights.make_weights(meta["layers"], meta["kind"], seed=0))
    for k, v in sd.items():
        if v.dtype == torch.float32 and not k.endswith(("running_mean", "running_var")):
            v.requires_grad_(True)
    x, gt, wt = (torch.from_numpy(a) for a in mg.train_case(tuple(meta["hw"]), meta["batch"]))
    with torch.enable_grad():
        saved = po.forward_train_keypoint(sd, meta["layers"], x)
        loss = po.keypoint_loss(saved, gt, wt)
        loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    for i, s in enumerate(saved):
        np.testing.assert_allclose(s.detach().numpy(), g["saved%d" % i], rtol=1e-4, atol=1e-5 * float(np.abs(g["saved%d" % i]).max()))
    for k in mg.TRAIN_GRAD_KEYS:
        want = g["grad:" + k]
        got = mg.sample_flat(sd[k].grad.numpy())
        assert got.shape == want.shape, k
        assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max(), k   # same graph, same torch: only summation order may differ


def test_prn_forward_oracle_vs_reference_golden(golden_dir):
    from oracle import make_goldens as mg
    g = _load(golden_dir, "prn_forward.npz")
    for tag in ("small", "prod"):
        meta = json.loads(str(g[tag + "_meta"]))
        sd = {k: torch.from_numpy(v) for k, v in mg.prn_weights(meta["nodes"], meta["coeff"]).items()}
        x = torch.from_numpy(mg.prn_case(17, meta["persons"], meta["coeff"]))
        out, saved = po.prn_forward(sd, x)
        assert out.shape == g[tag + "_out"].shape and saved[0] is out
        np.testing.assert_allclose(out.numpy(), g[tag + "_out"], rtol=2e-5, atol=1e-9)
        assert abs(float(out.sum()) - meta["persons"]) < 1e-3


def test_tta_oracle_vs_reference_golden(golden_dir):
    """f3 pinned: Tester._get_multiplier / crop_with_factor / _get_outputs / _handle_heat of the live reference with a stub model."""
    pytest.importorskip("cv2")
    from oracle import tta_oracle as to
    g = _load(golden_dir, "tta.npz")
    meta = json.loads(str(g["meta"]))
    img = to.test_image()
    mult = to.get_multiplier(img, meta["inp_size"])
    assert np.array_equal(np.array(mult), g["multiplier"])
    crop, sc, shp = to.crop_with_factor(img, 150.0, factor=32, pad_val=128)
    assert np.array_equal(crop, g["crop"]) and sc == float(g["crop_scale"]) and tuple(shp) == tuple(g["crop_shape"])
    hn, bn = to.get_outputs(to.stub_model, mult, img)
    hf, _ = to.get_outputs(to.stub_model, mult, img[:, ::-1, :])
    # cv2's cubic resize is not bit-reproducible between two calls on the same values (vector body vs scalar tail depends on
    # the buffer alignment, and the two differ in FMA use): ulp-level tolerance, like the device path
    tol = 1e-6 * np.abs(g["heat_normal"]).max()
    assert np.abs(hn - g["heat_normal"]).max() <= tol and np.abs(hf - g["heat_flipped"]).max() <= tol
    assert np.abs(to.handle_heat(hn, hf) - g["heat_avg"]).max() <= tol
    assert hn.dtype == np.float64
    assert bn == json.loads(str(g["bbox_normal"]))


def test_focal_loss_oracle_vs_reference_golden(golden_dir):
    """f4 pinned: the reference's own FocalLoss.forward values and autograd gradients (tests/golden/focal_loss.npz)."""
    from oracle import losses_oracle as lo
    g = _load(golden_dir, "focal_loss.npz")
    cls, reg, anchors, ann = (torch.from_numpy(a) for a in lo.focal_case(with_empty=False))
    cls.requires_grad_(True)
    reg.requires_grad_(True)
    with torch.enable_grad():
        cl, rl, _, _ = lo.focal_loss(cls, reg, anchors, ann)
        (cl.mean() + rl.mean()).backward()
    np.testing.assert_allclose(cl.detach().numpy(), g["cls_loss"], rtol=1e-6)
    np.testing.assert_allclose(rl.detach().numpy(), g["reg_loss"], rtol=1e-6)
    np.testing.assert_allclose(cls.grad.numpy(), g["dcls"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(reg.grad.numpy(), g["dreg"], rtol=1e-5, atol=1e-9)
    # an image without annotations contributes zero to both losses (losses.py:53-57) and gets zero gradients
    cls2, reg2, anchors2, ann2 = (torch.from_numpy(a) for a in lo.focal_case(with_empty=True))
    cls2.requires_grad_(True)
    with torch.enable_grad():
        cl2, rl2, per_c, per_r = lo.focal_loss(cls2, reg2, anchors2, ann2)
        (cl2.mean() + rl2.mean()).backward()
    assert float(per_c[1]) == 0.0 and float(per_r[1]) == 0.0 and float(cls2.grad[1].abs().max()) == 0.0
