"""Generate tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference):
    python -m oracle.make_goldens
Every array below is an output of the reference's own network/{posenet,fpn,anchors,utils}.py
executed by the container's torch (CPU fp32), with seeded inputs (numpy PCG64) and the seeded
weight sets of oracle/weights.py.  The entire_net ('both') branch needs a pth_nms; the
reference's cffi binary is unusable, so oracle/nms_oracle.c stands in (recorded in `meta`).
"""
import json
import os
import sys

import numpy as np
import torch

from . import refshim, weights
from .nms_oracle import nms_gpu_semantics, nms_cpu_semantics, nms_numpy_bruteforce

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def image(seed, shape):
    return np.random.Generator(np.random.PCG64(seed)).standard_normal(shape, dtype=np.float32)


def run_model(layers, kind, hw, batch, seed_img, full=True):
    w = weights.make_weights(layers, kind, seed=0)
    m = refshim.build_reference_model(layers, w)
    x = torch.from_numpy(image(seed_img, (batch, 3) + hw))
    out = {}
    with torch.no_grad():
        heat, saved = m([x, "keypoint_subnet"])
        out["kp_heat"] = heat.numpy()
        for i, s in enumerate(saved[:4]):
            out["kp_saved%d" % i] = s.numpy()
        _, (cls, reg, anc) = m([x, "detection_subnet"])
        out["det_cls"], out["det_reg"], out["det_anchors"] = cls.numpy(), reg.numpy(), anc.numpy()
        heat2, (sc, cl, bx) = m((x, "both"))
        out["both_heat"] = heat2.numpy()
        out["both_scores"], out["both_classes"], out["both_boxes"] = sc.numpy(), cl.numpy(), bx.numpy()
    if not full:  # 480x640: keep the file small -- strided samples of the big maps
        for k in ("kp_heat", "both_heat", "kp_saved0", "kp_saved1", "kp_saved2", "kp_saved3"):
            out[k] = np.ascontiguousarray(out[k][:, :, ::4, ::4])
        out["det_reg"] = np.ascontiguousarray(out["det_reg"][:, ::5])
        out["det_anchors"] = np.ascontiguousarray(out["det_anchors"][:, ::5])
    return out


def nms_cases():
    """Synthetic dets with the cfg3 recipe at small N plus edge cases."""
    cases = {}
    rng = np.random.Generator(np.random.PCG64(7))
    for n_gt, per in ((5, 8), (20, 13), (100, 41)):
        w = rng.uniform(20, 120, n_gt); h = rng.uniform(60, 300, n_gt)
        cx = rng.uniform(0, 640, n_gt); cy = rng.uniform(0, 480, n_gt)
        rows = []
        for g in range(n_gt):
            jw = w[g] * rng.uniform(0.9, 1.1, per); jh = h[g] * rng.uniform(0.9, 1.1, per)
            jx = cx[g] + w[g] * rng.uniform(-0.1, 0.1, per); jy = cy[g] + h[g] * rng.uniform(-0.1, 0.1, per)
            rows.append(np.stack([jx - jw / 2, jy - jh / 2, jx + jw / 2, jy + jh / 2], 1))
        b = np.concatenate(rows, 0)
        b[:, 0::2] = np.clip(b[:, 0::2], 0, 640); b[:, 1::2] = np.clip(b[:, 1::2], 0, 480)
        s = rng.permutation(np.linspace(0.0501, 0.9999, len(b)))
        cases["cfg3_%d" % len(b)] = np.concatenate([b, s[:, None]], 1).astype(np.float32)
    cases["single"] = np.array([[10, 10, 50, 60, 0.9]], np.float32)
    cases["identical"] = np.tile(np.array([[10, 10, 50, 60, 0.5]], np.float32), (70, 1))
    cases["identical"][:, 4] = np.linspace(0.9, 0.1, 70)
    cases["n65"] = cases["cfg3_260"][:65].copy()
    cases["n128"] = cases["cfg3_260"][:128].copy()
    return cases


PRN_CASES = [(0, {}), (1, dict(persons=8)), (2, dict(drop_joint=3)), (3, dict(persons=1, extra_boxes=0)),
             (4, dict(persons=12, noise_peaks=30)), (5, dict(persons=3, drop_joint=0, extra_boxes=2)),
             (6, dict(persons=0, extra_boxes=2, noise_peaks=0)), (7, dict(persons=4, hw=(200, 160), noise_peaks=40))]


def import_reference_tester():
    """evaluate/tester.py imports pycocotools and (through datasets/coco_data/prn_gaussian.py:2) skimage, neither installed
    here.  pycocotools is only used by Tester.val/coco_eval -> empty stub.  skimage.filters.gaussian(image) with the
    defaults prn_process uses is a one-line call of scipy.ndimage.gaussian_filter (skimage/filters/_gaussian.py of the
    pinned 0.13.1: image = img_as_float(image); return ndi.gaussian_filter(image, sigma, output=output, mode=mode,
    cval=cval, truncate=truncate)) -> shimmed with exactly that call."""
    import types
    import scipy.ndimage as ndi
    refshim.import_reference()

    def gaussian(image, sigma=1, output=None, mode="nearest", cval=0, multichannel=None, preserve_range=False, truncate=4.0):
        return ndi.gaussian_filter(np.asarray(image, dtype=np.float64), sigma, output=output, mode=mode, cval=cval, truncate=truncate)

    sk, skf = types.ModuleType("skimage"), types.ModuleType("skimage.filters")
    skf.gaussian, sk.filters = gaussian, skf
    sys.modules.setdefault("skimage", sk)
    sys.modules.setdefault("skimage.filters", skf)
    for n in ("pycocotools", "pycocotools.coco", "pycocotools.cocoeval"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["pycocotools.coco"].COCO = object
    sys.modules["pycocotools.cocoeval"].COCOeval = object
    sys.modules["lib"].__path__ = [os.path.join(refshim.REF_ROOT, "lib")]      # lib.utils.* from the checkout
    for k in [k for k in sys.modules if k == "evaluate" or k.startswith("evaluate.")]:
        if not (getattr(sys.modules[k], "__file__", "") or "").startswith(refshim.REF_ROOT):
            del sys.modules[k]
    import evaluate.tester as tester
    assert tester.__file__.startswith(refshim.REF_ROOT)
    return tester


def prn_goldens():
    """tests/golden/prn_assign.npz: Tester.prn_process of the reference (evaluate/tester.py:333-513) on seeded cases, with
    the PRN replaced by prn_oracle.synthetic_prn (so the records do not depend on 285 MB of weights)."""
    from . import prn_oracle
    tester = import_reference_tester()

    class Params(object):
        coeff, in_thres = 2, 0.21

    class Self(object):
        params = Params()

    g = {}
    for seed, kw in PRN_CASES:
        kps, boxes = prn_oracle.synthetic_case(seed, **kw)
        fn = prn_oracle.synthetic_prn(seed)
        me = Self()
        me.model = lambda args, fn=fn: (torch.from_numpy(fn(args[0].cpu().numpy())), None)
        rec = tester.Tester.prn_process(me, [list(k) for k in kps], [list(b) for b in boxes], "img%d" % seed, seed)
        g["case%d_keypoints" % seed] = np.array([r["keypoints"] for r in rec], dtype=np.float64).reshape(len(rec), 51)
        g["case%d_score" % seed] = np.array([r["score"] for r in rec], dtype=np.float64)
        g["case%d_bbox" % seed] = np.array([r["bbox"] for r in rec], dtype=np.float64).reshape(len(rec), 4)
    g["cases"] = np.array(json.dumps(PRN_CASES))
    g["versions"] = np.array(json.dumps({"numpy": np.__version__, "scipy": __import__("scipy").__version__}))
    np.savez_compressed(os.path.join(OUT, "prn_assign.npz"), **g)


TRAIN_GRAD_KEYS = ["fpn.conv1.weight", "fpn.bn1.weight", "fpn.bn1.bias", "fpn.layer1.0.conv1.weight", "fpn.layer1.0.bn3.weight",
                   "fpn.layer2.0.downsample.0.weight", "fpn.layer3.2.conv2.weight", "fpn.layer4.2.conv3.weight", "fpn.layer4.2.bn3.bias",
                   "fpn.toplayer.weight", "fpn.flatlayer2.weight", "fpn.smooth3.weight", "fpn.smooth3.bias", "convt1.weight", "convs4.bias",
                   "conv2.weight", "conv2.bias", "convfin.weight", "convfin.bias", "convfin_k3.weight", "convfin_k5.bias"]
TRAIN_STAT_KEYS = ["fpn.bn1.running_mean", "fpn.bn1.running_var", "fpn.layer3.1.bn2.running_mean", "fpn.layer4.2.bn3.running_var"]


def sample_flat(a, limit=16384):
    """Strided sample of a flattened array (keeps the golden file small); the same rule is applied by the tests."""
    f = np.ascontiguousarray(a).reshape(-1)
    return np.ascontiguousarray(f[:: max(1, -(-f.size // limit))])


def train_case(hw=(64, 96), batch=2):
    """The seeded training problem of tests/golden/train_step.npz (also rebuilt by the tests)."""
    x = image(41, (batch, 3) + hw)
    rng = np.random.Generator(np.random.PCG64(9))
    gt = rng.random((batch, 18, hw[0] // 4, hw[1] // 4), dtype=np.float32)
    wt = (rng.random((batch, 18, hw[0] // 4, hw[1] // 4)) > 0.2).astype(np.float32)
    return x, gt, wt


def train_golden():
    """tests/golden/train_step.npz: the reference's own training forward + loss + autograd backward for the keypoint subnet:
    model.train() (BatchNorm on batch statistics, trainer.py:170-174), model([img, 'keypoint_subnet']) (posenet.py:288-318),
    poseNet.build_loss -> build_keypoint_loss (posenet.py:352-403), loss.backward() (trainer.py:245-259)."""
    layers, hw, batch = 50, (64, 96), 2
    w = weights.make_weights(layers, "conditioned", seed=0)
    m = refshim.build_reference_model(layers, w)
    m.train()
    x, gt, wt = (torch.from_numpy(a) for a in train_case(hw, batch))
    with torch.enable_grad():
        out, saved = m([x, "keypoint_subnet"])
        loss, log = m.build_loss(saved, "keypoint_subnet", gt, wt)
        m.zero_grad()
        loss.backward()
    g = {"loss": np.array(float(loss), dtype=np.float64), "log_keys": np.array(json.dumps(list(log.keys()))),
         "log_vals": np.array([float(v) for v in log.values()], dtype=np.float64)}
    for i, s_ in enumerate(saved):
        g["saved%d" % i] = s_.detach().numpy()
    params = dict(m.named_parameters())
    for k in TRAIN_GRAD_KEYS:
        g["grad:" + k] = sample_flat(params[k].grad.numpy())
    bufs = dict(m.named_buffers())
    for k in TRAIN_STAT_KEYS:
        g["stat:" + k] = bufs[k].numpy().copy()
    g["meta"] = np.array(json.dumps({"layers": layers, "hw": hw, "batch": batch, "kind": "conditioned", "torch": torch.__version__,
                                     "problem": "oracle.make_goldens.train_case", "mode": "model.train(): BN batch statistics"}))
    np.savez_compressed(os.path.join(OUT, "train_step.npz"), **g)
    print("train_step", float(loss), {k: v.shape for k, v in g.items() if k.startswith("grad:")})


def prn_case(seed, persons, coeff):
    """Seeded PRN input: per person a [28*coeff, 18*coeff, 17] stack of sparse, blurred peaks (tester.py:393-404 shape)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    h, w_ = 28 * coeff, 18 * coeff
    x = np.zeros((persons, h, w_, 17), dtype=np.float32)
    for p in range(persons):
        for j in range(17):
            for _ in range(int(rng.integers(0, 3))):
                cy, cx = int(rng.integers(1, h - 1)), int(rng.integers(1, w_ - 1))
                x[p, cy - 1:cy + 2, cx - 1:cx + 2, j] += np.array([[.06, .12, .06], [.12, .25, .12], [.06, .12, .06]], np.float32) * float(rng.uniform(1, 4))
    return x


def prn_weights(node_count, coeff, seed=5):
    """Seeded PRN parameters (three Linear layers, posenet.py:130-141) in state_dict naming."""
    rng = np.random.Generator(np.random.PCG64(seed))
    d = 28 * coeff * 18 * coeff * 17
    out = {}
    for name, (o, i) in (("prn.dens1", (node_count, d)), ("prn.bneck", (node_count, node_count)), ("prn.dens2", (d, node_count))):
        out[name + ".weight"] = rng.standard_normal((o, i), dtype=np.float32) * np.float32(1.0 / np.sqrt(i))
        out[name + ".bias"] = rng.standard_normal((o,), dtype=np.float32) * np.float32(0.1)
    return out


def prn_forward_golden():
    """tests/golden/prn_forward.npz: the reference's own PRN (posenet.py:130-152 PRN.forward and :337-350 prn_forward) in eval
    mode, a small net (256 nodes, coeff 1) and the production shape (1024 nodes, coeff 2: 34272 -> 1024 -> 1024 -> 34272)."""
    ref = refshim.import_reference()
    g = {}
    for tag, nodes, coeff, persons in (("small", 256, 1, 5), ("prod", 1024, 2, 3)):
        pw = prn_weights(nodes, coeff)
        prn = ref.PRN(nodes, coeff)
        prn.load_state_dict({k[len("prn."):]: torch.from_numpy(v) for k, v in pw.items()})
        prn.eval()
        x = torch.from_numpy(prn_case(17, persons, coeff))
        with torch.no_grad():
            out = prn(x)
            model = ref.poseNet(50, prn_node_count=nodes, prn_coeff=coeff)
            model.prn.load_state_dict(prn.state_dict())
            model.eval()
            out2, saved = model([x, "prn_subnet"])
        assert torch.equal(out, out2) and saved[0] is out2
        g[tag + "_out"] = out.numpy()
        g[tag + "_meta"] = np.array(json.dumps({"nodes": nodes, "coeff": coeff, "persons": persons, "input": "prn_case(17, persons, coeff)",
                                                "weights": "prn_weights(nodes, coeff, seed=5)"}))
    np.savez_compressed(os.path.join(OUT, "prn_forward.npz"), **g)
    print("prn_forward", {k: v.shape for k, v in g.items()})


def tta_golden():
    """tests/golden/tta.npz: Tester._get_multiplier / _get_outputs / _handle_heat of the reference (evaluate/tester.py:256-331)
    run here on a seeded image with tta_oracle.stub_model in place of the network."""
    from . import tta_oracle
    tester = import_reference_tester()

    class Params(object):
        inp_size, gpus, subnet_name = 96, [0], "both"

    class Self(object):
        params = Params()

    me = Self()

    def model(args):
        heat, scores, classes, boxes = tta_oracle.stub_model(args[0].cpu().numpy())
        return torch.from_numpy(heat), [torch.from_numpy(scores), torch.from_numpy(classes), torch.from_numpy(boxes)]
    me.model = model
    img = tta_oracle.test_image()
    mult = tester.Tester._get_multiplier(me, img)
    heat_n, bbox_n = tester.Tester._get_outputs(me, mult, img)
    heat_f, bbox_f = tester.Tester._get_outputs(me, mult, img[:, ::-1, :])
    avg = tester.Tester._handle_heat(me, heat_n, heat_f)
    crop, sc, shp = tester.crop_with_factor(img, 150.0, factor=32, pad_val=128)
    assert heat_n.dtype == np.float64 and avg.dtype == np.float64
    # stored as float32 (the comparisons carry a 1e-6 tolerance: cv2's cubic resize is not bit-reproducible call to call)
    g = {"multiplier": np.array(mult), "heat_normal": heat_n.astype(np.float32), "heat_flipped": heat_f.astype(np.float32),
         "heat_avg": avg.astype(np.float32),
         "bbox_normal": np.array(json.dumps(bbox_n)), "crop": crop, "crop_scale": np.array(sc), "crop_shape": np.array(shp),
         "meta": np.array(json.dumps({"inp_size": 96, "image": "tta_oracle.test_image()", "model": "tta_oracle.stub_model",
                                      "cv2": __import__("cv2").__version__}))}
    np.savez_compressed(os.path.join(OUT, "tta.npz"), **g)
    print("tta", mult, heat_n.shape, [len(b) for b in bbox_n])


def focal_golden():
    """tests/golden/focal_loss.npz: the reference's own FocalLoss (network/losses.py:27-137) and its autograd gradients on
    losses_oracle.focal_case()."""
    from . import losses_oracle as lo
    refshim.import_reference()
    from network.losses import FocalLoss
    cls, reg, anchors, ann = (torch.from_numpy(a) for a in lo.focal_case(with_empty=False))
    cls.requires_grad_(True)
    reg.requires_grad_(True)
    # torch >= 1.2 compatibility shim (third shim beside refshim's two): losses.py:120 computes `1 - positive_indices` on what
    # torch 0.4 returned as a ByteTensor and torch 2.x returns as a bool tensor, for which `-` raises.  The value is never used
    # (negative_indices is dead code), so the shim only lets the line execute: bool operands of __rsub__ are viewed as uint8.
    real_rsub = torch.Tensor.__rsub__

    def rsub(self, other):
        return real_rsub(self.to(torch.uint8) if self.dtype == torch.bool else self, other)
    torch.Tensor.__rsub__ = rsub
    try:
        with torch.enable_grad():
            cl, rl = FocalLoss()(cls, reg, anchors, ann)
            total = cl.mean() + rl.mean()                                       # posenet.py:413-416
            total.backward()
    finally:
        torch.Tensor.__rsub__ = real_rsub
    g = {"cls_loss": cl.detach().numpy(), "reg_loss": rl.detach().numpy(), "dcls": cls.grad.numpy(), "dreg": reg.grad.numpy(),
         "meta": np.array(json.dumps({"case": "losses_oracle.focal_case(with_empty=False)", "torch": torch.__version__}))}
    np.savez_compressed(os.path.join(OUT, "focal_loss.npz"), **g)
    print("focal", float(cl), float(rl), int((cls.grad != 0).sum()), int((reg.grad != 0).sum()))


def main():
    os.makedirs(OUT, exist_ok=True)
    meta = {"torch": torch.__version__, "numpy": np.__version__, "reference": refshim.REF_ROOT,
            "nms_in_both_branch": "oracle/nms_oracle.c (reference cffi binary unusable)",
            "weights": "oracle.weights.make_weights(layers, kind, seed=0)", "image": "PCG64(seed).standard_normal"}
    ref = refshim.import_reference()
    # anchors straight from the reference module
    from network.anchors import Anchors
    anc = {}
    for hw in ((480, 640), (64, 96), (100, 130), (32, 32)):
        anc["%dx%d" % hw] = Anchors()(torch.zeros(1, 3, *hw)).numpy()[0]
    np.savez_compressed(os.path.join(OUT, "anchors.npz"), **anc)
    # decode + clip straight from the reference modules
    from network.utils import BBoxTransform, ClipBoxes
    rng = np.random.Generator(np.random.PCG64(3))
    a = torch.from_numpy(anc["64x96"])[None]
    d = torch.from_numpy(rng.standard_normal((2, a.shape[1], 4), dtype=np.float32) * 2)
    boxes = ClipBoxes()(BBoxTransform()(a, d), torch.zeros(2, 3, 64, 96))
    np.savez_compressed(os.path.join(OUT, "decode.npz"), anchors=a.numpy(), deltas=d.numpy(), boxes=boxes.numpy())
    # image normalisation straight from the reference function (datasets/coco_data/preprocessing.py:15-26)
    from datasets.coco_data.preprocessing import resnet_preprocess
    prng = np.random.Generator(np.random.PCG64(21))
    pimg = prng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    pfull = np.arange(256, dtype=np.uint8).repeat(3).reshape(16, 16, 3)  # every byte value in every channel
    np.savez_compressed(os.path.join(OUT, "preprocess.npz"), img=pimg, out=resnet_preprocess(pimg), full=pfull,
                        out_full=resnet_preprocess(pfull))
    # network forwards
    runs = [("r50_cond_64x96_b2", 50, "conditioned", (64, 96), 2, 11, True),
            ("r50_refinit_64x96_b1", 50, "refinit", (64, 96), 1, 12, True),
            ("r101_cond_64x96_b1", 101, "conditioned", (64, 96), 1, 13, True),
            ("r50_cond_480x640_b1", 50, "conditioned", (480, 640), 1, 14, False)]
    for name, layers, kind, hw, b, seed, full in runs:
        out = run_model(layers, kind, hw, b, seed, full)
        out["meta"] = np.array(json.dumps(dict(meta, layers=layers, kind=kind, hw=hw, batch=b, img_seed=seed, full=full)))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items() if k != "meta"}, "K=%d" % len(out["both_scores"]))
    # NMS cases: keep lists from the C restatement (cross-checked against brute-force numpy)
    nm = {}
    for k, dets in nms_cases().items():
        for thr in (0.5, 0.3):
            kg = nms_gpu_semantics(dets, thr)
            assert np.array_equal(kg, nms_numpy_bruteforce(dets, thr)), k
            kc = nms_cpu_semantics(dets, thr)
            assert np.array_equal(kc, nms_numpy_bruteforce(dets, thr, ge=True)), k
            nm["%s_dets" % k] = dets
            nm["%s_keep_gt_%g" % (k, thr)] = kg
            nm["%s_keep_ge_%g" % (k, thr)] = kc
    np.savez_compressed(os.path.join(OUT, "nms.npz"), **nm)
    # heat-map peaks: the reference's own get_joint_list (scipy maximum_filter + cv2.resize INTER_CUBIC of this container)
    from network.joint_utils import get_joint_list
    from . import peaks_oracle
    pk = {"cv2": np.array(__import__("cv2").__version__)}
    small = peaks_oracle.synthetic_heatmaps(5, C=18, H=24, W=32, persons=2)
    pk["small_heat"] = small
    pk["small_rows"] = get_joint_list(np.zeros((96, 128, 3), np.float32), {"thre1": 0.1}, np.ascontiguousarray(small.transpose(1, 2, 0)), 1.0)
    pk["small_rows_f2"] = get_joint_list(np.zeros((48, 64, 3), np.float32), {"thre1": 0.1}, np.ascontiguousarray(small.transpose(1, 2, 0)), 1.0)
    for seed in (1, 2):  # full-size maps are regenerated from the seed (peaks_oracle.synthetic_heatmaps); checksum pins them
        hm = peaks_oracle.synthetic_heatmaps(seed)
        pk["seed%d_sum" % seed] = np.array(hm.astype(np.float64).sum())
        pk["seed%d_rows" % seed] = get_joint_list(np.zeros((480, 640, 3), np.float32), {"thre1": 0.1},
                                                  np.ascontiguousarray(hm.transpose(1, 2, 0)), 1.0)
    np.savez_compressed(os.path.join(OUT, "peaks.npz"), **pk)
    prn_goldens()
    train_golden()
    prn_forward_golden()
    tta_golden()
    focal_golden()
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    if "prn" in sys.argv[1:]:   # only the PRN-assignment vectors
        sys.exit(prn_goldens())
    if "train" in sys.argv[1:]:   # only the training-step vectors
        sys.exit(train_golden())
    if "prn_forward" in sys.argv[1:]:
        sys.exit(prn_forward_golden())
    if "tta" in sys.argv[1:]:
        sys.exit(tta_golden())
    if "focal" in sys.argv[1:]:
        sys.exit(focal_golden())
    sys.exit(main())
