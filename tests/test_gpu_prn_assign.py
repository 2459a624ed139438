"""GPU parity of the PRN assignment (mpn_prn_build_inputs / mpn_prn_assign, evaluate/tester.py:333-513) against the numpy
restatement and the reference's own records: grid ownership and fp32 inputs bit-exact, assigned keypoints (float64) bit
for bit, batches of images equal to image-by-image calls."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prn_assign.npz")


class StubPRN(torch.nn.Module):
    """Stands where poseNet stands in prn_process: forward([inp, 'prn_subnet']) -> (out, [out]); the PRN itself is the
    oracle's seeded stand-in evaluated on the host (this test is about everything around the MLP)."""

    def __init__(self, seeds_by_rows):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))
        self.seeds_by_rows = seeds_by_rows          # one seed per person row (images may use different stand-ins)
        self.seen = None

    def forward(self, args):
        from oracle import prn_oracle as po
        inp, name = args
        assert name == "prn_subnet" and inp.is_cuda
        x = inp.cpu().numpy()
        self.seen = x
        out = np.concatenate([po.synthetic_prn(s)(x[i:i + 1]) for i, s in enumerate(self.seeds_by_rows)])
        o = torch.from_numpy(out).to(inp.device)
        return o, [o]


def _cases():
    g = np.load(GOLD)
    out = []
    for seed, kw in json.loads(str(g["cases"])):
        if "hw" in kw:
            kw["hw"] = tuple(kw["hw"])
        out.append((int(seed), kw))
    return g, out


def test_prn_process_vs_reference_goldens():
    from multiposenet.pytorch_b200.evaluate import prn_process
    from oracle import prn_oracle as po
    g, cases = _cases()
    for seed, kw in cases:
        kps, boxes = po.synthetic_case(seed, **kw)
        m = StubPRN([seed] * len(boxes)).cuda()
        rec = prn_process(m, kps, boxes, "img%d" % seed, seed)
        kp = np.array([r["keypoints"] for r in rec]).reshape(len(rec), 51)
        assert np.array_equal(kp, g["case%d_keypoints" % seed]), seed
        assert np.array_equal(np.array([r["score"] for r in rec]), g["case%d_score" % seed])
        assert np.array_equal(np.array([r["bbox"] for r in rec]).reshape(len(rec), 4), g["case%d_bbox" % seed])
        assert all(r["image_id"] == seed and r["file_name"] == "img%d" % seed and r["category_id"] == 1 for r in rec)
        # the PRN inputs the device built == the oracle's (scatter + scipy-order float64 gaussian -> fp32), bit for bit
        xy, ty = po.sort_peaks(kps)
        want = po.build_inputs(po.scatter(xy, ty, po.boxes_xywh(boxes)))
        assert np.array_equal(m.seen.view(np.uint32), want.view(np.uint32)), seed


def test_prn_process_batch_equals_per_image():
    from multiposenet.pytorch_b200.evaluate import prn_process_batch
    from oracle import prn_oracle as po
    specs = [(11, dict(persons=6)), (12, dict(persons=0, extra_boxes=0, noise_peaks=5)), (13, dict(persons=9, noise_peaks=20)),
             (14, dict(persons=2, drop_joint=5)), (15, dict(persons=0, extra_boxes=3, noise_peaks=0))]
    data = [po.synthetic_case(s, **kw) for s, kw in specs]
    seeds = [s for (s, _), (_, bx) in zip(specs, data) for _ in bx]
    m = StubPRN(seeds).cuda()
    got = prn_process_batch(m, [d[0] for d in data], [d[1] for d in data], ["f%d" % s for s, _ in specs], [s for s, _ in specs])
    assert len(got) == len(specs) and got[1] == []                                    # an image without boxes -> []
    for (s, _), (kps, boxes), rec in zip(specs, data, got):
        want = po.prn_process(kps, boxes, po.synthetic_prn(s), "f%d" % s, s)
        assert len(rec) == len(want)
        for a, b in zip(rec, want):
            assert a["keypoints"] == b["keypoints"] and a["score"] == b["score"] and a["bbox"] == b["bbox"] and a["image_id"] == b["image_id"]


def test_prn_scatter_quirks_and_window_sums():
    """Index wrap of the elif chain, last-writer-wins cells, and the numpy-order fp32 window sum on random PRN outputs."""
    from multiposenet.pytorch_b200 import ops
    from multiposenet.pytorch_b200.evaluate import prn_assign as pa
    from oracle import prn_oracle as po
    boxes = po.boxes_xywh([[100.0, 100.0, 140.0, 200.0], [90.0, 60.0, 200.0, 330.0]])
    xy = np.array([[147.0, 85.0], [95.0, 215.0], [95.0, 85.0], [120.0, 150.0], [120.5, 150.4], [130.0, 140.0], [101.0, 299.0]])
    ty = np.array([0, 1, 2, 3, 3, 3, 16], dtype=np.int32)
    dev = torch.device("cuda")
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(dev)
    pstart, bstart = t(np.array([0, len(ty)]), torch.int32), t(np.array([0, 2]), torch.int32)
    jstart = t(np.searchsorted(ty, np.arange(18))[None], torch.int32)
    ws = ops.prn_workspace(2, len(ty), len(ty), dev)
    owner, inp = ops.prn_build_inputs(t(xy, torch.float64), t(ty, torch.int32), pstart, t(boxes, torch.float64), t(np.zeros(2), torch.int32),
                                      (56, 36), 0.21, pa._gaussian_weights(), len(ty), ws)
    want_owner = po.scatter(xy, ty, boxes)
    assert np.array_equal(owner.cpu().numpy(), want_owner)
    assert np.array_equal(inp.cpu().numpy().view(np.uint32), po.build_inputs(want_owner).view(np.uint32))
    rng = np.random.default_rng(5)
    out = rng.random((2, 56, 36, 17), dtype=np.float32) ** 8                           # wide dynamic range: rounding order matters
    kp = ops.prn_assign(t(xy, torch.float64), pstart, jstart, t(boxes, torch.float64), t(np.zeros(2), torch.int32), bstart, owner,
                        torch.from_numpy(out).to(dev), len(ty), ws).cpu().numpy()
    assert np.array_equal(kp, po.assign(xy, ty, boxes, want_owner, out))


@pytest.mark.parametrize("precision", ["bf16x3", "f16f8"])
def test_prn_process_with_the_real_prn(precision):
    """End to end through poseNet's batched tcgen05 PRN: records equal the oracle's when it is fed the same PRN outputs."""
    from multiposenet.pytorch_b200 import poseNet
    from multiposenet.pytorch_b200.evaluate import prn_process
    from oracle import prn_oracle as po
    torch.manual_seed(0)
    m = poseNet(50, prn_node_count=128, prn_coeff=2, precision=precision).cuda().eval()
    kps, boxes = po.synthetic_case(21, persons=5)
    rec = prn_process(m, kps, boxes, "x", 3)

    def prn_fn(inp):
        with torch.no_grad():
            return m([torch.from_numpy(inp).cuda(), "prn_subnet"])[0].float().cpu().numpy()

    want = po.prn_process(kps, boxes, prn_fn, "x", 3)
    assert len(rec) == len(want) == len(boxes)
    for a, b in zip(rec, want):
        assert a["keypoints"] == b["keypoints"] and a["score"] == b["score"]


def test_prn_process_errors():
    from multiposenet.pytorch_b200.evaluate import prn_process
    m = StubPRN([0])
    with pytest.raises(RuntimeError):
        prn_process(m, [[1.0, 1.0, 1.0, 0.0, 0.0]], [[0.0, 0.0, 10.0, 10.0]], "f")     # CPU model: no fallback
    with pytest.raises(ZeroDivisionError):
        prn_process(m.cuda(), [], [[5.0, 5.0, 5.0, 9.0]], "f")                          # zero-width box (tester.py:371)
    assert prn_process(m.cuda(), [[1.0, 1.0, 1.0, 0.0, 0.0]], [], "f") == []
