"""CPU checks of the PRN-assignment restatement (oracle/prn_oracle.py, evaluate/tester.py:333-513): against the reference's
own Tester.prn_process records (tests/golden/prn_assign.npz), against the container's scipy / numpy for the two pieces of
third-party arithmetic it restates, and of the product's host-side regrouping against it."""
import json
import os

import numpy as np
import pytest

from oracle import prn_oracle as po


def _cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "prn_assign.npz"))
    return g, [(int(s), kw) for s, kw in json.loads(str(g["cases"]))]


def test_prn_oracle_vs_reference_goldens(golden_dir):
    g, cases = _cases(golden_dir)
    assigned = fallback = 0
    for seed, kw in cases:
        if "hw" in kw:
            kw["hw"] = tuple(kw["hw"])
        kps, boxes = po.synthetic_case(seed, **kw)
        rec = po.prn_process(kps, boxes, po.synthetic_prn(seed), "img%d" % seed, seed)
        kp = np.array([r["keypoints"] for r in rec]).reshape(len(rec), 51)
        assert np.array_equal(kp, g["case%d_keypoints" % seed])                       # float64, bit for bit
        assert np.array_equal(np.array([r["score"] for r in rec]), g["case%d_score" % seed])
        assert np.array_equal(np.array([r["bbox"] for r in rec]).reshape(len(rec), 4), g["case%d_bbox" % seed])
        assigned += int((kp[:, 2::3] > 0).sum())
        fallback += int(((kp[:, 2::3] == 0) & (kp[:, 0::3] != 0)).sum())
    assert assigned > 300 and fallback > 50                                           # both branches are exercised


def test_gaussian_restatement_vs_scipy():
    ndi = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(0)
    for _ in range(10):
        a = (rng.random((56, 36)) < 0.03).astype(np.float64)
        assert np.array_equal(po.gaussian_nearest(a), ndi.gaussian_filter(a, 1, mode="nearest", truncate=4.0))
    a = rng.random((28, 18))
    assert np.array_equal(po.gaussian_nearest(a), ndi.gaussian_filter(a, 1, mode="nearest", truncate=4.0))


def test_window_sum_restatement_vs_numpy():
    rng = np.random.default_rng(1)
    out = rng.random((56, 36, 17), dtype=np.float32)
    for gh, gw in ((56, 36), (28, 18)):
        pl = out[:gh, :gw, 5]
        for y in range(gh):
            for x in range(gw):
                r0, r1, c0, c1 = po.window_bounds(y, x, gh, gw)
                assert po.window_sum_f32(pl, r0, r1, c0, c1) == np.sum(pl[r0:r1, c0:c1])


def test_scatter_quirks():
    """The elif chain of tester.py:377-390 applies one correction only; what it leaves negative wraps (numpy indexing)."""
    boxes = po.boxes_xywh([[100.0, 100.0, 140.0, 200.0]])                             # w 40, h 100
    # right of the box (x0 >= w) and above it (y0 < 0): only x is clamped, y wraps to the bottom rows
    xy = np.array([[147.0, 85.0], [95.0, 215.0], [95.0, 85.0], [120.0, 150.0], [120.5, 150.4]])
    ty = np.array([0, 1, 2, 3, 3])
    own = po.scatter(xy, ty, boxes)
    assert own[0, 0, 56 - 8, 35] == 0                                                 # int(-15*0.56) = -8 -> row 48
    assert own[0, 1, 55, 36 - 4] == 1                                                 # y clamped, x0 = int(-4.5) = -4 wraps
    assert own[0, 2, 0, 0] == 2                                                       # both negative -> (0, 0)
    assert own[0, 3, 28, 18] == 4 and (own[0, 3] >= 0).sum() == 1                     # same cell: the later peak owns it


def test_host_regroup_matches_oracle():
    from multiposenet.pytorch_b200.evaluate import prn_assign as pp
    kps, _ = po.synthetic_case(4, persons=12, noise_peaks=30)
    rng = np.random.default_rng(0)
    kps = [kps[i] for i in rng.permutation(len(kps))]                                 # arbitrary arrival order
    xy, ty = pp._regroup(kps)
    oxy, oty = po.sort_peaks(kps)
    assert np.array_equal(xy, oxy) and np.array_equal(ty, oty)
    assert np.array_equal(pp._gaussian_weights(), po.gaussian_weights())
    assert pp._regroup([])[0].shape == (0, 2)
