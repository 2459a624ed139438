"""GPU parity of mpn_heatmap_peaks (joint_utils.py:19-32, 61-152) against the numpy restatement (bit-exact: same unfused
float32 operations in the same order) and the reference goldens (scores within 1e-6: cv2's SIMD build)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "peaks.npz")


def _rows(heat_np, factor=4, thre1=0.1, max_peaks=1024):
    from multiposenet.pytorch_b200 import ops
    h = torch.from_numpy(np.ascontiguousarray(heat_np)).cuda()
    rows, cnt = ops.heatmap_peaks(h, thre1=thre1, factor=factor, max_peaks=max_peaks, channels=heat_np.shape[1])
    torch.cuda.synchronize()
    return rows.cpu().numpy(), cnt.cpu().numpy()


@pytest.mark.parametrize("factor", [1, 2, 3, 4, 8])
def test_peaks_bit_exact_vs_oracle(factor):
    from oracle import peaks_oracle
    heat = np.stack([peaks_oracle.synthetic_heatmaps(s, C=18, H=40, W=56, persons=3) for s in (3, 4, 5)])
    rows, cnt = _rows(heat, factor=factor)
    for b in range(heat.shape[0]):
        want = peaks_oracle.joint_list(heat[b], 0.1, factor)
        assert cnt[b] == len(want)
        got = rows[b, : cnt[b]]
        assert np.array_equal(got[:, [0, 1, 3, 4]], want[:, [0, 1, 3, 4]].astype(np.float32))
        assert np.array_equal(got[:, 2].view(np.uint32), want[:, 2].astype(np.float32).view(np.uint32))


def test_peaks_vs_reference_goldens():
    from oracle import peaks_oracle
    g = np.load(GOLD)
    heat = np.stack([peaks_oracle.synthetic_heatmaps(1), peaks_oracle.synthetic_heatmaps(2)])   # [2, 18, 120, 160]
    rows, cnt = _rows(heat)
    for b, key in enumerate(("seed1_rows", "seed2_rows")):
        gold = g[key]
        assert cnt[b] == len(gold)
        got = rows[b, : cnt[b]].astype(np.float64)
        assert np.array_equal(got[:, [0, 1, 3, 4]], gold[:, [0, 1, 3, 4]])
        assert np.abs(got[:, 2] - gold[:, 2]).max() <= 1e-6
    small, cs = _rows(g["small_heat"][None])
    assert cs[0] == len(g["small_rows"]) and np.array_equal(small[0, : cs[0], :2].astype(np.float64), g["small_rows"][:, :2])


def test_peaks_channel_slice_capacity_and_empty():
    from oracle import peaks_oracle
    from multiposenet.pytorch_b200 import ops
    heat19 = np.concatenate([peaks_oracle.synthetic_heatmaps(6, C=18, H=24, W=32, persons=2),
                             np.ones((1, 24, 32), np.float32)])[None]                             # 19th channel must be ignored
    h = torch.from_numpy(heat19).cuda()
    rows, cnt = ops.heatmap_peaks(h, max_peaks=8)                                                 # capacity smaller than the count
    want = peaks_oracle.joint_list(heat19[0, :18], 0.1, 4)
    assert int(cnt[0]) == len(want) > 8
    assert np.array_equal(rows[0, :8, :2].cpu().numpy(), want[:8, :2].astype(np.float32))
    rows, cnt = ops.heatmap_peaks(torch.zeros(2, 18, 16, 16, device="cuda"))
    assert cnt.tolist() == [0, 0]


def test_get_joint_list_drop_in():
    from oracle import peaks_oracle
    from multiposenet.pytorch_b200.network.joint_utils import get_joint_list, joint_lists
    hm = peaks_oracle.synthetic_heatmaps(7, C=18, H=30, W=40, persons=3)
    want = peaks_oracle.joint_list(hm, 0.1, 4)
    want[:, :2] *= 1.5
    img = np.zeros((120, 160, 3), np.float32)
    got_np = get_joint_list(img, {"thre1": 0.1}, np.ascontiguousarray(hm.transpose(1, 2, 0)), 1.5)   # the reference's call
    got_cuda = get_joint_list(img, {"thre1": 0.1}, torch.from_numpy(hm).cuda(), 1.5)                 # without the round trip
    for got in (got_np, got_cuda):
        assert got.dtype == np.float64 and got.shape == want.shape
        assert np.array_equal(got[:, [0, 1, 3, 4]], want[:, [0, 1, 3, 4]])
        assert np.abs(got[:, 2] - want[:, 2]).max() <= 1e-7
    with pytest.raises(NotImplementedError):
        get_joint_list(np.zeros((100, 160, 3), np.float32), {"thre1": 0.1}, torch.from_numpy(hm).cuda(), 1.0)
    many = joint_lists(torch.from_numpy(np.stack([hm, hm])).cuda(), scales=[1.0, 2.0])
    assert len(many) == 2 and np.array_equal(many[1][:, :2], many[0][:, :2] * 2.0)
