"""GPU parity: anchors / decode / filter / sort / NMS through the C ABI vs the oracle and goldens (bit-exact)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_anchors_device_cache(golden_dir):
    from multiposenet.pytorch_b200 import ops
    g = _g(golden_dir, "anchors.npz")
    for key in g.files:
        h, w = map(int, key.split("x"))
        a = ops.anchors_for(h, w, torch.device("cuda")).cpu().numpy()[0]
        assert np.array_equal(a.view(np.uint32), g[key].view(np.uint32)), key


def test_decode_clip_vs_reference_golden(golden_dir):
    from multiposenet.pytorch_b200 import ops
    g = _g(golden_dir, "decode.npz")
    b = ops.decode_clip(torch.from_numpy(g["anchors"]).cuda(), torch.from_numpy(g["deltas"]).cuda(), 64, 96).cpu().numpy()
    # exp() differs between CPU SLEEF and CUDA libdevice by an ulp or two: float tolerance, stated
    np.testing.assert_allclose(b, g["boxes"], rtol=2e-6, atol=2e-5)


def test_nms_goldens_bit_exact(golden_dir):
    from multiposenet.pytorch_b200 import pth_nms
    g = _g(golden_dir, "nms.npz")
    names = sorted({k[: -len("_dets")] for k in g.files if k.endswith("_dets")})
    for n in names:
        dets = torch.from_numpy(g[n + "_dets"])
        for thr in (0.5, 0.3):
            k_gpu = pth_nms(dets.cuda(), thr)
            assert k_gpu.dtype == torch.int64 and k_gpu.is_cuda
            assert np.array_equal(k_gpu.cpu().numpy(), g["%s_keep_gt_%g" % (n, thr)]), (n, thr)
            k_cpu = pth_nms(dets, thr)  # CPU tensor -> reference CPU-branch semantics (>=)
            assert not k_cpu.is_cuda
            assert np.array_equal(k_cpu.numpy(), g["%s_keep_ge_%g" % (n, thr)]), (n, thr)


def test_nms_known_answers_and_edges():
    from multiposenet.pytorch_b200 import pth_nms
    d = torch.tensor([[0, 0, 9, 9, 0.9], [0, 0, 9, 19, 0.8], [100, 100, 120, 120, 0.7]])
    assert pth_nms(d.cuda(), 0.5).tolist() == [0, 1, 2]   # IoU == 0.5 exactly: '>' keeps it
    assert pth_nms(d, 0.5).tolist() == [0, 2]             # '>=' suppresses it
    assert pth_nms(d[[2, 0, 1]].cuda(), 0.5).tolist() == [1, 2, 0]
    assert pth_nms(torch.zeros(0, 5).cuda(), 0.5).shape == (0,)
    assert pth_nms(d[:1].cuda(), 0.5).tolist() == [0]
    with pytest.raises(RuntimeError):
        pth_nms(torch.zeros(3, 4).cuda(), 0.5)


@pytest.mark.parametrize("n,seed", [(1, 0), (63, 1), (64, 2), (65, 3), (500, 4), (4096, 5), (5000, 6)])
def test_nms_random_vs_oracle(n, seed):
    from multiposenet.pytorch_b200 import pth_nms
    from oracle import nms_oracle
    rng = np.random.default_rng(seed)
    ctr = rng.uniform(0, 600, (max(n // 30, 1), 2))
    which = rng.integers(0, len(ctr), n)
    xy = ctr[which] + rng.normal(0, 12, (n, 2))
    wh = rng.uniform(20, 150, (n, 2))
    sc = rng.permutation(n) / float(n) * 0.9 + 0.05
    dets = np.concatenate([xy - wh / 2, xy + wh / 2, sc[:, None]], 1).astype(np.float32)
    for thr, ge in ((0.5, False), (0.5, True), (0.3, False)):
        want = nms_oracle.nms_gpu_semantics(dets, thr, ge=ge)
        t = torch.from_numpy(dets)
        got = pth_nms(t, thr).numpy() if ge else pth_nms(t.cuda(), thr).cpu().numpy()
        assert np.array_equal(got, want), (n, thr, ge)
    # ties in score: stable order (documented divergence from torch's unstable sort)
    dets[:, 4] = np.round(dets[:, 4], 1)
    want = nms_oracle.nms_gpu_semantics(dets, 0.5)
    assert np.array_equal(pth_nms(torch.from_numpy(dets).cuda(), 0.5).cpu().numpy(), want)


def test_mask_stage_vs_reference_kernel():
    """oracle/_ref: the reference's own nms_kernel.cu compiled unchanged for sm_100a (mask-stage oracle)."""
    from multiposenet.pytorch_b200 import ops
    from oracle import nms_oracle
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_nms_kernel.so")
    rng = np.random.default_rng(11)
    n = 777
    xy = rng.uniform(0, 300, (n, 2)); wh = rng.uniform(10, 120, (n, 2))
    dets = np.concatenate([xy, xy + wh, np.sort(rng.uniform(0.05, 1, n))[::-1, None]], 1).astype(np.float32)
    t = torch.from_numpy(dets).cuda()
    ours = ops.nms_mask(t, 0.5).cpu().numpy().view(np.uint64)
    want = nms_oracle.nms_mask(dets, 0.5)
    cb = want.shape[1]
    for i in range(n):  # upper triangle (the part gpu_nms reads)
        assert np.array_equal(ours[i, i // 64:], want[i, i // 64:]), i
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libref_nms_kernel.so not built (needs /root/reference at build time)")
    ref = ctypes.CDLL(so)
    ref._nms.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float]
    ref._nms.restype = None
    m = torch.zeros((n, cb), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ref._nms(n, ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(m.data_ptr()), ctypes.c_float(0.5))
    torch.cuda.synchronize()
    refm = m.cpu().numpy().view(np.uint64)
    for i in range(n):
        assert np.array_equal(refm[i, i // 64:], want[i, i // 64:]), i
        assert np.array_equal(refm[i, i // 64:], ours[i, i // 64:]), i


def test_filter_sort_nms_batched_vs_oracle():
    from multiposenet.pytorch_b200 import ops
    from oracle import nms_oracle
    rng = np.random.default_rng(3)
    B, A = 3, 5000
    cls = rng.uniform(0, 0.2, (B, A, 1)).astype(np.float32)
    cls[1] *= 0.2  # image 1: nothing above 0.05 -> zero candidates
    xy = rng.uniform(0, 500, (B, A, 2)); wh = rng.uniform(10, 100, (B, A, 2))
    boxes = np.concatenate([xy, xy + wh], 2).astype(np.float32)
    det = ops.filter_sort_nms(torch.from_numpy(cls).cuda(), torch.from_numpy(boxes).cuda(), 0.05, 0.5, max_cand=4096)
    torch.cuda.synchronize()
    for b in range(B):
        m = cls[b, :, 0] > 0.05
        ns = int(m.sum())
        assert int(det.cand_cnt[b]) == ns
        assert np.array_equal(det.cand_idx[b, :ns].cpu().numpy(), np.nonzero(m)[0])
        d = np.concatenate([boxes[b][m], cls[b][m]], 1)
        want = nms_oracle.nms_gpu_semantics(d, 0.5)
        k = int(det.keep_cnt[b])
        assert k == len(want)
        assert np.array_equal(det.keep_idx[b, :k].cpu().numpy(), want)
        assert np.array_equal(det.scores[b, :k].cpu().numpy(), d[want, 4])
        assert np.array_equal(det.boxes[b, :k].cpu().numpy(), d[want, :4])
