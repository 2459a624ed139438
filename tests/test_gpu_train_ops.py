"""GPU parity of the training-step primitives vs torch autograd (fp32, TF32 off) -- SURVEY 8(a17)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _grad_enabled():
    with torch.enable_grad():
        yield


def _setup():
    from gpu_util import no_tf32
    no_tf32()


def _act(t, fmt, cpad=None):
    from multiposenet.pytorch_b200 import ops
    return ops.act_from_nchw(t, fmt, cstride=None) if cpad is None else ops.act_from_nchw(
        F.pad(t, (0, 0, 0, 0, 0, cpad - t.shape[1])), fmt)


WG_CASES = [  # N, H, W, Cin, Cout, R, stride, pad
    (2, 16, 24, 64, 64, 1, 1, 0),
    (2, 16, 24, 64, 128, 3, 1, 1),
    (1, 30, 40, 256, 256, 3, 1, 1),
    (2, 30, 40, 128, 128, 3, 2, 1),
    (2, 15, 21, 256, 512, 1, 2, 0),
    (2, 15, 20, 512, 128, 1, 1, 0),
    (3, 4, 5, 256, 256, 3, 1, 1),       # whole images per K block (TN > 1)
    (2, 24, 32, 256, 18, 1, 1, 0),      # head: Cout padded to 64 channels in dY
    (2, 12, 16, 128, 19, 1, 1, 0),
]


@pytest.mark.parametrize("fmt,tol", [(2, 2e-4), (1, 1e-2)])
@pytest.mark.parametrize("case", WG_CASES)
def test_wgrad_vs_autograd(case, fmt, tol):
    from gpu_util import nerr, round_fmt
    from multiposenet.pytorch_b200 import train_ops as T
    _setup()
    N, H, W, Cin, Cout, R, stride, pad = case
    g = torch.Generator().manual_seed(3)
    x = torch.randn(N, Cin, H, W, generator=g).cuda()
    OH, OW = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
    dy = torch.randn(N, Cout, OH, OW, generator=g).cuda()
    want = torch.nn.grad.conv2d_weight(round_fmt(x, fmt), (Cout, Cin, R, R), round_fmt(dy, fmt), stride=stride, padding=pad)
    cpad = (Cout + 63) // 64 * 64
    dw = T.conv_wgrad(_act(x, fmt), _act(dy, fmt, cpad), Cout, R, R, stride, pad)
    got = T.unpack_filter_grad(dw)
    torch.cuda.synchronize()
    assert got.shape == want.shape
    assert nerr(got, want) <= tol


@pytest.mark.parametrize("fmt,tol", [(2, 2e-4), (1, 1e-2)])
@pytest.mark.parametrize("case", WG_CASES)
def test_dgrad_vs_autograd(case, fmt, tol):
    from gpu_util import nerr, round_fmt
    from multiposenet.pytorch_b200 import train_ops as T
    _setup()
    N, H, W, Cin, Cout, R, stride, pad = case
    g = torch.Generator().manual_seed(4)
    w = (torch.randn(Cout, Cin, R, R, generator=g) / (Cin * R * R) ** 0.5).cuda()
    OH, OW = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
    dy = torch.randn(N, Cout, OH, OW, generator=g).cuda()
    want = torch.nn.grad.conv2d_input((N, Cin, H, W), round_fmt(w, fmt), round_fmt(dy, fmt), stride=stride, padding=pad)
    cpad = (Cout + 63) // 64 * 64
    res = torch.randn(N, Cin, H, W, generator=g).cuda()
    dx = T.conv_dgrad(_act(dy, fmt, cpad), w, stride, pad, (H, W), fmt, residual=_act(res, fmt)).to_nchw()
    torch.cuda.synchronize()
    assert nerr(dx, want + round_fmt(res, fmt)) <= tol


def test_stem_wgrad_vs_autograd():
    from gpu_util import nerr, round_fmt
    from multiposenet.pytorch_b200 import ops, train_ops as T
    _setup()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 64, 96, generator=g).cuda()
    dy = torch.randn(2, 64, 32, 48, generator=g).cuda()
    want = torch.nn.grad.conv2d_weight(x, (64, 3, 7, 7), round_fmt(dy, 2), stride=2, padding=3)
    xs = ops.stem_pack_input(x, 2)
    dw = T.conv_wgrad(xs, _act(dy, 2), 64, 4, 1, 1, 0)
    got = T.stem_unpack_filter_grad(dw)
    torch.cuda.synchronize()
    assert nerr(got, want) <= 2e-4


@pytest.mark.parametrize("fmt,tol", [(2, 1e-4), (1, 2e-2)])
@pytest.mark.parametrize("relu,res", [(True, False), (True, True), (False, False)])
def test_bn_train_forward_backward(fmt, tol, relu, res):
    from gpu_util import nerr, round_fmt
    from multiposenet.pytorch_b200 import train_ops as T
    _setup()
    g = torch.Generator().manual_seed(6)
    N, C, H, W = 3, 128, 10, 14
    y = (torch.randn(N, C, H, W, generator=g) * 2 + 0.5).cuda()
    r = torch.randn(N, C, H, W, generator=g).cuda() if res else None
    dz = torch.randn(N, C, H, W, generator=g).cuda()
    bn = torch.nn.BatchNorm2d(C).cuda()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.2)
    bn_ref = torch.nn.BatchNorm2d(C).cuda()
    bn_ref.load_state_dict(bn.state_dict())
    bn_ref.train()
    yr = round_fmt(y, fmt).requires_grad_(True)
    rr = round_fmt(r, fmt).requires_grad_(True) if res else None
    zr = bn_ref(yr)
    if res:
        zr = zr + rr
    if relu:
        zr = F.relu(zr)
    zr.backward(round_fmt(dz, fmt))
    ya, ra, dza = _act(y, fmt), (_act(r, fmt) if res else None), _act(dz, fmt)
    z, st = T.bn_train_forward(ya, bn, relu, residual=ra)
    dy, gmask, dgamma, dbeta = T.bn_train_backward(dza, st, bn, want_g=res)
    torch.cuda.synchronize()
    assert nerr(z.to_nchw(), zr.detach()) <= tol
    assert nerr(dy.to_nchw(), yr.grad) <= 5 * tol
    assert nerr(dgamma, bn_ref.weight.grad) <= 5 * tol and nerr(dbeta, bn_ref.bias.grad) <= 5 * tol
    if res:
        assert nerr(gmask.to_nchw(), rr.grad) <= tol
    assert nerr(bn.running_mean, bn_ref.running_mean) <= 1e-4 and nerr(bn.running_var, bn_ref.running_var) <= 1e-4
    assert int(bn.num_batches_tracked) == 1


def test_pool_upsample_loss_backward():
    from gpu_util import nerr, round_fmt
    from multiposenet.pytorch_b200 import ops, train_ops as T
    _setup()
    fmt = 2
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 64, 18, 22, generator=g).cuda()
    xr = round_fmt(x, fmt).requires_grad_(True)
    yr = F.max_pool2d(xr, 3, 2, 1)
    dy = torch.randn(yr.shape, generator=g).cuda()
    yr.backward(round_fmt(dy, fmt))
    dx = T.maxpool_backward(_act(x, fmt), _act(dy, fmt)).to_nchw()
    assert nerr(dx, xr.grad) <= 1e-4
    # nearest upsample backward (block sums), with a channel slice of a wider tensor
    fine = torch.randn(2, 96, 16, 24, generator=g).cuda()
    for r in (2, 4, 8):
        src = torch.zeros(2, 32, 16 // r, 24 // r, device="cuda", requires_grad=True)
        F.interpolate(src, scale_factor=r, mode="nearest").backward(round_fmt(fine[:, 32:64], fmt))
        got = T.block_sum(_act(fine, fmt), r, C=32, coffset=32).to_nchw()
        assert nerr(got, src.grad) <= 1e-4
    # weighted MSE heat-map loss (posenet.py:380-387) and its gradient
    B, H, W = 2, 12, 16
    pred = torch.randn(B, 19, H, W, generator=g).cuda().requires_grad_(True)
    gt = torch.rand(B, 18, H, W, generator=g).cuda()
    wt = (torch.rand(B, 18, H, W, generator=g) > 0.3).float().cuda()
    loss_ref = F.mse_loss(pred[:, :18] * wt, wt * gt)
    loss_ref.backward()
    acc = torch.zeros(1, dtype=torch.float64, device="cuda")
    d = T.mse_heatmap_loss(pred.detach(), gt, wt, acc, fmt, Cd=64)
    torch.cuda.synchronize()
    assert abs(float(acc) - float(loss_ref)) <= 1e-5 * max(1.0, float(loss_ref))
    dn = d.to_nchw()
    assert nerr(dn[:, :19], pred.grad) <= 1e-4 and float(dn[:, 19:].abs().max()) == 0.0
    # bias gradient = per-channel sum
    s = T.channel_sum(_act(fine, fmt))
    assert nerr(s, round_fmt(fine, fmt).sum(dim=(0, 2, 3))) <= 1e-4
