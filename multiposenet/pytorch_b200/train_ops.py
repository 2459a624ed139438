"""Host wrappers of the training-step entry points (include/mpn_b200.h, "training step" block).

Same conventions as ops.py: torch owns the memory, libmpn_b200 does the arithmetic.
"""
import ctypes

import torch

from . import _lib, ops
from ._lib import FMT_BF16, FMT_BF16X2, ConvDesc, ConvPtrs, check
from .ops import Act, PackedConv, _ptr, _stream


def _like(a, C=None, H=None, W=None, zero=False):
    return Act(a.fmt, a.N, a.H if H is None else H, a.W if W is None else W, a.C if C is None else C, a.hi.device, zero=zero)


def _dense(a):
    assert a.cstride == a.C and not a.wpitch, "dense NHWC activation expected"
    return a


def pack_dgrad_filter(weight, fmt, cout_pad=None):
    """OIHW fp32 -> PackedConv of the data-gradient conv (Cin' = padded Cout, Cout' = Cin, taps flipped)."""
    w = weight.detach().contiguous()
    Cout, Cin, R, S = w.shape
    cpad = cout_pad or (Cout + 63) // 64 * 64
    pc = PackedConv()
    pc.Cout, pc.Cin, pc.R, pc.S, pc.fmt, pc.cout_pad = Cin, cpad, R, S, fmt, Cin
    pc.w_hi = torch.empty((Cin, R, S, cpad), dtype=torch.bfloat16, device=w.device)
    pc.w_lo = torch.empty_like(pc.w_hi) if fmt == FMT_BF16X2 else None
    pc.scale = pc.bias = None
    check(_lib.lib().mpn_pack_filter_dgrad_bf16(_ptr(w), _ptr(pc.w_hi), _ptr(pc.w_lo), Cout, Cin, R, S, cpad, _stream()),
          "mpn_pack_filter_dgrad_bf16")
    return pc


def zero_insert2(dy, H, W):
    out = _like(_dense(dy), H=H, W=W)
    check(_lib.lib().mpn_zero_insert2(_ptr(dy.hi), _ptr(dy.lo), _ptr(out.hi), _ptr(out.lo), dy.N, dy.H, dy.W, dy.C, H, W, dy.fmt,
                                      _stream()), "mpn_zero_insert2")
    return out


def conv_dgrad(dy, weight, stride, pad, in_hw, fmt, residual=None, pc=None):
    """dX of y = conv2d(x, weight, stride, pad): dy Act [N,OH,OW,Cout(+pad to 64)], returns Act [N,H,W,Cin].
    residual (Act like dX) is added in the epilogue (gradient accumulation)."""
    H, W = in_hw
    R = weight.shape[2]
    if pc is None:
        pc = pack_dgrad_filter(weight, fmt, cout_pad=dy.C)
    assert pc.Cin == dy.C
    if stride == 2:
        dy = zero_insert2(dy, H, W)
    else:
        assert stride == 1
    dx = ops.conv2d(dy, pc, stride=1, pad=R - 1 - pad, residual=residual)
    assert (dx.H, dx.W) == (H, W), ((dx.H, dx.W), (H, W))
    return dx


def conv_wgrad(x, dy, Cout, R, S, stride, pad):
    """[Cout][R][S][Cin] fp32 weight gradient of y = conv2d(x, w, stride, pad); x / dy are Acts (dy may carry zero-padded
    channels beyond Cout)."""
    L = _lib.lib()
    d = ConvDesc()
    d.N, d.H, d.W, d.Cin = x.N, x.H, x.W, x.C
    d.Cout, d.R, d.S, d.stride, d.pad = Cout, R, S, stride, pad
    d.OH = (x.H + 2 * pad - R) // stride + 1
    d.OW = (x.W + 2 * pad - S) // stride + 1
    assert (dy.N, dy.H, dy.W) == (x.N, d.OH, d.OW) and dy.C >= Cout and dy.fmt == x.fmt
    d.fmt, d.in_cstride, d.in_wpitch, d.k_overlap = x.fmt, x.cstride, x.wpitch, x.k_overlap
    d.res_cstride = dy.cstride
    d.out_rep = 1
    p = ConvPtrs()
    p.x_hi, p.x_lo, p.res_hi, p.res_lo = _ptr(x.hi), _ptr(x.lo), _ptr(dy.hi), _ptr(dy.lo)
    dw = torch.empty((Cout, R, S, x.C), dtype=torch.float32, device=x.hi.device)
    check(L.mpn_conv2d_wgrad(ctypes.byref(d), ctypes.byref(p), _ptr(dw), _stream()), "mpn_conv2d_wgrad")
    return dw


def unpack_filter_grad(dw):
    Cout, R, S, Cin = dw.shape
    out = torch.empty((Cout, Cin, R, S), dtype=torch.float32, device=dw.device)
    check(_lib.lib().mpn_unpack_filter_grad(_ptr(dw), _ptr(out), Cout, Cin, R, S, _stream()), "mpn_unpack_filter_grad")
    return out


def stem_unpack_filter_grad(dw):
    Cout = dw.shape[0]
    out = torch.empty((Cout, 3, 7, 7), dtype=torch.float32, device=dw.device)
    check(_lib.lib().mpn_stem_unpack_filter_grad(_ptr(dw), _ptr(out), Cout, _stream()), "mpn_stem_unpack_filter_grad")
    return out


def channel_sum(a, C=None, coffset=0):
    """fp32 [C] per-channel sum over all pixels (bias gradient)."""
    C = a.C if C is None else C
    ws = torch.empty((C,), dtype=torch.float64, device=a.hi.device)
    out = torch.empty((C,), dtype=torch.float32, device=a.hi.device)
    L = _lib.lib()
    check(L.mpn_channel_sums(_ptr(a.hi), _ptr(a.lo), a.N * a.H * a.W, C, a.cstride, coffset, a.fmt, _ptr(ws), None, _stream()),
          "mpn_channel_sums")
    check(L.mpn_double_to_float(_ptr(ws), _ptr(out), C, 1.0, _stream()), "mpn_double_to_float")
    return out


class BNState(object):
    __slots__ = ("mean", "var", "y", "z", "relu", "eps")


def bn_train_forward(y, bn, relu, residual=None, update_running=True):
    """BatchNorm2d in training mode on the raw conv output y (+residual, +ReLU).  Returns (z, BNState)."""
    L = _lib.lib()
    y = _dense(y)
    C, pixels, dev = y.C, y.N * y.H * y.W, y.hi.device
    st = BNState()
    st.mean = torch.empty((C,), dtype=torch.float32, device=dev)
    st.var = torch.empty((C,), dtype=torch.float32, device=dev)
    ws = torch.empty((2 * C,), dtype=torch.float64, device=dev)
    check(L.mpn_bn_stats(_ptr(y.hi), _ptr(y.lo), pixels, C, y.fmt, _ptr(st.mean), _ptr(st.var), _ptr(ws), _stream()), "mpn_bn_stats")
    if update_running:
        check(L.mpn_bn_update_running(_ptr(st.mean), _ptr(st.var), _ptr(bn.running_mean), _ptr(bn.running_var), pixels,
                                      float(bn.momentum), C, _stream()), "mpn_bn_update_running")
        bn.num_batches_tracked += 1
    z = _like(y)
    coef = torch.empty((3 * C,), dtype=torch.float32, device=dev)
    check(L.mpn_bn_apply(_ptr(y.hi), _ptr(y.lo), _ptr(st.mean), _ptr(st.var), _ptr(bn.weight.detach()), _ptr(bn.bias.detach()),
                         float(bn.eps), _ptr(residual.hi) if residual is not None else None,
                         _ptr(residual.lo) if residual is not None else None, int(relu), _ptr(z.hi), _ptr(z.lo), pixels, C,
                         y.fmt, _ptr(coef), _stream()), "mpn_bn_apply")
    st.y, st.z, st.relu, st.eps = y, z, bool(relu), float(bn.eps)
    return z, st


def bn_train_backward(dz, st, bn, want_g=False):
    """Returns (dy, g or None, dgamma, dbeta) for z = act(bn(y) [+ residual])."""
    L = _lib.lib()
    y = st.y
    C, pixels, dev = y.C, y.N * y.H * y.W, y.hi.device
    dy = _like(y)
    g = _like(y) if want_g else None
    dgamma = torch.empty((C,), dtype=torch.float32, device=dev)
    dbeta = torch.empty((C,), dtype=torch.float32, device=dev)
    ws = torch.empty((2 * C,), dtype=torch.float64, device=dev)
    coef = torch.empty((3 * C,), dtype=torch.float32, device=dev)
    check(L.mpn_bn_backward(_ptr(dz.hi), _ptr(dz.lo), _ptr(st.z.hi), _ptr(st.z.lo), _ptr(y.hi), _ptr(y.lo), _ptr(st.mean),
                            _ptr(st.var), _ptr(bn.weight.detach()), st.eps, int(st.relu), pixels, C, y.fmt, _ptr(dy.hi), _ptr(dy.lo),
                            _ptr(g.hi) if g is not None else None, _ptr(g.lo) if (g is not None and g.lo is not None) else None,
                            _ptr(dgamma), _ptr(dbeta), _ptr(ws), _ptr(coef), _stream()), "mpn_bn_backward")
    return dy, g, dgamma, dbeta


def relu_backward(dz, z):
    out = _like(dz)
    check(_lib.lib().mpn_relu_backward(_ptr(dz.hi), _ptr(dz.lo), _ptr(z.hi), _ptr(z.lo), _ptr(out.hi), _ptr(out.lo), dz.hi.numel(),
                                       dz.fmt, _stream()), "mpn_relu_backward")
    return out


def add(a, b):
    out = _like(a)
    check(_lib.lib().mpn_add_act(_ptr(a.hi), _ptr(a.lo), _ptr(b.hi), _ptr(b.lo), _ptr(out.hi), _ptr(out.lo), a.hi.numel(), a.fmt,
                                 _stream()), "mpn_add_act")
    return out


def maxpool_backward(x, dy):
    dx = _like(x)
    check(_lib.lib().mpn_maxpool3x3s2_backward(_ptr(x.hi), _ptr(x.lo), _ptr(dy.hi), _ptr(dy.lo), _ptr(dx.hi), _ptr(dx.lo), x.N, x.H,
                                               x.W, x.C, x.fmt, _stream()), "mpn_maxpool3x3s2_backward")
    return dx


def block_sum(fine, r, C=None, coffset=0):
    """Backward of a nearest upsample by r: [N,H,W,*] -> [N,H/r,W/r,C] summing each r x r block of a channel slice."""
    C = fine.C if C is None else C
    assert fine.H % r == 0 and fine.W % r == 0
    out = Act(fine.fmt, fine.N, fine.H // r, fine.W // r, C, fine.hi.device)
    check(_lib.lib().mpn_block_sum(_ptr(fine.hi), _ptr(fine.lo), fine.cstride, coffset, _ptr(out.hi), _ptr(out.lo), fine.N, out.H,
                                   out.W, C, r, fine.fmt, _stream()), "mpn_block_sum")
    return out


def mse_heatmap_loss(pred, gt, weight, loss_acc, fmt, Cd=64, grad_scale=1.0):
    """Adds mean((pred[:, :18]*w - gt*w)^2) to loss_acc (fp64 [1]) and returns d loss / d pred as an NHWC Act with Cd channels."""
    B, Cp, H, W = pred.shape
    d = Act(fmt, B, H, W, Cd, pred.device)
    check(_lib.lib().mpn_mse_heatmap_loss(_ptr(pred.contiguous()), _ptr(gt.contiguous()), _ptr(weight.contiguous()), B, Cp, H, W,
                                          _ptr(loss_acc), _ptr(d.hi), _ptr(d.lo), Cd, fmt, float(grad_scale), _stream()),
          "mpn_mse_heatmap_loss")
    return d
