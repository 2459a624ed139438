"""Helpers shared by the -m gpu tests (import only inside GPU tests)."""
import numpy as np
import torch
import torch.nn.functional as F

from multiposenet.pytorch_b200 import ops
from multiposenet.pytorch_b200._lib import FMT_BF16, FMT_BF16X2, FMT_F16F8, FMT_F32, OUT_ACT, OUT_F32_NCHW, OUT_F32_NHWC


def no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")


def nerr(a, b):
    """max|a-b| / max|b|  (SURVEY section 7 'hard parts': the metric for the 1e-3 bar)."""
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def round_fmt(t, fmt):
    if fmt == FMT_BF16:
        return t.bfloat16().float()
    if fmt == FMT_BF16X2:
        hi = t.bfloat16().float()
        return hi + (t - hi).bfloat16().float()
    return t


def strip_h8(a):
    """The same f16f8 activation stored WITHOUT its e5m2 copy plane (a new, smaller lo buffer: a stray h8 access would fault
    or read another allocation, not silently work)."""
    b = ops.Act(a.fmt, a.N, a.H, a.W, a.C, a.hi.device, cstride=a.cstride, has_h8=False)
    b.hi.copy_(a.hi)
    b.lo[0].copy_(a.lo[0])
    return b


def conv_case(fmt, N, H, W, Cin, Cout, R, stride, pad, relu=False, sigmoid=False, residual=False, up=False, bn=False,
              bias=True, out_mode=OUT_ACT, rep=1, seed=0, coffset=0, ctotal=None, no_h8=False, derive=None):
    """Returns (ours NCHW fp32, reference NCHW fp32 with fmt-rounded operands, reference with fp32 operands).
    FMT_F16F8 operand variants: no_h8 = filter packed with the fp16 residual plane, tensors without the copy plane (MODE_F16F8B);
    derive = True: tensors without the copy plane, derived in shared memory (MODE_F16F8C); False: the stored plane is loaded."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    dev = "cuda"
    x = torch.randn(N, Cin, H, W, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, R, R, generator=g) / (Cin * R * R) ** 0.5).to(dev)
    b = torch.randn(Cout, generator=g).to(dev) if (bias and not bn) else None
    OH, OW = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
    bnp = None
    if bn:
        bnp = (torch.rand(Cout, generator=g).to(dev) + 0.5, torch.randn(Cout, generator=g).to(dev) * 0.1,
               torch.randn(Cout, generator=g).to(dev) * 0.1, torch.rand(Cout, generator=g).to(dev) + 0.5, 1e-5)
    res = torch.randn(N, Cout, OH, OW, generator=g).to(dev) if residual else None
    upt = torch.randn(N, Cout, (OH + 1) // 2, (OW + 1) // 2, generator=g).to(dev) if up else None

    def ref(xx, ww, rr, uu):
        y = F.conv2d(xx, ww, b, stride=stride, padding=pad)
        if bnp is not None:
            y = F.batch_norm(y, bnp[2], bnp[3], bnp[0], bnp[1], False, 0.0, bnp[4])
        if rr is not None:
            y = y + rr
        if uu is not None:
            y = F.interpolate(uu, size=(OH, OW), mode="nearest") + y
        if relu:
            y = F.relu(y)
        if sigmoid:
            y = torch.sigmoid(y)
        if rep > 1:
            y = F.interpolate(y, scale_factor=rep, mode="nearest")
        return y

    ref_exact = ref(x, w, res, upt)
    ref_rounded = ref(round_fmt(x, fmt), round_fmt(w, fmt), round_fmt(res, fmt) if res is not None else None,
                      round_fmt(upt, fmt) if upt is not None else None)
    pc = ops.pack_conv(w, b, bnp, fmt, in_no_h8=no_h8)
    xa = ops.act_from_nchw(x, fmt)
    ra = ops.act_from_nchw(res, fmt) if res is not None else None
    ua = ops.act_from_nchw(upt, fmt) if upt is not None else None
    if fmt == 3 and derive is None and not no_h8:
        derive = ops.DERIVE_H8
    if no_h8 or (derive and fmt == 3):   # input, shortcut and upsample source without the copy plane; the output too (unless it is a concat slice)
        xa, ra, ua = strip_h8(xa), (strip_h8(ra) if ra is not None else None), (strip_h8(ua) if ua is not None else None)
    slim_out = no_h8 or bool(derive and fmt == 3)
    if out_mode == OUT_ACT:
        out = None
        if ctotal is not None:
            out = ops.Act(fmt, N, OH * rep, OW * rep, ctotal, dev, zero=True)
        o = ops.conv2d(xa, pc, stride=stride, pad=pad, relu=relu, sigmoid=sigmoid, residual=ra, up=ua, out=out,
                       out_coffset=coffset, out_rep=rep, want_h8=not slim_out, derive=derive)
        assert o.has_h8 == (not slim_out or ctotal is not None)
        ours = o.to_nchw()
        if ctotal is not None:
            ours = ours[:, coffset:coffset + Cout]
    elif out_mode == OUT_F32_NCHW:
        ours = ops.conv2d(xa, pc, stride=stride, pad=pad, relu=relu, sigmoid=sigmoid, residual=ra, up=ua,
                          out_mode=out_mode, out_rep=rep, derive=derive)
    else:
        ours = ops.conv2d(xa, pc, stride=stride, pad=pad, relu=relu, sigmoid=sigmoid, residual=ra, up=ua,
                          out_mode=out_mode, out_rep=rep, derive=derive).permute(0, 3, 1, 2).contiguous()
    torch.cuda.synchronize()
    return ours, ref_rounded, ref_exact


def load_model(layers, kind, precision, device="cuda"):
    from multiposenet.pytorch_b200 import poseNet
    from oracle import weights
    w = weights.make_weights(layers, kind, seed=0)
    m = poseNet(layers, precision=precision)
    sd = m.state_dict()
    for k in sd:
        if k in w:
            sd[k] = torch.from_numpy(np.ascontiguousarray(w[k]))
    m.load_state_dict(sd)
    return m.to(device).eval(), w


def image(seed, shape, device="cuda"):
    return torch.from_numpy(np.random.Generator(np.random.PCG64(seed)).standard_normal(shape, dtype=np.float32)).to(device)
