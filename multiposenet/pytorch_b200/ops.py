"""Host-side wrappers: torch tensors own the device memory, the C ABI does the work.

torch is plumbing here (allocation, streams); every computation below is a libmpn_b200 call.
"""
import ctypes

import torch

from . import _lib
import math

from ._lib import (EPI_NO_H8, EPI_RELU, EPI_SIGMOID, FMT_BF16, FMT_BF16X2, FMT_F16F8, FMT_F32, IN_DERIVE_H8, IN_NO_H8, OUT_ACT, OUT_F32_NCHW, W_MERGED,
                   OUT_F32_NHWC, ConvDesc, ConvPtrs, check)

PRECISIONS = {"fp32": FMT_F32, "bf16": FMT_BF16, "bf16x3": FMT_BF16X2, "f16f8": FMT_F16F8}

# Launch accounting for bench.py: `launches` counts kernels launched through this module; when
# `conv_events` is a list, every tensor-core / CUDA-core conv launch is bracketed by CUDA events.
import os

stats = {"launches": 0, "conv_events": None, "range_check": os.environ.get("MPN_RANGE_CHECK", "0") == "1"}
F16_MAX = 65504.0


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _dtype(fmt):
    return torch.float32 if fmt == FMT_F32 else torch.float16 if fmt == FMT_F16F8 else torch.bfloat16


class Act(object):
    """NHWC activation: `hi` (and `lo` for the split formats) are [N,H,W,cstride] tensors; FMT_F16F8: hi fp16, lo = uint8
    [2,N,H,W,cstride] (the e5m2 residual plane and the e5m2 copy plane back to back, see mpn_b200.h), or [1,...] when the tensor
    is stored without the copy plane (has_h8 False: 3 bytes per element; only convolutions packed with in_no_h8 can read it)."""
    __slots__ = ("fmt", "N", "H", "W", "C", "cstride", "hi", "lo", "wpitch", "k_overlap", "has_h8")

    def __init__(self, fmt, N, H, W, C, device, cstride=None, zero=False, wpitch=0, k_overlap=0, has_h8=True):
        self.fmt, self.N, self.H, self.W, self.C = fmt, N, H, W, C
        self.cstride = C if cstride is None else cstride
        self.wpitch, self.k_overlap = wpitch, k_overlap  # see mpn_conv_desc.in_wpitch / k_overlap
        self.has_h8 = bool(has_h8) if fmt == FMT_F16F8 else True
        mk = torch.zeros if zero else torch.empty
        shape = (N, H, wpitch if wpitch else W, self.cstride)
        self.hi = mk(shape, dtype=_dtype(fmt), device=device)
        if fmt == FMT_F16F8:
            if self.cstride % 16:
                raise ValueError("FMT_F16F8 activations need a channel stride that is a multiple of 16 (got %d)" % self.cstride)
            self.lo = mk((2 if has_h8 else 1,) + shape, dtype=torch.uint8, device=device)
        else:
            self.lo = mk(shape, dtype=torch.bfloat16, device=device) if fmt == FMT_BF16X2 else None

    def to_nchw(self):
        out = torch.empty((self.N, self.C, self.H, self.W), dtype=torch.float32, device=self.hi.device)
        check(_lib.lib().mpn_nhwc_to_nchw(_ptr(self.hi), _ptr(self.lo), _ptr(out), self.N, self.C, self.H, self.W,
                                          self.cstride, self.fmt, _stream()), "mpn_nhwc_to_nchw")
        stats["launches"] += 1
        return out


def act_from_nchw(x, fmt, cstride=None):
    assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 4
    x = x.contiguous()
    N, C, H, W = x.shape
    a = Act(fmt, N, H, W, C, x.device, cstride=cstride)
    check(_lib.lib().mpn_nchw_to_nhwc(_ptr(x), _ptr(a.hi), _ptr(a.lo), N, C, H, W, a.cstride, fmt, _stream()), "mpn_nchw_to_nhwc")
    stats["launches"] += 1
    return a


class PackedConv(object):
    """Filter + per-channel epilogue constants of one nn.Conv2d (+ folded eval-mode BatchNorm2d)."""
    __slots__ = ("Cout", "Cin", "R", "S", "fmt", "w_hi", "w_lo", "cout_pad", "scale", "bias", "acc_scale", "in_no_h8", "w_merged")


# FMT_F16F8 convolutions come in two operand variants (mpn_b200.h MPN_IN_NO_H8): the default reads the input's e5m2 copy plane
# (8 tensor-core slots per K block); in_no_h8 multiplies the fp16 plane with an fp16 weight residual instead (10 slots) and lets
# the producer of its input drop that plane from HBM -- the choice for the HBM / epilogue-bound 1x1 convolutions.
NO_H8 = __import__("os").environ.get("MPN_NO_H8", "1") == "1"
# Third variant (MPN_IN_DERIVE_H8, opt-in with MPN_DERIVE_H8=1): the kernel derives the e5m2 copy in shared memory from the fp16
# tile (two converter warps per CTA), so NO activation needs the plane in HBM and every convolution runs the 8-slot schedule on
# ordinarily packed filters with one TMA box less per K block.  Measured (profiles/r02q_derive_h8_ab.txt): bit-compatible parity,
# shortcut convolutions 6 % faster, but the large 3x3 convolutions -- 92-94 % tensor-bound in cycles with the stored plane --
# wait ~260 cycles per K block for the converters: the step is 3-4 % slower, so the stored-plane plan stays the default.
DERIVE_H8 = __import__("os").environ.get("MPN_DERIVE_H8", "0") == "1"


# FMT_F16F8 filters: the two byte planes interleaved per K block (mpn_b200.h MPN_W_MERGED) -- one TMA box less per K block
W_MERGE = __import__("os").environ.get("MPN_W_MERGED", "1") == "1"


def _auto_h8(want_h8):
    """want_h8=None (the default of every producer): store the copy plane only when the kernels do not derive it."""
    return (not DERIVE_H8) if want_h8 is None else bool(want_h8)


def _input_flags(pc, x, derive):
    """Operand variant of an FMT_F16F8 convolution reading Act x with filter pc: 0 (copy plane loaded by TMA), IN_NO_H8 (filter
    packed with the fp16 residual plane) or IN_DERIVE_H8.  derive: None = derive when enabled or when x has no copy plane."""
    if getattr(pc, "in_no_h8", False):
        return IN_NO_H8
    wm = W_MERGED if getattr(pc, "w_merged", False) else 0
    if derive is None:
        derive = DERIVE_H8 or not getattr(x, "has_h8", True)
    if derive:
        return IN_DERIVE_H8 | wm
    if not getattr(x, "has_h8", True):
        raise ValueError("this activation was stored without its h8 plane: read it with derive=True or a filter packed with in_no_h8=True")
    return wm


def pack_conv(weight, bias, bn, fmt, fold_scale=True, in_no_h8=False):
    """weight OIHW fp32 cuda; bias fp32 or None; bn = (gamma, beta, mean, var, eps) or None.

    fold_scale (bf16 formats): the eval-mode BN scale is multiplied into the filter before the hi/lo split, so the
    epilogue adds only the folded bias (and a residual can be accumulated on the tensor core); False keeps scale in
    the epilogue."""
    L = _lib.lib()
    w = weight.detach().contiguous()
    assert w.is_cuda and w.dtype == torch.float32
    pc = PackedConv()
    pc.Cout, pc.Cin, pc.R, pc.S = w.shape
    pc.fmt = fmt
    pc.in_no_h8 = bool(in_no_h8) and fmt == FMT_F16F8
    dev = w.device
    npad = (pc.Cout + 63) // 64 * 64
    pc.scale = pc.bias = None
    if bn is not None:
        gamma, beta, mean, var, eps = bn
        pc.scale = torch.zeros(npad, dtype=torch.float32, device=dev)
        pc.bias = torch.zeros(npad, dtype=torch.float32, device=dev)
        check(L.mpn_fold_bn(_ptr(gamma.detach().contiguous()), _ptr(beta.detach().contiguous()), _ptr(mean.contiguous()),
                            _ptr(var.contiguous()), float(eps), _ptr(pc.scale), _ptr(pc.bias), pc.Cout, _stream()), "mpn_fold_bn")
        assert bias is None
    elif bias is not None:
        pc.bias = torch.zeros(npad, dtype=torch.float32, device=dev)
        pc.bias[: pc.Cout].copy_(bias.detach())
    pc.acc_scale = 0.0
    if fmt == FMT_F32:
        pc.cout_pad = npad
        pc.w_hi = torch.empty((pc.R, pc.S, pc.Cin, pc.cout_pad), dtype=torch.float32, device=dev)
        pc.w_lo = None
        check(L.mpn_pack_filter_f32(_ptr(w), _ptr(pc.w_hi), pc.Cout, pc.Cin, pc.R, pc.S, pc.cout_pad, _stream()), "mpn_pack_filter_f32")
    elif fmt == FMT_F16F8:
        pc.cout_pad = pc.Cout
        fold = fold_scale and pc.scale is not None
        _pack_f16f8(pc, w, pc.scale if fold else None, stem=False)
        if fold:
            pc.scale = None
    else:
        pc.cout_pad = pc.Cout
        pc.w_hi = torch.empty((pc.Cout, pc.R, pc.S, pc.Cin), dtype=torch.bfloat16, device=dev)
        pc.w_lo = torch.empty_like(pc.w_hi) if fmt == FMT_BF16X2 else None
        fold = fold_scale and pc.scale is not None
        check(L.mpn_pack_filter_bf16_scaled(_ptr(w), _ptr(pc.scale) if fold else None, _ptr(pc.w_hi), _ptr(pc.w_lo), pc.Cout, pc.Cin,
                                            pc.R, pc.S, _stream()), "mpn_pack_filter_bf16_scaled")
        if fold:
            pc.scale = None
    return pc


def _pack_f16f8(pc, w, scale, stem):
    """FMT_F16F8 filter planes: the filter (times the folded BN scale) is prescaled by the power of two 2^k that brings its
    largest magnitude into [2^14, 2^15) -- exact, undone by acc_scale = 2^-k in the epilogue -- so that the fp16 hi plane,
    the e4m3 residual plane and the e4m3 copy (times 2^-12, pairing with the activations' 2^12-scaled residual) all sit
    in range and the three products share one accumulator."""
    L = _lib.lib()
    amax = torch.zeros(1, dtype=torch.float32, device=w.device)
    per = w[0].numel()
    check(L.mpn_filter_absmax(_ptr(w), _ptr(scale), pc.Cout, per, _ptr(amax), _stream()), "mpn_filter_absmax")
    a = float(amax.item())
    k = 0 if not (a > 0.0 and math.isfinite(a)) else max(-100, min(100, 14 - math.frexp(a)[1] + 1))   # a * 2^k in [2^14, 2^15)
    shape = (pc.Cout, 4, 1, 64) if stem else (pc.Cout, pc.R, pc.S, pc.Cin)
    pc.w_hi = torch.empty(shape, dtype=torch.float16, device=w.device)
    lay16 = bool(getattr(pc, "in_no_h8", False))          # [lo16 fp16 plane][h8 plane] instead of [lo8 plane][h8 plane]
    Cout, Cin, R, S = w.shape
    pc.w_merged = W_MERGE and not lay16 and (stem or (Cin * R * S) % 64 == 0)   # [Cout][K/64][64 lo8 | 64 h8]
    pc.w_lo = torch.empty((3 if lay16 else 2,) + shape, dtype=torch.uint8, device=w.device)
    check(L.mpn_pack_filter_f16f8(_ptr(w), _ptr(scale), float(2.0 ** k), _ptr(pc.w_hi), _ptr(pc.w_lo), Cout, Cin, R, S,
                                  int(stem) | (2 if lay16 else 0) | (4 if pc.w_merged else 0), _stream()), "mpn_pack_filter_f16f8")
    pc.acc_scale = float(2.0 ** -k)


def conv2d(x, pc, stride=1, pad=0, relu=False, sigmoid=False, residual=None, up=None, out=None, out_mode=OUT_ACT,
           out_coffset=0, out_rep=1, out_tensor=None, out_elem_offset=0, out_cstride=None, out_nstride=0, f32_input=False,
           want_h8=None, gather=None, derive=None):
    """y = epilogue(conv2d(x, w)).  Returns the output Act (OUT_ACT) or the fp32 tensor written.

    out          : existing Act to write into (concat buffers), else a new one is allocated
    out_tensor   : fp32 tensor for OUT_F32_* (allocated if None); out_elem_offset shifts the base pointer
    f32_input    : x is an fp32 NHWC Act while the output/epilogue use pc.fmt (stem)
    want_h8      : FMT_F16F8 activation outputs: False stores the tensor without its e5m2 copy plane (for consumers that derive
                   it or are packed with in_no_h8); None = only when the kernels do not derive it (DERIVE_H8)
    derive       : FMT_F16F8 input: True = the kernel derives the copy plane in shared memory, False = TMA loads the stored one,
                   None = automatic (_input_flags)
    gather       : up to two (Act, shift) phase-class addends (mpn_conv_desc.gat_*): Act is the 9*Cout-channel class
                   convolution of a map 2^shift times smaller than the output (phase_class_filter below)
    """
    L = _lib.lib()
    fmt = pc.fmt
    d = ConvDesc()
    d.N, d.H, d.W, d.Cin = x.N, x.H, x.W, pc.Cin
    assert x.C == pc.Cin, (x.C, pc.Cin)
    d.Cout, d.R, d.S, d.stride, d.pad = pc.Cout, pc.R, pc.S, stride, pad
    d.OH = (x.H + 2 * pad - pc.R) // stride + 1
    d.OW = (x.W + 2 * pad - pc.S) // stride + 1
    d.fmt = fmt
    d.in_cstride = x.cstride
    d.in_wpitch, d.k_overlap = x.wpitch, x.k_overlap
    d.flags = (EPI_RELU if relu else 0) | (EPI_SIGMOID if sigmoid else 0)
    if fmt == FMT_F16F8 and not f32_input:
        d.flags |= _input_flags(pc, x, derive)
    d.out_mode, d.out_rep, d.out_coffset = out_mode, out_rep, out_coffset
    d.w_cout_pad = pc.cout_pad
    d.acc_scale = getattr(pc, "acc_scale", 0.0)
    p = ConvPtrs()
    p.x_hi, p.x_lo = _ptr(x.hi), _ptr(x.lo)
    p.w_hi, p.w_lo = _ptr(pc.w_hi), _ptr(pc.w_lo)
    p.scale, p.bias = _ptr(pc.scale), _ptr(pc.bias)
    if residual is not None:
        assert (residual.N, residual.H, residual.W, residual.C) == (x.N, d.OH, d.OW, pc.Cout) and residual.fmt == fmt
        d.res_cstride = residual.cstride
        p.res_hi, p.res_lo = _ptr(residual.hi), _ptr(residual.lo)
    if up is not None:
        assert up.N == x.N and up.C == pc.Cout and up.fmt == fmt
        d.up_h, d.up_w, d.up_cstride = up.H, up.W, up.cstride
        p.up_hi, p.up_lo = _ptr(up.hi), _ptr(up.lo)
    keep = []
    if gather:
        assert len(gather) <= 2 and out_mode == OUT_ACT and out_rep == 1
        d.gat_n = len(gather)
        for i, (g, sh) in enumerate(gather):
            assert g.fmt == fmt and g.N == x.N and g.C == 9 * pc.Cout and (g.H << sh, g.W << sh) == (d.OH, d.OW), (g.H, g.W, g.C, sh)
            d.gat_shift[i], d.gat_h[i], d.gat_w[i], d.gat_cstride[i] = sh, g.H, g.W, g.cstride
            p.gat_hi[i], p.gat_lo[i] = g.hi.data_ptr(), g.lo.data_ptr() if g.lo is not None else None
            keep.append(g)
    ret = None
    if out_mode == OUT_ACT:
        if out is None:
            out = Act(fmt, x.N, d.OH * out_rep, d.OW * out_rep, pc.Cout, x.hi.device, has_h8=_auto_h8(want_h8))
        assert out.fmt == fmt and out.N == x.N and out.H == d.OH * out_rep and out.W == d.OW * out_rep
        if fmt == FMT_F16F8 and not out.has_h8:
            d.flags |= EPI_NO_H8
        d.out_cstride = out.cstride
        p.y_hi, p.y_lo = _ptr(out.hi), _ptr(out.lo)
        ret = out
    else:
        if out_tensor is None:
            shape = ((x.N, pc.Cout, d.OH * out_rep, d.OW * out_rep) if out_mode == OUT_F32_NCHW
                     else (x.N, d.OH * out_rep, d.OW * out_rep, pc.Cout))
            out_tensor = torch.empty(shape, dtype=torch.float32, device=x.hi.device)
        assert out_tensor.dtype == torch.float32 and out_tensor.is_contiguous()
        d.out_cstride = pc.Cout if out_cstride is None else out_cstride
        d.out_nstride = out_nstride
        p.y_hi = ctypes.c_void_p(out_tensor.data_ptr() + 4 * out_elem_offset)
        ret = out_tensor
    fn = L.mpn_conv2d_fwd_f32in if f32_input else L.mpn_conv2d_fwd
    ev = stats["conv_events"]
    if ev is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(fn(ctypes.byref(d), ctypes.byref(p), _stream()), "mpn_conv2d_fwd")
    if ev is not None:
        e1.record()
        slots = {FMT_BF16: 1.0, FMT_BF16X2: 3.0, FMT_F16F8: 2.5 if (d.flags & IN_NO_H8) else 2.0}.get(fmt, 0.0)  # bf16-equivalent MMA passes
        ev.append((e0, e1, 2.0 * x.N * d.OH * d.OW * pc.Cout * pc.Cin * pc.R * pc.S, bool(f32_input or fmt == FMT_F32), slots))
    stats["launches"] += 1
    if stats.get("range_check") and fmt == FMT_F16F8 and out_mode == OUT_ACT and not torch.cuda.is_current_stream_capturing():
        amax = float(ret.hi.abs().max())  # debug mode: one reduction + host sync per conv
        if not amax < F16_MAX:
            raise _lib.MpnError("f16f8 range guard: a conv output (%dx%d, %d -> %d channels, %dx%d filter) reached |x| >= 65504 and was "
                                "clamped by the saturating fp16 store; run this model with precision='bf16x3'"
                                % (d.OH, d.OW, pc.Cin, pc.Cout, pc.R, pc.S))
    return ret


def phase_class_filter(weight):
    """3x3 pad-1 filter W [Cout, Cin, 3, 3] applied to a nearest-upsampled map == per output phase one of nine 3x3 pad-1
    filters on the low-resolution map (mpn_conv_desc.gat_*).  Returns V [9*Cout, Cin, 3, 3], class-major, class = rc*3 + cc:
    along one axis, class 0 (first row of a block) sends tap 0 to the previous low-resolution row and taps 1,2 to the same
    row; class 1 (inside) sends all three taps to the same row; class 2 (last row) sends taps 0,1 to the same row and tap 2
    to the next one.  Holds for every upsampling factor >= 2 (factor 2 has no class 1)."""
    w = weight.detach().double()
    assert w.dim() == 4 and w.shape[2:] == (3, 3)
    M = torch.zeros(3, 3, 3, dtype=torch.float64, device=w.device)   # [class][low-res tap i][tap r]
    M[0, 0, 0] = M[0, 1, 1] = M[0, 1, 2] = 1
    M[1, 1, :] = 1
    M[2, 1, 0] = M[2, 1, 1] = M[2, 2, 2] = 1
    v = torch.einsum("air,bjs,ocrs->abocij", M, M, w)
    return v.reshape(9 * w.shape[0], w.shape[1], 3, 3).float().contiguous()


def conv2d_multi(xs, pc, pad=0, relu=False, sigmoid=False, out_mode=OUT_ACT, out_tensor=None, out_elem_offsets=None,
                 out_cstride=None, out_nstride=0, want_h8=None, derive=None):
    """The same stride-1 convolution (shared packed filter) applied to several Acts in ONE persistent launch
    (mpn_conv2d_fwd_multi): a RetinaNet tower layer over the pyramid levels.  OUT_ACT: returns the list of output Acts;
    fp32 modes: every level writes at out_elem_offsets[i] of out_tensor (strides as in conv2d)."""
    L = _lib.lib()
    n = len(xs)
    fmt = pc.fmt
    D, Pp = (ConvDesc * n)(), (ConvPtrs * n)()
    outs = []
    flops = 0.0
    for i, x in enumerate(xs):
        d, p = D[i], Pp[i]
        assert x.C == pc.Cin and x.fmt == fmt and not x.wpitch and not x.k_overlap
        d.N, d.H, d.W, d.Cin = x.N, x.H, x.W, pc.Cin
        d.Cout, d.R, d.S, d.stride, d.pad = pc.Cout, pc.R, pc.S, 1, pad
        d.OH, d.OW = x.H + 2 * pad - pc.R + 1, x.W + 2 * pad - pc.S + 1
        d.fmt, d.in_cstride = fmt, x.cstride
        d.flags = (EPI_RELU if relu else 0) | (EPI_SIGMOID if sigmoid else 0)
        if fmt == FMT_F16F8:
            d.flags |= _input_flags(pc, x, derive)
        d.out_mode, d.out_rep, d.out_coffset = out_mode, 1, 0
        d.w_cout_pad, d.acc_scale = pc.cout_pad, getattr(pc, "acc_scale", 0.0)
        p.x_hi, p.x_lo = _ptr(x.hi), _ptr(x.lo)
        p.w_hi, p.w_lo, p.scale, p.bias = _ptr(pc.w_hi), _ptr(pc.w_lo), _ptr(pc.scale), _ptr(pc.bias)
        if out_mode == OUT_ACT:
            o = Act(fmt, x.N, d.OH, d.OW, pc.Cout, x.hi.device, has_h8=_auto_h8(want_h8))
            d.out_cstride = o.cstride
            if fmt == FMT_F16F8 and not o.has_h8:
                d.flags |= EPI_NO_H8
            p.y_hi, p.y_lo = _ptr(o.hi), _ptr(o.lo)
            outs.append(o)
        else:
            assert out_tensor is not None and out_tensor.dtype == torch.float32 and out_tensor.is_contiguous() and out_nstride > 0
            d.out_cstride = pc.Cout if out_cstride is None else out_cstride
            d.out_nstride = out_nstride
            p.y_hi = ctypes.c_void_p(out_tensor.data_ptr() + 4 * out_elem_offsets[i])
        flops += 2.0 * x.N * d.OH * d.OW * pc.Cout * pc.Cin * pc.R * pc.S
    ev = stats["conv_events"]
    if ev is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(L.mpn_conv2d_fwd_multi(D, Pp, n, _stream()), "mpn_conv2d_fwd_multi")
    if ev is not None:
        e1.record()
        slots = {FMT_BF16: 1.0, FMT_BF16X2: 3.0, FMT_F16F8: 2.5 if (D[0].flags & IN_NO_H8) else 2.0}.get(fmt, 0.0)
        ev.append((e0, e1, flops, False, slots))
    stats["launches"] += 1
    return outs if out_mode == OUT_ACT else out_tensor


def stem_pack_input(img, fmt, want_h8=None):
    """fp32 NCHW image -> zero-padded space-to-depth Act [N, H/2+3, W/2 (+3 pitch), 64-wide windows of 16 ch]
    (mpn_stem_pack_input); with pack_stem_filter the 7x7/2 stem becomes a tcgen05 conv (R=4, S=1, Cin=64)."""
    assert img.is_cuda and img.dtype == torch.float32 and img.dim() == 4 and img.shape[1] == 3
    img = img.contiguous()
    N, _, H, W = img.shape
    H2, W2 = (H + 1) // 2, (W + 1) // 2
    a = Act(fmt, N, H2 + 3, W2, 64, img.device, cstride=16, wpitch=W2 + 3, k_overlap=1, has_h8=_auto_h8(want_h8))
    check(_lib.lib().mpn_stem_pack_input(_ptr(img), _ptr(a.hi), _ptr(a.lo), N, H, W, fmt, 0 if a.has_h8 else EPI_NO_H8, _stream()),
          "mpn_stem_pack_input")
    stats["launches"] += 1
    return a


def resnet_preprocess_u8(img_u8):
    """uint8 [N,H,W,3] BGR (cuda) -> fp32 [N,3,H,W], bit-identical to datasets/coco_data/preprocessing.py:15-26."""
    assert img_u8.is_cuda and img_u8.dtype == torch.uint8 and img_u8.dim() == 4 and img_u8.shape[3] == 3
    img_u8 = img_u8.contiguous()
    N, H, W, _ = img_u8.shape
    out = torch.empty((N, 3, H, W), dtype=torch.float32, device=img_u8.device)
    check(_lib.lib().mpn_preprocess_u8_nchw(_ptr(img_u8), _ptr(out), N, H, W, _stream()), "mpn_preprocess_u8_nchw")
    stats["launches"] += 1
    return out


def stem_pack_input_u8(img_u8, fmt, want_h8=None):
    """uint8 [N,H,W,3] BGR image -> the tensor-core stem operand with resnet_preprocess fused in."""
    assert img_u8.is_cuda and img_u8.dtype == torch.uint8 and img_u8.dim() == 4 and img_u8.shape[3] == 3
    img_u8 = img_u8.contiguous()
    N, H, W, _ = img_u8.shape
    H2, W2 = (H + 1) // 2, (W + 1) // 2
    a = Act(fmt, N, H2 + 3, W2, 64, img_u8.device, cstride=16, wpitch=W2 + 3, k_overlap=1, has_h8=_auto_h8(want_h8))
    check(_lib.lib().mpn_stem_pack_input_u8(_ptr(img_u8), _ptr(a.hi), _ptr(a.lo), N, H, W, fmt, 0 if a.has_h8 else EPI_NO_H8, _stream()),
          "mpn_stem_pack_input_u8")
    stats["launches"] += 1
    return a


def pack_stem_filter(weight, bn, fmt, in_no_h8=False):
    L = _lib.lib()
    w = weight.detach().contiguous()
    assert tuple(w.shape[1:]) == (3, 7, 7)
    pc = PackedConv()
    pc.Cout, pc.Cin, pc.R, pc.S, pc.fmt, pc.cout_pad = w.shape[0], 64, 4, 1, fmt, w.shape[0]
    pc.acc_scale = 0.0
    pc.in_no_h8 = bool(in_no_h8) and fmt == FMT_F16F8
    if fmt == FMT_F16F8:
        _pack_f16f8(pc, w, None, stem=True)   # the BN scale stays in the epilogue, as for the bf16 stem
    else:
        pc.w_hi = torch.empty((pc.Cout, 4, 1, 64), dtype=torch.bfloat16, device=w.device)
        pc.w_lo = torch.empty_like(pc.w_hi) if fmt == FMT_BF16X2 else None
        check(L.mpn_stem_pack_filter(_ptr(w), _ptr(pc.w_hi), _ptr(pc.w_lo), pc.Cout, _stream()), "mpn_stem_pack_filter")
    gamma, beta, mean, var, eps = bn
    npad = (pc.Cout + 63) // 64 * 64
    pc.scale = torch.zeros(npad, dtype=torch.float32, device=w.device)
    pc.bias = torch.zeros(npad, dtype=torch.float32, device=w.device)
    check(L.mpn_fold_bn(_ptr(gamma.detach().contiguous()), _ptr(beta.detach().contiguous()), _ptr(mean.contiguous()),
                        _ptr(var.contiguous()), float(eps), _ptr(pc.scale), _ptr(pc.bias), pc.Cout, _stream()), "mpn_fold_bn")
    return pc


def add_softmax_rows(a, res):
    """softmax(a + res, dim=1): a = Act viewed as [P, D(+pad)] rows, res fp32 [P, D]."""
    P, D = res.shape
    out = torch.empty((P, D), dtype=torch.float32, device=res.device)
    check(_lib.lib().mpn_add_softmax_rows(_ptr(a.hi), _ptr(a.lo), _ptr(res.contiguous()), _ptr(out), P, D, a.cstride, a.fmt, _stream()),
          "mpn_add_softmax_rows")
    stats["launches"] += 1
    return out


def maxpool3x3s2(x, want_h8=None):
    OH, OW = (x.H + 2 - 3) // 2 + 1, (x.W + 2 - 3) // 2 + 1
    want_h8 = _auto_h8(want_h8)
    if x.fmt == FMT_F16F8 and x.C % 8:
        want_h8 = True   # the generic (non-vectorised) kernel keeps the full format
    y = Act(x.fmt, x.N, OH, OW, x.C, x.hi.device, has_h8=want_h8)
    assert x.cstride == x.C
    check(_lib.lib().mpn_maxpool3x3s2(_ptr(x.hi), _ptr(x.lo), _ptr(y.hi), _ptr(y.lo), x.N, x.H, x.W, x.C, x.fmt,
                                      0 if y.has_h8 else EPI_NO_H8, _stream()), "mpn_maxpool3x3s2")
    stats["launches"] += 1
    return y


def relu(x):
    y = Act(x.fmt, x.N, x.H, x.W, x.C, x.hi.device, cstride=x.cstride)
    check(_lib.lib().mpn_relu(_ptr(x.hi), _ptr(x.lo), _ptr(y.hi), _ptr(y.lo), x.hi.numel(), x.fmt, _stream()), "mpn_relu")
    stats["launches"] += 1
    return y


_ANCHOR_CACHE = {}


def anchors_for(H, W, device):
    """[1,A,4] fp32 anchors (bit-identical to network/anchors.py), generated once per (H, W, device)."""
    key = (int(H), int(W), str(device))
    t = _ANCHOR_CACHE.get(key)
    if t is None:
        L = _lib.lib()
        A = L.mpn_num_anchors(int(H), int(W))
        host = torch.empty((1, A, 4), dtype=torch.float32)
        check(L.mpn_generate_anchors(int(H), int(W), ctypes.c_void_p(host.data_ptr())), "mpn_generate_anchors")
        t = host.to(device)
        _ANCHOR_CACHE[key] = t
    return t


def decode_clip(anchors, reg, H, W):
    B, A = reg.shape[0], reg.shape[1]
    boxes = torch.empty((B, A, 4), dtype=torch.float32, device=reg.device)
    check(_lib.lib().mpn_decode_clip(_ptr(anchors), _ptr(reg), _ptr(boxes), B, A, int(H), int(W), _stream()), "mpn_decode_clip")
    stats["launches"] += 1
    return boxes


class Detections(object):
    __slots__ = ("cand_idx", "cand_cnt", "keep_idx", "keep_cnt", "scores", "boxes", "max_cand")


def filter_sort_nms(cls, boxes, score_thresh=0.05, iou_thresh=0.5, ge=False, max_cand=4096, stage_ms=None):
    """cls [B,A,1] or [B,A] fp32, boxes [B,A,4] fp32 -> Detections (device tensors, no host sync).
    stage_ms: a list -> the blocking profiling twin runs instead and appends the five stage times (ms): filter, sort, gather,
    mask, reduce (bench.py roofline_aux)."""
    L = _lib.lib()
    B, A = boxes.shape[0], boxes.shape[1]
    max_cand = int(min(max(64, max_cand), A))
    dev = boxes.device
    det = Detections()
    det.max_cand = max_cand
    det.cand_idx = torch.empty((B, max_cand), dtype=torch.int32, device=dev)
    det.cand_cnt = torch.empty((B,), dtype=torch.int32, device=dev)
    det.keep_idx = torch.empty((B, max_cand), dtype=torch.int64, device=dev)
    det.keep_cnt = torch.empty((B,), dtype=torch.int32, device=dev)
    det.scores = torch.empty((B, max_cand), dtype=torch.float32, device=dev)
    det.boxes = torch.empty((B, max_cand, 4), dtype=torch.float32, device=dev)
    ws_bytes = L.mpn_detect_workspace_bytes(B, A, max_cand)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    if stage_ms is not None:
        ms = (ctypes.c_float * 5)()
        check(L.mpn_filter_sort_nms_profile(_ptr(cls.contiguous()), _ptr(boxes.contiguous()), B, A, float(score_thresh), float(iou_thresh),
                                            int(bool(ge)), max_cand, _ptr(det.cand_idx), _ptr(det.cand_cnt), _ptr(det.keep_idx),
                                            _ptr(det.keep_cnt), _ptr(det.scores), _ptr(det.boxes), _ptr(ws), ws_bytes, _stream(),
                                            ctypes.cast(ms, ctypes.c_void_p)), "mpn_filter_sort_nms_profile")
        stage_ms.append([float(v) for v in ms])
    else:
        check(L.mpn_filter_sort_nms(_ptr(cls.contiguous()), _ptr(boxes.contiguous()), B, A, float(score_thresh), float(iou_thresh),
                                    int(bool(ge)), max_cand, _ptr(det.cand_idx), _ptr(det.cand_cnt), _ptr(det.keep_idx),
                                    _ptr(det.keep_cnt), _ptr(det.scores), _ptr(det.boxes), _ptr(ws), ws_bytes, _stream()),
              "mpn_filter_sort_nms")
    stats["launches"] += 6  # filter/compact, segments, radix sort (>=1), gather, mask, reduce
    return det


def nms(dets, thresh, ge=False):
    """dets [n,5] fp32 cuda -> int64 [K] indices into dets, descending score (pth_nms semantics)."""
    L = _lib.lib()
    assert dets.is_cuda and dets.dtype == torch.float32 and dets.dim() == 2 and dets.shape[1] == 5
    n = dets.shape[0]
    if n == 0:
        return torch.zeros((0,), dtype=torch.int64, device=dets.device)
    dets = dets.contiguous()
    keep = torch.empty((n,), dtype=torch.int64, device=dets.device)
    num = torch.zeros((1,), dtype=torch.int32, device=dets.device)
    ws_bytes = L.mpn_nms_workspace_bytes(n)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dets.device)
    check(L.mpn_nms(_ptr(dets), n, float(thresh), int(bool(ge)), _ptr(keep), _ptr(num), _ptr(ws), ws_bytes, _stream()), "mpn_nms")
    k = int(num.item())
    return keep[:k]


def nms_mask(sorted_dets, thresh, ge=False):
    n = sorted_dets.shape[0]
    cb = (n + 63) // 64
    mask = torch.zeros((n, cb), dtype=torch.int64, device=sorted_dets.device)
    check(_lib.lib().mpn_nms_mask(_ptr(sorted_dets.contiguous()), n, float(thresh), int(bool(ge)), _ptr(mask), _stream()), "mpn_nms_mask")
    return mask


def heatmap_peaks(heat, thre1=0.1, factor=4, max_peaks=1024, channels=18):
    """joint_utils.NMS + get_joint_list on the device (network/joint_utils.py:61-152, tester.py:215-221).

    heat: CUDA fp32 [B, C>=channels, H, W] heat maps.  Returns (rows fp32 [B, max_peaks, 5] = (x, y, score, id, joint_type)
    in the reference's order, count int32 [B]); coordinates are in the factor-times-upsampled grid (multiply by the
    caller's image scale, joint_utils.py:146-147)."""
    assert heat.is_cuda and heat.dtype == torch.float32 and heat.dim() == 4 and heat.shape[1] >= channels
    heat = heat.contiguous()
    B, C, H, W = heat.shape
    L = _lib.lib()
    rows = torch.zeros((B, max_peaks, 5), dtype=torch.float32, device=heat.device)
    count = torch.zeros((B,), dtype=torch.int32, device=heat.device)
    wsb = L.mpn_heatmap_peaks_workspace_bytes(B, channels)
    ws = torch.empty((max(wsb, 4),), dtype=torch.uint8, device=heat.device)
    check(L.mpn_heatmap_peaks(_ptr(heat), B, channels, H, W, C * H * W, float(thre1), int(factor), _ptr(rows), max_peaks, _ptr(count),
                              _ptr(ws), wsb, _stream()), "mpn_heatmap_peaks")
    stats["launches"] += 2
    return rows, count


# ------------------------------------------------------------------ PRN assignment (evaluate/tester.py:333-513)
def prn_workspace(P, n_peaks, kmax, device):
    wsb = _lib.lib().mpn_prn_workspace_bytes(P, n_peaks, kmax)
    return torch.empty((max(int(wsb), 8),), dtype=torch.uint8, device=device), int(wsb)


def prn_build_inputs(peak_xy, peak_type, peak_img_start, boxes_xywh, box_img, grid_hw, in_thres, gauss_w, kmax, workspace):
    """tester.py:363-403 on the device for every box of the batch.  peak_xy f64 [n,2], peak_type i32 [n], peak_img_start
    i32 [B+1], boxes_xywh f64 [P,4], box_img i32 [P] (CUDA tensors); gauss_w: 5 host doubles.  Returns (owner i32
    [P,17,h,w], inp fp32 [P,h,w,17])."""
    gh, gw = grid_hw
    P, n = boxes_xywh.shape[0], peak_type.shape[0]
    assert boxes_xywh.is_cuda and boxes_xywh.dtype == torch.float64 and box_img.dtype == torch.int32
    assert peak_xy.dtype == torch.float64 and peak_type.dtype == torch.int32 and peak_img_start.dtype == torch.int32
    owner = torch.empty((P, 17, gh, gw), dtype=torch.int32, device=boxes_xywh.device)
    inp = torch.empty((P, gh, gw, 17), dtype=torch.float32, device=boxes_xywh.device)
    gw_arr = (ctypes.c_double * 5)(*[float(v) for v in gauss_w])
    ws, wsb = workspace
    check(_lib.lib().mpn_prn_build_inputs(_ptr(peak_xy), _ptr(peak_type), _ptr(peak_img_start), n, _ptr(boxes_xywh), _ptr(box_img), P,
                                          gh, gw, float(in_thres), ctypes.cast(gw_arr, ctypes.c_void_p), _ptr(owner), _ptr(inp), _ptr(ws), wsb,
                                          int(kmax), _stream()), "mpn_prn_build_inputs")
    stats["launches"] += 2
    return owner, inp


def prn_assign(peak_xy, peak_img_start, joint_start, boxes_xywh, box_img, box_img_start, owner, output, kmax, workspace):
    """tester.py:412-483 on the device.  output: fp32 [P,h,w,17] PRN output.  Returns bbox_keypoints f64 [P,17,3]."""
    P, _, gh, gw = owner.shape
    B = box_img_start.shape[0] - 1
    assert output.is_cuda and output.dtype == torch.float32 and tuple(output.shape) == (P, gh, gw, 17)
    assert joint_start.dtype == torch.int32 and tuple(joint_start.shape) == (B, 18) and box_img_start.dtype == torch.int32
    output = output.contiguous()
    res = torch.empty((P, 17, 3), dtype=torch.float64, device=owner.device)
    ws, wsb = workspace
    check(_lib.lib().mpn_prn_assign(_ptr(peak_xy), _ptr(peak_img_start), _ptr(joint_start), peak_xy.shape[0], _ptr(boxes_xywh), _ptr(box_img),
                                    _ptr(box_img_start), P, B, gh, gw, _ptr(owner), _ptr(output), _ptr(res), _ptr(ws), wsb, int(kmax),
                                    _stream()), "mpn_prn_assign")
    stats["launches"] += 3
    return res


# ------------------------------------------------------------------ multi-scale TTA (evaluate/tester.py:298-304, 316-331)
def resize_cubic(src, sh, sw, dh, dw, scale_x, scale_y, dst=None, out_f64=False, div=1.0, mirror=False, plane_map=None):
    """cv2.resize(INTER_CUBIC) of the [sh, sw] top-left region of every plane of `src` (fp32 [..., Hs, Ws], leading dims = planes).
    dst None -> new fp32 [planes, dh, dw]; out_f64: dst (fp64 [planes, dh, dw]) += resized / div, optionally mirrored in x and
    written to plane plane_map[p]."""
    assert src.is_cuda and src.dtype == torch.float32 and src.is_contiguous()
    Hs, Ws = src.shape[-2], src.shape[-1]
    planes = src.numel() // (Hs * Ws)
    assert 0 < sh <= Hs and 0 < sw <= Ws
    L = _lib.lib()
    if out_f64:
        assert dst is not None and dst.dtype == torch.float64 and dst.is_contiguous() and dst.shape[-2:] == (dh, dw)
    else:
        dst = torch.empty((planes, dh, dw), dtype=torch.float32, device=src.device)
    wsb = L.mpn_resize_cubic_workspace_bytes(dh, dw)
    ws = torch.empty((wsb,), dtype=torch.uint8, device=src.device)
    check(L.mpn_resize_cubic(_ptr(src), Hs * Ws, Ws, sh, sw, _ptr(dst), dh * dw, dw, dh, dw, planes, float(scale_x), float(scale_y),
                             1 if out_f64 else 0, float(div), int(bool(mirror)), _ptr(plane_map), _ptr(ws), wsb, _stream()), "mpn_resize_cubic")
    stats["launches"] += 3
    return dst


def tta_combine(normal, flipped=None, want_f32=True):
    """(normal + flipped) / 2 in float64 (tester.py:329), returned as (fp64, fp32 or None)."""
    assert normal.dtype == torch.float64 and normal.is_contiguous()
    out = torch.empty_like(normal)
    out32 = torch.empty(normal.shape, dtype=torch.float32, device=normal.device) if want_f32 else None
    check(_lib.lib().mpn_tta_combine(_ptr(normal), _ptr(flipped), _ptr(out), _ptr(out32), normal.numel(), _stream()), "mpn_tta_combine")
    stats["launches"] += 1
    return out, out32


# ------------------------------------------------------------------ detection-subnet training loss (network/losses.py:5-137)
def focal_loss(cls, reg, anchors, annotations, want_grads=True, gscale_cls=1.0, gscale_reg=1.0):
    """cls [B,A,C] (after the sigmoid), reg [B,A,4], anchors [1,A,4] or [A,4], annotations [B,M,5] (class -1 = padding), CUDA fp32.
    Returns (cls_loss [B], reg_loss [B], dcls or None, dreg or None): the per-image losses of FocalLoss.forward and the gradients
    of gscale_cls * mean(cls_loss) + gscale_reg * mean(reg_loss)."""
    assert cls.is_cuda and cls.dtype == torch.float32 and reg.dtype == torch.float32 and annotations.dtype == torch.float32
    cls, reg, annotations = cls.contiguous(), reg.contiguous(), annotations.contiguous()
    anchors = anchors.reshape(-1, 4).contiguous()
    B, A, C = cls.shape
    M = annotations.shape[1]
    assert reg.shape == (B, A, 4) and anchors.shape[0] == A and annotations.shape == (B, M, 5)
    L = _lib.lib()
    dev = cls.device
    cl = torch.empty((B,), dtype=torch.float32, device=dev)
    rl = torch.empty((B,), dtype=torch.float32, device=dev)
    dcls = torch.empty_like(cls) if want_grads else None
    dreg = torch.empty_like(reg) if want_grads else None
    wsb = L.mpn_focal_loss_workspace_bytes(B, A)
    ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
    check(L.mpn_focal_loss(_ptr(cls), _ptr(reg), _ptr(anchors), _ptr(annotations), B, A, C, M, _ptr(cl), _ptr(rl), _ptr(dcls), _ptr(dreg),
                           float(gscale_cls), float(gscale_reg), _ptr(ws), wsb, _stream()), "mpn_focal_loss")
    stats["launches"] += 2
    return cl, rl, dcls, dreg
