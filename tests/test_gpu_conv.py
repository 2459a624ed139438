"""GPU parity of the conv primitive (CUDA-core fp32 and tcgen05 bf16 / bf16x3) vs torch fp32 conv2d."""
import pytest
import torch

pytestmark = pytest.mark.gpu

# (N, H, W, Cin, Cout, R, stride, pad, kwargs)
SIMT_CASES = [
    (2, 17, 23, 3, 64, 7, 2, 3, dict(bn=True, relu=True, bias=False)),           # stem
    (2, 12, 20, 64, 64, 1, 1, 0, dict(bn=True, relu=True, bias=False)),
    (1, 13, 9, 64, 128, 3, 2, 1, dict(bn=True, relu=True, bias=False)),
    (2, 8, 8, 128, 256, 1, 1, 0, dict(bn=True, residual=True, relu=True, bias=False)),
    (2, 10, 14, 96, 64, 1, 1, 0, dict(up=True)),
    (1, 7, 9, 32, 18, 1, 1, 0, dict(out_mode=2, rep=4)),
    (2, 6, 5, 64, 36, 3, 1, 1, dict(out_mode=1)),
    (2, 6, 5, 64, 9, 3, 1, 1, dict(out_mode=1, sigmoid=True)),
    (1, 5, 6, 64, 128, 3, 1, 1, dict(rep=2, coffset=128, ctotal=512)),
]

TC_CASES = [
    (2, 16, 32, 64, 64, 1, 1, 0, dict(bn=True, relu=True, bias=False)),
    (2, 16, 32, 64, 64, 3, 1, 1, dict(bn=True, relu=True, bias=False)),
    (1, 30, 40, 256, 256, 3, 1, 1, dict()),
    (2, 15, 20, 512, 128, 1, 1, 0, dict(bn=True, relu=True, bias=False)),
    (2, 30, 40, 128, 128, 3, 2, 1, dict(bn=True, relu=True, bias=False)),          # stride-2 phase views
    (2, 15, 20, 256, 512, 1, 2, 0, dict(bn=True, bias=False)),                      # projection shortcut
    (2, 15, 21, 128, 128, 3, 2, 1, dict()),                                         # odd sizes, stride 2
    (2, 8, 8, 128, 256, 1, 1, 0, dict(bn=True, residual=True, relu=True, bias=False)),
    (2, 10, 14, 128, 256, 1, 1, 0, dict(up=True)),
    (1, 16, 24, 256, 18, 1, 1, 0, dict(out_mode=2)),
    (1, 4, 6, 256, 19, 1, 1, 0, dict(out_mode=2, rep=4)),
    (2, 6, 5, 256, 36, 3, 1, 1, dict(out_mode=1)),
    (2, 6, 5, 256, 9, 3, 1, 1, dict(out_mode=1, sigmoid=True)),
    (1, 5, 6, 128, 128, 3, 1, 1, dict(rep=2, coffset=128, ctotal=512)),
    (3, 2, 3, 2048, 256, 3, 2, 1, dict()),                                          # conv6 on a tiny map
    (3, 1, 2, 256, 256, 3, 2, 1, dict()),                                           # conv7: empty phase views
    (2, 60, 80, 64, 256, 1, 1, 0, dict(bn=True, relu=True, bias=False)),            # many tiles / persistence
    (1, 24, 32, 512, 256, 3, 1, 1, dict(relu=True)),                                # long K loop (72 k-iters)
    # tile-plan / epilogue variants of the persistent kernel
    (16, 30, 40, 64, 256, 3, 1, 1, dict(bn=True, relu=True, bias=False)),           # 160 tiles: tail cut into 32-col sub-tiles, 2 K blocks/stage
    (8, 30, 40, 64, 256, 3, 1, 1, dict(bn=True, relu=True, bias=False)),            # 128-wide plan with a 32-col tail
    (16, 30, 40, 64, 256, 1, 1, 0, dict(bn=True, residual=True, relu=True, bias=False)),  # tensor-core residual, 120-row boxes, 64-col tail
    (4, 16, 32, 128, 512, 1, 1, 0, dict(bn=True, residual=True, relu=True, bias=False)),  # tensor-core residual, two channel tiles
    (2, 8, 8, 128, 320, 1, 1, 0, dict(bn=True, residual=True, bias=False)),         # residual box partly beyond Cout
    (2, 8, 8, 128, 96, 1, 1, 0, dict(bn=True, residual=True, relu=True, bias=False)),     # Cout % 64 != 0: residual through the LSU
    (2, 8, 8, 128, 128, 1, 1, 0, dict(residual=True)),                              # residual without BN / bias (data-gradient accumulation)
    (1, 10, 12, 128, 128, 3, 1, 1, dict(coffset=128, ctotal=512)),                  # TMA store into a channel slice of a wider tensor
    # CTA-pair kernel with the shortcut loaded by TMA (f16f8 / bf16; bf16x3 under MPN_RES_MMA=0)
    (3, 30, 40, 64, 320, 1, 1, 0, dict(bn=True, residual=True, relu=True, bias=False)),   # 30 M tiles, last channel tile 64 of 256 wide
    (5, 15, 20, 128, 256, 1, 1, 0, dict(bn=True, residual=True, relu=True, bias=False)),  # 15 M tiles: the last pair has a phantom half
    (2, 30, 40, 128, 128, 3, 1, 1, dict(residual=True, relu=True)),                       # 128-wide pair tiles, 3x3 taps
    (32, 30, 40, 64, 1024, 1, 1, 0, dict(bn=True, residual=True, relu=True, bias=False)), # layer3 expansion shape: many tiles per CTA
    # exact-2x nearest-upsample add through the TMA epilogue (FPN laterals): the tile's source pixels arrive as one half-resolution box
    (2, 32, 64, 128, 256, 1, 1, 0, dict(up=True)),                                  # 64x2 boxes
    (3, 30, 40, 256, 256, 1, 1, 0, dict(up=True)),                                  # 20x6 boxes, clipped bottom rows, odd tile count
    (2, 60, 80, 128, 128, 1, 1, 0, dict(up=True, relu=True)),                       # 128-wide pair tiles
    (6, 4, 6, 64, 256, 1, 1, 0, dict(up=True)),                                     # whole images per box (TN = 5), 2 tiles
]


def _check(fmt, case, tol_rounded, tol_exact):
    from gpu_util import conv_case, nerr, no_tf32
    no_tf32()
    N, H, W, Cin, Cout, R, stride, pad, kw = case
    ours, ref_r, ref_e = conv_case(fmt, N, H, W, Cin, Cout, R, stride, pad, **kw)
    assert ours.shape == ref_e.shape
    assert torch.isfinite(ours).all()
    er, ee = nerr(ours, ref_r), nerr(ours, ref_e)
    assert er <= tol_rounded and ee <= tol_exact, "err vs rounded-operand ref %.3g (tol %.3g), vs fp32 ref %.3g (tol %.3g)" % (
        er, tol_rounded, ee, tol_exact)


@pytest.mark.parametrize("case", SIMT_CASES)
def test_conv_fp32_cuda_core(case):
    _check(0, case, 2e-5, 2e-5)


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tcgen05_bf16x3(case):
    # operands carry ~16 mantissa bits; bf16 hi/lo re-split of the output costs 2^-17
    _check(2, case, 5e-5, 1e-4)


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tcgen05_bf16(case):
    # vs the same conv on bf16-rounded operands only accumulation order + bf16 output rounding differ
    _check(1, case, 6e-3, 3e-2)


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tcgen05_f16f8(case):
    # fp16 hi*hi on kind::f16 + the two fp8 cross terms on kind::f8f6f4: product error ~2^-15, output re-split 2^-15
    # (the default operand variant: MPN_IN_DERIVE_H8 unless MPN_DERIVE_H8=0)
    _check(3, case, 2e-4, 2e-4)


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tcgen05_f16f8_stored_copy_plane(case):
    """MODE_F16F8: the e5m2 copy plane of the input is stored in HBM and loaded by TMA."""
    N, H, W, Cin, Cout, R, stride, pad, kw = case
    _check(3, (N, H, W, Cin, Cout, R, stride, pad, dict(kw, derive=False)), 2e-4, 2e-4)


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tcgen05_f16f8_derived_copy_plane(case):
    """MODE_F16F8C: tensors without the copy plane, the converter warps derive it from the fp16 tile in shared memory; agrees with
    the stored-plane variant to the rounding of the 3-bit cross term."""
    from gpu_util import conv_case, nerr, no_tf32
    no_tf32()
    N, H, W, Cin, Cout, R, stride, pad, kw = case
    ours, ref_r, ref_e = conv_case(3, N, H, W, Cin, Cout, R, stride, pad, derive=True, **kw)
    assert ours.shape == ref_e.shape and torch.isfinite(ours).all()
    assert nerr(ours, ref_e) <= 2e-4
    full, _, _ = conv_case(3, N, H, W, Cin, Cout, R, stride, pad, derive=False, **kw)
    assert nerr(ours, full) <= 2e-4


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tcgen05_f16f8_without_h8_plane(case):
    """MPN_IN_NO_H8 / MPN_EPI_NO_H8: input, shortcut, upsample source and output stored as fp16 + e5m2 residual only (3 bytes per
    element); the weight-residual term runs as fp16 x fp16 (MODE_F16F8B).  At least as accurate as the default variant."""
    from gpu_util import conv_case, nerr, no_tf32
    no_tf32()
    N, H, W, Cin, Cout, R, stride, pad, kw = case
    ours, ref_r, ref_e = conv_case(3, N, H, W, Cin, Cout, R, stride, pad, no_h8=True, **kw)
    assert ours.shape == ref_e.shape and torch.isfinite(ours).all()
    assert nerr(ours, ref_e) <= 2e-4
    full, _, _ = conv_case(3, N, H, W, Cin, Cout, R, stride, pad, **kw)
    assert nerr(ours, full) <= 2e-4


@pytest.mark.parametrize("derive", [False, True])
@pytest.mark.parametrize("case", TC_CASES[::3])
def test_conv_tcgen05_f16f8_separate_filter_planes(case, derive, monkeypatch):
    """MPN_W_MERGED off: the filter's lo8 and h8 planes as two tensors of 64-byte rows (the round-1 layout) give the same result as
    the interleaved default to the bit."""
    from gpu_util import conv_case
    from multiposenet.pytorch_b200 import ops
    N, H, W, Cin, Cout, R, stride, pad, kw = case
    merged, _, _ = conv_case(3, N, H, W, Cin, Cout, R, stride, pad, derive=derive, **kw)
    monkeypatch.setattr(ops, "W_MERGE", False)
    planar, _, ref = conv_case(3, N, H, W, Cin, Cout, R, stride, pad, derive=derive, **kw)
    assert torch.equal(merged, planar)


def test_maxpool_without_h8_plane_and_guard():
    import torch.nn.functional as F
    from gpu_util import strip_h8
    from multiposenet.pytorch_b200 import ops
    x = torch.randn(2, 64, 17, 23, generator=torch.Generator().manual_seed(3)).cuda()
    xa = ops.act_from_nchw(x, 3)
    y = ops.maxpool3x3s2(strip_h8(xa), want_h8=False)
    assert not y.has_h8 and y.lo.shape[0] == 1
    assert torch.equal(y.to_nchw(), F.max_pool2d(xa.to_nchw(), 3, 2, 1))
    w = torch.randn(64, 64, 3, 3).cuda() * 0.05
    with pytest.raises(ValueError):   # the stored-plane variant must not read a tensor without its copy plane
        ops.conv2d(y, ops.pack_conv(w, None, None, 3), pad=1, derive=False)


@pytest.mark.parametrize("fmt,tol", [(2, 1e-4), (1, 2e-2), (3, 2e-4)])
@pytest.mark.parametrize("shape", [(2, 64, 96), (1, 50, 70), (2, 33, 47)])
def test_tensor_core_stem_vs_torch(fmt, tol, shape):
    """fpn.py:99: 7x7/2 conv + BN + ReLU as a tcgen05 conv over the space-to-depth image."""
    import torch.nn.functional as F
    from gpu_util import nerr, no_tf32
    from multiposenet.pytorch_b200 import ops
    no_tf32()
    N, H, W = shape
    g = torch.Generator().manual_seed(5)
    x = torch.randn(N, 3, H, W, generator=g).cuda()
    w = (torch.randn(64, 3, 7, 7, generator=g) / 12.0).cuda()
    bn = (torch.rand(64, generator=g).cuda() + 0.5, torch.randn(64, generator=g).cuda() * 0.1,
          torch.randn(64, generator=g).cuda() * 0.1, torch.rand(64, generator=g).cuda() + 0.5, 1e-5)
    ref = F.relu(F.batch_norm(F.conv2d(x, w, None, stride=2, padding=3), bn[2], bn[3], bn[0], bn[1], False, 0.0, 1e-5))
    pc = ops.pack_stem_filter(w, bn, fmt)
    y = ops.conv2d(ops.stem_pack_input(x, fmt), pc, relu=True).to_nchw()
    assert y.shape == ref.shape
    assert nerr(y, ref) <= tol


@pytest.mark.parametrize("fmt", [0, 1, 2, 3])
@pytest.mark.parametrize("shape", [(2, 64, 17, 23), (1, 8, 6, 6), (2, 4, 9, 8), (3, 64, 32, 40)])
def test_maxpool3x3s2_vs_torch(fmt, shape):
    """fpn.py:100 max_pool2d(3, 2, 1): the maximum of representable values is representable, so every format is exact
    (C % 8 == 0 takes the vectorised bf16 / f16f8 kernels, C = 4 the generic one)."""
    import torch.nn.functional as F
    from multiposenet.pytorch_b200 import ops
    N, C, H, W = shape
    if fmt == 3 and C % 16:
        C = 16  # f16f8 byte planes need 16-channel rows
    g = torch.Generator().manual_seed(11)
    x = torch.randn(N, C, H, W, generator=g)
    x = torch.where(x.abs() < 1e-2, torch.full_like(x, 0.5), x).cuda()  # keep clear of the fp16 / e5m2 subnormal range
    xa = ops.act_from_nchw(x, fmt)
    xr = xa.to_nchw()  # the values the format actually holds
    y = ops.maxpool3x3s2(xa).to_nchw()
    ref = F.max_pool2d(xr, 3, 2, 1)
    assert y.shape == ref.shape
    assert torch.equal(y, ref)


@pytest.mark.parametrize("fmt", [3, 2, 1])
@pytest.mark.parametrize("B,sizes", [(3, [(12, 16), (6, 8), (3, 4), (2, 2), (1, 1)]), (32, [(30, 40), (15, 20), (8, 10), (4, 5)]), (2, [(60, 80), (7, 9)])])
def test_multi_level_launch_equals_per_level_launches(fmt, B, sizes):
    """mpn_conv2d_fwd_multi (one tower layer over the pyramid levels, posenet.py:262-263): the concatenated tile list must give
    exactly the tensors of the per-level launches -- activation outputs (TMA-store epilogue, CTA pairs, odd tile counts) and the
    fp32 outputs written at the levels' anchor offsets."""
    from multiposenet.pytorch_b200 import ops
    from multiposenet.pytorch_b200._lib import OUT_F32_NHWC
    g = torch.Generator().manual_seed(1)
    xs = [ops.act_from_nchw(torch.randn(B, 256, h, w, generator=g).cuda(), fmt) for h, w in sizes]
    w1 = (torch.randn(256, 256, 3, 3, generator=g) / 48.0).cuda()
    b1 = torch.randn(256, generator=g).cuda()
    pc = ops.pack_conv(w1, b1, None, fmt)
    multi = ops.conv2d_multi(xs, pc, pad=1, relu=True)
    for x, o in zip(xs, multi):
        ref = ops.conv2d(x, pc, pad=1, relu=True)
        assert torch.equal(o.hi, ref.hi) and (o.lo is None or torch.equal(o.lo, ref.lo))
    # fp32 head outputs (36 = 9 anchors x 4) into the concatenated [B, A, 4] tensor, sigmoid variant with 9 channels
    for cout, sig in ((36, False), (9, True)):
        per = cout // 9
        w2 = (torch.randn(cout, 256, 3, 3, generator=g) / 48.0).cuda()
        b2 = torch.randn(cout, generator=g).cuda()
        pc2 = ops.pack_conv(w2, b2, None, fmt)
        cells = [h * w for h, w in sizes]
        A = 9 * sum(cells)
        offs = [9 * sum(cells[:i]) * per for i in range(len(sizes))]
        out_m = torch.zeros((B, A, per), dtype=torch.float32, device="cuda")
        out_s = torch.zeros_like(out_m)
        ops.conv2d_multi(multi, pc2, pad=1, sigmoid=sig, out_mode=OUT_F32_NHWC, out_tensor=out_m, out_elem_offsets=offs, out_cstride=cout,
                         out_nstride=A * per)
        for o, off in zip(multi, offs):
            ops.conv2d(o, pc2, pad=1, sigmoid=sig, out_mode=OUT_F32_NHWC, out_tensor=out_s, out_elem_offset=off, out_cstride=cout,
                       out_nstride=A * per)
        torch.cuda.synchronize()
        assert torch.equal(out_m, out_s)
        assert float(out_m.abs().min()) > 0.0 or sig   # every anchor slot was written


@pytest.mark.parametrize("shape", [(2, 64, 96), (1, 50, 70)])
def test_tensor_core_stem_without_h8_plane(shape):
    """The stem on the no-h8 operand variant (its 64 output channels make it A-traffic bound): space-to-depth operand packed without
    the e5m2 copy plane, filter packed with the fp16 residual plane, overlapping K windows; fp32 and uint8 inputs."""
    import numpy as np
    import torch.nn.functional as F
    from gpu_util import nerr, no_tf32
    from multiposenet.pytorch_b200 import ops
    no_tf32()
    N, H, W = shape
    g = torch.Generator().manual_seed(5)
    x = torch.randn(N, 3, H, W, generator=g).cuda()
    w = (torch.randn(64, 3, 7, 7, generator=g) / 12.0).cuda()
    bn = (torch.rand(64, generator=g).cuda() + 0.5, torch.randn(64, generator=g).cuda() * 0.1,
          torch.randn(64, generator=g).cuda() * 0.1, torch.rand(64, generator=g).cuda() + 0.5, 1e-5)
    ref = F.relu(F.batch_norm(F.conv2d(x, w, None, stride=2, padding=3), bn[2], bn[3], bn[0], bn[1], False, 0.0, 1e-5))
    pc = ops.pack_stem_filter(w, bn, 3, in_no_h8=True)
    xs = ops.stem_pack_input(x, 3, want_h8=False)
    assert not xs.has_h8 and xs.lo.shape[0] == 1
    y = ops.conv2d(xs, pc, relu=True, want_h8=False)
    assert not y.has_h8 and nerr(y.to_nchw(), ref) <= 2e-4
    full = ops.conv2d(ops.stem_pack_input(x, 3, want_h8=True), ops.pack_stem_filter(w, bn, 3), relu=True, derive=False)
    assert nerr(y.to_nchw(), full.to_nchw()) <= 2e-4
    der = ops.conv2d(ops.stem_pack_input(x, 3, want_h8=False), ops.pack_stem_filter(w, bn, 3), relu=True, derive=True)   # MODE_F16F8C
    assert nerr(der.to_nchw(), full.to_nchw()) <= 2e-4 and nerr(der.to_nchw(), ref) <= 2e-4
    u8 = torch.from_numpy(np.random.Generator(np.random.PCG64(2)).integers(0, 256, (N, H, W, 3), dtype=np.uint8)).cuda()
    a, b = ops.stem_pack_input_u8(u8, 3, want_h8=False), ops.stem_pack_input_u8(u8, 3)
    assert torch.equal(a.hi, b.hi) and torch.equal(a.lo[0], b.lo[0])


@pytest.mark.parametrize("fmt,tol", [(3, 2e-4), (2, 1e-4), (1, 2e-2)])
@pytest.mark.parametrize("shape", [(2, 32, 64), (3, 24, 40), (1, 8, 8), (32, 120, 160)])
def test_phase_class_addends_equal_the_replicated_concat(fmt, tol, shape):
    """Keypoint head conv2 (posenet.py:250-256) with the x8 / x4 quarters evaluated at low resolution as phase-class convolutions and
    added in the epilogue == the same 3x3 convolution over the replicated 512-channel concat."""
    import torch.nn.functional as F
    from gpu_util import nerr, no_tf32
    from multiposenet.pytorch_b200 import ops
    no_tf32()
    N, H, W = shape
    big = N * H * W > 100000
    Cout = 256
    g = torch.Generator().manual_seed(11)
    q5 = torch.randn(N, 128, H // 8, W // 8, generator=g).cuda()
    q4 = torch.randn(N, 128, H // 4, W // 4, generator=g).cuda()
    q32 = torch.randn(N, 256, H, W, generator=g).cuda()
    w = (torch.randn(Cout, 512, 3, 3, generator=g) / 68.0).cuda()
    b = torch.randn(Cout, generator=g).cuda()
    z5 = ops.conv2d(ops.act_from_nchw(q5, fmt), ops.pack_conv(ops.phase_class_filter(w[:, :128]), None, None, fmt), pad=1, want_h8=False)
    z4 = ops.conv2d(ops.act_from_nchw(q4, fmt), ops.pack_conv(ops.phase_class_filter(w[:, 128:256]), None, None, fmt), pad=1, want_h8=False)
    y = ops.conv2d(ops.act_from_nchw(q32, fmt), ops.pack_conv(w[:, 256:].contiguous(), b, None, fmt), pad=1, relu=True,
                   gather=[(z5, 3), (z4, 2)]).to_nchw()
    # the replicated-concat formulation on the device (the previous product path) and, for the small cases, torch fp32
    cat = ops.Act(fmt, N, H, W, 512, "cuda")
    eye = torch.eye(128).reshape(128, 128, 1, 1).cuda()
    pe = ops.pack_conv(eye, None, None, fmt)
    ops.conv2d(ops.act_from_nchw(q5, fmt), pe, out=cat, out_coffset=0, out_rep=8)
    ops.conv2d(ops.act_from_nchw(q4, fmt), pe, out=cat, out_coffset=128, out_rep=4)
    ops.conv2d(ops.act_from_nchw(q32, fmt), ops.pack_conv(torch.eye(256).reshape(256, 256, 1, 1).cuda(), None, None, fmt), out=cat, out_coffset=256)
    y_cat = ops.conv2d(cat, ops.pack_conv(w, b, None, fmt), pad=1, relu=True).to_nchw()
    torch.cuda.synchronize()
    assert nerr(y, y_cat) <= tol
    if not big:
        x = torch.cat([F.interpolate(q5, scale_factor=8, mode="nearest"), F.interpolate(q4, scale_factor=4, mode="nearest"), q32], 1)
        ref = F.relu(F.conv2d(x, w, b, padding=1))
        assert nerr(y, ref) <= tol
