"""Micro-benchmark of mpn_conv2d_fwd on the layer shapes that dominate the R101 step (CUDA events), with a parity check of
every shape against torch fp32 conv2d.  Used to iterate on the tcgen05 kernel and as the target of ncu captures:

    python scripts/bench_conv_shapes.py [--precision bf16x3] [--batch 32] [--iters 10] [--only SUBSTR] [--once] [--out FILE]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from multiposenet.pytorch_b200 import _lib, ops

# (name, H, W, Cin, Cout, k, stride, residual)
SHAPES = [
    ("l3.conv2   30x40 c256->256 k3", 30, 40, 256, 256, 3, 1, False),
    ("l3.conv3   30x40 c256->1024 k1 res", 30, 40, 256, 1024, 1, 1, True),
    ("l3.conv1   30x40 c1024->256 k1", 30, 40, 1024, 256, 1, 1, False),
    ("head       60x80 c256->256 k3", 60, 80, 256, 256, 3, 1, False),
    ("l1.conv3   120x160 c64->256 k1 res", 120, 160, 64, 256, 1, 1, True),
    ("l2.conv3   60x80 c128->512 k1 res", 60, 80, 128, 512, 1, 1, True),
    ("l2.conv2   60x80 c128->128 k3", 60, 80, 128, 128, 3, 1, False),
    ("head       15x20 c256->256 k3", 15, 20, 256, 256, 3, 1, False),
    ("l4.conv2   15x20 c512->512 k3", 15, 20, 512, 512, 3, 1, False),
    ("l4.conv3   15x20 c512->2048 k1 res", 15, 20, 512, 2048, 1, 1, True),
    ("l4.conv1   15x20 c2048->512 k1", 15, 20, 2048, 512, 1, 1, False),
    ("kp.convt   120x160 c512->256 k3", 120, 160, 512, 256, 3, 1, False),
    ("l1.conv2   120x160 c64->64 k3", 120, 160, 64, 64, 3, 1, False),
    ("head       8x10 c256->256 k3", 8, 10, 256, 256, 3, 1, False),
    ("l3.down    60x80 c512->1024 k1 s2", 60, 80, 512, 1024, 1, 2, False),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default=None)
    ap.add_argument("--once", action="store_true", help="one launch per shape, no timing loop (ncu target)")
    ap.add_argument("--out", default=None)
    ap.add_argument("--slim", type=int, default=1, help="f16f8: the engine's storage plan -- 1x1 convs read / write tensors without the "
                    "e5m2 copy plane (MPN_IN_NO_H8), the bottleneck 3x3 writes one without it")
    a = ap.parse_args()
    fmt = {"bf16x3": _lib.FMT_BF16X2, "bf16": _lib.FMT_BF16, "f16f8": _lib.FMT_F16F8}[a.precision]
    tol = {"bf16x3": 2e-4, "f16f8": 4e-4}.get(a.precision, 3e-2)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lines = ["precision %s batch %d" % (a.precision, a.batch),
             "%-40s %9s %9s %9s %9s" % ("shape", "us", "TFLOP/s", "us(L2cold)", "relerr")]
    bad = 0
    for name, H, W, Cin, Cout, k, stride, res in SHAPES:
        if a.only and not any(t in name for t in a.only.split(",")):
            continue
        B = a.batch
        x = torch.randn(B, Cin, H, W, device=dev, generator=g)
        w = torch.randn(Cout, Cin, k, k, device=dev, generator=g) * (2.0 / (Cin * k * k)) ** 0.5
        gamma = torch.rand(Cout, device=dev, generator=g) + 0.5
        beta = torch.randn(Cout, device=dev, generator=g) * 0.1
        mean = torch.randn(Cout, device=dev, generator=g) * 0.1
        var = torch.rand(Cout, device=dev, generator=g) + 0.5
        slim = bool(a.slim) and fmt == _lib.FMT_F16F8
        in_no_h8 = slim and k == 1 and not (a.slim == 2 and res)   # --slim 2: the shortcut convs read an input WITH the copy plane
        out_h8 = not (slim and (k == 1 or ".conv2" in name))
        xa = ops.act_from_nchw(x, fmt)
        pc = ops.pack_conv(w, None, (gamma, beta, mean, var, 1e-5), fmt, in_no_h8=in_no_h8)
        OH, OW = (H + 2 * (k // 2) - k) // stride + 1, (W + 2 * (k // 2) - k) // stride + 1
        ra = ops.act_from_nchw(torch.randn(B, Cout, OH, OW, device=dev, generator=g), fmt) if res else None
        def strip(t):
            u = ops.Act(t.fmt, t.N, t.H, t.W, t.C, dev, has_h8=False)
            u.hi.copy_(t.hi); u.lo[0].copy_(t.lo[0])
            return u
        if in_no_h8:
            xa = strip(xa)
        if slim and k == 1 and ra is not None:
            ra = strip(ra)
        out = ops.Act(fmt, B, OH, OW, Cout, dev, has_h8=out_h8)

        def run():
            return ops.conv2d(xa, pc, stride=stride, pad=k // 2, relu=True, residual=ra, out=out)

        run()
        torch.cuda.synchronize()
        # parity on 2 images (fp32 reference on the exactly representable hi(+lo) inputs)
        nb = min(2, B)
        xr = xa.to_nchw()[:nb]
        y = F.conv2d(xr, w, stride=stride, padding=k // 2)
        sc = gamma / torch.sqrt(var + 1e-5)
        y = y * sc[None, :, None, None] + (beta - mean * sc)[None, :, None, None]
        if res:
            y = y + ra.to_nchw()[:nb]
        y = torch.relu(y)
        got = out.to_nchw()[:nb]
        err = float((got - y).abs().max() / y.abs().max())
        bad += err > tol
        if a.once:
            lines.append("%-40s relerr %.2e" % (name, err))
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            run()
        e0.record()
        for _ in range(a.iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / a.iters
        cold = 0.0
        for _ in range(3):
            flush.fill_(1)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            cold += e0.elapsed_time(e1) * 1e3 / 3
        fl = 2.0 * B * OH * OW * Cout * Cin * k * k
        lines.append("%-40s %9.1f %9.1f %9.1f %9.2e%s" % (name, us, fl / us / 1e6, cold, err, "  PARITY FAIL" if err > tol else ""))
    txt = "\n".join(lines)
    print(txt)
    if a.out:
        open(a.out, "w").write(txt + "\n")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
