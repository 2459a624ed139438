"""CPU numerics study for the next precision mode of conv_tc_kernel: end-to-end error of the R50 / R101
entire_net graph when every convolution is evaluated with an emulated operand split.  TEST/DESIGN TOOL: it drives the
oracle graph (oracle/posenet_oracle.py) with a patched conv; nothing here is on the product path.

  bf16x3   : hi/lo bf16 planes, hi*hi + lo*hi + hi*lo                       (3 MMA units, the shipped parity mode)
  bf16     : one bf16 plane                                                 (1 unit, the fast mode)
  f16w2    : x fp16, w = fp16 hi + fp16 lo: x*wh + x*wl                     (2 units)
  f16f8s   : single-accumulator form of f16f8r (per-tensor weight prescale), the candidate for conv_tc_kernel
  f16f8r   : as f16f8 with the activations' fp8 planes in e5m2 (wider range, 3 significant bits)
  f16f8    : x = fp16 hi + e4m3 lo (scaled 2^12) + e4m3 copy; same for w (per-output-channel power-of-two prescale);
             hi*hi on kind::f16, lo8*hi8 + hi8*lo8 on kind::f8f6f4          (1 + 2 * 0.5 = 2 units)
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import torch.nn.functional as F

from oracle import posenet_oracle as po, weights

S_LO = 12


def bf16(t):
    return t.bfloat16().float()


def f16(t):
    return t.half().float()


def e4m3(t):
    return t.clamp(-448.0, 448.0).to(torch.float8_e4m3fn).float()


def e5m2(t):
    return t.clamp(-57344.0, 57344.0).to(torch.float8_e5m2).float()


def split_conv(mode):
    def conv(sd, name, x, stride=1, pad=0):
        w, b = sd[name + ".weight"], sd.get(name + ".bias")
        c = lambda a, ww: F.conv2d(a.double(), ww.double(), None, stride=stride, padding=pad)
        if mode == "bf16x3":
            xh, wh = bf16(x), bf16(w)
            xl, wl = bf16(x - xh), bf16(w - wh)
            y = c(xh, wh) + c(xl, wh) + c(xh, wl)
        elif mode == "bf16":
            y = c(bf16(x), bf16(w))
        elif mode == "f16w2":
            xh, wh = f16(x), f16(w)
            y = c(xh, wh) + c(xh, f16(w - wh))
        elif mode == "f16f8s":   # the shipped candidate: ONE accumulator.  Per-tensor weight prescale 2^k (max|w'| in [2^14, 2^15)) puts
            # all three terms on the same scale: xh16*wh16 + e5m2(xl*2^12)*e4m3(w'*2^-12) + e5m2(x)*e4m3(w' - wh16)
            k = torch.floor(torch.log2(32000.0 / w.abs().max().clamp_min(1e-30)))
            ws = w * torch.exp2(k)
            xh, wh = f16(x), f16(ws)
            y = (c(xh, wh) + c(e5m2((x - xh) * 2.0 ** S_LO), e4m3(ws * 2.0 ** -S_LO)) + c(e5m2(x), e4m3(ws - wh))) * float(torch.exp2(-k))
        elif mode == "f16f8r":   # range-robust variant: activations' fp8 planes in e5m2 (range 2^-16 .. 57344), weights' in e4m3
            amax = w.abs().flatten(1).max(1).values.clamp_min(1e-30)
            k = torch.floor(torch.log2(64.0 / amax)).view(-1, 1, 1, 1)
            ws = w * torch.exp2(k)
            xh, wh = f16(x), f16(ws)
            xl8 = e5m2((x - xh) * 2.0 ** S_LO)
            wl8 = e4m3((ws - wh) * 2.0 ** S_LO)
            xh8, wh8 = e5m2(x), e4m3(ws)
            y = (c(xh, wh) + (c(xl8, wh8) + c(xh8, wl8)) * 2.0 ** -S_LO) * torch.exp2(-k).view(1, -1, 1, 1).double()
        elif mode == "f16f8":
            amax = w.abs().flatten(1).max(1).values.clamp_min(1e-30)
            k = torch.floor(torch.log2(64.0 / amax)).view(-1, 1, 1, 1)        # max|w * 2^k| in [32, 64]
            ws = w * torch.exp2(k)
            xh, wh = f16(x), f16(ws)
            xl8 = e4m3((x - xh) * 2.0 ** S_LO)
            wl8 = e4m3((ws - wh) * 2.0 ** S_LO)
            xh8, wh8 = e4m3(x), e4m3(ws)
            y = (c(xh, wh) + (c(xl8, wh8) + c(xh8, wl8)) * 2.0 ** -S_LO) * torch.exp2(-k).view(1, -1, 1, 1).double()
        else:
            raise ValueError(mode)
        y = y.float()
        return y if b is None else y + b.view(1, -1, 1, 1)
    return conv


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=50)
    ap.add_argument("--hw", type=int, nargs=2, default=(64, 96))
    ap.add_argument("--modes", nargs="+", default=["bf16x3", "bf16", "f16w2", "f16f8"])
    ap.add_argument("--kind", default="conditioned")
    a = ap.parse_args()
    torch.set_grad_enabled(False)
    w = weights.make_weights(a.layers, a.kind, seed=0)
    sd = weights.to_torch_state_dict(w)
    x = torch.from_numpy(np.random.Generator(np.random.PCG64(1)).standard_normal((1, 3) + tuple(a.hw), dtype=np.float32))
    real = po._conv
    ref_heat, _ = po.forward(sd, a.layers, x, "keypoint_subnet")
    _, (rcls, rreg, _) = po.forward(sd, a.layers, x, "detection_subnet")
    nerr = lambda p, q: float((p.double() - q.double()).abs().max() / q.double().abs().max())
    print("R%d %s %dx%d: max|a-b|/max|b| vs the fp32 graph" % (a.layers, a.kind, a.hw[0], a.hw[1]))
    for mode in a.modes:
        po._conv = split_conv(mode)
        try:
            heat, _ = po.forward(sd, a.layers, x, "keypoint_subnet")
            _, (cls, reg, _) = po.forward(sd, a.layers, x, "detection_subnet")
        finally:
            po._conv = real
        print("  %-7s heat %.2e   cls %.2e   reg %.2e" % (mode, nerr(heat, ref_heat), nerr(cls, rcls), nerr(reg, rreg)))


if __name__ == "__main__":
    main()
