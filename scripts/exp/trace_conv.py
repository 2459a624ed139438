"""Timeline of one persistent CTA of conv_tc_kernel (debug experiment, not on the product path).

Needs the instrumented build:  python -m multiposenet.pytorch_b200.csrc.build --trace   (-> csrc/libmpn_b200_trace.so), then

    MPN_B200_LIB=multiposenet/pytorch_b200/csrc/libmpn_b200_trace.so python scripts/exp/trace_conv.py [--shape l3] [--out FILE]

The kernel stamps clock64 at: producer K-block issue (1), MMA issuer: accumulator free (2) / stage full (3), wide-epilogue group
issuer: before / after the accumulator-full wait (4, 5), shortcut box arrived (6), chunk math done (7), previous store read (8),
barrier A (9), barrier B (10), tile released (11).  Prints, per role, the mean time between consecutive events in steady state.
"""
import argparse
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from multiposenet.pytorch_b200 import _lib, ops

# (H, W, Cin, Cout, k, shortcut)
SHAPES = {"l3": (30, 40, 256, 1024, 1, True), "l2": (60, 80, 128, 512, 1, True), "l1": (120, 160, 64, 256, 1, True),
          "l4": (15, 20, 512, 2048, 1, True), "h60": (60, 80, 256, 256, 3, False), "kp": (120, 160, 512, 256, 3, False),
          "l3c1": (30, 40, 1024, 256, 1, False), "l3c2": (30, 40, 256, 256, 3, False)}
EV = {1: "P.issue", 2: "M.acc_free", 3: "M.stage_full", 4: "E.tile_start", 5: "E.acc_full", 6: "E.res_arrived", 7: "E.math_done",
      8: "E.store_read", 9: "E.barA", 10: "E.barB", 11: "E.tile_done", 12: "C.wait", 13: "C.a_landed", 14: "C.converted", 15: "C.arrived",
      16: "M.b_full", 17: "E.store_issued"}
CTAS, ROLES = 4, 8


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="l3")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--cap", type=int, default=4096)
    ap.add_argument("--out", default=None)
    ap.add_argument("--dump", default=None, help="write every event of CTA 0 (role, cycle, event, tile, arg) to this file")
    ap.add_argument("--out-h8", action="store_true", help="the output keeps its e5m2 copy plane (three store boxes per chunk)")
    ap.add_argument("--stored", action="store_true", help="MODE_F16F8: the input keeps its e5m2 copy plane and TMA loads it")
    ap.add_argument("--lo16", action="store_true", help="MODE_F16F8B (filter packed with the fp16 residual plane) instead of the derived copy plane")
    a = ap.parse_args()
    L = _lib.lib()
    if not hasattr(L, "mpn_debug_conv_trace"):
        raise SystemExit("not the trace build: set MPN_B200_LIB to libmpn_b200_trace.so")
    H, W, Cin, Cout, k, res = SHAPES[a.shape]
    fmt = _lib.FMT_F16F8
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(a.batch, Cin, H, W, device=dev, generator=g)
    w = torch.randn(Cout, Cin, k, k, device=dev, generator=g) * (2.0 / (Cin * k * k)) ** 0.5
    bn = (torch.rand(Cout, device=dev, generator=g) + 0.5, torch.randn(Cout, device=dev, generator=g) * 0.1,
          torch.randn(Cout, device=dev, generator=g) * 0.1, torch.rand(Cout, device=dev, generator=g) + 0.5, 1e-5)

    def strip(t):
        u = ops.Act(t.fmt, t.N, t.H, t.W, t.C, dev, has_h8=False)
        u.hi.copy_(t.hi); u.lo[0].copy_(t.lo[0])
        return u
    xa = ops.act_from_nchw(x, fmt)
    if not a.stored:
        xa = strip(xa)
    derive = False if a.stored else None
    ra = strip(ops.act_from_nchw(torch.randn(a.batch, Cout, H, W, device=dev, generator=g), fmt)) if res else None
    pc = ops.pack_conv(w, None, bn, fmt, in_no_h8=a.lo16)
    out = ops.Act(fmt, a.batch, H, W, Cout, dev, has_h8=a.out_h8)
    buf = torch.zeros(CTAS * ROLES * a.cap * 2, dtype=torch.int64, device=dev)
    L.mpn_debug_conv_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
    for _ in range(3):
        ops.conv2d(xa, pc, pad=k // 2, relu=True, residual=ra, out=out, derive=derive)
    torch.cuda.synchronize()
    assert L.mpn_debug_conv_trace(ctypes.c_void_p(buf.data_ptr()), a.cap) == 0
    ops.conv2d(xa, pc, pad=k // 2, relu=True, residual=ra, out=out, derive=derive)
    torch.cuda.synchronize()
    L.mpn_debug_conv_trace(None, 0)
    t = buf.cpu().numpy().reshape(CTAS, ROLES, a.cap, 2)
    lines = ["shape %s batch %d: %dx%d c%d->%d k%d%s" % (a.shape, a.batch, H, W, Cin, Cout, k, " res" if res else "")]
    mhz = 1.0
    for cta in range(CTAS):
        for role in range(ROLES):
            r = t[cta, role]
            n = int((r[:, 0] != 0).sum())
            if n < 8:
                continue
            clk = r[:n, 0].astype(np.int64)
            ev = (r[:n, 1] & 0xFF).astype(np.int64)
            tile = ((r[:n, 1] >> 8) & 0xFFFFFF).astype(np.int64)
            arg = (r[:n, 1] >> 32).astype(np.int64)
            tiles = sorted(set(tile.tolist()))
            lines.append("cta %d role %d (%s): %d events, %d tiles, span %d cycles, %.0f cycles / tile" %
                         (cta, role, "producer" if role == 0 else "mma" if role == 1 else "converter %d" % (role - 6) if role >= 6 else "epilogue group %d" % (role - 2), n, len(tiles),
                          clk[-1] - clk[0], (clk[-1] - clk[0]) / max(1, len(tiles) - 1)))
            # steady state: skip the first and the last tile
            keep = (tile != tiles[0]) & (tile != tiles[-1])
            d = np.diff(clk)
            trans = {}
            for i in range(n - 1):
                if keep[i] and keep[i + 1]:
                    trans.setdefault((int(ev[i]), int(ev[i + 1])), []).append(int(d[i]))
            for (e0, e1), v in sorted(trans.items()):
                lines.append("    %-14s -> %-14s  n %4d  mean %7.0f  min %6d  max %6d  (sum/tile %7.0f)" %
                             (EV[e0], EV[e1], len(v), np.mean(v), min(v), max(v), sum(v) / max(1, len(tiles) - 2)))
            if cta == 0 and role in (0, 1, 2, 6):   # raw timeline of steady-state tiles
                sel = [i for i in range(n) if tile[i] in (tiles[3:6] if len(tiles) > 6 else tiles[1:2])][:60]
                base = clk[sel[0]] if sel else 0
                lines.append("    raw: " + " ".join("%s@%d(t%d,%d)" % (EV[int(ev[i])].split(".")[1], clk[i] - base, tile[i], arg[i]) for i in sel))
    if a.dump:
        ev_all = []
        for role in range(ROLES):
            r = t[0, role]
            n = int((r[:, 0] != 0).sum())
            for i in range(n):
                ev_all.append((int(r[i, 0]), role, int(r[i, 1] & 0xFF), int((r[i, 1] >> 8) & 0xFFFFFF), int(r[i, 1] >> 32)))
        ev_all.sort()
        t0 = ev_all[0][0] if ev_all else 0
        with open(a.dump, "w") as f:
            for c, role, e, tl, ar in ev_all:
                f.write("%8d role %d %-16s tile %d arg %d\n" % (c - t0, role, EV[e], tl, ar))
    txt = "\n".join(lines)
    print(txt)
    if a.out:
        open(a.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
