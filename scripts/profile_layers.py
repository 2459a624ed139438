"""Per-launch table of one forward step (CUDA events around every conv launch): shape, ms, TFLOP/s."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from multiposenet.pytorch_b200 import ops, poseNet
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--layers", type=int, default=101)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import multiposenet.pytorch_b200.engine as E
    E.USE_STREAMS = False
    E.LEVEL_STREAMS = False  # serial launches: clean per-kernel durations
    dev = torch.device("cuda")
    m = poseNet(a.layers, precision=a.precision)
    bench.load_weights_into(m, a.layers)
    m = m.to(dev).eval()
    eng = m.engine()
    x = torch.randn(a.batch, 3, 480, 640, device=dev)
    names = []
    real = ops.conv2d

    def named(xa, pc, **kw):
        names.append("%dx%d c%d->%d k%d s%d%s%s%s" % (xa.H, xa.W, pc.Cin, pc.Cout, pc.R, kw.get("stride", 1),
                                                      " res" if kw.get("residual") is not None else "",
                                                      " up" if kw.get("up") is not None else "",
                                                      " rep%d" % kw["out_rep"] if kw.get("out_rep", 1) > 1 else ""))
        return real(xa, pc, **kw)

    ops.conv2d = named
    real_multi = ops.conv2d_multi

    def named_multi(xs, pc, **kw):
        names.append("multi[%s] c%d->%d k%d s1" % (",".join("%dx%d" % (t.H, t.W) for t in xs), pc.Cin, pc.Cout, pc.R))
        return real_multi(xs, pc, **kw)

    ops.conv2d_multi = named_multi
    for _ in range(2):
        eng.entire_forward_device(x, max_cand=8192)
    torch.cuda.synchronize()
    # whole step without per-launch events
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.entire_forward_device(x, max_cand=8192)
    e1.record()
    torch.cuda.synchronize()
    plain = e0.elapsed_time(e1)
    names.clear()
    ops.stats["conv_events"] = evs = []
    torch.cuda._sleep(int(7e7))   # let the host enqueue the whole step ahead of the GPU: intervals without launch gaps
    e0.record()
    eng.entire_forward_device(x, max_cand=8192)
    e1.record()
    torch.cuda.synchronize()
    ops.stats["conv_events"] = None
    rows = []
    for n, (s, e, fl, simt, _slots) in zip(names, evs):
        ms = s.elapsed_time(e)
        rows.append((n, ms, fl / ms / 1e9, simt))
    total = e0.elapsed_time(e1)
    conv_total = sum(r[1] for r in rows)
    lines = ["precision %s batch %d layers %d: step %.3f ms plain, %.3f ms with per-launch events; conv launches %.3f ms (%d)" % (
        a.precision, a.batch, a.layers, plain, total, conv_total, len(rows))]
    agg = {}
    for n, ms, tf, simt in rows:
        c = agg.setdefault((n, simt), [0, 0.0, 0.0])
        c[0] += 1; c[1] += ms; c[2] += tf * ms
    lines.append("%-44s %5s %9s %9s %8s" % ("launch", "count", "ms_total", "ms_each", "TFLOP/s"))
    for (n, simt), (c, ms, tfms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-44s %5d %9.3f %9.4f %8.1f%s" % (n, c, ms, ms / c, tfms / ms, "  [cuda-core]" if simt else ""))
    txt = "\n".join(lines)
    print(txt)
    if a.out:
        open(a.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
