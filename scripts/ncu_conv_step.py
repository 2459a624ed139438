"""Target for the ncu passes: pack the weights (one forward), then ONE eager forward step whose launches are profiled.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \\
        --clock-control none -k regex:conv_tc_kernel -s <launches of the first forward> -c <launches of one step> \\
        --csv --log-file gpurun_out/conv_traffic.csv python scripts/ncu_conv_step.py
The number of conv_tc_kernel launches per forward is printed (179 for R101 entire_net)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from multiposenet.pytorch_b200 import ops, poseNet


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--layers", type=int, default=101)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--names-out", default=None, help="write the layer shape of every conv launch of one step, in launch order")
    a = ap.parse_args()
    dev = torch.device("cuda")
    m = poseNet(a.layers, precision=a.precision)
    bench.load_weights_into(m, a.layers)
    m = m.to(dev).eval()
    with torch.no_grad():
        m.classificationModel.output.bias += bench.CLS_BIAS_SHIFT.get(a.layers, 0.0)
    import multiposenet.pytorch_b200.engine as E
    E.LEVEL_STREAMS = False  # serial launches
    eng = m.engine()
    x = torch.randn(a.batch, 3, bench.H, bench.W, device=dev)
    names, real = [], ops.conv2d

    def named(xa, pc, **kw):  # launch order -> layer shape, to label the rows of the ncu CSV (scripts/summarise_ncu_csv.py --names)
        names.append("%dx%d c%d->%d k%d s%d%s%s%s" % (xa.H, xa.W, pc.Cin, pc.Cout, pc.R, kw.get("stride", 1),
                                                      " res" if kw.get("residual") is not None else "",
                                                      " up" if kw.get("up") is not None else "",
                                                      " rep%d" % kw["out_rep"] if kw.get("out_rep", 1) > 1 else ""))
        return real(xa, pc, **kw)

    if a.names_out:
        ops.conv2d = named
        real_multi = ops.conv2d_multi

        def named_multi(xs, pc, **kw):
            names.append("multi[%s] c%d->%d k%d s1" % (",".join("%dx%d" % (t.H, t.W) for t in xs), pc.Cin, pc.Cout, pc.R))
            return real_multi(xs, pc, **kw)
        ops.conv2d_multi = named_multi
    for i in range(a.steps):
        ops.stats["conv_events"] = evs = []
        eng.entire_forward_device(x, max_cand=4096)
        torch.cuda.synchronize()
        print("forward %d: %d conv launches" % (i, len(evs)))
    ops.stats["conv_events"] = None
    if a.names_out:
        per = len(names) // a.steps
        open(a.names_out, "w").write("\n".join(names[-per:]) + "\n")


if __name__ == "__main__":
    main()
