"""Experiment: does splitting the batch over concurrent streams fill the tile-quantisation tails of the persistent conv kernels?
One CUDA graph of batch 32 against k graphs of batch 32/k replayed on k streams (separate engines, so no shared workspace)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from multiposenet.pytorch_b200 import poseNet


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--splits", default="1,2,4")
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda")
    res = {}
    for k in [int(s) for s in a.splits.split(",")]:
        b = a.batch // k
        engs, xs, streams = [], [], []
        for j in range(k):
            m = poseNet(101, precision=a.precision)
            bench.load_weights_into(m, 101)
            m = m.to(dev).eval()
            with torch.no_grad():
                m.classificationModel.output.bias += bench.CLS_BIAS_SHIFT[101]
            engs.append(m.engine())
            xs.append(torch.randn(b, 3, 480, 640, device=dev))
            streams.append(torch.cuda.Stream(device=dev))
        for j in range(k):  # capture
            with torch.cuda.stream(streams[j]):
                engs[j].graphed("entire", xs[j], max_cand=4096)
        torch.cuda.synchronize()

        def step():
            cur = torch.cuda.current_stream()
            for j in range(k):
                streams[j].wait_stream(cur)
                with torch.cuda.stream(streams[j]):
                    engs[j].graphed("entire", xs[j], max_cand=4096)
            for j in range(k):
                cur.wait_stream(streams[j])

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        res[k] = {"ms_per_step": ms, "images_per_s": a.batch / ms * 1e3}
        print("splits %d (batch %d each): %.3f ms/step, %.1f img/s" % (k, b, ms, a.batch / ms * 1e3), flush=True)
        del engs, xs
        torch.cuda.empty_cache()
    print(json.dumps({"precision": a.precision, "batch": a.batch, "results": res}))


if __name__ == "__main__":
    main()
