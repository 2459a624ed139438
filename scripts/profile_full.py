"""Stage timing of evaluate.process_batch (BASELINE config 5): wall time per stage with a device sync after each, plus a
cProfile of the host side of one step.  `python scripts/profile_full.py [--batch 64] [--persons 20]`"""
import argparse
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from multiposenet.pytorch_b200 import ops, poseNet, synthetic
from multiposenet.pytorch_b200.evaluate import pipeline, prn_assign


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--persons", type=int, default=20)
    ap.add_argument("--precision", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda")
    model = poseNet(101, precision=a.precision)
    bench.load_weights_into(model, 101)
    model = model.to(dev).eval()
    rng = np.random.Generator(np.random.PCG64(777))
    x = torch.from_numpy(rng.standard_normal((a.batch, 3, 480, 640), dtype=np.float32)).to(dev)
    probe = x[:4]
    with torch.no_grad():
        model.classificationModel.output.bias += bench.CLS_BIAS_SHIFT[101]  # the headline bench's detection load
    synthetic.calibrate_output_bias(model, probe, "heat", per_image=600, threshold=0.1)
    scales = [1.0] * a.batch
    for _ in range(2):
        pipeline.process_batch(model, x, scales, max_persons=a.persons, box_score_thresh=0.05)
    torch.cuda.synchronize()

    # wrap the stages
    marks = []
    real = {}

    def timed(mod, name):
        f = getattr(mod, name)
        real[(mod, name)] = f

        def w(*args, **kw):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = f(*args, **kw)
            torch.cuda.synchronize(); marks.append((name, (time.perf_counter() - t0) * 1e3))
            return r
        setattr(mod, name, w)

    eng = model.engine()
    timed(eng, "entire_forward_device")
    timed(pipeline.joint_utils, "joint_lists")
    timed(ops, "prn_build_inputs")
    timed(ops, "prn_assign")
    timed(eng, "prn_forward")
    timed(pipeline, "prn_process_batch")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pipeline.process_batch(model, x, scales, max_persons=a.persons, box_score_thresh=0.05)
    torch.cuda.synchronize(); total = (time.perf_counter() - t0) * 1e3
    for n, ms in marks:
        print("%-28s %8.2f ms" % (n, ms))
    print("%-28s %8.2f ms" % ("process_batch total", total))
    for (mod, name), f in real.items():
        setattr(mod, name, f)
    pr = cProfile.Profile()
    pr.enable()
    pipeline.process_batch(model, x, scales, max_persons=a.persons, box_score_thresh=0.05)
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)


if __name__ == "__main__":
    main()
