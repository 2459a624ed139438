"""compute-sanitizer target: a small but complete pass over every kernel family -- the CTA-pair /
TMA-shortcut / tail-split / LSU / fp32-head variants of conv_tc_kernel (one R50 entire_net forward at 96x128, batch 2, which
also runs decode / filter / sort / NMS), the PRN assignment kernels, heat-map peaks and (with --train) one training step
(wgrad_tc_kernel, BN / pool / loss backward).
    compute-sanitizer --tool memcheck  python scripts/sanitize_target.py --train
    compute-sanitizer --tool racecheck python scripts/sanitize_target.py"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from multiposenet.pytorch_b200 import ops, poseNet


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--train", action="store_true")
    ap.add_argument("--precision", default="f16f8")
    a = ap.parse_args()
    dev = torch.device("cuda")
    m = poseNet(50, precision=a.precision)
    bench.load_weights_into(m, 50)
    m = m.to(dev).eval()
    with torch.no_grad():
        m.classificationModel.output.bias += 2.0   # a few hundred candidates per image through the NMS kernels
    x = torch.from_numpy(np.random.Generator(np.random.PCG64(1)).standard_normal((2, 3, 96, 128), dtype=np.float32)).to(dev)
    with torch.no_grad():
        heat, (sc, cl, bx) = m((x, "both"))
        rows, cnt = ops.heatmap_peaks(heat + 0.5, thre1=0.1, factor=4)
        hk, saved = m((x, "keypoint_subnet"))
    torch.cuda.synchronize()
    det = m.engine().last_detections
    print("forward ok: candidates %s kept %s peaks %s" % (det.cand_cnt.tolist(), det.keep_cnt.tolist(), cnt.tolist()))
    if a.train:
        m.train()
        bench.freeze_for_keypoint_training(m)
        eng = m.train_engine()
        gt = torch.rand(2, 18, 24, 32, device=dev)
        wt = (torch.rand(2, 18, 24, 32, device=dev) > 0.2).float()
        with torch.enable_grad():
            loss, outs, grads = eng.forward_backward(x, gt, wt)
        torch.cuda.synchronize()
        print("train step ok: loss %.5f, %d gradients" % (float(loss), len(grads)))


if __name__ == "__main__":
    main()
