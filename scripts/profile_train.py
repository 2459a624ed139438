"""Kernel-time table of one training step (torch.profiler / CUPTI, eager launches)."""
import os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from multiposenet.pytorch_b200 import poseNet
from torch.profiler import profile, ProfilerActivity

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
dev = torch.device("cuda")
m = poseNet(101, precision=prec)
bench.load_weights_into(m, 101)
m = m.to(dev).train()
bench.freeze_for_keypoint_training(m)
eng = m.train_engine()
x = torch.randn(B, 3, 480, 640, device=dev)
gt = torch.rand(B, 18, 120, 160, device=dev)
wt = (torch.rand(B, 18, 120, 160, device=dev) > 0.2).float()
for _ in range(2):
    eng.forward_backward(x, gt, wt)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    eng.forward_backward(x, gt, wt)
    torch.cuda.synchronize()
rows = {}
for e in prof.events():
    if e.device_type.name != "CUDA":
        continue
    k = re.sub(r"\(anonymous namespace\)::|<unnamed>::|^void ", "", e.name)
    k = re.sub(r"\(.*$", "", k)[:70]
    r = rows.setdefault(k, [0, 0.0])
    r[0] += 1; r[1] += e.device_time
tot = sum(v[1] for v in rows.values())
print("batch %d %s: %d kernels, %.2f ms GPU busy" % (B, prec, sum(v[0] for v in rows.values()), tot / 1e3))
for k, (c, t) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:24]:
    print("%-72s %5d %9.3f ms %5.1f%%" % (k, c, t / 1e3, 100 * t / tot))
