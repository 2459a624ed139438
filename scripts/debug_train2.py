"""Gradient w.r.t. every bottleneck output: ours vs torch autograd (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from gpu_util import image, load_model, nerr, no_tf32
from oracle import posenet_oracle as po, weights

no_tf32()
hw, B = (128, 192), 4
m, w = load_model(50, "conditioned", "bf16x3")
m.train()
x = image(41, (B, 3) + hw)
g = torch.Generator().manual_seed(9)
gt = torch.rand(B, 18, hw[0] // 4, hw[1] // 4, generator=g).cuda()
wt = (torch.rand(B, 18, hw[0] // 4, hw[1] // 4, generator=g) > 0.2).float().cuda()
sd = {k: v.cuda() for k, v in weights.to_torch_state_dict(w).items()}
for k, v in sd.items():
    if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
        v.requires_grad_(True)
po.TRACE = {}
saved = po.forward_train_keypoint(sd, 50, x)
loss = po.keypoint_loss(saved, gt, wt)
loss.backward()
ref = dict(po.TRACE)
po.TRACE = None
eng = m.train_engine()
eng.trace = {}
l2, outs, grads = eng.forward_backward(x, gt, wt)
torch.cuda.synchronize()
names = list(ref.keys())
for n in reversed(names):
    r = ref[n]
    mine = eng.trace[n]
    # forward value of the block output as we saved it
    blk = [b for b in eng.last_saved.blocks if b.name == n][0]
    fwd = blk.st3.z.to_nchw()
    print("%-16s fwd err %.2e   d(out) err %.3e  |d|max %.2e  frac(out>0) ref %.3f ours %.3f" % (
        n, nerr(fwd, r.detach()), nerr(mine, r.grad), float(r.grad.abs().max()), float((r > 0).float().mean()), float((fwd > 0).float().mean())))
