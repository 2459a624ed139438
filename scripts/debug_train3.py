"""Run-to-run reproducibility of the training step: eager vs eager vs graph (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from gpu_util import image, load_model, nerr, no_tf32
no_tf32()
hw, B = (64, 96), 2
m, w = load_model(50, "conditioned", "bf16x3")
m.train()
x = image(41, (B, 3) + hw)
g = torch.Generator().manual_seed(9)
gt = torch.rand(B, 18, hw[0] // 4, hw[1] // 4, generator=g).cuda()
wt = (torch.rand(B, 18, hw[0] // 4, hw[1] // 4, generator=g) > 0.2).float().cuda()
eng = m.train_engine()
def run(graph):
    f = eng.graphed_forward_backward if graph else eng.forward_backward
    loss, outs, grads = f(x, gt, wt)
    torch.cuda.synchronize()
    masks = eng.relu_masks(eng.last_saved) if not graph else None
    return float(loss), [o.clone() for o in outs], {k: v.clone() for k, v in grads.items()}, masks
def cmp(a, b, name):
    errs = sorted(nerr(a[2][k], b[2][k]) for k in a[2] if float(b[2][k].abs().max()) > 0)
    outs_equal = all(torch.equal(p, q) for p, q in zip(a[1], b[1]))
    flips = None
    if a[3] is not None and b[3] is not None:
        flips = sum(int((p != q).sum()) for p, q in zip(a[3], b[3]))
    print("%-18s loss %.9f vs %.9f  outs bitwise equal %s  relu flips %s  grad err median %.2e p90 %.2e max %.2e" % (
        name, a[0], b[0], outs_equal, flips, errs[len(errs)//2], errs[int(len(errs)*0.9)], errs[-1]))
e1 = run(False); e2 = run(False); e3 = run(False)
cmp(e1, e2, "eager1 vs eager2"); cmp(e2, e3, "eager2 vs eager3")
g1 = run(True); g2 = run(True)
cmp(e1, g1, "eager1 vs graph1"); cmp(g1, g2, "graph1 vs graph2")
