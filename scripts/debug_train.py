"""Per-parameter gradient error listing of the training step vs torch autograd on the oracle (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from gpu_util import image, load_model, nerr, no_tf32
from oracle import posenet_oracle as po, weights

no_tf32()
for (hw, B, dtype) in (((64, 96), 2, torch.float32), ((128, 192), 4, torch.float32), ((128, 192), 4, torch.float64)):
    m, w = load_model(50, "conditioned", "bf16x3")
    m.train()
    x = image(41, (B, 3) + hw)
    g = torch.Generator().manual_seed(9)
    gt = torch.rand(B, 18, hw[0] // 4, hw[1] // 4, generator=g).cuda()
    wt = (torch.rand(B, 18, hw[0] // 4, hw[1] // 4, generator=g) > 0.2).float().cuda()
    sd = {k: v.cuda().to(dtype) if v.dtype == torch.float32 else v.cuda() for k, v in weights.to_torch_state_dict(w).items()}
    for k, v in sd.items():
        if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            v.requires_grad_(True)
    saved = po.forward_train_keypoint(sd, 50, x.to(dtype))
    loss = po.keypoint_loss(saved, gt.to(dtype), wt.to(dtype))
    loss.backward()
    eng = m.train_engine()
    l2, outs, grads = eng.forward_backward(x, gt, wt)
    torch.cuda.synchronize()
    print("==== hw", hw, "B", B, dtype, "loss ref %.8f ours %.8f" % (float(loss), float(l2)))
    print("forward errs:", ["%.2e" % nerr(a, b.float()) for a, b in zip(outs, saved)])
    rows = []
    for k, v in sd.items():
        if v.grad is None or k not in grads:
            continue
        if float(v.grad.abs().max()) == 0:
            continue
        rows.append((k, nerr(grads[k], v.grad.float()), float(v.grad.abs().max())))
    for k, e, mx in rows:
        flag = " <<<<" if e > 3e-3 else ""
        if e > 1e-3 or k.endswith(("conv1.weight", "conv2.weight")) and ".0." in k:
            print("%-40s err %.3e  |g|max %.3e%s" % (k, e, mx, flag))
    print("max err %.3e, n>3e-3: %d of %d" % (max(r[1] for r in rows), sum(r[1] > 3e-3 for r in rows), len(rows)))
