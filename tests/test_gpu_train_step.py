"""GPU parity of the whole keypoint-subnet training step vs torch autograd on the oracle restatement (BN train mode)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _grad_enabled():
    with torch.enable_grad():
        yield


def _problem(layers=50, hw=(64, 96), B=2):
    from gpu_util import image, load_model, no_tf32
    no_tf32()
    m, w = load_model(layers, "conditioned", "bf16x3")
    x = image(41, (B, 3) + hw)
    g = torch.Generator().manual_seed(9)
    gt = torch.rand(B, 18, hw[0] // 4, hw[1] // 4, generator=g).cuda()
    wt = (torch.rand(B, 18, hw[0] // 4, hw[1] // 4, generator=g) > 0.2).float().cuda()
    return m, w, x, gt, wt


def _reference_grads(w, layers, x, gt, wt, masks=None, pool_idx=None):
    """torch autograd on the oracle restatement (BN train mode).  `masks`: the ReLU sign patterns of the implementation
    under test -- the two forwards agree to ~1e-4, so a few pre-activations of magnitude ~1e-4 change sign; a single
    flipped unit changes max-norm gradient errors by O(10 %) without being an error of the backward pass."""
    from oracle import posenet_oracle as po, weights
    sd = {k: v.cuda() for k, v in weights.to_torch_state_dict(w).items()}
    for k, v in sd.items():
        if v.dtype == torch.float32 and not k.endswith(("running_mean", "running_var")):
            v.requires_grad_(True)
    po.RELU_MASKS = iter(masks) if masks is not None else None
    po.POOL_INDICES = pool_idx
    try:
        saved = po.forward_train_keypoint(sd, layers, x)
    finally:
        po.RELU_MASKS = None
        po.POOL_INDICES = None
    loss = po.keypoint_loss(saved, gt, wt)
    loss.backward()
    return float(loss.detach()), [s.detach() for s in saved], {k: v.grad for k, v in sd.items() if v.grad is not None}


def test_fused_train_step_vs_autograd():
    from gpu_util import nerr
    m, w, x, gt, wt = _problem()
    m.train()
    eng = m.train_engine()
    loss, outs, grads = eng.forward_backward(x, gt, wt)
    torch.cuda.synchronize()
    masks = eng.relu_masks(eng.last_saved)
    loss_ref, saved_ref, gref = _reference_grads(w, 50, x, gt, wt, masks, eng.pool_indices(eng.last_saved))
    _, _, gfree = _reference_grads(w, 50, x, gt, wt)  # unconstrained reference: how much do the sign flips alone move it?
    flips = max(nerr(gfree[k], gref[k]) for k in gref if float(gref[k].abs().max()) > 0)
    print("max-norm effect of the ReLU sign flips on the reference's own gradients: %.3g" % flips)
    for a, b in zip(outs, saved_ref):
        assert nerr(a, b) <= 1e-3
    assert abs(float(loss) - loss_ref) <= 1e-4 * max(1.0, abs(loss_ref))
    frozen = ("fpn.conv6", "fpn.conv7", "fpn.latlayer", "fpn.toplayer0", "fpn.toplayer1", "fpn.toplayer2",
              "regressionModel", "classificationModel", "prn")
    checked = 0
    worst = (0.0, None)
    for k, g in gref.items():
        if k.startswith(frozen):
            assert k not in grads
            continue
        assert k in grads, "missing gradient for %s" % k
        assert grads[k].shape == g.shape, k
        if float(g.abs().max()) == 0.0:
            continue
        e = nerr(grads[k], g)
        if e > worst[0]:
            worst = (e, k)
        checked += 1
    print("checked %d gradients, worst normalized error %.3g at %s" % (checked, worst[0], worst[1]))
    assert checked > 150 and worst[0] <= 3e-3, worst
    # BN running statistics moved exactly like torch's (momentum 0.1, unbiased variance)
    assert int(m.fpn.bn1.num_batches_tracked) == 1


def test_reference_training_loop_surface():
    """The reference's loop (trainer.py:245-259): forward, build_loss, zero_grad, backward, Adam step."""
    from gpu_util import nerr
    from multiposenet.pytorch_b200 import poseNet
    m, w, x, gt, wt = _problem()
    m.train()
    for name, mod in m.named_children():  # multipose_keypoint_train.py:78-89 freezes the detection subnet and the PRN
        if name in ("regressionModel", "classificationModel", "prn"):
            for p in mod.parameters():
                p.requires_grad = False
    for name, mod in m.fpn.named_children():
        if name in ("conv6", "conv7", "latlayer1", "latlayer2", "latlayer3", "toplayer0", "toplayer1", "toplayer2"):
            for p in mod.parameters():
                p.requires_grad = False
    opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=1e-4)
    out, saved = m([x, "keypoint_subnet"])
    loss, log = poseNet.build_loss(saved, "keypoint_subnet", gt, wt)
    opt.zero_grad()
    loss.backward()
    teng = m.train_engine()
    loss_ref, _, gref = _reference_grads(w, 50, x, gt, wt, teng.relu_masks(teng.last_saved), teng.pool_indices(teng.last_saved))
    assert abs(float(loss) - loss_ref) <= 1e-4 * max(1.0, abs(loss_ref))
    assert nerr(m.conv2.weight.grad, gref["conv2.weight"]) <= 3e-3
    assert nerr(m.fpn.layer1[0].conv1.weight.grad, gref["fpn.layer1.0.conv1.weight"]) <= 3e-3
    assert nerr(m.fpn.conv1.weight.grad, gref["fpn.conv1.weight"]) <= 3e-3
    assert m.regressionModel.conv1.weight.grad is None
    before = m.conv2.weight.detach().clone()
    opt.step()
    assert not torch.equal(before, m.conv2.weight)
    assert "heatmap_loss" in log and "max_ht" in log
    # second step runs with the updated weights (filters are re-packed every step)
    out2, saved2 = m([x, "keypoint_subnet"])
    loss2, _ = poseNet.build_loss(saved2, "keypoint_subnet", gt, wt)
    assert float(loss2) < float(loss)


def test_graphed_train_step_matches_eager():
    from gpu_util import nerr
    m, w, x, gt, wt = _problem()
    m.train()
    eng = m.train_engine()
    loss_e, outs_e, grads_e = eng.forward_backward(x, gt, wt)
    ref = {k: v.clone() for k, v in grads_e.items()}
    outs_ref = [o.clone() for o in outs_e]
    loss_e = float(loss_e)
    nb = int(m.fpn.bn1.num_batches_tracked)
    for _ in range(2):
        loss_g, outs_g, grads_g = eng.graphed_forward_backward(x, gt, wt)
    torch.cuda.synchronize()
    assert abs(float(loss_g) - loss_e) <= 1e-5 * max(1.0, abs(loss_e))
    assert set(grads_g) == set(ref)
    # The forward is reproducible (fp64 cross-thread accumulation of the batch statistics); the weight gradients carry
    # fp32 atomic-order noise only.  A graph replay must therefore agree with the eager step to ~1e-5.
    errs = sorted(nerr(grads_g[k], ref[k]) for k in ref if float(ref[k].abs().max()) > 0)
    print("graph vs eager: median max-norm err %.2e, worst %.2e" % (errs[len(errs) // 2], errs[-1]))
    assert all(torch.equal(a, b) for a, b in zip(outs_g, outs_ref)), "forward not reproducible"
    assert errs[-1] <= 1e-4, errs[-1]
    assert int(m.fpn.bn1.num_batches_tracked) == nb + 2  # the capture/warm-up runs did not count as steps


def test_overlapped_step_matches_single_graph_step():
    """train_step_overlapped: two graphs sharing a pool, gradients gathered into one flat buffer in two buckets -- same loss and
    gradients as the single-graph step, BatchNorm running statistics advanced once per step."""
    from gpu_util import nerr
    m, w, x, gt, wt = _problem()
    m.train()
    eng = m.train_engine()
    loss_e, outs_e, grads_e = eng.forward_backward(x, gt, wt)
    ref = {k: v.clone() for k, v in grads_e.items()}
    loss_e = float(loss_e)
    nb = int(m.fpn.bn1.num_batches_tracked)
    for _ in range(2):
        loss_o = eng.train_step_overlapped(x, gt, wt, 1)
    torch.cuda.synchronize()
    assert abs(float(loss_o) - loss_e) <= 1e-5 * max(1.0, abs(loss_e))
    st = eng._ov[(tuple(x.shape), str(x.device))]
    names = [n for n, _, _, _ in st.index]
    assert set(names) == set(ref) and 0 < st.n_first < st.flat.numel()
    assert st.n_first > 0.8 * st.flat.numel()                       # head + neck + layer4 + layer3 are the large bucket
    first = [n for n, _, o, _ in st.index if o < st.n_first]
    assert any(n.startswith("fpn.layer3.0.") for n in first) and not any(n.startswith(("fpn.layer2", "fpn.layer1", "fpn.conv1")) for n in first)
    params = dict(m.named_parameters())
    errs = sorted(nerr(params[n].grad, ref[n]) for n in names if float(ref[n].abs().max()) > 0)
    assert errs[-1] <= 1e-4, errs[-1]
    assert int(m.fpn.bn1.num_batches_tracked) == nb + 2             # the dry run and the capture did not count as steps
