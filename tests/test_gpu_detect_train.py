"""Detection-subnet training: the focal-loss kernel against the reference's own FocalLoss (golden made by
the live reference) and the whole training step against torch autograd on the oracle restatement."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _grad_enabled():
    with torch.enable_grad():
        yield


def test_focal_loss_kernel_vs_reference_golden(golden_dir):
    from multiposenet.pytorch_b200.network.losses import FocalLoss, calc_iou
    from oracle import losses_oracle as lo
    g = np.load(os.path.join(golden_dir, "focal_loss.npz"))
    cls, reg, anchors, ann = (torch.from_numpy(a).cuda() for a in lo.focal_case(with_empty=False))
    cls.requires_grad_(True)
    reg.requires_grad_(True)
    cl, rl = FocalLoss()(cls, reg, anchors, ann)
    assert cl.shape == (1,) and rl.shape == (1,)
    (cl.mean() + rl.mean()).backward()
    np.testing.assert_allclose(cl.detach().cpu().numpy(), g["cls_loss"], rtol=2e-6)
    np.testing.assert_allclose(rl.detach().cpu().numpy(), g["reg_loss"], rtol=2e-6)
    dc, dr = cls.grad.cpu().numpy(), reg.grad.cpu().numpy()
    assert np.array_equal(dc != 0, g["dcls"] != 0) and np.array_equal(dr != 0, g["dreg"] != 0)   # same assignment, same clamp mask
    np.testing.assert_allclose(dc, g["dcls"], rtol=2e-5, atol=1e-10)
    np.testing.assert_allclose(dr, g["dreg"], rtol=2e-5, atol=1e-10)
    # an image without annotations (losses.py:53-57) and upstream gradient scaling, against the oracle restatement
    c2, r2, a2, n2 = (torch.from_numpy(a).cuda() for a in lo.focal_case(seed=4, with_empty=True))
    c2.requires_grad_(True); r2.requires_grad_(True)
    cl2, rl2 = FocalLoss()(c2, r2, a2, n2)
    (3.0 * cl2.sum() + 0.5 * rl2.sum()).backward()
    c3, r3 = c2.detach().cpu().requires_grad_(True), r2.detach().cpu().requires_grad_(True)
    ocl, orl, _, _ = lo.focal_loss(c3, r3, a2.cpu(), n2.cpu())
    (3.0 * ocl.sum() + 0.5 * orl.sum()).backward()
    np.testing.assert_allclose(cl2.detach().cpu().numpy(), ocl.detach().numpy(), rtol=2e-6)
    np.testing.assert_allclose(c2.grad.cpu().numpy(), c3.grad.numpy(), rtol=2e-5, atol=1e-10)
    np.testing.assert_allclose(r2.grad.cpu().numpy(), r3.grad.numpy(), rtol=2e-5, atol=1e-10)
    assert float(c2.grad[1].abs().max()) == 0.0
    iou = calc_iou(a2[0, :50], n2[0, :2, :4])
    assert iou.shape == (50, 2) and float(iou.max()) <= 1.0


def _freeze_for_detection(m):
    """training/multipose_detection_train.py:64-79."""
    for name, mod in m.fpn.named_children():
        if name in ("conv1", "bn1", "layer1", "layer2", "layer3", "layer4", "toplayer", "flatlayer1", "flatlayer2", "flatlayer3",
                    "smooth1", "smooth2", "smooth3"):
            for p in mod.parameters():
                p.requires_grad = False
    for name, mod in m.named_children():
        if name in ("convt1", "convt2", "convt3", "convt4", "convs1", "convs2", "convs3", "convs4", "conv2", "convfin", "convfin_k2",
                    "convfin_k3", "convfin_k4", "convfin_k5", "prn"):
            for p in mod.parameters():
                p.requires_grad = False


def test_detection_training_step_vs_autograd():
    """The reference loop (trainer.py:245-259) for the detection subnet: forward, build_loss, backward, Adam step."""
    from gpu_util import image, load_model, nerr, no_tf32
    from multiposenet.pytorch_b200 import poseNet
    from oracle import losses_oracle as lo, posenet_oracle as po, weights
    no_tf32()
    layers, hw, B = 50, (96, 128), 2
    m, w = load_model(layers, "conditioned", "bf16x3")
    m.train()
    m.freeze_bn()                                                     # trainer.py:173-174
    _freeze_for_detection(m)
    x = image(51, (B, 3) + hw)
    _, _, _, ann = lo.focal_case(seed=7, hw=hw, batch=B, with_empty=False)
    ann = torch.from_numpy(ann).cuda()
    opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=1e-5)
    out, saved = m([x, "detection_subnet"])
    assert out == [] and saved[0].requires_grad and saved[1].requires_grad
    loss, log = poseNet.build_loss(saved, "detection_subnet", ann)
    opt.zero_grad()
    loss.backward()
    # torch autograd on the oracle graph with the same weights
    sd = {k: v.cuda() for k, v in weights.to_torch_state_dict(w).items()}
    trainable = [k for k in sd if k.startswith(("fpn.conv6", "fpn.conv7", "fpn.latlayer", "fpn.toplayer0", "fpn.toplayer1", "fpn.toplayer2",
                                                "regressionModel", "classificationModel"))]
    for k in trainable:
        sd[k].requires_grad_(True)
    _, c3, c4, c5 = po.backbone(sd, layers, x)
    cls, reg = po.detection_heads(sd, po.detection_neck(sd, c3, c4, c5))
    ocl, orl, _, _ = lo.focal_loss(cls, reg, saved[2], ann)
    oloss = ocl.mean() + orl.mean()
    oloss.backward()
    assert nerr(saved[0].detach(), cls.detach()) <= 1e-3 and nerr(saved[1].detach(), reg.detach()) <= 1e-3
    assert abs(float(loss) - float(oloss)) <= 2e-3 * abs(float(oloss)), (float(loss), float(oloss))
    assert set(log) == {"total_loss", "classification_loss", "regression_loss"}
    params = dict(m.named_parameters())
    worst_max, worst_l2, checked = (0.0, None), (0.0, None), 0
    for k in trainable:
        gref = sd[k].grad
        got = params[k].grad
        assert got is not None and got.shape == gref.shape, k
        if float(gref.abs().max()) == 0.0:
            continue
        worst_max = max(worst_max, (nerr(got, gref), k))
        worst_l2 = max(worst_l2, (float((got.double() - gref.double()).norm() / gref.double().norm()), k))
        checked += 1
    print("detection step: %d gradients, worst max-norm %.3g (%s), worst L2-relative %.3g (%s)" % ((checked,) + worst_max + worst_l2))
    assert checked == 36
    # free-running comparison (no ReLU pattern imposed on the oracle): a handful of sign flips in the towers bound the max-norm
    assert worst_l2[0] <= 2e-2 and worst_max[0] <= 0.1, (worst_max, worst_l2)
    assert m.fpn.layer1[0].conv1.weight.grad is None and m.conv2.weight.grad is None      # frozen parts untouched
    before = m.regressionModel.conv1.weight.detach().clone()
    opt.step()
    assert not torch.equal(before, m.regressionModel.conv1.weight)
    # a model whose trunk is not frozen, or whose BatchNorm is live, is refused (no silent partial training)
    m2, _ = load_model(layers, "conditioned", "bf16x3")
    m2.train(); m2.freeze_bn()
    with pytest.raises(NotImplementedError):
        m2([x, "detection_subnet"])
