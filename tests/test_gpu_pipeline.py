"""The batched evaluate pipeline (forward -> peaks -> boxes -> PRN assignment; evaluate/tester.py:200-243) against the
image-by-image composition of the reference-shaped public calls, and against the numpy restatements on the same heat
maps / detections."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(precision="bf16x3"):
    from gpu_util import image, load_model
    from multiposenet.pytorch_b200 import synthetic
    m, _ = load_model(50, "conditioned", precision)
    x = image(7, (3, 3, 96, 128))
    synthetic.calibrate_output_bias(m, x, "cls", per_image=40, threshold=0.5)     # ~40 anchors/image above the box filter
    synthetic.calibrate_output_bias(m, x, "heat", per_image=150, threshold=0.1)   # ~150 heat-map pixels/image above thre1
    return m, x


@pytest.mark.parametrize("precision", ["bf16x3", "f16f8"])
def test_pipeline_batch_equals_per_image_and_oracle(precision):
    from multiposenet.pytorch_b200.evaluate import pipeline, prn_process, process_batch
    from multiposenet.pytorch_b200.network.joint_utils import get_joint_list
    from oracle import peaks_oracle, prn_oracle as po
    m, x = _model(precision)
    scales = [1.5, 2.0, 1.0]
    names, ids = ["a", "b", "c"], [5, 6, 7]
    recs, heat, det = process_batch(m, x, scales, names, ids)
    assert len(recs) == 3 and sum(len(r) for r in recs) > 0
    n_assigned = 0
    cnt = det.keep_cnt.cpu().numpy()
    for b in range(3):
        with torch.no_grad():
            heat1, _ = m((x[b:b + 1], "both"))                                       # tester.py:213 runs batch 1
        assert float((heat1[0] - heat[b]).abs().max()) <= 1e-5 * float(heat[b].abs().max())   # batch-size independent forward
        heat1 = heat[b:b + 1]
        sc, bx = det.scores[b, :cnt[b]], det.boxes[b, :cnt[b]]
        img_resized = np.zeros((96, 128, 3), np.float32)                             # only its shape is used (joint_utils.py:143)
        jl = get_joint_list(img_resized, {"thre1": 0.1}, heat1[0, :18], scales[b])
        kps = pipeline.joints_for_prn(jl)
        boxes = pipeline.boxes_for_prn(sc.cpu().numpy(), bx.cpu().numpy(), scales[b])
        want = prn_process(m, kps.tolist(), boxes, names[b], ids[b])
        assert recs[b] == want
        # the same chain through the numpy restatements, fed with this heat map / these detections and the device PRN
        ojl = peaks_oracle.joint_list(heat1[0, :18].cpu().numpy(), 0.1, 4)
        ojl[:, :2] *= scales[b]
        assert np.array_equal(ojl[:, [0, 1, 3, 4]], jl[:, [0, 1, 3, 4]])

        def prn_fn(inp):
            with torch.no_grad():
                return m([torch.from_numpy(inp).cuda(), "prn_subnet"])[0].float().cpu().numpy()

        owant = po.prn_process(pipeline.joints_for_prn(ojl).tolist(), boxes, prn_fn, names[b], ids[b])
        assert len(owant) == len(want)
        for a, o in zip(want, owant):
            assert a["keypoints"] == o["keypoints"] and a["score"] == o["score"] and a["bbox"] == o["bbox"]
            n_assigned += sum(1 for v in a["keypoints"][2::3] if v > 0)
    assert n_assigned > 0


def test_pipeline_max_persons_and_no_boxes():
    from multiposenet.pytorch_b200.evaluate import process_batch
    m, x = _model()
    recs, _, _ = process_batch(m, x, [1.0] * 3, max_persons=2)
    assert all(len(r) <= 2 for r in recs)
    with torch.no_grad():
        m.classificationModel.output.bias -= 50.0                                    # nothing passes the score filter
    recs, _, det = process_batch(m, x, [1.0] * 3)
    assert recs == [[], [], []] and int(det.keep_cnt.max()) == 0
