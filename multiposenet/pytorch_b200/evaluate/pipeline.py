"""The per-image body of `Tester._process` (evaluate/tester.py:200-243) for a whole batch of images on the device:

    forward('both')  ->  get_joint_list  ->  neck row dropped, joint types shifted  ->  boxes with score > 0.5  ->  prn_process

The reference does this one image at a time (batch-1 forward, heat maps copied to the host, scipy/cv2 peak finding,
batch-1 PRN calls).  Here one batched forward + decode/filter/NMS, one `mpn_heatmap_peaks`, one batched PRN forward and
the assignment kernels serve the whole batch; the host only regroups the few hundred peak / box rows in between
(BASELINE config 5: "full posenet inference incl. PRN assignment").
"""
import numpy as np
import torch

from .. import ops
from ..network import joint_utils
from .prn_assign import prn_process_batch

NECK = 1  # joint type 1 of the 18 heat-map channels is the neck (tester.py:224-229 drops it)


def joints_for_prn(joint_rows):
    """tester.py:222-229: drop the neck rows, shift the later joint types down by one (type 0 stays 0)."""
    rows = np.asarray(joint_rows, dtype=np.float64).reshape(-1, 5)
    rows = rows[rows[:, 4].astype(np.int64) != NECK].copy()
    rows[:, 4] = np.maximum(0, rows[:, 4].astype(np.int64) - 1)
    return rows


def boxes_for_prn(scores, boxes, scale, score_thresh=0.5, limit=None):
    """tester.py:232-240: kept detections with score > 0.5, scaled back to the original image (class is always 0).
    One vectorised float32 multiply: the same products as the reference's per-box `boxes[i] * scale`.  limit = keep only
    the first `limit` selected boxes (they arrive in descending score order)."""
    scores = np.asarray(scores)
    sel = np.where(scores > score_thresh)[0]
    if limit is not None:
        sel = sel[:limit]
    return (np.asarray(boxes)[sel] * scale).tolist() if len(sel) else []


@torch.no_grad()
def process_batch(model, img_batch, scales, file_names=None, image_ids=None, thre1=0.1, box_score_thresh=0.5, coeff=2,
                  in_thres=0.21, max_cand=4096, max_persons=None):
    """img_batch: CUDA fp32 [B,3,H,W] (resnet_preprocess'ed, tester.py:208-212) or uint8 [B,H,W,3] BGR raw images;
    scales[b] = original size / network input size (tester.py:203).  Returns (records per image, heat maps, Detections).
    max_persons (None = all, the reference) caps the boxes per image at the best-scoring ones -- a synthetic-benchmark knob."""
    B = img_batch.shape[0]
    eng = model.engine()
    with torch.cuda.device(img_batch.device):
        heat, cls, reg, boxes, det = eng.entire_forward_device(img_batch, max_cand=max_cand)
        H = img_batch.shape[1] if img_batch.dtype == torch.uint8 else img_batch.shape[2]
        factor = H // heat.shape[2]                                               # joint_utils.py:143-144 (480/120)
        joints = joint_utils.joint_lists(heat, thre1=thre1, factor=factor, scales=scales)
        ncand = int(det.cand_cnt.max())
        if ncand > det.max_cand:  # rare: more survivors than the fast-path capacity -> redo with room (engine.entire_forward)
            det = ops.filter_sort_nms(cls, boxes, 0.05, 0.5, max_cand=ncand)
        cnt = det.keep_cnt.cpu().numpy()
        kmax = int(cnt.max()) if B else 0
        sc = det.scores[:, :kmax].cpu().numpy()
        bx = det.boxes[:, :kmax].cpu().numpy()
        kps, bboxes = [], []
        for b in range(B):
            kps.append(joints_for_prn(joints[b]))
            bboxes.append(boxes_for_prn(sc[b, :cnt[b]], bx[b, :cnt[b]], float(scales[b]), box_score_thresh, max_persons))   # descending score (pth_nms order)
        records = prn_process_batch(model, kps, bboxes, file_names, image_ids, coeff, in_thres)
    return records, heat, det
