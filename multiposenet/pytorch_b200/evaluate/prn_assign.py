"""`Tester.prn_process` (evaluate/tester.py:333-513) on the device, for one image or a whole batch of images.

The reference runs, per image and in Python: a peaks x boxes loop that scatters one-hot joints into a 56x36x17 grid per
person box, one skimage gaussian per (box, joint) plane, one batch-1 PRN forward per box, and list-comprehension table
logic that gives each box at most one peak per joint type.  Here the host only regroups the (tiny) peak and box lists;
`mpn_prn_build_inputs` -> batched PRN forward on the tcgen05 path (`poseNet([inp, 'prn_subnet'])`) -> `mpn_prn_assign`
do the work for every box of every image in five kernel launches plus the PRN's four, and one [P,17,3] float64 array
comes back.  Same arguments and same records as the reference method; there is no CPU path.
"""

import numpy as np
import torch

from .. import ops

NUM_JOINTS = 17


def _gaussian_weights(sigma=1.0, truncate=4.0):
    """skimage.filters.gaussian's defaults (sigma=1, truncate=4) -> scipy.ndimage._gaussian_kernel1d weights, index =
    distance from the centre (float64; the kernel multiplies with exactly these values)."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    phi = phi / phi.sum()
    return phi[radius:].copy()


def _regroup(kps):
    """tester.py:337-350: rows regrouped by joint type 0..16 in their original order (rows of other types are ignored)."""
    kps = np.asarray(kps, dtype=np.float64).reshape(-1, 5) if len(kps) else np.zeros((0, 5), dtype=np.float64)
    ty = kps[:, 4]
    keep = np.isin(ty, np.arange(NUM_JOINTS))
    kps, ty = kps[keep], ty[keep].astype(np.int32)
    order = np.argsort(ty, kind="stable")
    return np.ascontiguousarray(kps[order, :2]), np.ascontiguousarray(ty[order])


def _reference_would_divide_by_zero(xy, ty, bx, in_thres):
    """tester.py:371-372 / :476-477 divide by ceil(box width / height).  A box with width or height <= 0 contains no peak
    (the strict inside test :366-369 cannot hold), so the first division is never reached for it; the second one sits in the
    fallback branch (:473), taken when some joint type has no peak inside ANY box of the image, and runs over ALL boxes:
    it raises ZeroDivisionError exactly when that branch is reached and some box has ceil(w) == 0 or ceil(h) == 0."""
    if not (np.ceil(bx[:, 2:]) == 0).any():
        return False
    inside = ((xy[:, None, 0] > bx[None, :, 0] - bx[None, :, 2] * in_thres) & (xy[:, None, 1] > bx[None, :, 1] - bx[None, :, 3] * in_thres) &
              (xy[:, None, 0] < bx[None, :, 0] + bx[None, :, 2] * (1.0 + in_thres)) & (xy[:, None, 1] < bx[None, :, 1] + bx[None, :, 3] * (1.0 + in_thres)))
    hit = inside.any(axis=1)
    return any(not hit[ty == j].any() for j in range(NUM_JOINTS))


def prn_process_batch(model, kps_per_image, bboxes_per_image, file_names=None, image_ids=None, coeff=2, in_thres=0.21, strict=True):
    """kps_per_image[b]: joint rows (x, y, score, id, joint_type 0..16) of image b (tester.py:219-229);
    bboxes_per_image[b]: person boxes (x1, y1, x2, y2) (tester.py:232-240).  Returns one list of records per image, each
    list identical to Tester.prn_process(kps, bbox_list, file_name, image_id).

    Degenerate boxes (ClipBoxes can leave x2 < x1): handled per box like the reference -- they contain no peaks.  Only where
    the reference itself raises ZeroDivisionError for an image (see _reference_would_divide_by_zero) does this function
    raise (strict=True, after nothing was launched) or return [] for that image alone (strict=False)."""
    nimg = len(kps_per_image)
    assert len(bboxes_per_image) == nimg
    file_names = file_names if file_names is not None else [""] * nimg
    image_ids = image_ids if image_ids is not None else [0] * nimg
    gh, gw = int(28 * coeff), int(18 * coeff)                                    # :353-354
    results = [[] for _ in range(nimg)]
    cand = [b for b in range(nimg) if len(bboxes_per_image[b]) > 0]              # :360 (an image without boxes yields [])
    prep, live = {}, []
    for b in cand:
        xy, ty = _regroup(kps_per_image[b])
        bx = np.array([[bb[0], bb[1], bb[2] - bb[0], bb[3] - bb[1]] for bb in bboxes_per_image[b]], dtype=np.float64)   # :356-358
        if _reference_would_divide_by_zero(xy, ty, bx, in_thres):
            if strict:
                raise ZeroDivisionError("prn_process: image %d reaches the fallback branch with a box whose ceil(width) or "
                                        "ceil(height) is 0 (tester.py:476-477 divides by it)" % b)
            continue
        prep[b] = (xy, ty, bx)
        live.append(b)
    if not live:
        return results
    xy_l, ty_l, box_l, box_img, pstart, bstart, jstart = [], [], [], [], [0], [0], []
    for li, b in enumerate(live):
        xy, ty, bx = prep[b]
        jstart.append(pstart[-1] + np.searchsorted(ty, np.arange(NUM_JOINTS + 1), side="left"))
        xy_l.append(xy); ty_l.append(ty); box_l.append(bx)
        box_img += [li] * len(bx)
        pstart.append(pstart[-1] + len(ty)); bstart.append(bstart[-1] + len(bx))
    dev = next(model.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("prn_process runs on the device: move the model to CUDA (there is no CPU path)")

    def up(a, dt):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(dev)

    peak_xy = up(np.concatenate(xy_l).reshape(-1, 2), torch.float64)
    peak_type = up(np.concatenate(ty_l), torch.int32)
    boxes = up(np.concatenate(box_l), torch.float64)
    box_img_t, pstart_t, bstart_t = up(np.array(box_img), torch.int32), up(np.array(pstart), torch.int32), up(np.array(bstart), torch.int32)
    jstart_t = up(np.stack(jstart), torch.int32)
    kmax = max(1, int(np.diff(pstart).max()))
    ws = ops.prn_workspace(boxes.shape[0], peak_type.shape[0], kmax, dev)
    owner, inp = ops.prn_build_inputs(peak_xy, peak_type, pstart_t, boxes, box_img_t, (gh, gw), in_thres, _gaussian_weights(), kmax, ws)
    with torch.no_grad():
        output, _ = model([inp, "prn_subnet"])                                   # :400-408, all persons in one batch
    kp = ops.prn_assign(peak_xy, pstart_t, jstart_t, boxes, box_img_t, bstart_t, owner, output.float(), kmax, ws).cpu().numpy()
    kpl = kp.reshape(-1, 3 * NUM_JOINTS).tolist()                                # python floats: the same IEEE doubles
    for li, b in enumerate(live):                                                # :485-511
        bxl = box_l[li].tolist()
        for i in range(len(bxl)):
            k = kpl[bstart[li] + i]
            pose_score = 0
            for v in k[2::3]:                                                    # sequential sum over the 17 joints (:497-499)
                pose_score += v
            pose_score /= 17.0
            results[b].append({"image_id": image_ids[b], "file_name": file_names[b], "category_id": 1,
                               "bbox": bxl[i], "score": float(pose_score), "keypoints": k})
    return results


def prn_process(model, kps, bbox_list, file_name, image_id=0, coeff=2, in_thres=0.21):
    """Drop-in for `Tester.prn_process(self, kps, bbox_list, file_name, image_id)` (self.model -> model,
    self.params.coeff / in_thres -> keyword arguments)."""
    return prn_process_batch(model, [kps], [bbox_list], [file_name], [image_id], coeff, in_thres)[0]
