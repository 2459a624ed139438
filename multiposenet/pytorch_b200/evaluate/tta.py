"""Multi-scale / flip test-time augmentation of `Tester.coco_eval` (evaluate/tester.py:131-152, 256-331) around the device path.

The reference runs, per image, 5 scales x {original, mirrored} = 10 batch-1 forwards; each heat map goes to the host, is
upsampled x4 and then resized to the original image size with cv2.INTER_CUBIC, and the ten results are averaged in numpy.
Here the original and the mirrored image of a scale (same padded size) form ONE batch-2 forward, the heat maps never leave
the device (`mpn_resize_cubic`: both cubic resizes, the `/ len(multiplier)` and the float64 accumulation, the mirror and the
left/right channel swap of the flipped pass fused into the second one; `mpn_tta_combine`: the final average), and the result
feeds `get_joint_list` / `mpn_heatmap_peaks` directly.  Host side (image decode side, cv2): the bilinear rescale + pad of the
input image (`crop_with_factor`, tester.py:38-81) and `resnet_preprocess`.

Same function names / arguments as the reference methods (`self` -> `model`, `self.params.inp_size` -> keyword).
"""
import numpy as np
import torch

from .. import ops

SCALE_SEARCH = (0.5, 1., 1.5, 2, 2.5)                                                       # tester.py:261
SWAP_HEAT = (0, 1, 5, 6, 7, 2, 3, 4, 11, 12, 13, 8, 9, 10, 15, 14, 17, 16)                  # tester.py:325-326


def get_multiplier(img, inp_size=480):
    """tester.py:256-262."""
    return [x * inp_size / float(img.shape[0]) for x in SCALE_SEARCH]


def _factor_closest(num, factor, is_ceil=True):
    num = float(num) / factor
    num = np.ceil(num) if is_ceil else np.floor(num)
    return int(num) * factor


def crop_with_factor(im, dest_size, factor=32, pad_val=0, basedon="min"):
    """tester.py:38-81: scale so that the `basedon` side becomes dest_size (cv2 bilinear), pad to a multiple of `factor`.
    Returns (padded image, scale, shape of the unpadded scaled image)."""
    import cv2
    im_size_min, im_size_max = np.min(im.shape[0:2]), np.max(im.shape[0:2])
    im_base = {"min": im_size_min, "max": im_size_max, "w": im.shape[1], "h": im.shape[0]}
    im_scale = float(dest_size) / im_base.get(basedon, im_size_min)
    im = cv2.resize(im, None, fx=im_scale, fy=im_scale)
    h, w = im.shape[:2]
    new_h, new_w = _factor_closest(h, factor), _factor_closest(w, factor)
    new_shape = [new_h, new_w] if im.ndim < 3 else [new_h, new_w, im.shape[-1]]
    im_padded = np.full(new_shape, fill_value=pad_val, dtype=im.dtype)
    im_padded[0:h, 0:w] = im
    return im_padded, im_scale, im.shape


def _resnet_preprocess(image):
    """datasets/coco_data/preprocessing.py:15-26 on the host for float images (the uint8 path is fused on the device)."""
    x = image.astype(np.float32) / 255.
    x = x.copy()[:, :, ::-1]
    for i, (m, s) in enumerate(zip((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))):
        x[:, :, i] = x[:, :, i] - m
        x[:, :, i] = x[:, :, i] / s
    return np.ascontiguousarray(x.transpose((2, 0, 1))).astype(np.float32)


def _boxes(scores, classes, boxes, im_scale):
    """tester.py:306-313: detections with score > 0.5 of class 0, scaled back to the original image."""
    out = []
    idxs = np.where(scores > 0.5)
    for j in range(idxs[0].shape[0]):
        bbox = boxes[idxs[0][j], :] / im_scale
        if int(classes[idxs[0][j]]) == 0:
            out.append(bbox.tolist())
    return out


def _accumulate(heat_b, pad_hw, real_shape, out_hw, acc, nscales, mirror=False, plane_map=None):
    """tester.py:296-304 for one image's heat maps (fp32 [C>=18, h, w] on the device) into the fp64 accumulator [18, H0, W0]."""
    sh, sw = int(pad_hw[0] / 4), int(pad_hw[1] / 4)                                          # :297
    src = heat_b[:18].contiguous()
    up = ops.resize_cubic(src, sh, sw, int(round(sh * 4.0)), int(round(sw * 4.0)), 1.0 / 4.0, 1.0 / 4.0)   # :298-299 (fx = fy = 4)
    rh, rw = min(int(real_shape[0]), up.shape[1]), min(int(real_shape[1]), up.shape[2])     # :300 (slicing clamps)
    H0, W0 = out_hw
    ops.resize_cubic(up, rh, rw, H0, W0, 1.0 / (float(W0) / rw), 1.0 / (float(H0) / rh), dst=acc, out_f64=True, div=float(nscales),
                     mirror=mirror, plane_map=plane_map)                                   # :301-304


def _forward(model, batch_np, device):
    x = torch.from_numpy(batch_np).to(device)
    with torch.no_grad():
        heat, (scores, classes, boxes) = model([x, "both"])
    return heat, scores, classes, boxes


def get_outputs(model, multiplier, img, device=None, as_numpy=True):
    """Drop-in for Tester._get_outputs (tester.py:264-314): (heatmap_avg [H0, W0, 18] float64, bbox_all per scale)."""
    device = device or next(model.parameters()).device
    H0, W0 = img.shape[0], img.shape[1]
    acc = torch.zeros((18, H0, W0), dtype=torch.float64, device=device)
    bbox_all = []
    for scale in multiplier:
        im_cropped, im_scale, real_shape = crop_with_factor(img, scale * img.shape[0], factor=32, pad_val=128)
        heat, scores, classes, boxes = _forward(model, _resnet_preprocess(im_cropped)[None], device)
        _accumulate(heat[0], im_cropped.shape, real_shape, (H0, W0), acc, len(multiplier))
        bbox_all.append(_boxes(scores.cpu().numpy(), classes.cpu().numpy(), boxes.cpu().numpy(), im_scale))
    if as_numpy:
        return acc.permute(1, 2, 0).cpu().numpy(), bbox_all
    return acc, bbox_all


def handle_heat(normal_heat, flipped_heat):
    """tester.py:316-331 on host arrays (numpy [H, W, 18]); the device path fuses this into multi_scale_flip."""
    swap = np.array(SWAP_HEAT)
    return (normal_heat + flipped_heat[:, ::-1, :][:, :, swap]) / 2.


def multi_scale_flip(model, img, inp_size=480, device=None):
    """tester.py:143-152 for one image: the 5 scales x {original, mirrored} as 5 batch-2 forwards, heat maps averaged on the
    device.  Returns (heat fp64 [18, H0, W0], heat fp32 [1, 18, H0, W0] for get_joint_list / mpn_heatmap_peaks, bbox_all of
    the original image per scale, bbox_all of the mirrored image per scale)."""
    device = device or next(model.parameters()).device
    multiplier = get_multiplier(img, inp_size)
    H0, W0 = img.shape[0], img.shape[1]
    swapped = img[:, ::-1, :]                                                                # :147
    acc_n = torch.zeros((18, H0, W0), dtype=torch.float64, device=device)
    acc_f = torch.zeros((18, H0, W0), dtype=torch.float64, device=device)
    # flipped[:, ::-1, :][:, :, swap]: destination channel c takes flipped channel swap[c]; SWAP_HEAT is an involution, so
    # source plane p lands in destination plane SWAP_HEAT[p]
    plane_map = torch.tensor(SWAP_HEAT, dtype=torch.int32, device=device)
    bbox_n, bbox_f = [], []
    eng = model.module.engine() if hasattr(model, "module") else model.engine()
    for scale in multiplier:
        a, im_scale, real_shape = crop_with_factor(img, scale * img.shape[0], factor=32, pad_val=128)
        b, _, real_b = crop_with_factor(swapped, scale * img.shape[0], factor=32, pad_val=128)
        batch = np.stack([_resnet_preprocess(a), _resnet_preprocess(b)])
        heat, scores, classes, boxes = _forward(model, batch, device)
        _accumulate(heat[0], a.shape, real_shape, (H0, W0), acc_n, len(multiplier))
        _accumulate(heat[1], b.shape, real_b, (H0, W0), acc_f, len(multiplier), mirror=True, plane_map=plane_map)
        det = eng.last_detections
        if det is None or int(det.cand_cnt[0]) == 0:
            bbox_n.append([])
        else:
            bbox_n.append(_boxes(scores.cpu().numpy(), classes.cpu().numpy(), boxes.cpu().numpy(), im_scale))
        if det is not None and det.keep_cnt.shape[0] > 1 and int(det.cand_cnt[1]) > 0:
            k = int(det.keep_cnt[1])
            bbox_f.append(_boxes(det.scores[1, :k].cpu().numpy(), np.zeros(k, np.int64), det.boxes[1, :k].cpu().numpy(), im_scale))
        else:
            bbox_f.append([])
    heat64, heat32 = ops.tta_combine(acc_n, acc_f)
    return heat64, heat32.view(1, 18, H0, W0), bbox_n, bbox_f
