"""Host restatement of the reference's multi-scale / flip test-time augmentation (TEST INFRASTRUCTURE).

Follows evaluate/tester.py: `_factor_closest` :19-35, `crop_with_factor` :38-81, `Tester._get_multiplier` :256-262,
`Tester._get_outputs` :264-314, `Tester._handle_heat` :316-331 -- with cv2 doing the resizes exactly as the reference calls
it (the third-party arithmetic, OpenCV 4.x here vs the reference's pinned opencv 3.4) and `model_fn` in place of
`self.model([im_data, subnet_name])`.  Pinned by tests/golden/tta.npz = the outputs of the reference's own methods run in
this container with a stub model (oracle/make_goldens.py tta_golden)."""
import numpy as np

SCALE_SEARCH = [0.5, 1., 1.5, 2, 2.5]
SWAP_HEAT = np.array((0, 1, 5, 6, 7, 2, 3, 4, 11, 12, 13, 8, 9, 10, 15, 14, 17, 16))


def factor_closest(num, factor, is_ceil=True):
    num = float(num) / factor
    num = np.ceil(num) if is_ceil else np.floor(num)
    return int(num) * factor


def crop_with_factor(im, dest_size, factor=32, pad_val=0, basedon="min"):
    import cv2
    im_size_min, im_size_max = np.min(im.shape[0:2]), np.max(im.shape[0:2])
    im_base = {"min": im_size_min, "max": im_size_max, "w": im.shape[1], "h": im.shape[0]}
    im_scale = float(dest_size) / im_base.get(basedon, im_size_min)
    im = cv2.resize(im, None, fx=im_scale, fy=im_scale)
    h, w = im.shape[:2]
    new_h, new_w = factor_closest(h, factor), factor_closest(w, factor)
    new_shape = [new_h, new_w] if im.ndim < 3 else [new_h, new_w, im.shape[-1]]
    im_padded = np.full(new_shape, fill_value=pad_val, dtype=im.dtype)
    im_padded[0:h, 0:w] = im
    return im_padded, im_scale, im.shape


def get_multiplier(img, inp_size=480):
    return [x * inp_size / float(img.shape[0]) for x in SCALE_SEARCH]


def get_outputs(model_fn, multiplier, img):
    """model_fn(im_data float32 [1,3,h,w]) -> (heat [1,C>=18,h/4,w/4], scores [K], classes [K], boxes [K,4]) numpy arrays."""
    import cv2
    from .preprocess_oracle import resnet_preprocess
    heatmap_avg = np.zeros((img.shape[0], img.shape[1], 18))
    bbox_all = []
    for scale in multiplier:
        inp_size = scale * img.shape[0]
        im_cropped, im_scale, real_shape = crop_with_factor(img, inp_size, factor=32, pad_val=128)
        im_data = np.expand_dims(resnet_preprocess(im_cropped), 0)
        heatmaps, scores, classification, boxes = model_fn(im_data)
        heatmaps = np.asarray(heatmaps).transpose(0, 2, 3, 1)
        heatmap = heatmaps[0, :int(im_cropped.shape[0] / 4), :int(im_cropped.shape[1] / 4), :18]
        heatmap = cv2.resize(np.ascontiguousarray(heatmap), None, fx=4, fy=4, interpolation=cv2.INTER_CUBIC)
        heatmap = heatmap[0:real_shape[0], 0:real_shape[1], :]
        heatmap = cv2.resize(heatmap, (img.shape[1], img.shape[0]), interpolation=cv2.INTER_CUBIC)
        heatmap_avg = heatmap_avg + heatmap / len(multiplier)
        idxs = np.where(scores > 0.5)
        bboxs = []
        for j in range(idxs[0].shape[0]):
            bbox = boxes[idxs[0][j], :] / im_scale
            if int(classification[idxs[0][j]]) == 0:
                bboxs.append(bbox.tolist())
        bbox_all.append(bboxs)
    return heatmap_avg, bbox_all


def handle_heat(normal_heat, flipped_heat):
    return (normal_heat + flipped_heat[:, ::-1, :][:, :, SWAP_HEAT]) / 2.


def stub_model(im_data):
    """A deterministic stand-in for the network (golden generation): 18 smooth heat maps derived from the input image by
    4x4 average pooling, three boxes whose scores depend on the image mean."""
    x = np.asarray(im_data, dtype=np.float32)[0]
    h4, w4 = x.shape[1] // 4, x.shape[2] // 4
    pooled = x[:, :h4 * 4, :w4 * 4].reshape(3, h4, 4, w4, 4).mean(axis=(2, 4))
    yy, xx = np.mgrid[0:h4, 0:w4].astype(np.float32)
    heat = np.stack([np.tanh(pooled[c % 3] * np.float32(0.3 + 0.05 * c)) * np.float32(0.5)
                     + np.float32(0.25) * np.sin(yy * np.float32(0.37 + 0.01 * c) + xx * np.float32(0.23 - 0.005 * c)) for c in range(18)]).astype(np.float32)
    m = float(x.mean())
    scores = np.array([0.9, 0.55 + 0.01 * m, 0.3], dtype=np.float32)
    classes = np.zeros(3, dtype=np.int64)
    boxes = np.array([[4, 6, 0.5 * x.shape[2], 0.7 * x.shape[1]], [10, 12, 60, 90], [1, 2, 3, 4]], dtype=np.float32)
    return heat[None], scores, classes, boxes


def test_image(seed=3, hw=(45, 61)):
    """A smooth seeded BGR float32 test image in 0..255 (what cv2.imread(...).astype(np.float32) yields)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    yy, xx = np.mgrid[0:hw[0], 0:hw[1]].astype(np.float64)
    img = np.stack([127 + 100 * np.sin(yy * rng.uniform(0.05, 0.3) + xx * rng.uniform(0.05, 0.3) + rng.uniform(0, 6)) for _ in range(3)], -1)
    return np.clip(img + rng.normal(0, 6, img.shape), 0, 255).round().astype(np.float32)
