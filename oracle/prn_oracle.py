"""numpy restatement of the reference's pose-residual-network assignment step (TEST INFRASTRUCTURE).

Follows evaluate/tester.py:333-513 (`Tester.prn_process`): the joint peaks of one image are scattered into one
[56, 36, 17] grid per detected person box (:344-393), every plane is blurred with `skimage.filters.gaussian` (:396-398),
the PRN MLP runs on each grid (:400-408, network/posenet.py:337-350), every scattered peak is scored by the sum of a
15x15 window of the PRN output around its cell (:412-430, datasets/coco_data/prn_gaussian.py:122-146 `crop`), and a greedy
per-joint-type table walk gives each box at most one peak of each type (:432-470), with an arg-max fallback (:471-483).

Third-party arithmetic on this path, absent from /root/reference:
  * skimage.filters.gaussian (pinned scikit-image 0.13.1) with its defaults sigma=1, mode='nearest', truncate=4.0 is a
    call of scipy.ndimage.gaussian_filter (pinned scipy 1.1.0; this container has 1.18.1): a separable correlation,
    axis 0 then axis 1, radius int(4*sigma + 0.5) = 4, weights exp(-x^2/2)/sum in float64, evaluated by NI_Correlate1D's
    symmetric branch as  c*w[0] + sum_{j=4..1} (a[-j] + a[+j])*w[j]  in that order.  `gaussian_nearest` restates it;
    tests/test_oracle.py checks it bit-for-bit against the container's scipy.
  * np.sum over a float32 window (pinned numpy 1.14.3; 2.3.5 here): the strided 2-D window is ravelled row-major and
    summed by numpy's pairwise routine in float32 (n < 8: sequential; 8 <= n <= 128: eight accumulators, tree-combined,
    then the remainder sequentially; n > 128: split at (n/2 rounded down to a multiple of 8), recurse).
    `window_sum_f32` restates it; tests/test_oracle.py checks it against np.sum for every window position.
  * Python set iteration order (tester.py:438 `list(set(...))`) and np.argsort's tie order (:455,:460,:468) are
    implementation details; this restatement orders the table columns by ascending peak id and breaks score ties by the
    lower index (kind='stable').  Only exact ties between positive scores can make that visible.

Pinned by tests/golden/prn_assign.npz: outputs of the reference's own Tester.prn_process run in this container on seeded
cases (oracle/make_goldens.py; skimage is not installed, so its one-line call of scipy is shimmed there).
"""
import math

import numpy as np

NUM_JOINTS = 17


def grid_size(coeff=2):
    """tester.py:353-354 -> (h, w) = (28*coeff, 18*coeff)."""
    return int(28 * coeff), int(18 * coeff)


def sort_peaks(kps):
    """tester.py:337-350: peaks regrouped by joint type, id = running index in that order.
    kps rows (x, y, score, id, joint_type 0..16).  Returns float64 [n, 2] (x, y), int [n] types, in id order."""
    kps = np.asarray(kps, dtype=np.float64).reshape(-1, 5) if len(kps) else np.zeros((0, 5))
    xy, ty = [], []
    for j in range(NUM_JOINTS):
        for k in kps:
            if k[-1] == j:
                xy.append((k[0], k[1]))
                ty.append(j)
    return np.array(xy, dtype=np.float64).reshape(-1, 2), np.array(ty, dtype=np.int64)


def boxes_xywh(bbox_list):
    """tester.py:356-358: (x1, y1, x2 - x1, y2 - y1) in float64."""
    return np.array([[b[0], b[1], b[2] - b[0], b[3] - b[1]] for b in bbox_list], dtype=np.float64).reshape(-1, 4)


def scatter(xy, ty, boxes, coeff=2, in_thres=0.21):
    """tester.py:363-393.  Returns owner int32 [P, 17, h, w]: id of the peak whose one-hot sits in the cell (the last
    writer in id order), -1 if none."""
    h, w = grid_size(coeff)
    owner = np.full((len(boxes), NUM_JOINTS, h, w), -1, dtype=np.int32)
    for k in range(len(xy)):
        p_x, p_y = xy[k]
        for bi, b in enumerate(boxes):
            inside = (p_x > b[0] - b[2] * in_thres and p_y > b[1] - b[3] * in_thres and
                      p_x < b[0] + b[2] * (1.0 + in_thres) and p_y < b[1] + b[3] * (1.0 + in_thres))
            if not inside:
                continue
            x_scale = float(w) / math.ceil(b[2])
            y_scale = float(h) / math.ceil(b[3])
            x0 = int((p_x - b[0]) * x_scale)
            y0 = int((p_y - b[1]) * y_scale)
            # :377-390 -- an elif chain: only ONE of the corrections is applied
            if x0 >= w and y0 >= h:
                x0, y0 = w - 1, h - 1
            elif x0 >= w:
                x0 = w - 1
            elif y0 >= h:
                y0 = h - 1
            elif x0 < 0 and y0 < 0:
                x0, y0 = 0, 0
            elif x0 < 0:
                x0 = 0
            elif y0 < 0:
                y0 = 0
            # an index left negative by the chain wraps (numpy indexing), :392
            owner[bi, ty[k], y0 % h if y0 < 0 else y0, x0 % w if x0 < 0 else x0] = k
    return owner


def gaussian_weights(sigma=1.0, truncate=4.0):
    """scipy.ndimage._gaussian_kernel1d: float64 weights, index = distance from the centre."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    phi = phi / phi.sum()
    return phi[radius:].copy()


def _correlate_nearest(a, axis, wts):
    r = len(wts) - 1
    a = np.moveaxis(a, axis, 0)
    n = a.shape[0]
    pad = np.concatenate([np.repeat(a[:1], r, 0), a, np.repeat(a[-1:], r, 0)], 0)
    out = pad[r:r + n] * wts[0]
    for j in range(r, 0, -1):                       # NI_Correlate1D: jj = -size1 .. -1
        out = out + (pad[r - j:r - j + n] + pad[r + j:r + j + n]) * wts[j]
    return np.moveaxis(out, 0, axis)


def gaussian_nearest(plane, sigma=1.0):
    """skimage.filters.gaussian(plane) == scipy.ndimage.gaussian_filter(plane, 1, mode='nearest', truncate=4.0)."""
    wts = gaussian_weights(sigma)
    a = np.asarray(plane, dtype=np.float64)
    return _correlate_nearest(_correlate_nearest(a, 0, wts), 1, wts)


def build_inputs(owner):
    """tester.py:396-403: one-hot planes -> gaussian -> float32 [P, h, w, 17] (the `.float()` PRN input)."""
    P, J, h, w = owner.shape
    inp = np.zeros((P, h, w, J), dtype=np.float32)
    for p in range(P):
        for t in range(J):
            inp[p, :, :, t] = gaussian_nearest((owner[p, t] >= 0).astype(np.float64)).astype(np.float32)
    return inp


def window_bounds(y, x, h, w, N=15):
    """prn_gaussian.py:122-146 `crop(img, (y, x), N)`: rows [r0, r1), columns [c0, c1)."""
    hh = (N - 1) / 2
    r0, c0 = int(y - hh), int(x - hh)
    r1, c1 = int(y + hh) + 1, int(x + hh) + 1
    r0, c0 = max(r0, 0), max(c0, 0)
    if r1 > h - 1:
        r1 = h
    if c1 > w - 1:
        c1 = w
    return r0, r1, c0, c1


def pairwise_sum_f32(a):
    """numpy's pairwise summation (umath loops `pairwise_sum`) of a 1-D float32 sequence."""
    f = np.float32
    n = len(a)
    if n < 8:
        res = f(0.0)
        for v in a:
            res = f(res + v)
        return res
    if n <= 128:
        r = [f(v) for v in a[:8]]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = f(r[j] + a[i + j])
            i += 8
        res = f(f(f(r[0] + r[1]) + f(r[2] + r[3])) + f(f(r[4] + r[5]) + f(r[6] + r[7])))
        while i < n:
            res = f(res + a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return f(pairwise_sum_f32(a[:n2]) + pairwise_sum_f32(a[n2:]))


def window_sum_f32(plane, r0, r1, c0, c1):
    """np.sum(plane[r0:r1, c0:c1]) of a strided float32 view: numpy ravels the window (row-major copy) and runs its
    pairwise sum over the rows*cols <= 225 values."""
    return pairwise_sum_f32(np.ascontiguousarray(plane[r0:r1, c0:c1]).ravel())


def scores(owner, output, n_peaks):
    """tester.py:412-430: S[p, k] = float64(np.sum(window of output[p, :, :, t] around the cell of peak k)) for every cell
    that holds a peak (kp_score == 1), hit[p, k] = True there.  output float32 [P, h, w, 17]."""
    P, J, h, w = owner.shape
    S = np.zeros((P, n_peaks), dtype=np.float64)
    hit = np.zeros((P, n_peaks), dtype=bool)
    for p, t, y, x in np.argwhere(owner >= 0):
        r0, r1, c0, c1 = window_bounds(y, x, h, w)
        k = owner[p, t, y, x]
        S[p, k] = float(np.sum(output[p, r0:r1, c0:c1, t]))
        hit[p, k] = True
    return S, hit


def assign(xy, ty, boxes, owner, output, coeff=2):
    """tester.py:432-483.  Returns bbox_keypoints float64 [P, 17, 3]."""
    h, w = grid_size(coeff)
    P = len(boxes)
    S, hit = scores(owner, output, len(ty))
    out = np.zeros((P, NUM_JOINTS, 3), dtype=np.float64)
    for i in range(NUM_JOINTS):
        cols = [k for k in range(len(ty)) if ty[k] == i and hit[:, k].any()]   # kp_ids (:438), ascending id
        if cols:
            table = np.where(hit[:, cols], S[:, cols], 0.0)                    # :440-449
            for bbox in range(P):
                row = np.argsort(-table[bbox], kind="stable")                  # :455
                if table[bbox, row[0]] > 0:
                    for r in row:
                        if table[bbox, r] > 0:
                            column = np.argsort(-table[:, r], kind="stable")   # :460
                            if bbox == column[0]:
                                out[bbox, i] = (xy[cols[r], 0], xy[cols[r], 1], 1.0)
                                break
                            row2 = np.argsort(table[column[0]], kind="stable")  # :468
                            if row2[0] == r:
                                out[bbox, i] = (xy[cols[r], 0], xy[cols[r], 1], 1.0)
                                break
        else:                                                                  # :471-483
            for j in range(P):
                b = boxes[j]
                x_scale = float(w) / math.ceil(b[2])
                y_scale = float(h) / math.ceil(b[3])
                for t in range(NUM_JOINTS):
                    if not (owner[j, t] >= 0).any():
                        plane = output[j, :, :, t]
                        my, mx = np.argwhere(plane == np.max(plane))[0]
                        out[j, t] = (mx / x_scale + b[0], my / y_scale + b[1], 0.0)
    return out


def results(bbox_keypoints, boxes, file_name="", image_id=0):
    """tester.py:485-511: the per-person records."""
    res = []
    for i in range(bbox_keypoints.shape[0]):
        k = np.zeros(51)
        k[0::3], k[1::3], k[2::3] = bbox_keypoints[i, :, 0], bbox_keypoints[i, :, 1], bbox_keypoints[i, :, 2]
        pose_score = 0
        for f in range(NUM_JOINTS):
            pose_score += bbox_keypoints[i, f, 2]
        pose_score /= 17.0
        res.append({"image_id": image_id, "file_name": file_name, "category_id": 1, "bbox": [float(v) for v in boxes[i]],
                    "score": float(pose_score), "keypoints": k.tolist()})
    return res


def prn_process(kps, bbox_list, prn_fn, file_name="", image_id=0, coeff=2, in_thres=0.21):
    """The whole of Tester.prn_process; prn_fn maps float32 [P, h, w, 17] inputs to the PRN outputs of the same shape."""
    xy, ty = sort_peaks(kps)
    boxes = boxes_xywh(bbox_list)
    if len(boxes) == 0:                                                         # :360 (len(peaks) is always 17)
        return []
    owner = scatter(xy, ty, boxes, coeff, in_thres)
    output = np.asarray(prn_fn(build_inputs(owner)), dtype=np.float32)
    return results(assign(xy, ty, boxes, owner, output, coeff), boxes, file_name, image_id)


def synthetic_case(seed, persons=5, hw=(480, 480), coeff=2, extra_boxes=1, drop_joint=None, noise_peaks=6):
    """Seeded people: a box and 17 joints inside it (some missing, some shared cells, some outside every box)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    H, W = hw
    kps, boxes = [], []
    for _ in range(persons):
        bw, bh = rng.uniform(40, 160), rng.uniform(90, 300)
        x1, y1 = rng.uniform(0, max(W - bw, 1.0)), rng.uniform(0, max(H - bh, 1.0))
        boxes.append([x1, y1, x1 + bw, y1 + bh])
        for j in range(NUM_JOINTS):
            if j == drop_joint or rng.random() < 0.15:
                continue
            kps.append([float(np.floor(rng.uniform(x1 - 0.2 * bw, x1 + 1.2 * bw))), float(np.floor(rng.uniform(y1 - 0.2 * bh, y1 + 1.2 * bh))),
                        rng.uniform(0.1, 1.0), 0.0, float(j)])
    for _ in range(extra_boxes):                                   # a box with (almost) no joints of its own
        bw, bh = rng.uniform(30, 80), rng.uniform(40, 120)
        x1, y1 = rng.uniform(0, max(W - bw, 1.0)), rng.uniform(0, max(H - bh, 1.0))
        boxes.append([x1, y1, x1 + bw, y1 + bh])
    for _ in range(noise_peaks):
        j = int(rng.integers(0, NUM_JOINTS))
        if j != drop_joint:
            kps.append([float(rng.integers(0, W)), float(rng.integers(0, H)), rng.uniform(0.1, 1.0), 0.0, float(j)])
    order = rng.permutation(len(kps))                              # the joint list arrives grouped by type; keep it general
    kps = [kps[i] for i in order]
    kps.sort(key=lambda r: r[-1])
    for i, r in enumerate(kps):
        r[3] = float(i)
    return kps, boxes


def synthetic_prn(seed, sharp=6.0):
    """A deterministic, batch-size independent stand-in for the PRN: per person, softmax over the whole grid of
    (sharp * input + seeded noise) -- the same noise field for every person."""
    def fn(inp):
        inp = np.asarray(inp, dtype=np.float32)
        noise = np.random.Generator(np.random.PCG64(seed)).standard_normal(inp.shape[1:], dtype=np.float32)
        z = (np.float32(sharp) * inp + noise[None]).reshape(len(inp), -1)
        e = np.exp(z - z.max(1, keepdims=True))
        return (e / e.sum(1, keepdims=True)).astype(np.float32).reshape(inp.shape)
    return fn
