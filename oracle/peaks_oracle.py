"""numpy restatement of the reference heat-map peak extraction (TEST INFRASTRUCTURE).

Follows network/joint_utils.py: find_peaks (:19-32: 3x3-cross maximum filter == value, value > thre1), NMS (:61-138:
per joint type, every peak's <=5x5 patch is upsampled x upsampFactor with cv2.INTER_CUBIC, the arg-max of the upsampled
patch refines the location and gives the score, ids count up over joint types) and get_joint_list (:141-152: rows
(x, y, score, id, joint_type)), as called by evaluate/tester.py:219 with upsampFactor = 480/120 = 4.

cv2.resize is a third-party dependency (the reference's environment pins opencv 3.4.x; this container has 4.13): its
INTER_CUBIC algorithm for float32 is restated here from the OpenCV sources -- half-pixel centres
(fx = (dx + 0.5)/factor - 0.5), Keys cubic with A = -0.75, clamped (replicated) border taps, horizontal pass then
vertical pass -- in plain unfused float32 arithmetic.  The container's cv2 build (SIMD/IPP dispatch) differs from this
restatement by <= 4e-7 absolute on [0,1] maps; tests/golden/peaks.npz (made by running the reference's own
get_joint_list here, oracle/make_goldens.py) pins the restatement to that tolerance: coordinates, ids and joint types
equal, scores within 1e-6.
"""
import numpy as np

f32 = np.float32


def find_peaks(img, thre1):
    """joint_utils.py:19-32.  scipy's maximum_filter uses mode='reflect', so an out-of-range neighbour repeats the
    border pixel itself: a pixel is a peak iff it is >= its in-range 4-neighbours and > thre1.  Returns [[x, y], ...]
    in np.nonzero order (row-major)."""
    pad = np.pad(img, 1, mode="edge")
    m = np.maximum.reduce([pad[1:-1, 1:-1], pad[:-2, 1:-1], pad[2:, 1:-1], pad[1:-1, :-2], pad[1:-1, 2:]])
    binary = (m == img) & (img > thre1)
    ys, xs = np.nonzero(binary)
    return np.stack([xs, ys], 1) if len(xs) else np.zeros((0, 2), dtype=np.int64)


def cubic_coeffs(x):
    """OpenCV interpolateCubic (imgproc/resize.cpp), float32, A = -0.75."""
    x = f32(x)
    A, one = f32(-0.75), f32(1)
    c0 = ((A * (x + one) - f32(5) * A) * (x + one) + f32(8) * A) * (x + one) - f32(4) * A
    c1 = ((A + f32(2)) * x - (A + f32(3))) * x * x + one
    c2 = ((A + f32(2)) * (one - x) - (A + f32(3))) * (one - x) * (one - x) + one
    c3 = one - c0 - c1 - c2
    return f32(c0), f32(c1), f32(c2), f32(c3)


def resize_tables(n_src, factor):
    """Per destination index: (first tap index sx - 1, 4 coefficients)."""
    n_dst = int(round(n_src * factor))
    tab = []
    for d in range(n_dst):
        fx = f32((d + 0.5) * (1.0 / factor) - 0.5)  # double expression, then float (resize.cpp)
        sx = int(np.floor(fx))
        tab.append((sx - 1, cubic_coeffs(f32(fx - f32(sx)))))
    return tab


def resize_cubic(patch, factor):
    """cv2.resize(patch, None, fx=factor, fy=factor, interpolation=cv2.INTER_CUBIC) for a small float32 patch."""
    patch = np.asarray(patch, dtype=np.float32)
    h, w = patch.shape
    tx, ty = resize_tables(w, factor), resize_tables(h, factor)
    hr = np.zeros((h, len(tx)), dtype=np.float32)
    for y in range(h):
        for d, (s0, c) in enumerate(tx):
            S = [patch[y, min(max(s0 + j, 0), w - 1)] for j in range(4)]
            hr[y, d] = f32(f32(f32(f32(S[0] * c[0]) + f32(S[1] * c[1])) + f32(S[2] * c[2])) + f32(S[3] * c[3]))
    up = np.zeros((len(ty), len(tx)), dtype=np.float32)
    for e, (s0, b) in enumerate(ty):
        rows = [hr[min(max(s0 + k, 0), h - 1)] for k in range(4)]
        for d in range(len(tx)):
            up[e, d] = f32(f32(f32(f32(rows[0][d] * b[0]) + f32(rows[1][d] * b[1])) + f32(rows[2][d] * b[2])) + f32(rows[3][d] * b[3]))
    return up


def joint_list(heatmaps, thre1=0.1, factor=4, win_size=2):
    """heatmaps [C, H, W] float32 -> float64 [P, 5] rows (x, y, score, id, joint_type): NMS (:61-138) with
    bool_refine_center=True, bool_gaussian_filt=False, followed by get_joint_list's flattening (:148-150, before the
    caller's `* scale`)."""
    rows = []
    cnt = 0
    for joint in range(heatmaps.shape[0]):
        m = np.asarray(heatmaps[joint], dtype=np.float32)
        h, w = m.shape
        for (px, py) in find_peaks(m, thre1):
            x_min, y_min = max(0, px - win_size), max(0, py - win_size)            # :100
            x_max, y_max = min(w - 1, px + win_size), min(h - 1, py + win_size)    # :101-102
            up = resize_cubic(m[y_min:y_max + 1, x_min:x_max + 1], factor)          # :106-108
            iy, ix = np.unravel_index(up.argmax(), up.shape)                        # :117-118 (first maximum)
            # :121-126 and :131-132: (p + 0.5)*f - 0.5 + (i - ((p - p_min + 0.5)*f - 0.5)) == p_min*f + i exactly
            rows.append((float(x_min * factor + ix), float(y_min * factor + iy), float(up[iy, ix]), float(cnt), float(joint)))
            cnt += 1
    return np.array(rows, dtype=np.float64).reshape(-1, 5)


def synthetic_heatmaps(seed, C=18, H=120, W=160, persons=6):
    """Sums of Gaussians at sub-pixel centres (some on the border), exact plateaus and a bump exactly at the threshold."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
    heat = np.zeros((C, H, W), dtype=np.float32)
    for c in range(C):
        for _ in range(persons):
            cx, cy = rng.uniform(-1, W), rng.uniform(-1, H)
            amp, sig = rng.uniform(0.05, 1.0), rng.uniform(1.0, 2.5)
            heat[c] += (amp * np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / (2 * sig * sig))).astype(np.float32)
        heat[c] += (rng.random((H, W), dtype=np.float32) * np.float32(0.004))
    heat[0, 0, 0] = 0.9                                            # corner peak: 3x3 patch
    heat[1, H - 1, 7] = 0.8                                        # bottom-edge peak
    heat[2, H // 3, W // 3] = heat[2, H // 3, W // 3 + 1] = 0.75   # two-pixel plateau: both are peaks
    heat[3, H // 2, W - 1] = 0.7                                   # right-edge peak
    heat[4, 10, 10] = np.float32(0.1)                              # exactly at the threshold: not a peak (strict >)
    return heat
