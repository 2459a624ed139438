"""numpy restatement of the reference anchor generator (TEST INFRASTRUCTURE).

Follows network/anchors.py:21-37 (Anchors.forward), :39-70 (generate_anchors) and
:106-126 (shift): pyramid levels 3..7, stride 2^l, base size 2^(l+2), ratios {.5,1,2},
scales {1, 2^(1/3), 2^(2/3)}, float64 arithmetic, cast to float32 at the end, ordering
(level, cell row-major, anchor).  Pinned by tests/golden/anchors_*.npz (made from the
reference module itself).
"""
import numpy as np

LEVELS = (3, 4, 5, 6, 7)
RATIOS = (0.5, 1.0, 2.0)
SCALES = (2.0 ** 0, 2.0 ** (1.0 / 3.0), 2.0 ** (2.0 / 3.0))


def base_anchors(base_size):
    """anchors.py:39-70: 9 (x1,y1,x2,y2) windows centred on 0, ratio-major, scale-minor."""
    out = np.zeros((9, 4), dtype=np.float64)
    k = 0
    for r in RATIOS:
        for s in SCALES:
            side = base_size * s                 # :55   w = h = base*scale
            area = side * side                   # :58
            w = np.sqrt(area / r)                # :61
            h = w * r                            # :62
            out[k] = (0.0 - w * 0.5, 0.0 - h * 0.5, 0.0 + w * 0.5, 0.0 + h * 0.5)  # :65-66
            k += 1
    return out


def anchors_for_image(height, width):
    """anchors.py:21-37 -> float32 [A,4]."""
    chunks = []
    for lvl in LEVELS:
        stride = 2 ** lvl
        fh = (height + stride - 1) // stride     # :25
        fw = (width + stride - 1) // stride
        base = base_anchors(2 ** (lvl + 2))
        sx = (np.arange(fw, dtype=np.float64) + 0.5) * stride   # :107
        sy = (np.arange(fh, dtype=np.float64) + 0.5) * stride   # :108
        gx, gy = np.meshgrid(sx, sy)                            # row-major cells
        shifts = np.stack([gx.ravel(), gy.ravel(), gx.ravel(), gy.ravel()], axis=1)
        chunks.append((shifts[:, None, :] + base[None, :, :]).reshape(-1, 4))  # :122-124
    return np.concatenate(chunks, axis=0).astype(np.float32)


def level_cells(height, width):
    return [((height + 2 ** l - 1) // 2 ** l, (width + 2 ** l - 1) // 2 ** l) for l in LEVELS]
