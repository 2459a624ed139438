"""ctypes front-end of oracle/nms_oracle.c (TEST INFRASTRUCTURE) restating lib/nms/pth_nms.py.

`nms_gpu_semantics`  <- pth_nms.py:25-44 (sort desc, gather, gpu_nms, order[keep]); IoU > thresh
`nms_cpu_semantics`  <- pth_nms.py:9-24  (areas, sort desc, cpu_nms); ovr >= thresh
Ties in score: the reference's torch sort is unstable; the restatement (and the product)
use a STABLE descending sort, i.e. equal scores keep their original order.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libnms_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "nms_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-fPIC", "-shared", src, "-o", _SO, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        i64p, f32p, u64p = (ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint64))
        _lib.oracle_cpu_nms.argtypes = [i64p, i64p, f32p, ctypes.c_int64, ctypes.c_int64, i64p, f32p, ctypes.c_float]
        _lib.oracle_gpu_nms.argtypes = [i64p, i64p, f32p, ctypes.c_int, ctypes.c_float, ctypes.c_int]
        _lib.oracle_nms_mask.argtypes = [f32p, ctypes.c_int, ctypes.c_float, ctypes.c_int, u64p]
        _lib.oracle_nms_mask.restype = None
        _lib.oracle_reduce_mask.argtypes = [i64p, i64p, u64p, ctypes.c_int]
        _lib.oracle_iou.argtypes = [f32p, f32p]
        _lib.oracle_iou.restype = ctypes.c_float
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def stable_desc_order(scores):
    return np.argsort(-scores.astype(np.float32), kind="stable").astype(np.int64)


def nms_gpu_semantics(dets, thresh, ge=False):
    """dets float32 [N,5] (x1,y1,x2,y2,score) -> int64 [K] indices into dets, descending score."""
    dets = np.ascontiguousarray(dets, dtype=np.float32).reshape(-1, 5)
    n = dets.shape[0]
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    order = stable_desc_order(dets[:, 4])
    sdets = np.ascontiguousarray(dets[order])
    keep = np.zeros(n, dtype=np.int64)
    num = np.zeros(1, dtype=np.int64)
    rc = lib().oracle_gpu_nms(_p(keep, ctypes.c_int64), _p(num, ctypes.c_int64), _p(sdets, ctypes.c_float), n,
                              ctypes.c_float(thresh), int(bool(ge)))
    assert rc == 1
    return order[keep[: num[0]]]


def nms_cpu_semantics(dets, thresh):
    dets = np.ascontiguousarray(dets, dtype=np.float32).reshape(-1, 5)
    n = dets.shape[0]
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    one = np.float32(1)
    areas = np.ascontiguousarray((dets[:, 2] - dets[:, 0] + one) * (dets[:, 3] - dets[:, 1] + one))
    order = stable_desc_order(dets[:, 4])
    keep = np.zeros(n, dtype=np.int64)
    num = np.zeros(1, dtype=np.int64)
    rc = lib().oracle_cpu_nms(_p(keep, ctypes.c_int64), _p(num, ctypes.c_int64), _p(dets, ctypes.c_float), n, 5,
                              _p(order, ctypes.c_int64), _p(areas, ctypes.c_float), ctypes.c_float(thresh))
    assert rc == 1
    return keep[: num[0]].copy()


def nms_mask(sorted_dets, thresh, ge=False):
    sd = np.ascontiguousarray(sorted_dets, dtype=np.float32).reshape(-1, 5)
    n = sd.shape[0]
    cb = (n + 63) // 64
    mask = np.zeros((max(n, 1), max(cb, 1)), dtype=np.uint64)
    lib().oracle_nms_mask(_p(sd, ctypes.c_float), n, ctypes.c_float(thresh), int(bool(ge)), _p(mask, ctypes.c_uint64))
    return mask[:n, :cb]


def reduce_mask(mask, n):
    mask = np.ascontiguousarray(mask, dtype=np.uint64)
    keep = np.zeros(max(n, 1), dtype=np.int64)
    num = np.zeros(1, dtype=np.int64)
    lib().oracle_reduce_mask(_p(keep, ctypes.c_int64), _p(num, ctypes.c_int64), _p(mask, ctypes.c_uint64), n)
    return keep[: num[0]].copy()


def nms_numpy_bruteforce(dets, thresh, ge=False):
    """Independent pure-numpy greedy NMS (small N) used to cross-check the C restatement."""
    dets = np.asarray(dets, dtype=np.float32).reshape(-1, 5)
    order = stable_desc_order(dets[:, 4])
    one = np.float32(1)
    alive = np.ones(len(order), dtype=bool)
    keep = []
    b = dets[order]
    area = (b[:, 2] - b[:, 0] + one) * (b[:, 3] - b[:, 1] + one)
    for i in range(len(order)):
        if not alive[i]:
            continue
        keep.append(order[i])
        w = np.maximum(np.minimum(b[i, 2], b[i + 1:, 2]) - np.maximum(b[i, 0], b[i + 1:, 0]) + one, np.float32(0))
        h = np.maximum(np.minimum(b[i, 3], b[i + 1:, 3]) - np.maximum(b[i, 1], b[i + 1:, 1]) + one, np.float32(0))
        inter = w * h
        iou = inter / (area[i] + area[i + 1:] - inter)
        alive[i + 1:] &= ~((iou >= thresh) if ge else (iou > thresh))
    return np.asarray(keep, dtype=np.int64)
