"""numpy restatement of the reference image normalisation (TEST INFRASTRUCTURE).

Follows datasets/coco_data/preprocessing.py:15-26 (resnet_preprocess): float32(image)/255, BGR -> RGB, subtract the
ImageNet mean, divide by the std (all float32 ops), HWC -> CHW.  Pinned by tests/golden/preprocess.npz (made by
importing the reference function).
"""
import numpy as np

MEANS = (0.485, 0.456, 0.406)
STDS = (0.229, 0.224, 0.225)


def resnet_preprocess(image):
    """image: uint8 or float [H,W,3] BGR -> float32 [3,H,W]."""
    x = image.astype(np.float32) / np.float32(255.0)           # :16
    x = x[:, :, ::-1].copy()                                   # :20  BGR -> RGB
    for i in range(3):                                         # :21-23
        x[:, :, i] = (x[:, :, i] - np.float32(MEANS[i])) / np.float32(STDS[i])
    return np.ascontiguousarray(x.transpose(2, 0, 1)).astype(np.float32)   # :25
