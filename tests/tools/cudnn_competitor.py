"""The honest GPU competitor ("not required"): the reference graph itself (oracle restatement = torch
functional conv2d / batch_norm / max_pool2d, i.e. cuDNN) in eager mode on the same GPU, in fp32 (TF32 off), TF32 and bf16
autocast, for the keypoint + detection forward up to cls / reg (cfg2: no NMS, the oracle's NMS is a host loop).
Checker-side tool (imports oracle/): `python tests/tools/cudnn_competitor.py [--batch 32] [--layers 101] [--device cuda]`.
Prints one JSON line per mode: images/s, ms per forward, and max|a-b|/max|b| of the heat maps against the fp32 run."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--layers", type=int, default=101)
    ap.add_argument("--hw", type=int, nargs=2, default=[480, 640])
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--modes", default="fp32,tf32,bf16")
    a = ap.parse_args()
    from oracle import posenet_oracle as po, weights
    dev = torch.device(a.device)
    w = weights.make_weights(a.layers, "conditioned", seed=0)
    sd = {k: v.to(dev) for k, v in weights.to_torch_state_dict(w).items()}
    x = torch.from_numpy(np.random.Generator(np.random.PCG64(0)).standard_normal((a.batch, 3, a.hw[0], a.hw[1]), dtype=np.float32)).to(dev)

    def graph(xx):  # entire_net up to the head outputs (posenet.py:236-263)
        c2, c3, c4, c5 = po.backbone(sd, a.layers, xx)
        heat = po.keypoint_head(sd, *po.keypoint_neck(sd, c2, c3, c4, c5))
        cls, reg = po.detection_heads(sd, po.detection_neck(sd, c3, c4, c5))
        return heat, cls, reg

    def sync():
        if dev.type == "cuda":
            torch.cuda.synchronize()

    ref = None
    for mode in a.modes.split(","):
        tf32 = mode == "tf32"
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True
        ctx = torch.autocast(dev.type, dtype=torch.bfloat16) if mode == "bf16" else torch.autocast(dev.type, enabled=False)
        with torch.no_grad(), ctx:
            for _ in range(2):
                out = graph(x)
            sync()
            t0 = time.perf_counter()
            for _ in range(a.iters):
                out = graph(x)
            sync()
            dt = (time.perf_counter() - t0) / a.iters
        heat = out[0].float()
        if ref is None:
            ref = heat
        err = float((heat - ref).abs().max() / ref.abs().max())
        print(json.dumps({"competitor": "torch eager (cuDNN) on the oracle graph", "mode": mode, "batch": a.batch, "layers": a.layers,
                          "images_per_s": a.batch / dt, "ms_per_forward": dt * 1e3, "heat_err_vs_first_mode": err,
                          "note": "no NMS stage; weights / activations NCHW as in the reference"}), flush=True)


if __name__ == "__main__":
    main()
