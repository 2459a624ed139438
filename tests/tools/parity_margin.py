"""Parity margin table: max|a-b|/max|b| of every network output against the fp32 oracle restatement run on the same GPU
(cuDNN fp32, TF32 off) for each precision mode, R50/R101 at 480x640.  Checker-side script (imports oracle/)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", default="50,101")
    ap.add_argument("--precisions", default="bf16x3,f16f8,bf16")
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--kinds", default="conditioned,refinit")
    a = ap.parse_args()
    from gpu_util import image, load_model, nerr, no_tf32
    from oracle import posenet_oracle as po, weights
    no_tf32()
    out = []
    for layers in [int(s) for s in a.layers.split(",")]:
        for kind in a.kinds.split(","):
            x = image(7, (a.batch, 3, 480, 640))
            ref = None
            for prec in a.precisions.split(","):
                try:
                    m, w = load_model(layers, kind, prec)
                except Exception as e:  # unknown weight kind name
                    print("skip", layers, kind, prec, e)
                    break
                if ref is None:
                    sd = {k: v.cuda() for k, v in weights.to_torch_state_dict(w).items()}
                    with torch.no_grad():
                        oheat, osaved = po.forward(sd, layers, x, "keypoint_subnet")
                        _, (ocls, oreg, _) = po.forward(sd, layers, x, "detection_subnet")
                    ref = (oheat, osaved, ocls, oreg)
                with torch.no_grad():
                    heat, saved = m([x, "keypoint_subnet"])
                    _, (cls, reg, anc) = m([x, "detection_subnet"])
                row = {"layers": layers, "weights": kind, "precision": prec, "heat": nerr(heat, ref[0]),
                       "intermediate_max": max(nerr(p, q) for p, q in zip(saved[:4], ref[1][:4])),
                       "cls": nerr(cls, ref[2]), "reg": nerr(reg, ref[3])}
                out.append(row)
                print(json.dumps(row), flush=True)
                del m
                torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
