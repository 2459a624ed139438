"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's algorithm for the MultiPoseNet hot path
(ResNet-FPN backbone + keypoint head + RetinaNet heads + anchors/decode/filter/NMS).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import anything from here, and only as the checker / CPU baseline.
The product (multiposenet.pytorch_b200) never imports this package.

Parity pinning: the reference ships no golden vectors or tests.
The network restatement (posenet_oracle.py) is pinned against the reference's own
Python modules imported in the build container (oracle/make_goldens.py writes
tests/golden/*.npz from /root/reference); the NMS restatement (nms_oracle.c) follows
lib/nms/src/nms.c, nms_cuda.c and cuda/nms_kernel.cu and is additionally checked on
the GPU box against the reference's own nms_kernel.cu compiled unchanged for sm_100a
(oracle/_ref/, built by oracle/build.py).  The arithmetic of conv/BN itself lives in
PyTorch (third-party, pinned pytorch=0.4.0 in multipose_environment.yaml:6); the
container's torch 2.11 fp32 CPU kernels stand in for it => "parity unpinned" at that
boundary (no reference-held known-answer vector exists).
"""
