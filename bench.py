#!/usr/bin/env python
"""bench.py -- images/sec of the MultiPoseNet hot path (R101-FPN backbone -> keypoint + RetinaNet heads ->
decode / filter / NMS) on synthetic 3x480x640 batches, BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--precision f16f8|bf16x3|bf16|fp32]
    (N > 1: launched by torch.distributed.run, one rank per GPU; inference shards by image, no collective)

One "step" = one forward of batch 32/GPU through poseNet's entire_net graph incl. NMS.
  value : images/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e   : images/s through the public call model((img, 'both')) with pinned HOST input, H2D + D2H in the region
  roofline      : conv_tc_kernel (all tcgen05 conv launches of a step): algorithmic conv FLOPs / summed duration
  cpu_baseline  : the oracle port of the reference graph on the host cores (bounded sample)
--impl reference times that CPU port alone (the reference's Python cannot travel to the GPU box).

The same JSON line also carries (unless --no-extras), so that the one command the driver runs measures every BASELINE config:
  parity        : max|a-b|/max|b| of heat / cls / reg (and the kept-box count) of image 0 of the timed batch against the fp32
                  CPU oracle output the cpu_baseline leg computes anyway; the run FAILS if one exceeds 1e-3
  cpu_cfg1      : BASELINE configs[0]: R50 keypoint_subnet, batch 1, CPU oracle port (+ the same call on the GPU, host in/out)
  roofline_aux  : decode / filter / sort / gather / IoU mask / greedy reduction (bench load and the cfg3 100-person feed) and the
                  batched PRN forward: algorithmic bytes / CUDA-event time against the measured HBM copy bandwidth
  train_step    : BASELINE configs[3]: keypoint-subnet training step, batch 16/GPU, NCCL gradient allreduce at N ranks
  full_pipeline : BASELINE configs[4]: entire_net + NMS + heat-map peaks + PRN assignment, batch 64/GPU, host images in
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

LAYERS = 101
H, W = 480, 640
METRIC = "images/sec (3x480x640) full PoseNet fwd+NMS"
# Class-head bias shift that puts ~3700 of the 57600 anchors per N(0,1) image above the 0.05 score filter for the
# seeded R101 "conditioned" weights (SURVEY 8(d) cfg3 regime).  Calibrated once with calibrate_cls_bias() on the
# B200 (profiles/r01d_bench_n1.json: cls_bias_shift) and frozen so both arms of the bench use identical weights.
CLS_BIAS_SHIFT = {101: -0.7387505}


def load_weights_into(model, layers):
    from multiposenet.pytorch_b200 import synthetic  # seeded synthetic weights (data only)
    w = synthetic.make_weights(layers, "conditioned", seed=0)
    sd = model.state_dict()
    for k in sd:
        if k in w:
            sd[k] = torch.from_numpy(np.ascontiguousarray(w[k]))
    model.load_state_dict(sd)
    return w


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def conv_traffic_record():
    """roofline.traffic: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per conv_tc_kernel launch, averaged over
    the launches of one forward step, from the newest committed ncu pass (scripts/ncu_conv_step.py + summarise_ncu_csv.py)."""
    import glob
    recs = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_conv_traffic.json")))
    if not recs:
        return None, None
    d = json.load(open(recs[-1]))
    return d.get("traffic_bytes_per_launch"), "profiles/%s (%d launches under ncu)" % (os.path.basename(recs[-1]), d.get("launches", 0))


class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def available_cpus():
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:  # cgroup v2 quota
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(float(q) / float(per))))
    except Exception:
        pass
    return n


def pick_cpu_threads(sd):
    """The oracle port runs on torch's CPU (oneDNN) kernels; on a many-core host the best thread count is not
    always 'all of them' (oversubscription).  Probe a small forward at a few counts and keep the fastest."""
    from oracle import posenet_oracle as po
    navail = available_cpus()
    cands = sorted({navail, max(1, navail // 2), max(1, navail // 4), min(navail, 32), min(navail, 16), min(navail, 8)}, reverse=True)
    x = torch.randn(1, 3, 160, 224)
    best = (None, 1e30)
    for t in cands:
        torch.set_num_threads(t)
        with torch.no_grad():
            po.forward(sd, LAYERS, x, "keypoint_subnet")
            t0 = time.perf_counter()
            po.forward(sd, LAYERS, x, "keypoint_subnet")
            dt = time.perf_counter() - t0
        if dt < best[1]:
            best = (t, dt)
    torch.set_num_threads(best[0])
    return best[0], navail


def calibrate_cls_bias(model, dev, target=3600):
    """Synthetic-weight construction: shift the class-head bias so that ~`target` of the
    57600 anchors per image pass the 0.05 score filter on N(0,1) images."""
    import math
    x = torch.from_numpy(np.random.Generator(np.random.PCG64(99)).standard_normal((2, 3, H, W), dtype=np.float32)).to(dev)
    with torch.no_grad():
        _, (cls, _, _) = model((x, "detection_subnet"))
    p = cls[:, :, 0].double().clamp(1e-12, 1 - 1e-12)
    logit = torch.log(p / (1 - p))
    q = 1.0 - float(target) / logit.shape[1]
    cut = float(torch.quantile(logit.flatten()[:: max(1, logit.numel() // 2000000)].float(), q))
    shift = math.log(0.05 / 0.95) - cut
    with torch.no_grad():
        model.classificationModel.output.bias += shift
    return shift


def cpu_reference_step(sd, x):
    """The reference graph (oracle port) on the host: entire_net incl. NMS for the images in x."""
    from oracle import posenet_oracle as po
    with torch.no_grad():
        return po.forward(sd, LAYERS, x, "both")


def run_reference(args, rank, world):
    from oracle import weights
    if rank != 0:
        return
    w = weights.make_weights(LAYERS, "conditioned", seed=0)
    sd = weights.to_torch_state_dict(w)
    sd["classificationModel.output.bias"] = sd["classificationModel.output.bias"] + CLS_BIAS_SHIFT[LAYERS]
    nthreads, navail = pick_cpu_threads(sd)
    nimg = 1
    x = torch.from_numpy(np.random.Generator(np.random.PCG64(0)).standard_normal((nimg, 3, H, W), dtype=np.float32))
    steps, warm = max(1, min(args.steps, 8)), max(1, min(args.warmup, 2))
    for _ in range(warm):
        cpu_reference_step(sd, x)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter(); cpu_reference_step(sd, x); ts.append(time.perf_counter() - t0)
    med = float(np.median(ts))
    v = nimg / med
    sample = "%d step(s) of %d image(s) (R101 entire_net + NMS, batch %d; full workload is batch 32/GPU); %d of %d host threads (fastest of a probe)" % (steps, nimg, nimg, nthreads, navail)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": "R101-FPN entire_net fwd + decode/filter/NMS, 3x480x640", "batch_per_step": nimg},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def freeze_for_keypoint_training(model):
    """training/multipose_keypoint_train.py:78-89: detection neck / heads and the PRN are frozen."""
    for name, mod in model.fpn.named_children():
        if name in ("conv6", "conv7", "latlayer1", "latlayer2", "latlayer3", "toplayer0", "toplayer1", "toplayer2"):
            for p in mod.parameters():
                p.requires_grad = False
    for name, mod in model.named_children():
        if name in ("regressionModel", "classificationModel", "prn"):
            for p in mod.parameters():
                p.requires_grad = False


def train_leg(args, rank, world, local, steps, warmup, precision=None):
    """BASELINE config 4: keypoint-subnet training step (fwd + bwd + NCCL gradient allreduce + Adam), batch 16/GPU.
    Collective: every rank calls it; returns the record on rank 0 (None elsewhere)."""
    from multiposenet.pytorch_b200 import poseNet, shard
    dev = torch.device("cuda", local)
    precision = precision or ("bf16x3" if args.precision in (None, "f16f8", "fp32") else args.precision)
    B = args.train_batch
    model = poseNet(args.layers, precision=precision)
    load_weights_into(model, args.layers)
    model = model.to(dev).train()
    freeze_for_keypoint_training(model)
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-4)
    eng = model.train_engine()
    rng = np.random.Generator(np.random.PCG64(4321 + rank))
    x = torch.from_numpy(rng.standard_normal((B, 3, H, W), dtype=np.float32)).to(dev)
    gt = torch.from_numpy(rng.random((B, 18, H // 4, W // 4), dtype=np.float32)).to(dev)
    wt = torch.from_numpy((rng.random((B, 18, H // 4, W // 4)) > 0.2).astype(np.float32)).to(dev)
    comm_ms = []

    def step():
        if hasattr(eng, "train_step_overlapped") and args.graph and args.overlap:
            loss = eng.train_step_overlapped(x, gt, wt, world)   # bucketed allreduce under the tail of the backward
        else:
            if args.graph:
                loss, outs, grads = eng.graphed_forward_backward(x, gt, wt)  # one cudaGraphLaunch for fwd + bwd
            else:
                loss, outs, grads = eng.forward_backward(x, gt, wt)
            eng.assign_grads(grads, world)   # one NCCL allreduce over the flat fp32 gradient (no-op at world 1)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    losses = []
    for _ in range(max(warmup, 3)):
        losses.append(step().clone())
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        losses.append(step().clone())
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    value = shard.whole_job_rate(B * steps, ms, dev)
    ms = shard.max_over_ranks(ms, dev)
    nparam = sum(p.numel() for p in model.parameters() if p.requires_grad)
    # the allreduce alone (same flat buffer size), timed after the step loop: what an un-overlapped collective costs
    ar_ms = None
    if world > 1:
        import torch.distributed as dist
        flat = torch.zeros(nparam, dtype=torch.float32, device=dev)
        for _ in range(2):
            dist.all_reduce(flat)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            dist.all_reduce(flat)
        e1.record()
        torch.cuda.synchronize()
        ar_ms = shard.max_over_ranks(e0.elapsed_time(e1) / 5, dev)
        del flat
    rec = None
    if rank == 0:
        pk, pk_src = peaks()
        alg = 3.0 * 198.23e9 if args.layers == 101 else 3.0 * 152.78e9  # fwd + dgrad + wgrad of the keypoint sub-graph
        ach = alg * B * world * steps / (ms / 1e3) / 1e12 / world
        rec = {
            "mode": "train", "metric": "images/sec keypoint-subnet training step (fwd+bwd+allreduce+Adam), 3x480x640", "value": value,
            "unit": "images/s", "n_gpus": world, "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms / steps,
            "wall_ms_per_step": wall * 1e3 / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": precision, "data": "synthetic",
            "config": {"workload": "R%d keypoint_subnet train step, batch %d/GPU, BN train mode, Adam lr 1e-4" % (args.layers, B),
                       "global_batch": B * world, "cuda_graph": bool(args.graph),
                       "allreduce_overlapped": bool(hasattr(eng, "train_step_overlapped") and args.graph and args.overlap and world > 1),
                       "parallelism": "dp%d, NCCL allreduce of %d fp32 gradients (%.1f MB) per step" % (world, nparam, nparam * 4 / 1e6)},
            "allreduce_alone_ms": ar_ms, "allreduce_bytes": nparam * 4,
            "loss_first": float(losses[0]), "loss_last": float(losses[-1]), "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": ach, "peak": pk.get("bf16_tflops_sustained"), "unit": "TFLOP/s",
                         "frac": ach / pk.get("bf16_tflops_sustained"), "note": "per-GPU algorithmic FLOPs = 3 x forward of the trainable graph"}}
    del eng, opt, model
    torch.cuda.empty_cache()
    return rec


def run_train(args, rank, world, local):
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.batch != 32:
        args.train_batch = args.batch
    rec = train_leg(args, rank, world, local, args.steps, args.warmup, precision=args.precision)
    if rank == 0:
        print(json.dumps(rec))
    if world > 1:
        dist.destroy_process_group()


def full_leg(args, rank, world, local, steps, warmup, precision=None):
    """BASELINE config 5: full inference incl. heat-map peaks and PRN assignment (the body of Tester._process,
    evaluate/tester.py:200-243) for a batch of 64 images per GPU, through evaluate.process_batch with HOST images in and
    per-person records out.  Synthetic-weight construction: the class-head bias of the headline bench (~3700 candidates per
    image), the best `--persons` NMS survivors per image as person boxes, and a convfin bias shift that lets ~170 heat-map
    peaks per image pass thre1.  Every rank calls it; the record comes back on rank 0."""
    from multiposenet.pytorch_b200 import ops, poseNet, shard, synthetic
    from multiposenet.pytorch_b200.evaluate import process_batch
    dev = torch.device("cuda", local)
    precision = precision or args.precision
    B = args.full_batch
    model = poseNet(args.layers, precision=precision)
    load_weights_into(model, args.layers)
    model = model.to(dev).eval()
    rng = np.random.Generator(np.random.PCG64(777 + rank))
    host = [torch.from_numpy(rng.standard_normal((B, 3, H, W), dtype=np.float32)).pin_memory() for _ in range(2)]
    probe = host[0][:4].to(dev)
    # Detection load = the headline bench's (cfg3 regime: ~3700 anchors per image above the 0.05 filter, ~490 kept by NMS);
    # the best `--persons` kept boxes per image go to the PRN (random weights have no meaningful 0.5 score level, and a
    # bias that lifts 80 anchors above 0.5 lifts > 20000 above 0.05, i.e. benchmarks a 20000-box NMS: r01p/r01t, 420 img/s).
    if args.layers in CLS_BIAS_SHIFT:
        cls_shift = CLS_BIAS_SHIFT[args.layers]
        with torch.no_grad():
            model.classificationModel.output.bias += cls_shift
    else:
        cls_shift = calibrate_cls_bias(model, dev)
    shifts = {"cls": cls_shift,
              "heat": synthetic.calibrate_output_bias(model, probe, "heat", per_image=600, threshold=0.1)}
    scales = [1.0] * B
    stats = {}

    copy_stream = torch.cuda.Stream(device=dev)
    pending = {}

    def prefetch(i):
        """H2D of batch i from pinned host memory on a copy stream: overlaps the previous batch's kernels and host work."""
        with torch.cuda.stream(copy_stream):
            xb = host[i % 2].to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return xb, ev

    def step(i):
        x, ev = pending.pop(i, None) or prefetch(i)
        torch.cuda.current_stream().wait_event(ev)
        x.record_stream(torch.cuda.current_stream())
        pending[i + 1] = prefetch(i + 1)
        recs, heat, det = process_batch(model, x, scales, max_persons=args.persons, box_score_thresh=0.05)
        stats["candidates"] = float(det.cand_cnt.float().mean())
        stats["persons"] = sum(len(r) for r in recs) / float(B)
        stats["assigned"] = sum(1 for r in recs for q in r for v in q["keypoints"][2::3] if v > 0) / float(B)
        return recs

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(warmup, 3)):
        step(i)
    barrier()
    pending.clear()  # every timed step uploads its own batch inside the timed region
    l0 = ops.stats["launches"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    value = shard.whole_job_rate(B * steps, ms, dev)
    ms = shard.max_over_ranks(ms, dev)
    rec = None
    if rank == 0:
        rec = {
            "mode": "full", "metric": "images/sec (3x480x640) full inference incl. peaks + PRN assignment, host images in, records out",
            "value": value, "unit": "images/s", "n_gpus": world, "steps": steps, "warmup": max(warmup, 3),
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": precision,
            "data": "synthetic",
            "config": {"workload": "R%d entire_net + NMS + heat-map peaks + PRN assignment, batch %d/GPU, 3x480x640" % (args.layers, B),
                       "global_batch": B * world, "persons_per_image": stats.get("persons"), "assigned_joints_per_image": stats.get("assigned"),
                       "candidates_per_image": stats.get("candidates"), "box_score_thresh": 0.05,
                       "max_persons": args.persons, "bias_shifts": shifts, "parallelism": "dp%d (image shards, no collective)" % world},
            "h2d_bytes_per_step": host[0].numel() * 4, "gpu_launches": ops.stats["launches"] - l0, "clocks": clocks}
    pending.clear()
    del model, host
    torch.cuda.empty_cache()
    return rec


def run_full(args, rank, world, local):
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.batch != 32:
        args.full_batch = args.batch
    rec = full_leg(args, rank, world, local, args.steps, args.warmup)
    if rank == 0:
        print(json.dumps(rec))
    if world > 1:
        dist.destroy_process_group()


def cpu_cfg1_leg(dev):
    """BASELINE configs[0]: ResNet50-FPN keypoint subnet forward, batch 1, 3x480x640 on the host cores (the oracle port of the
    evaluate/multipose_keypoint_val.py forward, seeded weights), and the same call through the public API on the GPU with the
    host copies inside the timed region."""
    from multiposenet.pytorch_b200 import poseNet
    from oracle import posenet_oracle as po, weights   # checker-side: the CPU baseline leg
    w = weights.make_weights(50, "conditioned", seed=0)
    sd = weights.to_torch_state_dict(w)
    x = torch.from_numpy(np.random.Generator(np.random.PCG64(50)).standard_normal((1, 3, H, W), dtype=np.float32))
    with torch.no_grad():
        ref_heat, _ = po.forward(sd, 50, x, "keypoint_subnet")
        ts = []
        t_start = time.perf_counter()
        while len(ts) < 5 and time.perf_counter() - t_start < 15:
            t0 = time.perf_counter(); po.forward(sd, 50, x, "keypoint_subnet"); ts.append(time.perf_counter() - t0)
    cpu_ms = float(np.median(ts)) * 1e3
    m = poseNet(50)
    sdm = m.state_dict()
    for k in sdm:
        if k in w:
            sdm[k] = torch.from_numpy(np.ascontiguousarray(w[k]))
    m.load_state_dict(sdm)
    m = m.to(dev).eval()
    xp = x.pin_memory()
    out_host = torch.empty((1, 18, H // 4, W // 4), dtype=torch.float32).pin_memory()

    def once():
        with torch.no_grad():
            heat, saved = m((xp.to(dev, non_blocking=True), "keypoint_subnet"))
        out_host.copy_(heat, non_blocking=True)
        torch.cuda.synchronize()
        return heat
    for _ in range(3):
        heat = once()
    err = float((heat.cpu() - ref_heat).abs().max() / ref_heat.abs().max())
    t0 = time.perf_counter()
    for _ in range(20):
        once()
    gpu_ms = (time.perf_counter() - t0) / 20 * 1e3
    del m
    torch.cuda.empty_cache()
    return {"workload": "R50 keypoint_subnet forward, batch 1, 3x480x640 (BASELINE configs[0])", "cpu_ms": cpu_ms,
            "cpu_images_per_s": 1e3 / cpu_ms, "cpu_threads": torch.get_num_threads(), "cpu_kind": "port (oracle restatement, torch CPU fp32)",
            "gpu_ms_e2e": gpu_ms, "gpu_images_per_s_e2e": 1e3 / gpu_ms, "gpu_precision": "f16f8",
            "gpu_note": "public model((img,'keypoint_subnet')) call, pinned host image in, heat map back, wall clock incl. sync",
            "heat_err_vs_cpu": err, "speedup_e2e": cpu_ms / gpu_ms}


def aux_roofline_leg(model, eng, dev, cls, boxes, reg, n_s, n_k, B):
    """HBM-bound kernels of the path against the measured copy bandwidth: algorithmic bytes / CUDA-event
    time.  decode: its own launch; filter .. reduce: events between the stages (mpn_filter_sort_nms_profile); twice: the
    bench's own detection load and the cfg3 feed (100 persons x 41 jittered candidates per image)."""
    from multiposenet.pytorch_b200 import ops, synthetic
    pk, pk_src = peaks()
    hbm = pk.get("hbm_gbs")
    A = boxes.shape[1]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    anchors = ops.anchors_for(H, W, dev)
    for _ in range(2):
        ops.decode_clip(anchors, reg, H, W)
    e0.record()
    for _ in range(10):
        ops.decode_clip(anchors, reg, H, W)
    e1.record()
    torch.cuda.synchronize()
    out = {"peak_gbs": hbm, "peak_source": pk_src, "unit": "GB/s", "kernels": []}

    def add(regime, name, ms, nbytes, note):
        out["kernels"].append({"regime": regime, "kernel": name, "ms": ms, "bytes": nbytes, "achieved": nbytes / (ms / 1e3) / 1e9 if ms > 0 else None,
                               "frac": (nbytes / (ms / 1e3) / 1e9 / hbm) if (ms > 0 and hbm) else None, "bytes_model": note})
    add("bench", "decode_clip_kernel", e0.elapsed_time(e1) / 10, B * A * 32.0, "16 B reg read + 16 B box written per anchor (anchors stay in L2)")

    def stages(regime, c, bx, ns, nk, max_cand):
        st = []
        for _ in range(3):
            ops.filter_sort_nms(c, bx, 0.05, 0.5, max_cand=max_cand, stage_ms=st)
        ms = np.median(np.array(st), axis=0)
        cb = (ns + 63) // 64
        add(regime, "filter_compact_kernel", float(ms[0]), B * (A * 4.0 + ns * 12.0), "4 B score per anchor + 12 B per candidate written")
        add(regime, "cub segmented radix sort (+segments)", float(ms[1]), B * ns * 16.0, "one ideal pass: 8 B key+rank read and written per candidate")
        add(regime, "gather_sorted_kernel", float(ms[2]), B * ns * 48.0, "key, rank, index, 16 B box read, 20 B row written per candidate")
        add(regime, "nms_mask_kernel", float(ms[3]), B * (ns * 20.0 + ns * cb * 8.0 / 2), "N_s*20 B read + upper triangle of N_s x ceil(N_s/64) u64 written")
        add(regime, "nms_reduce_kernel", float(ms[4]), B * (ns * cb * 8.0 / 2 + nk * 36.0), "upper-triangle mask rows staged once + kept rows written")
        return float(ms.sum())
    tot_bench = stages("bench (~%d candidates, ~%d kept per image)" % (n_s, n_k), cls, boxes, int(n_s), int(n_k), 4096)
    c3, b3 = synthetic.cfg3_detections(B, seed=3)
    c3, b3 = torch.from_numpy(c3).to(dev), torch.from_numpy(b3).to(dev)
    det3 = ops.filter_sort_nms(c3, b3, 0.05, 0.5, max_cand=4224)
    ns3, nk3 = float(det3.cand_cnt.float().mean()), float(det3.keep_cnt.float().mean())
    tot_cfg3 = stages("cfg3 (100 persons x 41 candidates: N_s %d, kept %.0f per image)" % (ns3, nk3), c3, b3, int(ns3), int(nk3), 4224)
    out["post_process_ms"] = {"bench": tot_bench, "cfg3": tot_cfg3, "batch": B}
    # batched PRN forward (weights are the traffic): 20 persons per image of the batch
    P = 20 * B
    xin = torch.rand(P, 56, 36, 17, device=dev) * 0.2
    with torch.no_grad():
        for _ in range(2):
            eng.prn_forward(xin)
        e0.record()
        for _ in range(5):
            eng.prn_forward(xin)
        e1.record()
    torch.cuda.synchronize()
    nw = sum(p.numel() for p in model.prn.parameters())
    wbytes = nw * (4.0 if eng.precision in ("f16f8", "bf16x3") else 2.0)
    add("prn", "PRN forward (3 FCs on conv_tc_kernel + add_softmax_rows), %d persons" % P, e0.elapsed_time(e1) / 5,
        wbytes + P * 34272 * 4.0 * 4, "packed weights read once (%d parameters) + input, FC3 output, residual and softmax output rows" % nw)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=None,
                    help="default: $MPN_PRECISION, else f16f8 for the inference modes (parity mode with the widest margin per "
                         "ms, profiles/r01s_parity_margin.txt) and bf16x3 for --mode train (the weight-gradient kernel runs on bf16 planes)")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--layers", type=int, default=LAYERS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fast", action="store_true")
    ap.add_argument("--mode", default="infer", choices=["infer", "train", "full"],
                    help="train = BASELINE config 4, full = config 5 incl. peaks + PRN assignment (extra lines, not the headline)")
    ap.add_argument("--persons", type=int, default=20, help="--mode full: person boxes per image fed to the PRN")
    ap.add_argument("--streams", type=int, default=None, help="branch-level side streams (default: engine default = on)")
    ap.add_argument("--graph", type=int, default=1, help="replay the step as a CUDA graph (0 = eager launches)")
    ap.add_argument("--overlap", type=int, default=1, help="training: bucketed allreduce under the backward (0 = one blocking allreduce)")
    ap.add_argument("--no-extras", action="store_true", help="skip parity / cpu_cfg1 / roofline_aux / train_step / full_pipeline keys")
    ap.add_argument("--train-batch", type=int, default=16)
    ap.add_argument("--full-batch", type=int, default=64)
    ap.add_argument("--extra-steps", type=int, default=5, help="timed steps of the train_step / full_pipeline legs of the default line")
    args = ap.parse_args()
    if args.precision is None:
        args.precision = os.environ.get("MPN_PRECISION") or ("bf16x3" if args.mode == "train" else "f16f8")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if args.mode == "train":
        return run_train(args, rank, world, local)
    if args.mode == "full":
        return run_full(args, rank, world, local)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    from multiposenet.pytorch_b200 import ops, poseNet, shard, synthetic

    B = args.batch
    model = poseNet(args.layers, precision=args.precision)
    w = load_weights_into(model, args.layers)
    model = model.to(dev).eval()
    if args.layers in CLS_BIAS_SHIFT:
        bias_shift = CLS_BIAS_SHIFT[args.layers]
        with torch.no_grad():
            model.classificationModel.output.bias += bias_shift
    else:
        bias_shift = calibrate_cls_bias(model, dev)
    eng = model.engine()
    flops_img = synthetic.conv_flops_entire(args.layers, H, W)
    import multiposenet.pytorch_b200.engine as engine_mod
    engine_mod.USE_GRAPHS = bool(args.graph)  # the public forward() replays a captured graph as well
    if args.streams is not None:
        engine_mod.USE_STREAMS = bool(args.streams)
    streams_default = engine_mod.USE_STREAMS

    # synthetic input: 3 distinct batches rotated so no step re-reads the previous step's input from L2
    rng = np.random.Generator(np.random.PCG64(1234 + rank))
    host = [torch.from_numpy(rng.standard_normal((B, 3, H, W), dtype=np.float32)).pin_memory() for _ in range(2)]
    devin = [h.to(dev) for h in host]

    def step_device(i):
        if args.graph:  # one cudaGraphLaunch replays the ~250 launches of the step
            return eng.graphed("entire", devin[i % len(devin)], max_cand=MAXC)
        return eng.entire_forward_device(devin[i % len(devin)], max_cand=MAXC)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # eager probe: launch count per step and the candidate capacity the NMS stage needs
    n0 = ops.stats["launches"]
    heat, cls, reg, boxes, det = eng.entire_forward_device(devin[0], max_cand=8192)
    launches_per_step = ops.stats["launches"] - n0
    torch.cuda.synchronize()
    worst = max(int(eng.entire_forward_device(d_, max_cand=8192)[4].cand_cnt.max()) for d_ in devin)
    MAXC = 4096 if worst <= 4096 else 8192
    for _ in range(args.warmup):
        out = step_device(_)
    barrier()
    heat, cls, reg, boxes, det = out
    torch.cuda.synchronize()
    n_s = det.cand_cnt.float().mean().item()
    n_k = det.keep_cnt.float().mean().item()
    assert int(det.cand_cnt.max()) <= MAXC, "candidate capacity exceeded"

    # ---- timed region: device-resident inputs
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_device(i)
    e1.record()
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    value = shard.whole_job_rate(B * args.steps, elapsed_ms, dev)
    elapsed_ms = shard.max_over_ranks(elapsed_ms, dev)

    # ---- end to end through the public API: pinned host -> device -> model((img,'both')) -> host
    heat_host = torch.empty((B, 18, H // 4, W // 4), dtype=torch.float32).pin_memory()

    copy_stream = torch.cuda.Stream(device=dev)

    def prefetch(i):
        """H2D of batch i from pinned host memory on a copy stream (overlaps the previous batch's compute)."""
        with torch.cuda.stream(copy_stream):
            x = host[i % len(host)].to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return x, ev

    d2h_stream = torch.cuda.Stream(device=dev)

    def read_heat(hm):
        """Device->host read of the step's heat maps (caller-owned tensor) on a side stream, so that the 44 MB copy
        overlaps the next step's kernels instead of sitting between two graph replays."""
        d2h_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(d2h_stream):
            heat_host.copy_(hm, non_blocking=True)
        hm.record_stream(d2h_stream)

    det_host = {}

    def read_detections(d):
        """Device->host read of the step's detections (keep counts, scores, boxes of the whole batch): the engine's tensors are
        static graph outputs the next replay overwrites, so they are cloned on the device (2.6 MB) and the clones go to pinned
        host buffers on the side stream, like the heat maps -- the compute stream never waits for PCIe."""
        srcs = [d.keep_cnt.clone(), d.scores.clone(), d.boxes.clone()]
        d2h_stream.wait_stream(torch.cuda.current_stream())
        outs = []
        with torch.cuda.stream(d2h_stream):
            for i, t in enumerate(srcs):
                hb = det_host.get((i, tuple(t.shape)))
                if hb is None:
                    hb = det_host[(i, tuple(t.shape))] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
                hb.copy_(t, non_blocking=True)
                t.record_stream(d2h_stream)
                outs.append(hb)
        return outs

    def run_e2e(nsteps):
        outs = None
        nxt = prefetch(0)
        for i in range(nsteps):
            x, ev = nxt
            torch.cuda.current_stream().wait_event(ev)
            x.record_stream(torch.cuda.current_stream())
            if i + 1 < nsteps:
                nxt = prefetch(i + 1)
            with torch.no_grad():
                hm, (sc, cl, bx) = model((x, "both"))  # the public call; syncs on the candidate counts
            read_heat(hm)
            outs = read_detections(eng.last_detections)
        torch.cuda.synchronize()
        return outs

    outs = run_e2e(2)
    d2h = heat_host.numel() * 4 + sum(o.numel() * o.element_size() for o in outs) + 4 * B
    h2d = host[0].numel() * 4
    barrier()
    e0.record()
    run_e2e(args.steps)
    e1.record()
    barrier()
    e2e_value = shard.whole_job_rate(B * args.steps, e0.elapsed_time(e1), dev)

    # ---- extension: the same end-to-end loop fed with raw uint8 BGR images (cv2 layout); the
    # reference's resnet_preprocess runs fused in the stem packing kernel, so 4x fewer bytes cross PCIe
    # the uint8 images are the fp32 batches pushed back through the inverse of resnet_preprocess (BGR bytes), so that both loops
    # see the same image statistics and hence the same detection load (uniform random bytes put > 4096 candidates per image
    # through the NMS stage and its capacity-retry path: the r01 / r02a-m "u8 slower than fp32" artefact)
    def to_u8(xf):
        mean = np.array([0.485, 0.456, 0.406], np.float32).reshape(1, 3, 1, 1)
        std = np.array([0.229, 0.224, 0.225], np.float32).reshape(1, 3, 1, 1)
        rgb = np.clip(np.rint((xf.numpy() * std + mean) * 255.0), 0, 255).astype(np.uint8)
        return torch.from_numpy(np.ascontiguousarray(rgb[:, ::-1].transpose(0, 2, 3, 1))).pin_memory()
    host_u8 = [to_u8(h) for h in host]

    def run_e2e_u8(nsteps):
        with torch.cuda.stream(copy_stream):
            nxt = (host_u8[0].to(dev, non_blocking=True), torch.cuda.Event())
            nxt[1].record(copy_stream)
        for i in range(nsteps):
            x, ev = nxt
            torch.cuda.current_stream().wait_event(ev)
            x.record_stream(torch.cuda.current_stream())
            if i + 1 < nsteps:
                with torch.cuda.stream(copy_stream):
                    nxt = (host_u8[(i + 1) % 2].to(dev, non_blocking=True), torch.cuda.Event())
                    nxt[1].record(copy_stream)
            with torch.no_grad():
                hm, _ = model((x, "both"))
            read_heat(hm)
            read_detections(eng.last_detections)
        torch.cuda.synchronize()

    run_e2e_u8(2)
    barrier()
    e0.record()
    run_e2e_u8(args.steps)
    e1.record()
    barrier()
    e2e_u8 = {"value": shard.whole_job_rate(B * args.steps, e0.elapsed_time(e1), dev), "unit": "images/s",
              "h2d_bytes_per_step": host_u8[0].numel(), "d2h_bytes_per_step": d2h,
              "note": "uint8 BGR input, resnet_preprocess fused on the device (API extension, not the reference's fp32 interface)"}

    # ---- roofline of the dominant kernel: every tcgen05 conv launch of a step, CUDA events per launch
    roof = None
    if rank == 0:
        pk, pk_src = peaks()
        nprof = 2
        engine_mod.USE_STREAMS = False  # serial launches: clean per-kernel durations
        level_default, engine_mod.LEVEL_STREAMS = engine_mod.LEVEL_STREAMS, False
        ops.stats["conv_events"] = evs = []
        for i in range(nprof):
            # eager: events around each launch.  The GPU first spins for ~40 ms so that the host enqueues the whole step ahead
            # of it: otherwise the start event of a short kernel is reached while the stream is empty and the interval
            # includes the host's launch gap (~10 us x 139 launches), not just the kernel
            torch.cuda._sleep(int(7e7))
            eng.entire_forward_device(devin[i % len(devin)], max_cand=MAXC)
            torch.cuda.synchronize()
        torch.cuda.synchronize()
        ops.stats["conv_events"] = None
        engine_mod.USE_STREAMS = streams_default
        engine_mod.LEVEL_STREAMS = level_default
        tc = [(a.elapsed_time(b), f, sl) for a, b, f, simt, sl in evs if not simt]
        conv_ms = sum(t for t, _, _ in tc) / nprof
        nconv = len(tc) // nprof
        # bf16-equivalent MMA passes issued per algorithmic MAC, FLOP-weighted over the launches (f16f8: 2, or 2.5 for the
        # convolutions whose input is stored without its e5m2 copy plane; bf16x3: 3; bf16: 1)
        passes = sum(f * sl for _, f, sl in tc) / max(1.0, sum(f for _, f, _ in tc))
        stem_flops = 2.0 * 64 * 3 * 49 * (H // 2) * (W // 2)
        on_cuda_cores = any(simt for _, _, _, simt, _ in evs)  # fp32 mode / MPN_TC_STEM=0: the stem is not a tensor-core launch
        alg = (flops_img - (stem_flops if on_cuda_cores else 0.0)) * B
        traffic, traffic_src = conv_traffic_record()
        achieved = alg / (conv_ms / 1e3) / 1e12
        # FLOPs the launches actually issue (2*MAC of every launch as launched): fewer than the reference graph's where the
        # algebra was changed (keypoint head conv2 without the replicated concat), more where operands are padded (stem K window)
        executed = sum(f for _, f, _ in tc) / nprof / (conv_ms / 1e3) / 1e12
        issued = sum(f * sl for _, f, sl in tc) / nprof / (conv_ms / 1e3) / 1e12
        peak = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
        roof = {"bound": "tensor", "kernel": "conv_tc_kernel (%d launches/step)" % nconv, "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": pk_src + " (sustained cuBLAS bf16)",
                "kernel_ms_per_step": conv_ms, "share_of_step": conv_ms / (elapsed_ms / args.steps),
                "mma_passes": passes,
                "executed_tflops": executed, "executed_frac": executed / peak,
                "tensor_pipe_frac": issued / peak,
                "note": "achieved = algorithmic conv FLOPs of the REFERENCE graph (2*MAC, fp32-equivalent) / conv kernel time; executed_* = the FLOPs "
                        "of the launches as issued (the keypoint head's conv2 is evaluated without its replicated concat: fewer MACs than "
                        "the reference graph), and tensor_pipe_frac = issued bf16-equivalent MMA passes of those launches / peak; bf16x3 issues 3 MMAs per "
                        "MAC (hi*hi + lo*hi + hi*lo); f16f8 issues 1 fp16 MMA + 2 fp8 MMAs at twice the rate = 2 bf16-equivalent passes (2.5 for the 1x1 "
                        "convolutions that read a tensor stored without its e5m2 copy plane: 2 fp16 MMAs + 1 fp8 MMA); mma_passes is the "
                        "FLOP-weighted mean over the launches of a step"}

    # ---- optional: single-pass bf16 throughput (not the parity mode; reported beside the headline)
    def side_mode(prec):
        feng = model.engine(prec)
        fstep = (lambda i: feng.graphed("entire", devin[i % 2], max_cand=MAXC)) if args.graph else \
                (lambda i: feng.entire_forward_device(devin[i % 2], max_cand=MAXC))
        for i in range(3):
            fstep(i)
        barrier()
        e0.record()
        for i in range(args.steps):
            fstep(i)
        e1.record()
        barrier()
        return shard.whole_job_rate(B * args.steps, e0.elapsed_time(e1), dev)

    fast = alt = None
    if not args.no_fast and args.precision in ("bf16x3", "f16f8"):
        fast = {"precision": "bf16 single pass (fails the 1e-3 parity bar, ~1e-2)", "value": side_mode("bf16"), "unit": "images/s"}
        other = "bf16x3" if args.precision == "f16f8" else "f16f8"
        alt = {"precision": other + " (the other parity mode, same 1e-3 bar)", "value": side_mode(other), "unit": "images/s"}

    # ---- CPU baseline (rank 0, N == 1): oracle port of the reference graph on the host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import weights
        sd = weights.to_torch_state_dict(w)
        sd["classificationModel.output.bias"] = model.classificationModel.output.bias.detach().cpu().clone()
        nthreads, navail = pick_cpu_threads(sd)
        x1 = host[0][:1].clone()
        cpu_out = cpu_reference_step(sd, x1)
        ts = []
        t_start = time.perf_counter()
        while len(ts) < 5 and time.perf_counter() - t_start < 25:
            t0 = time.perf_counter(); cpu_reference_step(sd, x1); ts.append(time.perf_counter() - t0)
        cpu = {"value": 1.0 / float(np.median(ts)), "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": "%d timed forwards of 1 image (R%d entire_net + NMS) after 1 warm-up; full step is %d images; %d of %d host threads" % (len(ts), args.layers, B, nthreads, navail)}

    # ---- parity of the timed configuration: image 0 of the timed batch against the fp32 CPU oracle (same weights, same image)
    parity = None
    if cpu is not None:
        with torch.no_grad():
            gheat, gcls, greg, gboxes, gdet = eng.entire_forward_device(devin[0], max_cand=MAXC)
        torch.cuda.synchronize()
        oheat, (osc, ocl, obx), oaux = cpu_out
        nerr = lambda a_, b_: float((a_.double().cpu() - b_.double()).abs().max() / b_.double().abs().max())
        parity = {"heat": nerr(gheat[:1], oheat), "cls": nerr(gcls[:1], oaux["cls"]), "reg": nerr(greg[:1], oaux["reg"]),
                  "kept_boxes": [int(gdet.keep_cnt[0]), int(len(osc))], "bar": 1e-3, "metric": "max|a-b|/max|b| per output tensor",
                  "against": "fp32 oracle port on the host (oracle/posenet_oracle.py), image 0 of the timed batch, R%d %s" % (args.layers, args.precision)}
        parity["ok"] = bool(max(parity["heat"], parity["cls"], parity["reg"]) <= (1e-3 if args.precision != "bf16" else 1.0))

    extras = {}
    if not args.no_extras:
        def guarded(name, fn):
            try:
                extras[name] = fn()
            except Exception as e:  # an auxiliary leg must not lose the headline line
                import traceback
                extras[name] = {"error": "%s: %s" % (type(e).__name__, e), "trace": traceback.format_exc()[-1500:]}
        if rank == 0:
            guarded("roofline_aux", lambda: aux_roofline_leg(model, eng, dev, cls, boxes, reg, n_s, n_k, B))
            if world == 1 and not args.no_cpu_baseline:
                guarded("cpu_cfg1", lambda: cpu_cfg1_leg(dev))
    # free the inference engines before the other configurations run
    del out, heat, cls, reg, boxes, det
    model.invalidate_engines()
    model.__dict__["_engines"].clear()
    del eng, model
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    if not args.no_extras:
        barrier()
        try:
            extras["train_step"] = train_leg(args, rank, world, local, args.extra_steps, 3)
        except Exception as e:
            import traceback
            extras["train_step"] = {"error": "%s: %s" % (type(e).__name__, e), "trace": traceback.format_exc()[-1500:]}
        gc.collect(); torch.cuda.empty_cache()
        barrier()
        try:
            extras["full_pipeline"] = full_leg(args, rank, world, local, args.extra_steps, 3)
        except Exception as e:
            import traceback
            extras["full_pipeline"] = {"error": "%s: %s" % (type(e).__name__, e), "trace": traceback.format_exc()[-1500:]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16x3": "bf16x3 (hi/lo split, fp32 accumulate)", "bf16": "bf16", "fp32": "f32",
                      "f16f8": "f16f8 (fp16 hi*hi + two fp8 cross terms, fp32 accumulate)"}[args.precision],
            "data": "synthetic",
            "config": {"workload": "R%d-FPN entire_net fwd (keypoint + RetinaNet heads) + decode/filter/NMS, batch %d/GPU, 3x480x640"
                                   % (args.layers, B), "global_batch": B * world, "parallelism": "dp%d (image shards, no collective)" % world,
                       "precision": args.precision, "candidates_per_image": n_s, "kept_per_image": n_k, "max_cand": MAXC, "cuda_graph": bool(args.graph), "branch_streams": streams_default, "level_streams": engine_mod.LEVEL_STREAMS,
                       "cls_bias_shift": bias_shift,
                       "l2": "2 rotating input batches; activations >5 GB/step >> 126 MB L2, no explicit flush",
                       "gflop_per_image": flops_img / 1e9},
            "clocks": clocks, "gpu_launches": launches_per_step * args.steps,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "e2e_u8_input": e2e_u8, "roofline": roof, "cpu_baseline": cpu, "fast_mode": fast, "alt_parity_mode": alt,
            "parity": parity,
        }
        line.update(extras)
        print(json.dumps(line))
        if parity is not None and not parity["ok"]:
            sys.exit("parity check failed: %s" % json.dumps(parity))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
