"""Training step of the keypoint subnet on the C ABI (BASELINE config 4).

Reference: `Trainer._train_one_epoch` (training/trainer.py:239-262) runs
    output, saved_for_loss = model([img, 'keypoint_subnet'])        # BN in TRAIN mode (trainer.py:170-174)
    loss, log = model.module.build_loss(saved_for_loss, 'keypoint_subnet', heat, heat_weight)   (posenet.py:367-403)
    loss.backward(); optimizer.step()
with the detection neck / heads / PRN frozen (training/multipose_keypoint_train.py:78-89).

Here the forward saves what the backward needs (raw conv outputs y, batch statistics, activations z), and the backward
is an explicit reverse schedule of libmpn_b200 launches:
    BN backward (two per-channel reductions + one elementwise pass), ReLU masks folded into it
    data gradients  = the forward tcgen05 conv kernel on the flipped / transposed filter (stride 2: zero-inserted dY)
    weight gradients = mpn_conv2d_wgrad (tcgen05, MN-major operands, split-K fp32 atomics)
    bias gradients  = per-channel sums; max-pool / nearest-upsample / concat-slice backward kernels
Gradient accumulation where a tensor has two consumers rides on the conv epilogue's residual input.
`KeypointTrainFunction` exposes this as a torch.autograd.Function so the reference's loop above works unchanged;
`TrainEngine.train_step` is the fused variant (loss kernel included) used by bench.py.  Data parallel: one process per
GPU, per-rank batch statistics (nn.DataParallel semantics, no SyncBN), one NCCL allreduce over the flat fp32 gradient.
"""
import torch

from . import ops
from . import train_ops as T
from ._lib import FMT_F32, OUT_ACT, OUT_F32_NCHW


class _Saved(object):
    pass


class TrainEngine(object):
    def __init__(self, model, precision="bf16x3"):
        if precision not in ("bf16x3", "bf16"):
            raise ValueError("training runs on the tcgen05 path: precision must be bf16x3 or bf16")
        self.model = model
        self.precision = precision
        self.fmt = ops.PRECISIONS[precision]
        self.grads = {}
        self.trace = None  # dict -> records d loss / d (block output) during backward (diagnostics)
        self._bucket_hook = None  # called once inside backward() when the gradients of head + neck + layer4 + layer3 are complete
        self._ov = {}

    # ------------------------------------------------------------------ helpers
    def _pc(self, conv):
        """Unfolded filter (+bias) of the CURRENT weights (they change every step)."""
        return ops.pack_conv(conv.weight, conv.bias, None, self.fmt)

    def _conv(self, x, conv, **kw):
        return ops.conv2d(x, self._pc(conv), stride=conv.stride[0], pad=conv.padding[0], **kw)

    def _conv_grads(self, name, conv, x, dy, need_dx=True, residual=None, dx_hw=None):
        """Accumulate dW / db of `conv` (x -> y, dy given, possibly with zero-padded extra channels); return dX."""
        Cout, _, R, S = conv.weight.shape
        stride, pad = conv.stride[0], conv.padding[0]
        dw = T.conv_wgrad(x, dy, Cout, R, S, stride, pad)
        self.grads[name + ".weight"] = T.unpack_filter_grad(dw)
        if conv.bias is not None:
            self.grads[name + ".bias"] = T.channel_sum(dy, C=Cout)
        if not need_dx:
            return None
        return T.conv_dgrad(dy, conv.weight, stride, pad, dx_hw or (x.H, x.W), self.fmt, residual=residual)

    def _bn_grads(self, name, dgamma, dbeta):
        self.grads[name + ".weight"] = dgamma
        self.grads[name + ".bias"] = dbeta

    # ------------------------------------------------------------------ forward (train mode)
    def forward(self, img):
        m, fpn = self.model, self.model.fpn
        S = _Saved()
        S.img_hw = (img.shape[2], img.shape[3])
        # stem (fpn.py:99-100)
        S.xs = ops.stem_pack_input(img, self.fmt)
        spc = ops.pack_stem_filter(fpn.conv1.weight, (torch.ones_like(fpn.bn1.weight), torch.zeros_like(fpn.bn1.bias),
                                                      torch.zeros_like(fpn.bn1.running_mean), torch.ones_like(fpn.bn1.running_var), 0.0),
                                   self.fmt)
        spc.scale = spc.bias = None
        y = ops.conv2d(S.xs, spc)
        S.stem_z, S.stem_st = T.bn_train_forward(y, fpn.bn1, relu=True)
        x = ops.maxpool3x3s2(S.stem_z)
        S.pool = x
        S.blocks = []
        feats = []
        for li in range(1, 5):
            for bi, blk in enumerate(getattr(fpn, "layer%d" % li)):
                b = _Saved()
                b.name, b.blk, b.x = "fpn.layer%d.%d" % (li, bi), blk, x
                y1 = self._conv(x, blk.conv1)
                b.z1, b.st1 = T.bn_train_forward(y1, blk.bn1, relu=True)
                y2 = self._conv(b.z1, blk.conv2)
                b.z2, b.st2 = T.bn_train_forward(y2, blk.bn2, relu=True)
                y3 = self._conv(b.z2, blk.conv3)
                if len(blk.downsample) > 0:
                    ys = self._conv(x, blk.downsample[0])
                    s, b.std = T.bn_train_forward(ys, blk.downsample[1], relu=False)
                else:
                    s, b.std = x, None
                x, b.st3 = T.bn_train_forward(y3, blk.bn3, relu=True, residual=s)
                S.blocks.append(b)
            feats.append(x)
        S.c2, S.c3, S.c4, S.c5 = feats
        # keypoint neck (fpn.py:117-124)
        S.fp5 = self._conv(S.c5, fpn.toplayer)
        S.fp4 = self._conv(S.c4, fpn.flatlayer1, up=S.fp5)
        S.fp3 = self._conv(S.c3, fpn.flatlayer2, up=S.fp4)
        S.fp2 = self._conv(S.c2, fpn.flatlayer3, up=S.fp3)
        S.s4 = self._conv(S.fp4, fpn.smooth1)
        S.s3 = self._conv(S.fp3, fpn.smooth2)
        S.s2 = self._conv(S.fp2, fpn.smooth3)
        ps = {2: S.s2, 3: S.s3, 4: S.s4, 5: S.fp5}
        # intermediate supervision (posenet.py:296-299)
        outs = []
        for k, rep in ((2, 1), (3, 2), (4, 4), (5, 8)):
            outs.append(self._conv(ps[k], getattr(m, "convfin_k%d" % k), out_mode=OUT_F32_NCHW, out_rep=rep))
        # head (posenet.py:302-315)
        S.cat = ops.Act(self.fmt, S.s2.N, S.s2.H, S.s2.W, 512, S.s2.hi.device)
        S.q = {}
        for i, (k, rep, off) in enumerate(((5, 8, 0), (4, 4, 128), (3, 2, 256), (2, 1, 384)), start=1):
            S.q[i] = self._conv(ps[k], getattr(m, "convt%d" % i))
            self._conv(S.q[i], getattr(m, "convs%d" % i), out=S.cat, out_coffset=off, out_rep=rep)
        S.h = self._conv(S.cat, m.conv2, relu=True)
        outs.append(self._conv(S.h, m.convfin, out_mode=OUT_F32_NCHW))
        S.ps = ps
        return outs, S

    # ------------------------------------------------------------------ backward
    def backward(self, S, douts):
        """douts: 5 Acts [B, H/4, W/4, 64] (channels >= 19/18 zero) = d loss / d (k2, k3, k4, k5, heat)."""
        m, fpn = self.model, self.model.fpn
        self.grads = {}
        # head: convfin <- conv2(+ReLU) <- concat
        d_h = self._conv_grads("convfin", m.convfin, S.h, douts[4])
        d_hpre = T.relu_backward(d_h, S.h)
        d_cat = self._conv_grads("conv2", m.conv2, S.cat, d_hpre)
        dp = {}
        for i, (k, rep, off) in enumerate(((5, 8, 0), (4, 4, 128), (3, 2, 256), (2, 1, 384)), start=1):
            d_cs = T.block_sum(d_cat, rep, C=128, coffset=off)
            d_q = self._conv_grads("convs%d" % i, getattr(m, "convs%d" % i), S.q[i], d_cs)
            dp[k] = self._conv_grads("convt%d" % i, getattr(m, "convt%d" % i), S.ps[k], d_q)
        # intermediate heads add into the same neck gradients (residual epilogue)
        for j, (k, rep) in enumerate(((2, 1), (3, 2), (4, 4), (5, 8))):
            d_small = T.block_sum(douts[j], rep) if rep > 1 else douts[j]
            dp[k] = self._conv_grads("convfin_k%d" % k, getattr(m, "convfin_k%d" % k), S.ps[k], d_small, residual=dp[k])
        # neck: smooth convs, then the top-down pathway
        d_fp2 = self._conv_grads("fpn.smooth3", fpn.smooth3, S.fp2, dp[2])
        d_fp3 = self._conv_grads("fpn.smooth2", fpn.smooth2, S.fp3, dp[3])
        d_fp4 = self._conv_grads("fpn.smooth1", fpn.smooth1, S.fp4, dp[4])
        d_c2 = self._conv_grads("fpn.flatlayer3", fpn.flatlayer3, S.c2, d_fp2)
        d_fp3 = T.add(d_fp3, T.block_sum(d_fp2, 2))
        d_c3 = self._conv_grads("fpn.flatlayer2", fpn.flatlayer2, S.c3, d_fp3)
        d_fp4 = T.add(d_fp4, T.block_sum(d_fp3, 2))
        d_c4 = self._conv_grads("fpn.flatlayer1", fpn.flatlayer1, S.c4, d_fp4)
        d_fp5 = T.add(dp[5], T.block_sum(d_fp4, 2))
        d_c5 = self._conv_grads("fpn.toplayer", fpn.toplayer, S.c5, d_fp5)
        # backbone, last block first; the neck's contribution joins at each stage output
        extra = {id(S.c4): d_c4, id(S.c3): d_c3, id(S.c2): d_c2}
        d = d_c5
        for b in reversed(S.blocks):
            blk, name = b.blk, b.name
            if self.trace is not None:
                self.trace[name] = d.to_nchw()  # gradient w.r.t. this block's output
            dy3, g, dg, db = T.bn_train_backward(d, b.st3, blk.bn3, want_g=True)
            self._bn_grads(name + ".bn3", dg, db)
            dz2 = self._conv_grads(name + ".conv3", blk.conv3, b.z2, dy3)
            dy2, _, dg, db = T.bn_train_backward(dz2, b.st2, blk.bn2)
            self._bn_grads(name + ".bn2", dg, db)
            dz1 = self._conv_grads(name + ".conv2", blk.conv2, b.z1, dy2)
            dy1, _, dg, db = T.bn_train_backward(dz1, b.st1, blk.bn1)
            self._bn_grads(name + ".bn1", dg, db)
            if b.std is not None:
                dys, _, dg, db = T.bn_train_backward(g, b.std, blk.downsample[1])
                self._bn_grads(name + ".downsample.1", dg, db)
                dsc = self._conv_grads(name + ".downsample.0", blk.downsample[0], b.x, dys)
            else:
                dsc = g
            d = self._conv_grads(name + ".conv1", blk.conv1, b.x, dy1, residual=dsc)
            if id(b.x) in extra:  # b.x is a stage output (c2 / c3 / c4) that also feeds a lateral conv
                d = T.add(d, extra[id(b.x)])
            if self._bucket_hook is not None and name == "fpn.layer3.0":
                # ~88 % of the trainable parameters (head, neck, layer4, layer3) have their gradients now, ~40 % of the backward's
                # time is still ahead (layer2, layer1 and the stem work on the large maps): the allreduce of this bucket overlaps it
                self._bucket_hook()
        # stem
        d_z = T.maxpool_backward(S.stem_z, d)
        dy, _, dg, db = T.bn_train_backward(d_z, S.stem_st, fpn.bn1)
        self._bn_grads("fpn.bn1", dg, db)
        dw = T.conv_wgrad(S.xs, dy, 64, 4, 1, 1, 0)
        self.grads["fpn.conv1.weight"] = T.stem_unpack_filter_grad(dw)
        return self.grads

    # ------------------------------------------------------------------ fused step
    def forward_backward(self, img, heat_gt, heat_weight):
        """One forward + backward with the fused weighted-MSE loss kernel (posenet.py:367-403: sum of 5 MSE terms).
        Returns (loss fp64 [1] on device, outs, grads dict name -> tensor)."""
        outs, S = self.forward(img)
        self.last_saved = S
        loss = torch.zeros(1, dtype=torch.float64, device=img.device)
        douts = [T.mse_heatmap_loss(o, heat_gt, heat_weight, loss, self.fmt, Cd=64) for o in outs]
        grads = self.backward(S, douts)
        return loss, outs, grads

    def relu_masks(self, S):
        """The {0,1} ReLU patterns of a saved forward, in the order the graph applies them (fp32 NCHW tensors):
        stem, then (relu1, relu2, relu_out) per bottleneck, then the head's conv2.  Test support."""
        acts = [S.stem_z]
        for b in S.blocks:
            acts += [b.z1, b.z2, b.st3.z]
        acts.append(S.h)
        return [(a.to_nchw() > 0).float() for a in acts]

    def pool_indices(self, S):
        """Flat arg-max indices (into H*W of the stem activation) chosen by the 3x3/2 max-pool of the saved forward."""
        z = S.stem_z.to_nchw()
        return torch.nn.functional.max_pool2d(z, 3, 2, 1, return_indices=True)[1]

    def graphed_forward_backward(self, img, heat_gt, heat_weight):
        """forward_backward replayed from a CUDA graph (captured on first use per input shape): the ~2000 launches
        of a step become one driver call, which matters because the eager step is bound by Python launch overhead.
        The filters are re-packed from the live weights INSIDE the graph, so optimizer updates are seen.
        Returned tensors are the graph's static outputs (overwritten by the next replay)."""
        key = (tuple(img.shape), str(img.device))
        g = getattr(self, "_graphs", {}).get(key)
        if g is None:
            if not hasattr(self, "_graphs"):
                self._graphs = {}
            sx, sg, sw = torch.empty_like(img), torch.empty_like(heat_gt), torch.empty_like(heat_weight)
            sx.copy_(img); sg.copy_(heat_gt); sw.copy_(heat_weight)
            bn_state = [(m, m.running_mean.clone(), m.running_var.clone(), m.num_batches_tracked.clone())
                        for m in self.model.modules() if isinstance(m, torch.nn.BatchNorm2d)]
            side = torch.cuda.Stream(device=img.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self.forward_backward(sx, sg, sw)  # warm-up off the capture (function attributes, allocator pools)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            for m, rm, rv, nb in bn_state:  # the warm-up must not count as a training step
                m.running_mean.copy_(rm); m.running_var.copy_(rv); m.num_batches_tracked.copy_(nb)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.forward_backward(sx, sg, sw)
            for m, rm, rv, nb in bn_state:  # capture does not execute, but keep the invariant explicit
                m.running_mean.copy_(rm); m.running_var.copy_(rv); m.num_batches_tracked.copy_(nb)
            g = (graph, sx, sg, sw, out)
            self._graphs[key] = g
        graph, sx, sg, sw, out = g
        sx.copy_(img, non_blocking=True); sg.copy_(heat_gt, non_blocking=True); sw.copy_(heat_weight, non_blocking=True)
        graph.replay()
        return out

    # ------------------------------------------------------------------ overlapped data-parallel step
    def _build_overlapped(self, img, heat_gt, heat_weight):
        """Two CUDA graphs that share one memory pool: graph 1 = forward + loss + backward down to the first block of layer3,
        graph 2 = the rest of the backward.  Each ends by copying its gradients into its slice of ONE flat fp32 buffer, so the
        NCCL allreduce of the first (large) bucket can run on a side stream while graph 2 replays."""
        st = _Saved()
        st.sx, st.sg, st.sw = torch.empty_like(img), torch.empty_like(heat_gt), torch.empty_like(heat_weight)
        st.sx.copy_(img); st.sg.copy_(heat_gt); st.sw.copy_(heat_weight)
        bn_state = [(m, m.running_mean.clone(), m.running_var.clone(), m.num_batches_tracked.clone())
                    for m in self.model.modules() if isinstance(m, torch.nn.BatchNorm2d)]

        def restore():
            for m, rm, rv, nb in bn_state:
                m.running_mean.copy_(rm); m.running_var.copy_(rv); m.num_batches_tracked.copy_(nb)
        # eager dry run: warms allocator pools / function attributes and records which gradients exist at the hook
        at_hook = []
        self._bucket_hook = lambda: at_hook.extend(self.grads.keys())
        side = torch.cuda.Stream(device=img.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            _, _, grads = self.forward_backward(st.sx, st.sg, st.sw)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._bucket_hook = None
        restore()
        named = [(n, p) for n, p in self.trainable_parameters() if n in grads]
        first = set(at_hook)
        order = [(n, p) for n, p in named if n in first] + [(n, p) for n, p in named if n not in first]
        st.index, off = [], 0
        for n, p in order:
            st.index.append((n, p, off, p.numel()))
            off += p.numel()
        st.n_first = sum(k for n, _, _, k in st.index if n in first)
        st.flat = torch.zeros(off, dtype=torch.float32, device=img.device)
        st.flat_a, st.flat_b = st.flat[:st.n_first], st.flat[st.n_first:]

        def gather(keys):   # inside the capture: this bucket's gradients -> the flat buffer
            for n, p, o, k in st.index:
                if (n in first) == keys:
                    st.flat[o:o + k].copy_(self.grads[n].reshape(-1))
        st.g1, st.g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        pool = torch.cuda.graph_pool_handle()
        cap = torch.cuda.Stream(device=img.device)
        cap.wait_stream(torch.cuda.current_stream())

        def split():
            gather(True)
            st.g1.capture_end()
            st.g2.capture_begin(pool=pool)
        self._bucket_hook = split
        try:
            with torch.cuda.stream(cap):
                st.g1.capture_begin(pool=pool)
                st.loss, st.outs, st.grads = self.forward_backward(st.sx, st.sg, st.sw)
                gather(False)
                st.g2.capture_end()
        finally:
            self._bucket_hook = None
        torch.cuda.current_stream().wait_stream(cap)
        torch.cuda.synchronize()
        restore()
        st.comm = torch.cuda.Stream(device=img.device)
        return st

    def train_step_overlapped(self, img, heat_gt, heat_weight, world_size=1):
        """forward + backward + gradient allreduce of one data-parallel step (reference: the DataParallel reduce-add of
        training/trainer.py:170, 245-259), with the allreduce of the first gradient bucket overlapped with the tail of the
        backward.  Leaves the averaged gradients in param.grad (views of one flat buffer) and returns the loss (fp64 [1])."""
        import torch.distributed as dist
        key = (tuple(img.shape), str(img.device))
        st = self._ov.get(key)
        if st is None:
            st = self._ov[key] = self._build_overlapped(img, heat_gt, heat_weight)
        st.sx.copy_(img, non_blocking=True); st.sg.copy_(heat_gt, non_blocking=True); st.sw.copy_(heat_weight, non_blocking=True)
        multi = world_size > 1 and dist.is_available() and dist.is_initialized()
        st.g1.replay()
        if multi:
            cur = torch.cuda.current_stream()
            ev = torch.cuda.Event()
            ev.record(cur)
            with torch.cuda.stream(st.comm):
                st.comm.wait_event(ev)
                w1 = dist.all_reduce(st.flat_a, async_op=True)      # runs under graph 2
        st.g2.replay()
        if multi:
            w2 = dist.all_reduce(st.flat_b, async_op=True)
            w1.wait()
            w2.wait()
            st.flat.div_(world_size)
        for n, p, o, k in st.index:
            p.grad = st.flat[o:o + k].view_as(p)
        return st.loss

    def trainable_parameters(self):
        return [(n, p) for n, p in self.model.named_parameters() if p.requires_grad]

    def assign_grads(self, grads, world_size=1):
        """Copy the step's gradients into param.grad; with several ranks, ONE allreduce (NCCL) over the flat fp32
        buffer first (reference: DataParallel's reduce-add onto GPU 0, training/trainer.py:170)."""
        from . import shard
        named = [(n, p) for n, p in self.trainable_parameters() if n in grads]
        flat, index = shard.flatten_grads([(n, grads[n]) for n, _ in named])
        if world_size > 1:
            shard.allreduce_mean_(flat)
        by_name = shard.unflatten_grads(flat, index)
        for n, p in named:
            p.grad = by_name[n].view_as(p)
        return flat


class KeypointTrainFunction(torch.autograd.Function):
    """model([img, 'keypoint_subnet']) under torch.enable_grad() in train mode: the five supervised maps come out as
    autograd-tracked fp32 tensors, so `build_loss(...).backward()` of the reference loop reaches the kernels above."""

    @staticmethod
    def forward(ctx, engine, img, *params):
        outs, S = engine.forward(img)
        engine.last_saved = S
        ctx.engine, ctx.saved, ctx.names = engine, S, [n for n, _ in engine.trainable_parameters()]
        ctx.out_shapes, ctx.device = [tuple(o.shape) for o in outs], img.device
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        eng, S = ctx.engine, ctx.saved
        acts = []
        for o_grad, o in zip(douts, ctx.out_shapes):
            if o_grad is None:
                o_grad = torch.zeros(o, dtype=torch.float32, device=ctx.device)
            g = o_grad.contiguous().float()
            pad = 64 - g.shape[1]
            acts.append(ops.act_from_nchw(torch.nn.functional.pad(g, (0, 0, 0, 0, 0, pad)), eng.fmt))
        grads = eng.backward(S, acts)
        out = [None, None]
        for n in ctx.names:
            out.append(grads.get(n))
        return tuple(out)
