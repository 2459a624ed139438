"""Data-parallel plumbing of the path: one process per GPU, images sharded by rank, NO data-path
collective for inference (the reference's nn.DataParallel scatter/gather, evaluate/tester.py:126, becomes
"each rank keeps its shard").  The only exchanges are control-plane: a barrier and the max-over-ranks of
the device-timed duration used to report whole-job throughput.
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous, balanced [begin, end) of `total` images for `rank` (first total % world ranks get +1)."""
    if world <= 0 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def max_over_ranks(value, device=None):
    """Max of a python float over all ranks (identity when torch.distributed is not initialised)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if device is not None and torch.device(device).type == "cuda":
        t = t.float()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    on_cuda = device is not None and torch.device(device).type == "cuda"
    t = torch.tensor([float(value)], dtype=torch.float32 if on_cuda else torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def whole_job_rate(images_this_rank, elapsed_ms_local, device=None):
    """images/s of the whole job: all ranks' images over the slowest rank's device time."""
    slowest = max_over_ranks(elapsed_ms_local, device)
    return sum_over_ranks(images_this_rank, device) / (slowest / 1e3)


def allreduce_mean_(flat):
    """In-place mean over ranks of a flat gradient buffer: the ONE collective of the training step.
    NCCL on CUDA tensors, gloo on CPU tensors; identity without an initialised process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat /= dist.get_world_size()
    return flat


def flatten_grads(named_grads):
    """[(name, tensor)] -> (flat fp32 buffer, [(name, shape, offset)])."""
    parts, index, off = [], [], 0
    for n, g in named_grads:
        parts.append(g.reshape(-1).float())
        index.append((n, tuple(g.shape), off))
        off += g.numel()
    return torch.cat(parts) if parts else torch.zeros(0), index


def unflatten_grads(flat, index):
    out = {}
    for n, shape, off in index:
        k = 1
        for d in shape:
            k *= d
        out[n] = flat[off:off + k].view(shape)
    return out
