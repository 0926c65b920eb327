"""Data-parallel plumbing for the hot path: one process per GPU, torch.distributed.

The path shards by independent units (tensors of a quantize sweep, evaluation windows); there is NO collective on
the data path.  What crosses ranks is scalars only: MAX of elapsed times, SUM of (nll, count), and for a
fine-tune step an all-reduce(SUM)/world of the trainable gradients.
"""
import torch
import torch.distributed as dist

__all__ = ["shard_indices", "windows_for", "reduce_max", "reduce_sum", "allreduce_grads_"]


def shard_indices(n_items: int, world: int, rank: int):
    """Round-robin assignment of `n_items` independent units: rank r gets r, r + world, ...  Every unit is
    assigned exactly once; shard sizes differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of size {world}")
    return list(range(rank, n_items, world))


def windows_for(n_tokens: int, max_length: int, stride: int):
    """Sliding evaluation windows (begin, end, target_len) over a token stream, as the reference's perplexity
    script builds them (examples/language_modeling/wikitext.py:146-165): windows of `max_length` every `stride`
    tokens, each scoring only the tokens not scored by the previous one."""
    out, prev_end = [], 0
    for begin in range(0, n_tokens, stride):
        end = min(begin + max_length, n_tokens)
        out.append((begin, end, end - prev_end))
        prev_end = end
        if end == n_tokens:
            break
    return out


def _device_for(group=None):
    backend = dist.get_backend(group) if dist.is_initialized() else None
    return torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")


def reduce_max(value: float, group=None) -> float:
    """MAX over ranks of a host scalar (timings are reported as the slowest rank's)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=_device_for(group))
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def reduce_sum(values, group=None):
    """SUM over ranks of a list of host scalars (nll sums, token counts, units processed)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [float(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.float64, device=_device_for(group))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return [float(v) for v in t.tolist()]


def allreduce_grads_(params, group=None, bucket_bytes=32 << 20):
    """Average the gradients of the TRAINABLE parameters over ranks, in flat buckets (one NCCL all-reduce per
    bucket; with LoRA adapters the whole model is a single ~2 MB bucket, i.e. latency-bound over NVLink).
    Activation-gradient quantization is rank-local in the reference, so nothing else is exchanged."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    grads = [p.grad for p in params if p.requires_grad and p.grad is not None]
    buckets, cur, cur_bytes = [], [], 0
    for g in grads:
        nbytes = g.numel() * g.element_size()
        if cur and (cur_bytes + nbytes > bucket_bytes or g.dtype != cur[0].dtype):
            buckets.append(cur)
            cur, cur_bytes = [], 0
        cur.append(g)
        cur_bytes += nbytes
    if cur:
        buckets.append(cur)
    for b in buckets:
        flat = torch.cat([g.reshape(-1) for g in b])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        off = 0
        for g in b:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
    return len(buckets)
