"""Data-parallel plumbing for the hot path: one process per GPU, torch.distributed.

The path shards by independent units (tensors of a quantize sweep, evaluation windows); there is NO collective on
the data path.  What crosses ranks is scalars only: MAX of elapsed times, SUM of (nll, count), and for a
fine-tune step an all-reduce(SUM)/world of the trainable gradients.
"""
import torch
import torch.distributed as dist

__all__ = ["shard_indices", "windows_for", "reduce_max", "reduce_sum", "allreduce_grads_", "GradReducer"]


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


class GradReducer:
    """All-reduce(SUM)/world of the TRAINABLE gradients, overlapped with the backward pass (SURVEY.md §8e row 3;
    the reference's loop: run_glue_no_trainer.py:658-668 -- there single-process, here one process per GPU).

    The trainable parameters are packed once into flat buckets in REVERSE registration order (the order backward
    produces their gradients); each parameter's `.grad` is a view into its bucket, so nothing is copied.  A
    post-accumulate-grad hook counts a bucket's gradients as they arrive; when the last one lands the bucket's NCCL
    all-reduce is launched asynchronously (on NCCL's own stream) while backward continues to compute the
    gradients of the earlier layers.  `finish()` -- call it before optimizer.step() -- waits for the outstanding
    handles and divides by the world size.  With LoRA adapters the buckets are a few hundred KB: the collective is
    latency-bound over NVLink and disappears behind the backward of the frozen layers below."""

    def __init__(self, params, group=None, bucket_bytes=1 << 20):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.buckets, self._handles, self._hooks = [], [], []
        # NCCL averages inside the collective; gloo (CPU tests) sums and finish() divides
        self._avg = dist.is_initialized() and dist.get_backend(group) == "nccl"
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            nbytes = p.numel() * p.element_size()
            if cur and (cur_bytes + nbytes > bucket_bytes or p.dtype != cur[0].dtype or p.device != cur[0].device):
                self._make_bucket(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self._make_bucket(cur)
        for b in self.buckets:
            for p in b["params"]:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(b)))

    def _make_bucket(self, plist):
        flat = torch.zeros(sum(p.numel() for p in plist), dtype=plist[0].dtype, device=plist[0].device)
        views, off = [], 0
        for p in plist:
            views.append(flat[off:off + p.numel()].view_as(p))
            p.grad = views[-1]                                # gradients accumulate straight into the bucket
            off += p.numel()
        self.buckets.append({"flat": flat, "params": plist, "views": views, "pending": len(plist), "launched": False})

    def _make_hook(self, bucket):
        def hook(param):
            bucket["pending"] -= 1
            if bucket["pending"] == 0 and not bucket["launched"]:
                self._launch(bucket)
        return hook

    def _launch(self, bucket):
        bucket["launched"] = True
        if self.world > 1:
            op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
            self._handles.append(dist.all_reduce(bucket["flat"], op=op, group=self.group, async_op=True))

    def zero_grad(self):
        """Zero the buckets in place (keeps `.grad` the bucket views; do NOT call optimizer.zero_grad(set_to_none=True))."""
        for b in self.buckets:
            b["flat"].zero_()
            b["pending"], b["launched"] = len(b["params"]), False
            for p, v in zip(b["params"], b["views"]):
                if p.grad is not v:
                    p.grad = v

    def finish(self):
        """Launch whatever did not fire (parameters unused this step), wait, average.  Returns the bucket count."""
        for b in self.buckets:
            if not b["launched"]:
                self._launch(b)
        for h in self._handles:
            h.wait()
        self._handles = []
        if self.world > 1 and not self._avg:
            for b in self.buckets:
                b["flat"].div_(self.world)
        return len(self.buckets)

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
