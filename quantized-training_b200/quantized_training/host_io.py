"""Fake-quantize tensors that live in HOST memory.

`fake_quantize_host(mod, x_host)` streams a (pinned) host tensor through the GPU in chunks on several CUDA
streams, so the upload of chunk c+1, the kernel of chunk c and the download of chunk c-1 overlap (PCIe is
full duplex; the kernel itself is ~100x faster than the link).  Results are identical to `mod(x_host.cuda())`:
the delayed scale is updated once per call, and every chunk max-accumulates |x| into the same history slot.
"""
import torch

from . import _C
from .fake_quantize import FusedAmaxObsFakeQuantize, _channel_view

__all__ = ["fake_quantize_host", "HostPipeline"]


class HostPipeline:
    """Reusable staging buffers + streams for one device / dtype / chunk size."""

    def __init__(self, device, dtype=torch.bfloat16, chunk_elems=1 << 23, depth=3):
        self.device = torch.device(device)
        self.chunk = int(chunk_elems)
        self.streams = [torch.cuda.Stream(self.device) for _ in range(depth)]
        self.xin = [torch.empty(self.chunk, dtype=dtype, device=self.device) for _ in range(depth)]
        self.yout = [torch.empty(self.chunk, dtype=dtype, device=self.device) for _ in range(depth)]

    def _prepare(self, mod, n):
        """Observer step of one module (once per call, like mod.forward): (quantize?, observe?, amax slot)."""
        if mod.scale.device != self.device:
            mod.to(self.device)
        observe, quantize = mod._flags()
        if mod.is_per_channel or getattr(mod, "is_block_scaled", False):
            raise NotImplementedError("host streaming handles per-tensor and bare specs (chunks cut across channels)")
        amax_slot = None
        if observe:
            if n == 0:
                raise RuntimeError("amax(): cannot observe an empty tensor")
            if mod.amax_history.numel() == 0:
                mod.amax_history.resize_((mod.amax_history_len,)).fill_(0.0)
                mod.scale.resize_(()).fill_(1.0)
            _C.scale_update(mod.amax_history, mod.amax_history_len, 1, mod.scale, mod.quant_max,
                            mod.force_scale_power_of_two)
            amax_slot = mod.amax_history
        return quantize, observe, amax_slot

    def run(self, mod: FusedAmaxObsFakeQuantize, x_host: torch.Tensor, out: torch.Tensor = None, wait: bool = True):
        quantize = mod._flags()[1]
        if out is None:
            out = torch.empty_like(x_host).pin_memory() if quantize else x_host
        return self.run_many([mod], x_host, [out], wait=wait)[0]

    def run_many(self, mods, x_host: torch.Tensor, outs, wait: bool = True):
        """Several fake-quantizers over the SAME host tensor (a format sweep, or the activation quantizers of sibling
        consumers): every chunk is uploaded once and each module's result is downloaded into its own `outs[k]`.
        wait=True (default): returns when the results ARE in `outs` -- the host blocks on one event per side stream,
        recorded after its last download -- so the caller may read `outs` and reuse `x_host` immediately, like
        `mod(x.cuda()).cpu()`.  wait=False: returns right after enqueueing; the copies are ordered before later work
        on the current CUDA stream only, and the caller must `self.synchronize()` (or synchronize the device)
        before touching `outs` or modifying `x_host` on the host."""
        assert not x_host.is_cuda and x_host.is_contiguous() and x_host.dtype == self.xin[0].dtype
        assert len(mods) == len(outs)
        n = x_host.numel()
        xf = x_host.view(-1)
        main = torch.cuda.current_stream(self.device)
        plans = [self._prepare(m, n) for m in mods]
        ofs = [o.view(-1) for o in outs]
        ready = torch.cuda.Event()
        ready.record(main)
        for c, start in enumerate(range(0, n, self.chunk)):
            i = c % len(self.streams)
            m = min(self.chunk, n - start)
            with torch.cuda.stream(self.streams[i]):
                self.streams[i].wait_event(ready)
                xin, yout = self.xin[i][:m], self.yout[i][:m]
                xin.copy_(xf[start:start + m], non_blocking=True)
                for mod, of, (quantize, observe, amax_slot) in zip(mods, ofs, plans):
                    if quantize:
                        # same stream: the download of the previous module's chunk has drained `yout` before this
                        # kernel overwrites it (the kernel is ~100x faster than the link, nothing is lost)
                        _C.fq_forward(xin, yout, 1, 1, m, mod._fmt, mod.scale.reshape(1), amax_slot, mod.lut)
                        of[start:start + m].copy_(yout, non_blocking=True)
                    elif observe:
                        _C.amax(xin, 1, 1, m, amax_slot)
        self._done = []
        for s in self.streams:
            main.wait_stream(s)
            ev = torch.cuda.Event()
            ev.record(s)
            self._done.append(ev)
        if wait:
            self.synchronize()
        return outs

    def synchronize(self):
        """Block the host until every download of the last run()/run_many() call has landed in host memory."""
        for ev in getattr(self, "_done", ()):
            ev.synchronize()
        self._done = []


def fake_quantize_host(mod, x_host, out=None, device="cuda:0", chunk_elems=1 << 23, depth=3):
    """One-shot convenience wrapper (allocates the staging buffers each call; keep a HostPipeline to reuse them).
    Synchronous: the result is in host memory when it returns."""
    return HostPipeline(device, x_host.dtype, chunk_elems, depth).run(mod, x_host, out, wait=True)
