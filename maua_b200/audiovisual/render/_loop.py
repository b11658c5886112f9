"""Shared frame loop of the renderers: pinned host staging -> async H2D -> synthesizer -> uint8 frames."""
from __future__ import annotations

import torch


def frame_batches(synthesizer, inputs, batch_size, device, out_fmt="f32"):
    """Yield (start, frames) per batch; inputs is the reference's dict of [T,...] tensors (host tensors are staged in pinned
    memory once and copied batch by batch, asynchronously).  out_fmt "f32": the synthesizer's raw ~[-1,1] output; "f32_unit":
    (x + 1) / 2, the value the reference's renderers hand to postprocess (fused into the network's last kernel)."""
    keys = list(inputs.keys())
    T = len(next(iter(inputs.values())))
    staged = {k: (v.detach() if v.is_cuda else v.detach().cpu().contiguous().pin_memory()) for k, v in inputs.items()}
    synthesizer = synthesizer.to(device)
    for i in range(0, T, batch_size):
        batch = {k: staged[k][i:i + batch_size].to(device, non_blocking=True) for k in keys}
        fmt = out_fmt() if callable(out_fmt) else out_fmt     # a callable is asked per batch (the renderer may switch formats)
        yield i, (synthesizer(**batch) if fmt == "f32" else synthesizer(**batch, out_fmt=fmt))


def to_uint8(frames01):
    """clamp -> *255 -> round -> uint8 (tensor2bytes, maua/ops/io.py:47-70)."""
    return frames01.clamp(0, 1).mul(255).round().to(torch.uint8)


def frames_to_rgb24(frames, out=None, value_range=(0, 1)):
    """tensor2bytes (maua/ops/io.py:47-70) up to the host copy, as ONE kernel: CUDA float [B,C,H,W] in `value_range` ->
    uint8 [B,H,W,C] (written into `out[:B]` when given)."""
    from ... import _lib

    if not frames.is_cuda:
        raise RuntimeError("frames_to_rgb24: CUDA frames only (no CPU fallback)")
    x = frames.detach().to(torch.float32).contiguous()
    B, C, H, W = x.shape
    if out is None:
        out = torch.empty(B, H, W, C, device=x.device, dtype=torch.uint8)
    dst = out[:B]
    if tuple(dst.shape) != (B, H, W, C) or dst.dtype != torch.uint8 or not dst.is_contiguous():
        raise ValueError("frames_to_rgb24: `out` must be a contiguous uint8 [>=B, H, W, C] CUDA tensor")
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().mb_frames_to_rgb24(_lib.ptr(x), _lib.ptr(dst), B, C, H, W, float(value_range[0]), float(value_range[1]),
                                                  _lib.stream_ptr()))
    return dst


class AsyncFrameDownloader:
    """Finished uint8 frame batches -> pinned host ring on a side stream, so the device->host copy of batch i
    overlaps the synthesis of batch i+1 (the reference copies each batch synchronously, render/memmap.py:31-33).

        buf = dl.device_buffer(i)          # waits (on the compute stream) until copy i-depth has drained
        net(ws, out_fmt="u8", out=buf)
        dl.download(i)                     # enqueue the copy behind the kernels of batch i
        frames = dl.host(i)                # blocks the host until batch i is in pinned memory
    """

    def __init__(self, batch_shape, device, depth=2):
        self.depth = depth
        self.dev = [torch.empty(batch_shape, dtype=torch.uint8, device=device) for _ in range(depth)]
        self.pinned = [torch.empty(batch_shape, dtype=torch.uint8).pin_memory() for _ in range(depth)]
        self.copy_stream = torch.cuda.Stream(device=device)
        self.ready = [torch.cuda.Event() for _ in range(depth)]   # kernels of the batch are done
        self.done = [None] * depth                                 # copy of the batch is done

    def device_buffer(self, i):
        k = i % self.depth
        if self.done[k] is not None:
            torch.cuda.current_stream().wait_event(self.done[k])
        return self.dev[k]

    def download(self, i):
        k = i % self.depth
        self.ready[k].record(torch.cuda.current_stream())
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.ready[k])
            self.pinned[k].copy_(self.dev[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.done[k] = ev

    def host(self, i):
        k = i % self.depth
        self.done[k].synchronize()
        return self.pinned[k]

    def synchronize(self):
        self.copy_stream.synchronize()
