"""MemMap renderer: mirror of maua/audiovisual/render/memmap.py:12-34 (frames -> uint8 .npy -> np.memmap).

Same call signature and result ([T,3,H,W] uint8 memmap handed to ``postprocess``); frames are produced in
batches on the GPU and land in the .npy through a pinned double buffer instead of one D2H copy per frame.
"""
import os

import numpy as np
import torch

from . import Renderer
from ._loop import frame_batches


class MemMap(Renderer):
    def __init__(self, cache_file="workspace/frames_memmap.npy", batch_size=8) -> None:
        super().__init__()
        self.cache_file, self.batch_size = cache_file, batch_size

    def __call__(self, synthesizer, inputs, postprocess):
        T = len(next(iter(inputs.values())))
        os.makedirs(os.path.dirname(os.path.abspath(self.cache_file)), exist_ok=True)
        if os.path.exists(self.cache_file):
            os.remove(self.cache_file)
        out = None
        pinned = None
        pending = None  # (start, count, buffer index)
        for start, frames in frame_batches(synthesizer, inputs, self.batch_size, self.device, out_fmt="f32_01"):
            u8 = frames.mul(255).to(torch.uint8)  # (x+1)/2 .clamp(0,1) fused into the network; the reference truncates (astype), memmap.py:31
            if out is None:
                out = np.lib.format.open_memmap(self.cache_file, mode="w+", dtype=np.uint8, shape=(T,) + tuple(u8.shape[1:]))
                pinned = [torch.empty((self.batch_size,) + tuple(u8.shape[1:]), dtype=torch.uint8).pin_memory() for _ in range(2)]
            if pending is not None:
                torch.cuda.current_stream().synchronize()
                s, n, b = pending
                out[s:s + n] = pinned[b][:n].numpy()
            b = (start // self.batch_size) % 2
            pinned[b][: u8.shape[0]].copy_(u8, non_blocking=True)
            pending = (start, u8.shape[0], b)
        if pending is not None:
            torch.cuda.current_stream().synchronize()
            s, n, b = pending
            out[s:s + n] = pinned[b][:n].numpy()
        out.flush()
        del out
        frames = np.load(self.cache_file, mmap_mode="r")
        return postprocess(frames)
