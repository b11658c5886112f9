"""Shared frame loop of the renderers: pinned host staging -> async H2D -> synthesizer -> uint8 frames."""
from __future__ import annotations

import torch


def frame_batches(synthesizer, inputs, batch_size, device):
    """Yield (start, frames_f32[B,3,H,W] in ~[-1,1]) per batch; inputs is the reference's dict of [T,...] tensors."""
    keys = list(inputs.keys())
    T = len(next(iter(inputs.values())))
    staged = {k: (v.detach() if v.is_cuda else v.detach().cpu().contiguous().pin_memory()) for k, v in inputs.items()}
    synthesizer = synthesizer.to(device)
    for i in range(0, T, batch_size):
        batch = {k: staged[k][i:i + batch_size].to(device, non_blocking=True) for k in keys}
        yield i, synthesizer(**batch)


def to_uint8(frames01):
    """clamp -> *255 -> round -> uint8 (tensor2bytes, maua/ops/io.py:47-70)."""
    return frames01.clamp(0, 1).mul(255).round().to(torch.uint8)
