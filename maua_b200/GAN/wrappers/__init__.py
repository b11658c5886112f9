"""Generator-wrapper API: mirrors maua/GAN/wrappers/__init__.py:20-112 (MauaMapper / MauaSynthesizer /
MauaGenerator.render / get_generator_class) without the reference's import-time global torch flags.

``render`` keeps the reference contract -- ``inputs`` is a dict of ``[T, ...]`` tensors, frames come back
as ``[B,3,H,W]`` in [0,1] -- but feeds batches from pinned host memory straight into the sm_100a
synthesis pipeline instead of a DataLoader + fp16 module cast.
"""
from typing import Generator

import torch


class MauaMapper(torch.nn.Module):
    """latent_z (+ conditioning, truncation) -> W+; implemented per architecture."""

    def forward(self):
        raise NotImplementedError()


class MauaSynthesizer(torch.nn.Module):
    """W+ (+ per-frame controls) -> images.  ``_hook_handles`` lists what change_output_resolution installed (objects with
    ``remove()``: torch hook handles in the reference, resize handles of the native network here)."""

    _hook_handles = []

    def forward(self):
        raise NotImplementedError()

    def change_output_resolution(self):
        raise NotImplementedError()

    def refresh_model_hooks(self):
        installed, self._hook_handles = self._hook_handles, []
        for handle in installed:
            handle.remove()


def _batches(inputs, batch_size, device):
    """Pinned host staging + async H2D per batch (reference: TensorDataset/DataLoader, __init__.py:63-75)."""
    keys = list(inputs.keys())
    T = len(next(iter(inputs.values())))
    staged = {}
    for k in keys:
        v = inputs[k]
        if len(v) != T:
            raise ValueError("all render inputs must share the leading (frame) dimension")
        v = v.detach()
        if not v.is_cuda:
            v = v.cpu().contiguous()
            if torch.cuda.is_available():
                v = v.pin_memory()
        staged[k] = v
    for i in range(0, T, batch_size):
        yield {k: staged[k][i:i + batch_size].to(device, non_blocking=True) for k in keys}


class MauaGenerator(torch.nn.Module):
    MapperCls = None
    SynthesizerCls = None

    def __init__(self, mapper_kwargs={}, synthesizer_kwargs={}) -> None:
        super().__init__()
        cls = type(self)
        self.mapper = cls.MapperCls(**mapper_kwargs)
        self.synthesizer = cls.SynthesizerCls(**synthesizer_kwargs)

    def forward(self):
        raise NotImplementedError()

    def render(
        self,
        inputs,
        batch_size=32,
        postprocess_fn=lambda x: x,
        device=torch.device("cuda" if torch.cuda.is_available() else "cpu"),
        fp16=True,
        batched=True,
        verbose=False,
    ) -> Generator[torch.Tensor, None, None]:
        # fp16 is accepted for signature compatibility: every layer already runs fp16 operands with
        # fp32 accumulation on the tensor cores (what the reference's force_half does, :78-86).
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("maua_b200 renders on CUDA only (no CPU fallback)")
        self.to(device)
        it = _batches(inputs, batch_size, device)
        if verbose:
            from tqdm import tqdm

            it = tqdm(it, smoothing=0.8, unit_scale=batch_size, unit="img")
        for batch in it:
            # (x + 1) / 2 .clamp(0, 1) of the reference (:93) is fused into the network's last kernel (MB_OUT_F32_NCHW_01)
            frame_batch = self.synthesizer.forward(**batch, out_fmt="f32_01")
            frame_batch = postprocess_fn(frame_batch)
            if batched:
                yield frame_batch
            else:
                for frame in frame_batch:
                    yield frame[None]


_ARCHITECTURES = {"stylegan3": (".stylegan3", "StyleGAN3"), "stylegan2": (".stylegan2", "StyleGAN2")}


def get_generator_class(architecture: str) -> MauaGenerator:
    """"stylegan2" | "stylegan3" -> generator class (imported on first use)."""
    import importlib

    if architecture not in _ARCHITECTURES:
        raise Exception(f"Architecture not found: {architecture}")
    module, name = _ARCHITECTURES[architecture]
    return getattr(importlib.import_module(module, __name__), name)
