"""Host-side StyleGAN2 generator modules backed by libmaua_b200.

Mirrors the module / parameter / attribute surface of the reference's in-tree inference network
(maua/GAN/wrappers/inference/stylegan2.py:29-436: FullyConnectedLayer, MappingNetwork, SynthesisLayer, ToRGBLayer,
SynthesisBlock, SynthesisNetwork with ``bs`` ModuleList) so its state dicts load unchanged and the wrapper's
accesses (``G_synth.bs[i].conv1.noise_const``, ``.block_resolutions``, ``.num_ws``; wrappers/stylegan2.py:40-52,83-96)
keep working.  The torch modules only HOLD parameters; the synthesis arithmetic (modulated_conv2d / upfirdn2d /
bias_act of inference/ops.py) runs in hand-written sm_100a kernels behind ``mb_net_forward`` (csrc/sg2.cu).
"""
from __future__ import annotations

import ctypes as C
from math import sqrt

import numpy as np
import torch

from ... import _lib
from ._native import NativeNet


def setup_filter(f=(1, 3, 3, 1)):
    """inference/ops.py:236-256 (buffer kept so reference state dicts load with strict=True)."""
    f = torch.as_tensor(f, dtype=torch.float32)
    f = torch.outer(f, f)
    return f / f.sum()


class FullyConnectedLayer(torch.nn.Module):
    """inference/stylegan2.py:29-58, including its quirk: the non-linear branch multiplies by ``w.T`` (:57)."""

    def __init__(self, in_features, out_features, bias=True, activation="linear", lr_multiplier=1.0, bias_init=0.0):
        super().__init__()
        self.in_features, self.out_features, self.activation = in_features, out_features, activation
        self.weight = torch.nn.Parameter(torch.randn([out_features, in_features]) / lr_multiplier)
        self.bias = torch.nn.Parameter(torch.full([out_features], np.float32(bias_init))) if bias else None
        self.weight_gain = lr_multiplier / sqrt(in_features)
        self.bias_gain = lr_multiplier
        self.standard_matmul = False

    def forward(self, x):
        # Off the render hot path (mapping network: once per key latent); plain library GEMM.
        w = self.weight.to(x.dtype) * self.weight_gain
        b = self.bias
        if b is not None and self.bias_gain != 1.0:
            b = b * self.bias_gain
        if self.activation == "linear":
            return torch.nn.functional.linear(x, w, b)
        # the reference's inference network multiplies by w itself here (:57, square layers only); checkpoints in the
        # training layout were trained with x @ w.T and set standard_matmul (GAN/load.py::_finish_sg2)
        y = torch.nn.functional.linear(x, w if self.standard_matmul else w.T, None)
        if b is not None:
            y = y + b.to(y.dtype)
        return torch.nn.functional.leaky_relu(y, 0.2) * sqrt(2)


class MappingNetwork(torch.nn.Module):
    """inference/stylegan2.py:116-192.  z -> w, 8 FC layers; off the hot path, plain torch."""

    def __init__(self, z_dim, c_dim, w_dim, num_ws, num_layers=8, embed_features=None, layer_features=None,
                 activation="lrelu", lr_multiplier=0.01, w_avg_beta=0.998):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim, self.num_ws, self.num_layers = z_dim, c_dim, w_dim, num_ws, num_layers
        if embed_features is None:
            embed_features = w_dim
        if c_dim == 0:
            embed_features = 0
        if layer_features is None:
            layer_features = w_dim
        features = [z_dim + embed_features] + [layer_features] * (num_layers - 1) + [w_dim]
        if c_dim > 0:
            self.embed = FullyConnectedLayer(c_dim, embed_features)
        self.fcs = torch.nn.ModuleList([
            FullyConnectedLayer(i, o, activation=activation, lr_multiplier=lr_multiplier)
            for i, o in zip(features[:-1], features[1:])])
        self.register_buffer("w_avg", torch.zeros([w_dim]))

    def forward(self, z, c=None, truncation_psi=1.0, truncation_cutoff=None):
        x = z.to(torch.float32)
        x = x * (x.square().mean(1, keepdim=True) + 1e-8).rsqrt()
        if self.c_dim > 0 and c is not None:
            y = self.embed(c.to(torch.float32))
            y = y * (y.square().mean(1, keepdim=True) + 1e-8).rsqrt()
            x = torch.cat([x, y], dim=1)
        for fc in self.fcs:
            x = fc(x)
        x = x.unsqueeze(1).repeat([1, self.num_ws, 1])
        if truncation_psi != 1.0:
            if truncation_cutoff is None:
                x = self.w_avg.lerp(x, truncation_psi)
            else:
                x[:, :truncation_cutoff] = self.w_avg.lerp(x[:, :truncation_cutoff], truncation_psi)
        return x


class SynthesisLayer(torch.nn.Module):
    """Parameter holder, init order of inference/stylegan2.py:195-232."""

    def __init__(self, in_channels, out_channels, w_dim, resolution, kernel_size=3, up=1, conv_clamp=256.0):
        super().__init__()
        self.in_channels, self.out_channels, self.w_dim = in_channels, out_channels, w_dim
        self.resolution, self.up, self.use_noise, self.activation, self.conv_clamp = resolution, up, True, "lrelu", conv_clamp
        self.register_buffer("resample_filter", setup_filter())
        self.padding = kernel_size // 2
        self.act_gain = sqrt(2)
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]))
        self.register_buffer("noise_const", torch.randn([resolution, resolution]))
        # scale of the noise map (constant or swapped in per frame): 1 = the reference's inference network, which adds it
        # unscaled (inference/ops.py:184); the loaders set the trained strength for training-layout checkpoints
        self.noise_strength = 1.0
        self.noise_adjusted = False
        self.bias = torch.nn.Parameter(torch.zeros([out_channels]))


class ToRGBLayer(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, kernel_size=1, conv_clamp=256.0):
        super().__init__()
        self.in_channels, self.out_channels, self.w_dim, self.conv_clamp = in_channels, out_channels, w_dim, conv_clamp
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]))
        self.bias = torch.nn.Parameter(torch.zeros([out_channels]))
        self.weight_gain = 1 / sqrt(in_channels * (kernel_size ** 2))


class SynthesisBlock(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, resolution, img_channels, is_last, architecture="skip",
                 conv_clamp=256.0, use_fp16=False):
        super().__init__()
        if architecture != "skip":
            raise ValueError("maua_b200 builds the 'skip' architecture (the reference's default, stylegan2.py:280)")
        self.in_channels, self.w_dim, self.resolution, self.img_channels = in_channels, w_dim, resolution, img_channels
        self.is_last, self.architecture, self.use_fp16 = is_last, architecture, use_fp16
        self.register_buffer("resample_filter", setup_filter())
        self.num_conv, self.num_torgb = 0, 0
        self.const = None
        if in_channels == 0:
            self.const = torch.nn.Parameter(torch.randn([out_channels, resolution, resolution]))
        self.conv0 = None
        if in_channels != 0:
            self.conv0 = SynthesisLayer(in_channels, out_channels, w_dim, resolution, up=2, conv_clamp=conv_clamp)
            self.num_conv += 1
        self.conv1 = SynthesisLayer(out_channels, out_channels, w_dim, resolution, conv_clamp=conv_clamp)
        self.num_conv += 1
        self.torgb = ToRGBLayer(out_channels, img_channels, w_dim, conv_clamp=conv_clamp)
        self.num_torgb += 1
        self.skip = None


class SynthesisNetwork(NativeNet):
    """``SynthesisNetwork(w_dim=512, img_resolution=R, img_channels=3)`` as built at maua/GAN/wrappers/stylegan2.py:34-36;
    forward(ws [B,num_ws,w_dim], noise_mode='const') -> float32 [B,3,R,R]."""

    def __init__(self, w_dim, img_resolution, img_channels, channel_base=32768, channel_max=512, num_fp16_res=0,
                 **block_kwargs):
        super().__init__()
        if block_kwargs.get("conv_clamp", 256.0) != 256.0:
            raise ValueError("conv_clamp is fixed at 256 (the reference's SynthesisBlock default)")
        self.w_dim, self.img_resolution, self.img_channels = w_dim, img_resolution, img_channels
        self.img_resolution_log2 = int(np.log2(img_resolution))
        self.num_fp16_res = num_fp16_res
        self.channel_base, self.channel_max = channel_base, channel_max
        self.block_resolutions = [2 ** i for i in range(2, self.img_resolution_log2 + 1)]
        channels = {res: min(channel_base // res, channel_max) for res in self.block_resolutions}
        self.num_ws = 0
        bs = []
        for res in self.block_resolutions:
            block = SynthesisBlock(channels[res // 2] if res > 4 else 0, channels[res], w_dim=w_dim, resolution=res,
                                   img_channels=img_channels, is_last=res == img_resolution,
                                   architecture=block_kwargs.get("architecture", "skip"))
            self.num_ws += block.num_conv
            if block.is_last:
                self.num_ws += block.num_torgb
            bs.append(block)
        self.bs = torch.nn.ModuleList(bs)
        self._init_native()

    def _create(self, lib, handle_ref):
        _lib.check(lib.mb_sg2_create(self.w_dim, self.img_resolution, self.img_channels, self.channel_base,
                                     self.channel_max, handle_ref))

    def _is_volatile(self, name, t):
        # per-frame noise maps are swapped in on every call by the wrapper (wrappers/stylegan2.py:83-96): always upload
        return name.endswith("noise_const") and t.ndim != 2

    def _noise_strength(self, name):
        return float(getattr(self.get_submodule(name.rsplit(".", 1)[0]), "noise_strength", 1.0))

    def _param_key(self, name, t):
        key = super()._param_key(name, t)
        return key + (self._noise_strength(name),) if name.endswith("noise_const") else key

    def _prepare_param(self, name, d):
        if name.endswith("noise_const"):
            strength = self._noise_strength(name)
            if strength != 1.0:
                d = d * strength
        return d

    # ---- output-size hook (maua/GAN/wrappers/stylegan2.py:104-151) ---------------------------------------------------
    RESIZE_MODES = {"stretch": 0, "constant": 1, "reflect": 2, "replicate": 3, "circular": 4}

    def set_resize(self, layer, mode, target_hw, pads=(0, 0), value=0.0, noise=None, stats=None):
        """Resize the output of layer_names[layer] to target_hw = (h, w) in every following forward (mb_sg2_set_resize).
        mode: a RESIZE_MODES key; pads = (top, left) leading pads of the pad modes; noise: CUDA float32 [C, h, w] added to the
        resized features (layer 0: the resized constant input itself); stats: CUDA float32 [2, C] receiving mean / std of the
        resized features.  layer=None clears the hook."""
        lib = _lib.load()
        if layer is None:
            _lib.check(lib.mb_sg2_set_resize(self._handle(), -1, 0, 1, 1, 0, 0, 0.0, None, None))
            self._resize = None
        else:
            h, w = int(target_hw[0]), int(target_hw[1])
            for t in (noise, stats):
                if t is not None and not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                    raise ValueError("set_resize: noise / stats must be contiguous float32 CUDA tensors")
            _lib.check(lib.mb_sg2_set_resize(self._handle(), int(layer), self.RESIZE_MODES[mode], h, w, int(pads[0]), int(pads[1]),
                                             float(value), None if noise is None else _lib.ptr(noise),
                                             None if stats is None else _lib.ptr(stats)))
            self._resize = (int(layer), mode, (h, w), noise, stats)    # keeps the device tensors alive while the hook is set
        self._workspace = {}   # the workspace size follows the layer geometry

    def output_hw(self):
        """(height, width) of the image the next forward writes."""
        h, w = C.c_int32(), C.c_int32()
        _lib.check(_lib.load().mb_net_output_shape(self._handle(), C.byref(h), C.byref(w)))
        return h.value, w.value

    def layer_resolution(self, layer):
        """Resolution of the wrapper's layer_names[layer] (maua/GAN/wrappers/stylegan2.py:48-51: entries 0 and 1 are
        bs.0.conv1, entry 2i / 2i+1 are block i's conv0 / conv1)."""
        return self.block_resolutions[layer // 2]

    def forward(self, ws, noise_mode="const", out_fmt="f32", out=None, warps=None, **unused):
        """warps: optional list of (layer, inv_mats float [B,2,3]) feature-map warps applied in list order to the output
        of layer_names[layer] (the reference's kornia forward hooks, wrappers/stylegan2.py:153-194)."""
        if noise_mode != "const":
            raise ValueError("only noise_mode='const' is built (what the reference wrapper passes, wrappers/stylegan2.py:98)")
        if not ws.is_cuda:
            raise RuntimeError("maua_b200 SynthesisNetwork.forward needs CUDA latents: there is no CPU path")
        lib = _lib.load()
        device = ws.device
        with torch.cuda.device(device):
            self._sync_params(device)
            ws32 = ws.detach().to(torch.float32).contiguous()
            B = ws32.shape[0]
            if ws32.ndim == 3 and ws32.shape[1] > self.num_ws and ws32.shape[2] == self.w_dim:
                ws32 = ws32[:, :self.num_ws].contiguous()  # the reference's mapper always emits num_ws=18
            if tuple(ws32.shape[1:]) != (self.num_ws, self.w_dim):
                raise ValueError(f"ws must be [B,{self.num_ws},{self.w_dim}], got {tuple(ws32.shape)}")
            oh, ow = self.output_hw()
            if out_fmt in ("f32", "f32_01", "f32_unit"):
                fmt = {"f32": _lib.MB_OUT_F32_NCHW, "f32_01": _lib.MB_OUT_F32_NCHW_01, "f32_unit": _lib.MB_OUT_F32_NCHW_UNIT}[out_fmt]
                shape, dtype = (B, self.img_channels, oh, ow), torch.float32
            elif out_fmt == "u8":
                fmt = _lib.MB_OUT_U8_NHWC
                shape, dtype = (B, oh, ow, self.img_channels), torch.uint8
            else:
                raise ValueError("out_fmt must be 'f32', 'f32_01', 'f32_unit' or 'u8'")
            if out is None:
                out = torch.empty(shape, device=device, dtype=dtype)
            elif tuple(out.shape) != shape or out.dtype != dtype or not out.is_contiguous():
                raise ValueError(f"out must be a contiguous {dtype} tensor of shape {shape}")
            warps = list(warps or [])
            if bool(warps) != getattr(self, "_warps_on", False):
                self._workspace = {}  # the warp ping-pong buffers change the workspace size
                self._warps_on = bool(warps)
            mats = None
            if warps:
                layers = (C.c_int32 * len(warps))(*[int(l) for l, _ in warps])
                mats = torch.stack([m.detach().to(device=device, dtype=torch.float32).reshape(-1, 2, 3).expand(B, 2, 3)
                                    for _, m in warps]).contiguous()
                _lib.check(lib.mb_sg2_set_warps(self._handle(), len(warps), layers, _lib.ptr(mats), B))
            try:
                wsb, off, nbytes = self._get_workspace(B, device)
                _lib.check(lib.mb_net_forward(self._handle(), _lib.ptr(ws32), None, B, _lib.ptr(out), fmt,
                                              C.c_void_p(wsb.data_ptr() + off), nbytes, _lib.stream_ptr()))
            finally:
                if warps:
                    _lib.check(lib.mb_sg2_set_warps(self._handle(), 0, None, None, 0))
                    mats.record_stream(torch.cuda.current_stream(device))
        return out


class Generator(torch.nn.Module):
    """Mapping + synthesis pair the checkpoint loaders construct (maua/GAN/wrappers/inference/stylegan2.py:439-472)."""

    def __init__(self, z_dim, c_dim, w_dim, img_resolution, img_channels, mapping_kwargs={}, **synthesis_kwargs):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim = z_dim, c_dim, w_dim
        self.img_resolution, self.img_channels = img_resolution, img_channels
        self.synthesis = SynthesisNetwork(w_dim=w_dim, img_resolution=img_resolution, img_channels=img_channels,
                                          **synthesis_kwargs)
        self.num_ws = self.synthesis.num_ws
        self.mapping = MappingNetwork(z_dim=z_dim, c_dim=c_dim, w_dim=w_dim, num_ws=self.num_ws, **mapping_kwargs)

    def forward(self, z, c=None, truncation_psi=1.0, truncation_cutoff=None, noise_mode="const", **synthesis_kwargs):
        ws = self.mapping(z, c, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff)
        return self.synthesis(ws, noise_mode, **synthesis_kwargs)
