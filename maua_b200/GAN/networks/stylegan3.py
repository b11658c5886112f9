"""Host-side StyleGAN3 generator modules backed by libmaua_b200.

Mirrors the module / parameter / attribute surface of the network the reference imports as
``from ..nv.networks import stylegan3`` (maua/GAN/wrappers/stylegan3.py:12; un-vendored submodule
maua-maua-maua/nvGAN @ 7809c05 = fork of NVlabs/stylegan3 training/networks_stylegan3.py) so state
dicts, ``G_synth.input.affine.bias.data.add_()`` (wrappers/stylegan3.py:56) and ``.layer_names`` /
``.out_size`` accesses (:75,:107) work unchanged.  The torch modules only HOLD parameters; the
synthesis arithmetic runs in hand-written sm_100a kernels behind ``mb_net_forward``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.signal
import scipy.special
import torch

from ... import _lib
from ._native import NativeNet


def sg3_cfg(w_dim=512, img_resolution=1024, img_channels=3, channel_base=32768, channel_max=512, num_layers=14,
            num_critical=2, first_cutoff=2, first_stopband=2 ** 2.1, last_stopband_rel=2 ** 0.3, margin_size=10,
            output_scale=0.25, conv_kernel=3, filter_size=6, lrelu_upsampling=2, use_radial_filters=False,
            conv_clamp=256, num_fp16_res=4, **unused):
    if num_fp16_res != 4:
        raise ValueError("maua_b200 computes every layer with fp16 operands / fp32 accumulation; num_fp16_res is fixed")
    if unused:
        raise TypeError(f"unsupported SynthesisNetwork arguments: {sorted(unused)}")
    return _lib.SG3Cfg(w_dim, img_resolution, img_channels, channel_base, channel_max, num_layers, num_critical,
                       conv_kernel, filter_size, lrelu_upsampling, int(bool(use_radial_filters)), margin_size,
                       float(first_cutoff), float(first_stopband), float(last_stopband_rel), float(output_scale),
                       float(conv_clamp))


def sg3_geometry(cfg):
    """Per-layer geometry from the library (pure host arithmetic, works without a GPU)."""
    lib = _lib.load()
    layers = (_lib.SG3Layer * (cfg.num_layers + 1))()
    ch, size = C.c_int32(), C.c_int32()
    sr, bw = C.c_double(), C.c_double()
    _lib.check(lib.mb_sg3_geometry(C.byref(cfg), layers, C.byref(ch), C.byref(size), C.byref(sr), C.byref(bw)))
    out = []
    for g in layers:
        d = {name: getattr(g, name) for name, _ in _lib.SG3Layer._fields_}
        d["name"] = g.name.decode()
        out.append(d)
    return dict(layers=out, input=dict(channels=ch.value, size=size.value, sampling_rate=sr.value, bandwidth=bw.value))


def design_lowpass_filter(numtaps, cutoff, width, fs, radial=False):
    """Kaiser-windowed low-pass design of upstream SynthesisLayer.design_lowpass_filter (host, scipy)."""
    if numtaps == 1:
        return None
    if not radial:
        return torch.as_tensor(scipy.signal.firwin(numtaps=numtaps, cutoff=cutoff, width=width, fs=fs),
                               dtype=torch.float32)
    x = (np.arange(numtaps) - (numtaps - 1) / 2) / fs
    r = np.hypot(*np.meshgrid(x, x))
    f = scipy.special.j1(2 * cutoff * (np.pi * r)) / (np.pi * r)
    beta = scipy.signal.kaiser_beta(scipy.signal.kaiser_atten(numtaps, width / (fs / 2)))
    w = np.kaiser(numtaps, beta)
    f *= np.outer(w, w)
    f /= np.sum(f)
    return torch.as_tensor(f, dtype=torch.float32)


class FullyConnectedLayer(torch.nn.Module):
    """Parameter holder with upstream's init order (weight drawn first, then bias)."""

    def __init__(self, in_features, out_features, activation="linear", bias=True, lr_multiplier=1.0, weight_init=1.0,
                 bias_init=0.0):
        super().__init__()
        self.in_features, self.out_features, self.activation = in_features, out_features, activation
        self.weight = torch.nn.Parameter(torch.randn([out_features, in_features]) * (weight_init / lr_multiplier))
        b = np.broadcast_to(np.asarray(bias_init, dtype=np.float32), [out_features])
        self.bias = torch.nn.Parameter(torch.from_numpy(b / lr_multiplier)) if bias else None
        self.weight_gain = lr_multiplier / np.sqrt(in_features)
        self.bias_gain = lr_multiplier

    def forward(self, x):
        # Only used off the hot path (mapping network, avg_shift); plain library GEMM.
        w = self.weight.to(x.dtype) * self.weight_gain
        y = x.matmul(w.t())
        if self.bias is not None:
            y = y + self.bias.to(x.dtype) * self.bias_gain
        if self.activation == "lrelu":
            y = torch.nn.functional.leaky_relu(y, 0.2) * np.sqrt(2)
        return y


class MappingNetwork(torch.nn.Module):
    """z -> w (2 FC layers).  Off the render hot path: runs once per key latent, plain torch."""

    def __init__(self, z_dim, c_dim, w_dim, num_ws, num_layers=2, lr_multiplier=0.01, w_avg_beta=0.998):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim, self.num_ws, self.num_layers = z_dim, c_dim, w_dim, num_ws, num_layers
        self.embed = FullyConnectedLayer(c_dim, w_dim) if c_dim > 0 else None
        features = [z_dim + (w_dim if c_dim > 0 else 0)] + [w_dim] * num_layers
        for idx, (i, o) in enumerate(zip(features[:-1], features[1:])):
            setattr(self, f"fc{idx}", FullyConnectedLayer(i, o, activation="lrelu", lr_multiplier=lr_multiplier))
        self.register_buffer("w_avg", torch.zeros([w_dim]))

    def forward(self, z, c=None, truncation_psi=1, truncation_cutoff=None, update_emas=False):
        x = z.to(torch.float32)
        x = x * (x.square().mean(1, keepdim=True) + 1e-8).rsqrt()
        if self.c_dim > 0:
            y = self.embed(c.to(torch.float32))
            y = y * (y.square().mean(1, keepdim=True) + 1e-8).rsqrt()
            x = torch.cat([x, y], dim=1)
        for idx in range(self.num_layers):
            x = getattr(self, f"fc{idx}")(x)
        x = x.unsqueeze(1).repeat([1, self.num_ws, 1])
        if truncation_psi != 1:
            x[:, :truncation_cutoff] = self.w_avg.lerp(x[:, :truncation_cutoff], truncation_psi)
        return x


class SynthesisInput(torch.nn.Module):
    def __init__(self, w_dim, channels, size, sampling_rate, bandwidth):
        super().__init__()
        self.w_dim, self.channels = w_dim, channels
        self.size = np.broadcast_to(np.asarray(size), [2])
        self.sampling_rate, self.bandwidth = sampling_rate, bandwidth
        freqs = torch.randn([channels, 2])
        radii = freqs.square().sum(dim=1, keepdim=True).sqrt()
        freqs /= radii * radii.square().exp().pow(0.25)
        freqs *= bandwidth
        phases = torch.rand([channels]) - 0.5
        self.weight = torch.nn.Parameter(torch.randn([channels, channels]))
        self.affine = FullyConnectedLayer(w_dim, 4, weight_init=0, bias_init=[1, 0, 0, 0])
        self.register_buffer("transform", torch.eye(3, 3))
        self.register_buffer("freqs", freqs)
        self.register_buffer("phases", phases)


class SynthesisLayer(torch.nn.Module):
    def __init__(self, w_dim, g, conv_clamp=256):
        super().__init__()
        self.w_dim, self.is_torgb, self.conv_clamp = w_dim, bool(g["is_torgb"]), conv_clamp
        self.is_critically_sampled = bool(g["is_critically_sampled"])
        self.use_fp16 = bool(g["use_fp16"])
        self.in_channels, self.out_channels = g["in_channels"], g["out_channels"]
        self.in_size = np.broadcast_to(np.asarray(g["in_size"]), [2])
        self.out_size = np.broadcast_to(np.asarray(g["out_size"]), [2])
        self.in_sampling_rate, self.out_sampling_rate = g["in_sampling_rate"], g["out_sampling_rate"]
        self.tmp_sampling_rate = g["tmp_sampling_rate"]
        self.in_cutoff, self.out_cutoff = g["in_cutoff"], g["out_cutoff"]
        self.in_half_width, self.out_half_width = g["in_half_width"], g["out_half_width"]
        self.conv_kernel, self.up_factor, self.down_factor = g["conv_kernel"], g["up"], g["down"]
        self.up_taps, self.down_taps = g["up_taps"], g["down_taps"]
        self.padding = [g["pad_lo"], g["pad_hi"], g["pad_lo"], g["pad_hi"]]
        self.affine = FullyConnectedLayer(w_dim, self.in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(
            torch.randn([self.out_channels, self.in_channels, self.conv_kernel, self.conv_kernel]))
        self.bias = torch.nn.Parameter(torch.zeros([self.out_channels]))
        self.register_buffer("magnitude_ema", torch.ones([]))
        self.register_buffer("up_filter", design_lowpass_filter(
            self.up_taps, self.in_cutoff, self.in_half_width * 2, self.tmp_sampling_rate))
        self.register_buffer("down_filter", design_lowpass_filter(
            self.down_taps, self.out_cutoff, self.out_half_width * 2, self.tmp_sampling_rate,
            radial=bool(g["down_radial"])))


class SynthesisNetwork(NativeNet):
    """``SynthesisNetwork(w_dim=512, img_resolution=1024, img_channels=3)`` as built at
    maua/GAN/wrappers/stylegan3.py:33; forward(ws [B,num_ws,w_dim]) -> float32 [B,3,H,W]."""

    def __init__(self, w_dim, img_resolution, img_channels, **synthesis_kwargs):
        super().__init__()
        self.w_dim, self.img_resolution, self.img_channels = w_dim, img_resolution, img_channels
        self._cfg = sg3_cfg(w_dim=w_dim, img_resolution=img_resolution, img_channels=img_channels, **synthesis_kwargs)
        geo = sg3_geometry(self._cfg)
        self.geometry = geo
        self.num_layers = self._cfg.num_layers
        self.num_ws = self.num_layers + 2
        self.output_scale = self._cfg.output_scale
        self.input = SynthesisInput(w_dim=w_dim, **geo["input"])
        self.layer_names = []
        for g in geo["layers"]:
            setattr(self, g["name"], SynthesisLayer(w_dim, g, conv_clamp=self._cfg.conv_clamp))
            self.layer_names.append(g["name"])
        self._init_native()

    def _create(self, lib, handle_ref):
        _lib.check(lib.mb_sg3_create(C.byref(self._cfg), handle_ref))

    # ---- output-size hook ---------------------------------------------------------------
    def set_resize(self, module, strategy=None, a=0, b=0):
        """Resize the output of one synthesis module (0 = ``input``, i = ``layer_names[i-1]``) inside the forward:
        what the reference does with ``register_forward_hook(get_hook(...))`` (maua/GAN/wrappers/stylegan3.py:77-78,
        :96-117).  strategy "stretch": (a, b) = target (height, width), bicubic; "pad-zero": (a, b) = zero rows / columns
        added on each side (negative crops); None removes the hook."""
        code = {None: _lib.MB_RESIZE_NONE, "stretch": _lib.MB_RESIZE_STRETCH, "pad-zero": _lib.MB_RESIZE_PAD_ZERO}
        if strategy not in code:
            raise Exception(f"Resize strategy not found: {strategy}")
        self._resize = (int(module), code[strategy], int(a), int(b))
        self._workspace = {}
        if self._net is not None:  # otherwise applied when the library handle is created (first forward)
            _lib.check(_lib.load().mb_net_set_resize(self._net, *self._resize))

    def _handle(self):
        fresh = self._net is None
        net = super()._handle()
        if fresh and getattr(self, "_resize", None) is not None:
            _lib.check(_lib.load().mb_net_set_resize(net, *self._resize))
        return net

    def output_hw(self):
        """(height, width) of the image the next forward writes."""
        h, w = C.c_int32(), C.c_int32()
        module, code, a, b = getattr(self, "_resize", None) or (0, _lib.MB_RESIZE_NONE, 0, 0)
        _lib.check(_lib.load().mb_sg3_resized_output(C.byref(self._cfg), module, code, a, b, C.byref(h), C.byref(w)))  # host only
        return h.value, w.value

    # ---- forward ------------------------------------------------------------------------
    def forward(self, ws, out_fmt="f32", out=None, transforms=None, **unused):
        """transforms: optional float [B,3,3], one ``input.transform`` per frame (extension of the reference, whose
        buffer is shared by the batch: maua/GAN/wrappers/stylegan3.py:59)."""
        if not ws.is_cuda:
            raise RuntimeError("maua_b200 SynthesisNetwork.forward needs CUDA latents: there is no CPU path")
        lib = _lib.load()
        device = ws.device
        with torch.cuda.device(device):
            self._sync_params(device)
            ws32 = ws.detach().to(torch.float32).contiguous()
            B = ws32.shape[0]
            if ws32.ndim == 3 and ws32.shape[1] > self.num_ws and ws32.shape[2] == self.w_dim:
                # the reference's StyleGANMapper always emits num_ws=18 (wrappers/stylegan.py:18)
                ws32 = ws32[:, :self.num_ws].contiguous()
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
            elif tuple(out.shape) != shape or out.dtype != dtype or not out.is_contiguous() or out.device != device:
                raise ValueError(f"out must be a contiguous {dtype} tensor of shape {shape} on {device}")
            wsb, off, nbytes = self._get_workspace(B, device)
            if transforms is None:
                _lib.check(lib.mb_net_forward(self._handle(), _lib.ptr(ws32), None, B, _lib.ptr(out), fmt,
                                              C.c_void_p(wsb.data_ptr() + off), nbytes, _lib.stream_ptr()))
            else:
                xf = transforms.detach().to(device=device, dtype=torch.float32).contiguous()
                if tuple(xf.shape) != (B, 3, 3):
                    raise ValueError(f"transforms must be [B,3,3] = {(B, 3, 3)}, got {tuple(xf.shape)}")
                _lib.check(lib.mb_net_forward_xf(self._handle(), _lib.ptr(ws32), _lib.ptr(xf), B, _lib.ptr(out), fmt,
                                                 C.c_void_p(wsb.data_ptr() + off), nbytes, _lib.stream_ptr()))
        return out


class Generator(torch.nn.Module):
    """Mapping + synthesis pair the checkpoint loaders construct (upstream networks_stylegan3.py Generator (built by maua/GAN/load.py:131-139))."""

    def __init__(self, z_dim, c_dim, w_dim, img_resolution, img_channels, mapping_kwargs={}, **synthesis_kwargs):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim = z_dim, c_dim, w_dim
        self.img_resolution, self.img_channels = img_resolution, img_channels
        self.synthesis = SynthesisNetwork(w_dim=w_dim, img_resolution=img_resolution, img_channels=img_channels,
                                          **synthesis_kwargs)
        self.num_ws = self.synthesis.num_ws
        self.mapping = MappingNetwork(z_dim=z_dim, c_dim=c_dim, w_dim=w_dim, num_ws=self.num_ws, **mapping_kwargs)

    def forward(self, z, c=None, truncation_psi=1.0, truncation_cutoff=None, **synthesis_kwargs):
        ws = self.mapping(z, c, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff)
        return self.synthesis(ws, **synthesis_kwargs)


SG3_R_KWARGS = dict(conv_kernel=1, channel_base=65536, channel_max=1024, use_radial_filters=True)
