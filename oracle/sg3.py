"""Oracle: StyleGAN3 synthesis network, fp32 PyTorch on CPU.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The reference imports this network from an un-vendored git
submodule (``maua/GAN/wrappers/stylegan3.py:12`` ``from ..nv.networks import
stylegan3``; submodule maua-maua-maua/nvGAN @ 7809c05ff37f68db0d367df8aa52ce663b950953,
a fork of NVlabs/stylegan3 ``training/networks_stylegan3.py`` and
``torch_utils/ops/{bias_act,upfirdn2d,filtered_lrelu}.py``).  ``/root/reference/maua/GAN/nv``
is an empty directory, so this file restates the *published* upstream algorithm
(Karras et al. 2021, "Alias-Free Generative Adversarial Networks", official
implementation) and anchors it on what the reference does pin:

  * ctor call            maua/GAN/wrappers/stylegan3.py:33
                         ``SynthesisNetwork(w_dim=512, img_resolution=1024, img_channels=3)``
  * attribute surface    .input.affine / .input.transform (stylegan3.py:56-59),
                         .layer_names (:75), .<layer>.out_size (:107),
                         .img_resolution / .w_dim / .num_ws (:38-40)
  * geometry             ``layer_multipliers`` (stylegan3.py:15-19) must equal
                         img_resolution / (layer out_size - 20) -> tests/test_sg3_geometry.py

Every op is written with plain torch fp32 ops (conv2d / pad / slicing), the
"reference semantics" path of upstream (``_filtered_lrelu_ref``, ``_upfirdn2d_ref``,
``_bias_act_ref``), never a fused kernel.
"""
from __future__ import annotations

import numpy as np
import scipy.signal
import scipy.special
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# ops (upstream torch_utils/ops/*_ref)


def bias_act_ref(x, b=None, act="linear", alpha=None, gain=None, clamp=None):
    """upstream bias_act._bias_act_ref: x+b -> act -> *gain -> clamp."""
    def_alpha, def_gain = {"linear": (0.0, 1.0), "lrelu": (0.2, float(np.sqrt(2)))}[act]
    alpha = float(alpha if alpha is not None else def_alpha)
    gain = float(gain if gain is not None else def_gain)
    clamp = float(clamp if clamp is not None else -1)
    if b is not None:
        x = x + b.reshape([-1 if i == 1 else 1 for i in range(x.ndim)])
    if act == "lrelu":
        x = F.leaky_relu(x, alpha)
    if gain != 1:
        x = x * gain
    if clamp >= 0:
        x = x.clamp(-clamp, clamp)
    return x


def upfirdn2d_ref(x, f, up=1, down=1, padding=(0, 0, 0, 0), flip_filter=False, gain=1.0):
    """upstream upfirdn2d._upfirdn2d_ref.  padding = [x0, x1, y0, y1]; f is 1-D (separable) or 2-D."""
    if f is None:
        f = torch.ones([1, 1], dtype=torch.float32)
    B, C, H, W = x.shape
    padx0, padx1, pady0, pady1 = [int(p) for p in padding]
    # zero-insert
    x = x.reshape([B, C, H, 1, W, 1])
    x = F.pad(x, [0, up - 1, 0, 0, 0, up - 1])
    x = x.reshape([B, C, H * up, W * up])
    # pad or crop
    x = F.pad(x, [max(padx0, 0), max(padx1, 0), max(pady0, 0), max(pady1, 0)])
    x = x[:, :, max(-pady0, 0): x.shape[2] - max(-pady1, 0), max(-padx0, 0): x.shape[3] - max(-padx1, 0)]
    # filter
    f = f * (gain ** (f.ndim / 2))
    f = f.to(x.dtype)
    if not flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f[None, None].repeat([C, 1] + [1] * f.ndim)
    if f.ndim == 4:
        x = F.conv2d(x, f, groups=C)
    else:
        x = F.conv2d(x, f.unsqueeze(2), groups=C)
        x = F.conv2d(x, f.unsqueeze(3), groups=C)
    return x[:, :, ::down, ::down]


def filtered_lrelu_ref(x, fu=None, fd=None, b=None, up=1, down=1, padding=(0, 0, 0, 0),
                       gain=float(np.sqrt(2)), slope=0.2, clamp=None):
    """upstream filtered_lrelu._filtered_lrelu_ref."""
    x = bias_act_ref(x, b)
    x = upfirdn2d_ref(x, fu, up=up, padding=padding, gain=up ** 2)
    x = bias_act_ref(x, act="lrelu", alpha=slope, gain=gain, clamp=clamp)
    x = upfirdn2d_ref(x, fd, down=down)
    return x


def modulated_conv2d_ref(x, w, s, demodulate=True, padding=0, input_gain=None):
    """upstream networks_stylegan3.modulated_conv2d (grouped-conv formulation)."""
    B = x.shape[0]
    O, I, kh, kw = w.shape
    if demodulate:
        w = w * w.square().mean([1, 2, 3], keepdim=True).rsqrt()
        s = s * s.square().mean().rsqrt()
    w = w.unsqueeze(0) * s.unsqueeze(1).unsqueeze(3).unsqueeze(4)
    if demodulate:
        dcoefs = (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()
        w = w * dcoefs.unsqueeze(2).unsqueeze(3).unsqueeze(4)
    if input_gain is not None:
        input_gain = input_gain.expand(B, I)
        w = w * input_gain.unsqueeze(1).unsqueeze(3).unsqueeze(4)
    x = x.reshape(1, -1, *x.shape[2:])
    w = w.reshape(-1, I, kh, kw)
    x = F.conv2d(x, w.to(x.dtype), padding=padding, groups=B)
    return x.reshape(B, -1, *x.shape[2:])


# --------------------------------------------------------------------------------------
# filter design (upstream SynthesisLayer.design_lowpass_filter)


def design_lowpass_filter(numtaps, cutoff, width, fs, radial=False):
    if numtaps == 1:
        return None
    if not radial:
        f = scipy.signal.firwin(numtaps=numtaps, cutoff=cutoff, width=width, fs=fs)
        return torch.as_tensor(f, dtype=torch.float32)
    x = (np.arange(numtaps) - (numtaps - 1) / 2) / fs
    r = np.hypot(*np.meshgrid(x, x))
    f = scipy.special.j1(2 * cutoff * (np.pi * r)) / (np.pi * r)
    beta = scipy.signal.kaiser_beta(scipy.signal.kaiser_atten(numtaps, width / (fs / 2)))
    w = np.kaiser(numtaps, beta)
    f *= np.outer(w, w)
    f /= np.sum(f)
    return torch.as_tensor(f, dtype=torch.float32)


# --------------------------------------------------------------------------------------
# geometry (upstream SynthesisNetwork.__init__)


def sg3_geometry(img_resolution=1024, img_channels=3, channel_base=32768, channel_max=512,
                 num_layers=14, num_critical=2, first_cutoff=2, first_stopband=2 ** 2.1,
                 last_stopband_rel=2 ** 0.3, margin_size=10, conv_kernel=3, filter_size=6,
                 lrelu_upsampling=2, use_radial_filters=False, num_fp16_res=4):
    """Per-layer table of the alias-free generator: list of dicts, index 0..num_layers (last = torgb)."""
    last_cutoff = img_resolution / 2
    last_stopband = last_cutoff * last_stopband_rel
    exponents = np.minimum(np.arange(num_layers + 1) / (num_layers - num_critical), 1)
    cutoffs = first_cutoff * (last_cutoff / first_cutoff) ** exponents
    stopbands = first_stopband * (last_stopband / first_stopband) ** exponents
    sampling_rates = np.exp2(np.ceil(np.log2(np.minimum(stopbands * 2, img_resolution))))
    half_widths = np.maximum(stopbands, sampling_rates / 2) - cutoffs
    sizes = sampling_rates + margin_size * 2
    sizes[-2:] = img_resolution
    channels = np.rint(np.minimum((channel_base / 2) / cutoffs, channel_max))
    channels[-1] = img_channels

    layers = []
    for idx in range(num_layers + 1):
        prev = max(idx - 1, 0)
        is_torgb = idx == num_layers
        is_crit = idx >= num_layers - num_critical
        in_sr, out_sr = int(sampling_rates[prev]), int(sampling_rates[idx])
        tmp_sr = max(in_sr, out_sr) * (1 if is_torgb else lrelu_upsampling)
        k = 1 if is_torgb else conv_kernel
        up = int(np.rint(tmp_sr / in_sr))
        down = int(np.rint(tmp_sr / out_sr))
        up_taps = filter_size * up if up > 1 and not is_torgb else 1
        down_taps = filter_size * down if down > 1 and not is_torgb else 1
        in_size, out_size = int(sizes[prev]), int(sizes[idx])
        pad_total = (out_size - 1) * down + 1
        pad_total -= (in_size + k - 1) * up
        pad_total += up_taps + down_taps - 2
        pad_lo = (pad_total + up) // 2
        pad_hi = pad_total - pad_lo
        layers.append(dict(
            idx=idx, name=f"L{idx}_{out_size}_{int(channels[idx])}", is_torgb=is_torgb,
            is_critically_sampled=is_crit,
            use_fp16=bool(sampling_rates[idx] * (2 ** num_fp16_res) > img_resolution),
            in_channels=int(channels[prev]), out_channels=int(channels[idx]),
            in_size=in_size, out_size=out_size, in_sampling_rate=in_sr, out_sampling_rate=out_sr,
            tmp_sampling_rate=tmp_sr, in_cutoff=float(cutoffs[prev]), out_cutoff=float(cutoffs[idx]),
            in_half_width=float(half_widths[prev]), out_half_width=float(half_widths[idx]),
            conv_kernel=k, up=up, down=down, up_taps=up_taps, down_taps=down_taps,
            down_radial=bool(use_radial_filters and not is_crit),
            padding=[int(pad_lo), int(pad_hi), int(pad_lo), int(pad_hi)],
        ))
    return dict(layers=layers, input=dict(channels=int(channels[0]), size=int(sizes[0]),
                                          sampling_rate=float(sampling_rates[0]),
                                          bandwidth=float(cutoffs[0])))


# --------------------------------------------------------------------------------------
# modules (parameter names = upstream state-dict keys)


class FullyConnectedLayer(torch.nn.Module):
    def __init__(self, in_features, out_features, activation="linear", bias=True, lr_multiplier=1.0,
                 weight_init=1.0, bias_init=0.0):
        super().__init__()
        self.in_features, self.out_features, self.activation = in_features, out_features, activation
        self.weight = torch.nn.Parameter(torch.randn([out_features, in_features]) * (weight_init / lr_multiplier))
        bias_init = np.broadcast_to(np.asarray(bias_init, dtype=np.float32), [out_features])
        self.bias = torch.nn.Parameter(torch.from_numpy(bias_init / lr_multiplier)) if bias else None
        self.weight_gain = lr_multiplier / np.sqrt(in_features)
        self.bias_gain = lr_multiplier

    def forward(self, x):
        w = self.weight.to(x.dtype) * self.weight_gain
        b = self.bias
        if b is not None:
            b = b.to(x.dtype)
            if self.bias_gain != 1:
                b = b * self.bias_gain
        if self.activation == "linear" and b is not None:
            return torch.addmm(b.unsqueeze(0), x, w.t())
        x = x.matmul(w.t())
        return bias_act_ref(x, b, act=self.activation)


class MappingNetwork(torch.nn.Module):
    def __init__(self, z_dim, c_dim, w_dim, num_ws, num_layers=2, lr_multiplier=0.01, w_avg_beta=0.998):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim, self.num_ws, self.num_layers = z_dim, c_dim, w_dim, num_ws, num_layers
        self.embed = FullyConnectedLayer(c_dim, w_dim) if c_dim > 0 else None
        features = [z_dim + (w_dim if c_dim > 0 else 0)] + [w_dim] * num_layers
        for idx, (i, o) in enumerate(zip(features[:-1], features[1:])):
            setattr(self, f"fc{idx}", FullyConnectedLayer(i, o, activation="lrelu", lr_multiplier=lr_multiplier))
        self.register_buffer("w_avg", torch.zeros([w_dim]))

    def forward(self, z, c=None, truncation_psi=1, truncation_cutoff=None):
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

    def forward(self, w):
        transforms = self.transform.unsqueeze(0)
        freqs = self.freqs.unsqueeze(0)
        phases = self.phases.unsqueeze(0)
        t = self.affine(w)
        t = t / t[:, :2].norm(dim=1, keepdim=True)
        m_r = torch.eye(3).unsqueeze(0).repeat([w.shape[0], 1, 1])
        m_r[:, 0, 0] = t[:, 0]
        m_r[:, 0, 1] = -t[:, 1]
        m_r[:, 1, 0] = t[:, 1]
        m_r[:, 1, 1] = t[:, 0]
        m_t = torch.eye(3).unsqueeze(0).repeat([w.shape[0], 1, 1])
        m_t[:, 0, 2] = -t[:, 2]
        m_t[:, 1, 2] = -t[:, 3]
        transforms = m_r @ m_t @ transforms
        phases = phases + (freqs @ transforms[:, :2, 2:]).squeeze(2)
        freqs = freqs @ transforms[:, :2, :2]
        amplitudes = (1 - (freqs.norm(dim=2) - self.bandwidth) / (self.sampling_rate / 2 - self.bandwidth)).clamp(0, 1)
        theta = torch.eye(2, 3)
        theta[0, 0] = 0.5 * self.size[0] / self.sampling_rate
        theta[1, 1] = 0.5 * self.size[1] / self.sampling_rate
        grids = F.affine_grid(theta.unsqueeze(0), [1, 1, int(self.size[1]), int(self.size[0])], align_corners=False)
        x = (grids.unsqueeze(3) @ freqs.permute(0, 2, 1).unsqueeze(1).unsqueeze(2)).squeeze(3)
        x = x + phases.unsqueeze(1).unsqueeze(2)
        x = torch.sin(x * (np.pi * 2))
        x = x * amplitudes.unsqueeze(1).unsqueeze(2)
        weight = self.weight / np.sqrt(self.channels)
        x = x @ weight.t()
        return x.permute(0, 3, 1, 2)


class SynthesisLayer(torch.nn.Module):
    def __init__(self, w_dim, g, conv_clamp=256):
        super().__init__()
        self.g = g
        self.w_dim, self.is_torgb, self.conv_clamp = w_dim, g["is_torgb"], conv_clamp
        self.in_channels, self.out_channels = g["in_channels"], g["out_channels"]
        self.in_size = np.broadcast_to(np.asarray(g["in_size"]), [2])
        self.out_size = np.broadcast_to(np.asarray(g["out_size"]), [2])
        self.conv_kernel, self.up_factor, self.down_factor = g["conv_kernel"], g["up"], g["down"]
        self.padding = g["padding"]
        self.use_fp16 = g["use_fp16"]
        self.affine = FullyConnectedLayer(w_dim, self.in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn([self.out_channels, self.in_channels, self.conv_kernel, self.conv_kernel]))
        self.bias = torch.nn.Parameter(torch.zeros([self.out_channels]))
        self.register_buffer("magnitude_ema", torch.ones([]))
        self.register_buffer("up_filter", design_lowpass_filter(
            g["up_taps"], g["in_cutoff"], g["in_half_width"] * 2, g["tmp_sampling_rate"]))
        self.register_buffer("down_filter", design_lowpass_filter(
            g["down_taps"], g["out_cutoff"], g["out_half_width"] * 2, g["tmp_sampling_rate"], radial=g["down_radial"]))

    def forward(self, x, w):
        input_gain = self.magnitude_ema.rsqrt()
        styles = self.affine(w)
        if self.is_torgb:
            styles = styles * (1 / np.sqrt(self.in_channels * (self.conv_kernel ** 2)))
        x = modulated_conv2d_ref(x.to(torch.float32), self.weight, styles, padding=self.conv_kernel - 1,
                                 demodulate=not self.is_torgb, input_gain=input_gain)
        gain = 1.0 if self.is_torgb else float(np.sqrt(2))
        slope = 1.0 if self.is_torgb else 0.2
        return filtered_lrelu_ref(x, fu=self.up_filter, fd=self.down_filter, b=self.bias.to(x.dtype),
                                  up=self.up_factor, down=self.down_factor, padding=self.padding,
                                  gain=gain, slope=slope, clamp=self.conv_clamp)


class SynthesisNetwork(torch.nn.Module):
    """Restatement of upstream SynthesisNetwork; ctor args as at maua/GAN/wrappers/stylegan3.py:33."""

    def __init__(self, w_dim, img_resolution, img_channels, output_scale=0.25, conv_clamp=256, **geom_kwargs):
        super().__init__()
        self.w_dim, self.img_resolution, self.img_channels = w_dim, img_resolution, img_channels
        self.output_scale = output_scale
        geo = sg3_geometry(img_resolution=img_resolution, img_channels=img_channels, **geom_kwargs)
        self.geometry = geo
        self.num_layers = len(geo["layers"]) - 1
        self.num_ws = self.num_layers + 2
        self.input = SynthesisInput(w_dim=w_dim, **geo["input"])
        self.layer_names = []
        for g in geo["layers"]:
            setattr(self, g["name"], SynthesisLayer(w_dim, g, conv_clamp=conv_clamp))
            self.layer_names.append(g["name"])

    def forward(self, ws, return_activations=False):
        ws = ws.to(torch.float32).unbind(dim=1)
        x = self.input(ws[0])
        acts = [x]
        for name, w in zip(self.layer_names, ws[1:]):
            x = getattr(self, name)(x, w)
            if return_activations:
                acts.append(x)
        if self.output_scale != 1:
            x = x * self.output_scale
        x = x.to(torch.float32)
        return (x, acts) if return_activations else x


SG3_R_KWARGS = dict(conv_kernel=1, channel_base=65536, channel_max=1024, use_radial_filters=True)


def make_synthesis(config="T", img_resolution=1024, seed=0, **kw):
    """Random-init generator exactly as the reference builds it for model_file=None (stylegan3.py:33)."""
    torch.manual_seed(seed)
    extra = dict(SG3_R_KWARGS) if config.upper() == "R" else {}
    extra.update(kw)
    return SynthesisNetwork(w_dim=512, img_resolution=img_resolution, img_channels=3, **extra).eval().requires_grad_(False)
