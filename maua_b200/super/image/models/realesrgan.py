"""RealESRGAN x4 upscaler: mirror of maua/super/image/models/realesrgan.py:12-49 on the sm_100a RRDBNet (csrc/rrdb.cu).

Kept: ``URLS``, ``load_model(model_name, device)`` returning an object with ``enhance(img)`` (what the reference gets from the
third-party ``RealESRGANer(scale=4, model_path, model, tile=0, half=True)``), and the ``upscale(images, model)`` generator.
basicsr / realesrgan are absent, un-pinned packages: ``RRDBNet`` carries the published architecture's parameter names so
their checkpoints (``params_ema`` / ``params``) load, ``RealESRGANer.enhance`` restates its pre / post-processing (reflect
pre-pad of 10 pixels on the right and bottom, BGR <-> RGB flips, clamp, * 255, round).  PARITY UNPINNED (oracle/rrdb.py).
There is no download (no network) and no CPU path: a missing checkpoint or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from .... import _lib

URLS = {
    "x4plus": "https://github.com/xinntao/Real-ESRGAN/releases/download/v0.1.0/RealESRGAN_x4plus.pth",
    "x4plus-anime": "https://github.com/xinntao/Real-ESRGAN/releases/download/v0.2.2.4/RealESRGAN_x4plus_anime_6B.pth",
    "xsx4-animevideo": "https://github.com/xinntao/Real-ESRGAN/releases/download/v0.2.3.0/RealESRGANv2-animevideo-xsx4.pth",
    "pbaylies-wikiart": "https://archive.org/download/hr-painting-upscaling/wikiart_g.pth",
    "pbaylies-hr-paintings": "https://archive.org/download/hr-painting-upscaling/hr-paintings_g.pth",
}


class _Conv(torch.nn.Module):
    """Parameter holder with torch.nn.Conv2d's names and default initialisation (kaiming_uniform(a=sqrt(5)))."""

    def __init__(self, cin, cout, scale=1.0):
        super().__init__()
        ref = torch.nn.Conv2d(cin, cout, 3, 1, 1)
        self.weight = torch.nn.Parameter(ref.weight.detach() * scale)
        self.bias = torch.nn.Parameter(ref.bias.detach() * (0.0 if scale != 1.0 else 1.0))


class ResidualDenseBlock(torch.nn.Module):
    def __init__(self, num_feat=64, num_grow_ch=32):
        super().__init__()
        for k in range(5):   # basicsr: default_init_weights([conv1..conv5], 0.1): kaiming_normal * 0.1, zero bias
            conv = _Conv(num_feat + k * num_grow_ch, num_grow_ch if k < 4 else num_feat)
            torch.nn.init.kaiming_normal_(conv.weight)
            conv.weight.data.mul_(0.1)
            conv.bias.data.zero_()
            setattr(self, f"conv{k + 1}", conv)


class RRDB(torch.nn.Module):
    def __init__(self, num_feat, num_grow_ch=32):
        super().__init__()
        self.rdb1, self.rdb2, self.rdb3 = (ResidualDenseBlock(num_feat, num_grow_ch) for _ in range(3))


class RRDBNet(torch.nn.Module):
    """``RRDBNet(num_in_ch=3, num_out_ch=3, num_feat=64, num_block=23, num_grow_ch=32, scale=4)`` as built at
    realesrgan.py:33-39; forward(x float [B, 3, H, W] in [0, 1] or uint8 [B, H, W, 3], CUDA) -> [B, 3, 4H, 4W]."""

    def __init__(self, num_in_ch=3, num_out_ch=3, scale=4, num_feat=64, num_block=23, num_grow_ch=32):
        super().__init__()
        if scale != 4 or num_feat != 64 or num_grow_ch != 32:
            raise NotImplementedError("RRDBNet: scale 4, num_feat 64, num_grow_ch 32 (the RealESRGAN x4 generators the reference loads)")
        self.num_in_ch, self.num_out_ch, self.scale, self.num_block = num_in_ch, num_out_ch, scale, num_block
        self.conv_first = _Conv(num_in_ch, num_feat)
        self.body = torch.nn.ModuleList([RRDB(num_feat, num_grow_ch) for _ in range(num_block)])
        self.conv_body = _Conv(num_feat, num_feat)
        self.conv_up1, self.conv_up2 = _Conv(num_feat, num_feat), _Conv(num_feat, num_feat)
        self.conv_hr, self.conv_last = _Conv(num_feat, num_feat), _Conv(num_feat, num_out_ch)
        self._net, self._uploaded, self._workspace = None, {}, None

    def half(self):          # RealESRGANer(half=True) calls model.half(): the device network is fp16 already
        return self

    def _handle(self):
        if self._net is None:
            h = C.c_void_p()
            _lib.check(_lib.load().mb_rrdb_create(self.num_in_ch, self.num_out_ch, 64, self.num_block, 32, self.scale, C.byref(h)))
            self._net = h
        return self._net

    def __del__(self):
        try:
            if self._net is not None:
                _lib.load().mb_rrdb_destroy(self._net)
                self._net = None
        except Exception:
            pass

    def _sync_params(self, device):
        lib, net, changed, keep = _lib.load(), self._handle(), False, []
        for name, t in self.named_parameters():
            try:
                version = t._version
            except RuntimeError:     # inference-mode tensors carry no version counter
                version = None
            key = (t.data_ptr(), version, str(t.device))
            if self._uploaded.get(name) == key:
                continue
            d = t.detach().to(device=device, dtype=torch.float32).contiguous()
            keep.append(d)
            shape = (C.c_int64 * d.ndim)(*d.shape)
            _lib.check(lib.mb_rrdb_set_param(net, name.encode(), _lib.ptr(d), shape, d.ndim, _lib.stream_ptr()))
            self._uploaded[name] = key
            changed = True
        if changed:
            _lib.check(lib.mb_rrdb_finalize(net, _lib.stream_ptr()))
        del keep

    def forward(self, x, out_fmt="f32", out=None):
        """out_fmt: "f32" raw network output, "f32_01" clamped to [0, 1], "u8" rgb24 [B, 4H, 4W, 3]."""
        if not x.is_cuda:
            raise RuntimeError("maua_b200 RRDBNet.forward needs CUDA input: there is no CPU path")
        lib, device = _lib.load(), x.device
        with torch.cuda.device(device):
            self._sync_params(device)
            if x.dtype == torch.uint8:
                x = x.contiguous()
                B, H, W, Cc = x.shape
                in_fmt = _lib.MB_OUT_U8_NHWC
            else:
                x = x.detach().to(torch.float32).contiguous()
                B, Cc, H, W = x.shape
                in_fmt = _lib.MB_OUT_F32_NCHW
            if Cc != self.num_in_ch:
                raise ValueError(f"RRDBNet: expected {self.num_in_ch} channels, got {Cc}")
            fmt = {"f32": _lib.MB_OUT_F32_NCHW, "f32_01": _lib.MB_OUT_F32_NCHW_01, "u8": _lib.MB_OUT_U8_NHWC}[out_fmt]
            shape = (B, 4 * H, 4 * W, self.num_out_ch) if out_fmt == "u8" else (B, self.num_out_ch, 4 * H, 4 * W)
            dtype = torch.uint8 if out_fmt == "u8" else torch.float32
            if out is None:
                out = torch.empty(shape, device=device, dtype=dtype)
            elif tuple(out.shape) != shape or out.dtype != dtype or not out.is_contiguous():
                raise ValueError(f"out must be a contiguous {dtype} tensor of shape {shape}")
            nbytes = lib.mb_rrdb_workspace_bytes(self._handle(), B, H, W)
            if self._workspace is None or self._workspace.numel() < nbytes + 1024 or self._workspace.device != device:
                self._workspace = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
            off = (-self._workspace.data_ptr()) % 1024
            _lib.check(lib.mb_rrdb_forward(self._handle(), _lib.ptr(x), in_fmt, B, H, W, _lib.ptr(out), fmt,
                                           C.c_void_p(self._workspace.data_ptr() + off), nbytes, _lib.stream_ptr()))
        return out

    def last_launch_count(self):
        return _lib.load().mb_rrdb_last_launch_count(self._handle())


class RealESRGANer:
    """The part of realesrgan.RealESRGANer the reference uses (``enhance`` with tile=0): numpy HWC image in (uint8 / 0..255
    floats, treated as BGR like cv2 images), numpy HWC uint8 out at 4x."""

    def __init__(self, scale, model_path, model=None, tile=0, tile_pad=10, pre_pad=10, half=False, device=None):
        if tile != 0:
            raise NotImplementedError("RealESRGANer: tile=0 only (what the reference passes, realesrgan.py:41)")
        self.scale, self.pre_pad, self.half = scale, pre_pad, half
        self.device = torch.device("cuda") if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("RealESRGANer: CUDA device required (no CPU path)")
        if model_path is not None:
            loadnet = torch.load(model_path, map_location="cpu", weights_only=True)
            key = "params_ema" if "params_ema" in loadnet else ("params" if "params" in loadnet else None)
            model.load_state_dict(loadnet[key] if key else loadnet, strict=True)
        self.model = model.eval().to(self.device)

    @torch.inference_mode()
    def enhance(self, img, outscale=None, alpha_upsampler="realesrgan"):
        img = np.asarray(img).astype(np.float32)
        max_range = 65535 if np.max(img) > 256 else 255          # 16-bit images
        img = img / max_range
        if img.ndim == 2 or img.shape[2] != 3:
            raise NotImplementedError("RealESRGANer.enhance: 3-channel images (gray / alpha inputs take cv2 paths of the third-party class)")
        x = torch.from_numpy(np.ascontiguousarray(img[:, :, ::-1].transpose(2, 0, 1)))[None].to(self.device)   # BGR -> RGB, HWC -> CHW
        if self.pre_pad != 0:
            x = torch.nn.functional.pad(x, (0, self.pre_pad, 0, self.pre_pad), "reflect")
        y = self.model(x)
        if self.pre_pad != 0:
            _, _, h, w = y.shape
            y = y[:, :, 0: h - self.pre_pad * self.scale, 0: w - self.pre_pad * self.scale]
        out = y[0].float().clamp_(0, 1).cpu().numpy()
        out = np.transpose(out[[2, 1, 0], :, :], (1, 2, 0))                                                      # RGB -> BGR, CHW -> HWC
        out = (out * 255.0).round().astype(np.uint8) if max_range == 255 else (out * 65535.0).round().astype(np.uint16)
        return out, "RGB"


def load_image(im):
    """maua/ops/io.py:17-18,37-38: tensors pass through, PIL images / paths become float [1, 3, H, W] in [0, 1]."""
    if isinstance(im, torch.Tensor):
        return im
    from PIL import Image

    pil = im if isinstance(im, Image.Image) else Image.open(im)
    arr = np.asarray(pil.convert("RGB"), dtype=np.float32) / 255.0
    return torch.from_numpy(arr).permute(2, 0, 1).unsqueeze(0)


def load_model(model_name="pbaylies-hr-paintings", device=torch.device("cuda"), checkpoint=None):
    """realesrgan.py:22-41.  ``checkpoint``: explicit path (default modelzoo/RealESRGAN_<name>.pth, which must exist: no download)."""
    checkpoint = checkpoint or f"modelzoo/RealESRGAN_{model_name}.pth"
    if not os.path.exists(checkpoint):
        raise FileNotFoundError(f"{checkpoint} not found (no network in this build: fetch {URLS.get(model_name, '?')} yourself)")
    if model_name == "xsx4-animevideo":
        raise NotImplementedError("SRVGGNetCompact (xsx4-animevideo) is not built")
    model = RRDBNet(num_in_ch=3, num_out_ch=3, num_feat=64, num_block=6 if model_name == "x4plus-anime" else 23, num_grow_ch=32, scale=4)
    return RealESRGANer(scale=4, model_path=checkpoint, model=model.eval(), tile=0, half=True, device=device)


@torch.inference_mode()
def upscale(images, model):
    """realesrgan.py:44-49: every image (tensor [1,3,H,W] / [3,H,W] in [0, 1], PIL image or path) -> float tensor [1, 3, 4H, 4W]."""
    for img in images:
        inp = load_image(img).detach().squeeze().permute(1, 2, 0).mul(255).cpu().numpy()
        large = model.enhance(inp)[0]
        yield torch.from_numpy(large).permute(2, 0, 1).unsqueeze(0).float().div(255)
