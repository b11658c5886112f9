"""Shared host-side plumbing of the networks backed by an ``mb_net`` handle of libmaua_b200 (StyleGAN3, StyleGAN2):
handle creation, incremental parameter upload (state-dict key -> mb_net_set_param), caller-owned workspace."""
from __future__ import annotations

import ctypes as C

import torch

from ... import _lib


class NativeNet(torch.nn.Module):
    def _init_native(self):
        self._net = None          # mb_net* handle
        self._uploaded = {}       # state-dict key -> (data_ptr, version, device)
        self._workspace = {}      # batch -> uint8 tensor
        self._options = {}

    def _create(self, lib, handle_ref):
        raise NotImplementedError

    def _is_volatile(self, name, t):
        """True for tensors that must be re-uploaded on every forward (their identity does not reveal a change)."""
        return name in self._POKED

    def _prepare_param(self, name, d):
        """Last host-side touch of a float32 device copy before it is handed to mb_net_set_param."""
        return d

    # ---- library handle management ----------------------------------------------------
    def _handle(self):
        if self._net is None:
            lib = _lib.load()
            h = C.c_void_p()
            self._create(lib, C.byref(h))
            self._net = h
            for k, v in self._options.items():
                _lib.check(lib.mb_net_set_option(self._net, k.encode(), int(v)))
        return self._net

    def set_option(self, key, value):
        """Test / tuning knobs of the library (see include/maua_b200.h mb_net_set_option)."""
        self._options[key] = int(value)
        if self._net is not None:
            _lib.check(_lib.load().mb_net_set_option(self._net, key.encode(), int(value)))

    def __del__(self):
        try:
            if self._net is not None:
                _lib.load().mb_net_destroy(self._net)
                self._net = None
        except Exception:
            pass

    # tensors the reference wrapper edits in place between forwards (wrappers/stylegan3.py:54-59)
    _POKED = ("input.affine.bias", "input.affine.weight", "input.transform")

    def _param_key(self, name, t):
        """Identity of a parameter's current value: storage pointer + autograd version counter (inference-mode tensors
        carry none).  Neither sees an edit through ``.data`` (the reference's stabilisation trick, wrappers/stylegan3.py:54-55):
        the three tiny tensors the wrapper pokes are therefore volatile -- re-uploaded on every forward (three asynchronous
        device copies, no synchronisation, no finalize)."""
        try:
            version = t._version
        except RuntimeError:
            version = None
        return (t.data_ptr(), version, str(t.device), tuple(t.shape))

    def _sync_params(self, device):
        """Upload every parameter / buffer whose storage or version changed since the last forward."""
        lib = _lib.load()
        net = self._handle()
        changed = False
        keep = []
        for name, t in list(self.named_parameters()) + list(self.named_buffers()):
            if t is None:
                continue
            key = self._param_key(name, t)
            volatile = self._is_volatile(name, t)
            if not volatile and self._uploaded.get(name) == key:
                continue
            d = self._prepare_param(name, t.detach().to(device=device, dtype=torch.float32)).contiguous()
            keep.append(d)
            shape = (C.c_int64 * max(d.ndim, 1))(*d.shape)
            _lib.check(lib.mb_net_set_param(net, name.encode(), _lib.ptr(d), shape, d.ndim, _lib.stream_ptr()))
            self._uploaded[name] = None if volatile else key
            changed = changed or not volatile
        if changed:
            _lib.check(lib.mb_net_finalize(net, _lib.stream_ptr()))  # synchronises the stream
        del keep

    def _get_workspace(self, batch, device):
        key = (batch, str(device))
        ws = self._workspace.get(key)
        if ws is None:
            nbytes = _lib.load().mb_net_workspace_bytes(self._handle(), batch)
            ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
            self._workspace = {key: ws}  # keep only the latest batch size resident
        off = (-ws.data_ptr()) % 1024
        return ws, off, ws.numel() - 1024

    def read_activation(self, batch):
        """float32 [B,C,H,W] of the last activation the previous forward produced (x * next style)."""
        lib = _lib.load()
        c, h, w = C.c_int32(), C.c_int32(), C.c_int32()
        _lib.check(lib.mb_net_activation_shape(self._handle(), C.byref(c), C.byref(h), C.byref(w)))
        out = torch.empty(batch, c.value, h.value, w.value, device="cuda", dtype=torch.float32)
        _lib.check(lib.mb_net_read_activation(self._handle(), 0, batch, _lib.ptr(out), _lib.stream_ptr()))
        return out

    def profile_read(self):
        """[(kind, layer, ms)] of the last forward (needs set_option('profile', 1) and a stream sync)."""
        cap = 8192
        ms, kind, layer = (C.c_float * cap)(), (C.c_int32 * cap)(), (C.c_int32 * cap)()
        n = _lib.load().mb_net_profile_read(self._handle(), ms, kind, layer, cap)
        return [(kind[i], layer[i], ms[i]) for i in range(n)]

    def last_launch_count(self):
        return _lib.load().mb_net_last_launch_count(self._handle())


