"""Host-tensor adapter for the ``maua.*`` import surface.

The reference's envelope / latent functions (audioreactive/signal.py, latent.py) take and return HOST tensors, and patch
files written against it do plain torch arithmetic on the results.  This build's functions compute on the GPU and refuse host
tensors (there is no CPU implementation to fall back to).  ``host_io`` bridges the two for the alias namespace only: when no
argument lives on a CUDA device, tensor arguments are moved to the GPU, the device function runs, and tensor results are moved
back -- the arithmetic is still the library's kernels; without a GPU the call raises.
"""
import functools

import numpy as np
import torch


def _map(obj, fn):
    if torch.is_tensor(obj):
        return fn(obj)
    if isinstance(obj, (list, tuple)):
        return type(obj)(_map(o, fn) for o in obj)
    if isinstance(obj, dict):
        return {k: _map(v, fn) for k, v in obj.items()}
    return obj


def _any_cuda(obj):
    found = []
    _map(obj, lambda t: found.append(t.is_cuda) or t)
    return any(found)


def host_io(fn):
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        args = tuple(torch.from_numpy(a) if isinstance(a, np.ndarray) else a for a in args)
        if _any_cuda((args, kwargs)):
            return fn(*args, **kwargs)
        if not torch.cuda.is_available():
            raise RuntimeError(f"maua.audiovisual.audioreactive.{fn.__name__}: needs a CUDA device (the arithmetic runs in "
                               "libmaua_b200's kernels; there is no CPU implementation)")
        dev = torch.device("cuda")
        out = fn(*_map(args, lambda t: t.to(dev)), **_map(kwargs, lambda t: t.to(dev)))
        return _map(out, lambda t: t.cpu())

    return wrapper
