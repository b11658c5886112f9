"""CPU restatement of the reference's quantile (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Follows maua/audiovisual/audioreactive/selfsupervised/features/efficient_quantile/__init__.py:6-7 and
efficient_quantile.cpp:86-208: ``quantile(t, q)`` flattens, drops NaNs, and calls the C++ routine with the quantile as a
**float32** one-element tensor and interpolation method 3 ("mid point").  What that computes:

    qs   = double(float32(q))                      the float32 rounding of q is part of the result (:6, .cpp:106)
    qm   = qs * (size - 1);  ql = trunc(qm);  qu = ceil(qm)         (.cpp:155-160, 193-195)
    out  = lerp(v[ql], v[qu], 0.5 if qu > ql else 0) in double, cast back to the input dtype   (.cpp:72-83)

with v the ascending order statistics (std::nth_element, .cpp:14).  torch.lerp evaluates weight >= 0.5 as
``end - (end - start) * (1 - weight)``.  Pinned against the reference's own compiled routine (oracle/_ref, built by
oracle/build_ref.py) in tests/test_oracle_quantile.py and through tests/golden/quantile.pt.
"""
import numpy as np
import torch


def quantile(t, q):
    v = t.detach().cpu().flatten()
    dtype = v.dtype
    v = v.double().numpy()
    v = v[~np.isnan(v)]
    size = v.shape[0]
    if size == 0:
        return torch.tensor(float("nan"), dtype=dtype)
    qs = float(np.float32(q))
    qm = qs * (size - 1)
    ql, qu = int(qm), int(np.ceil(qm))
    part = np.partition(v, sorted({ql, qu}))
    lo, hi = part[ql], part[qu]
    out = hi - (hi - lo) * 0.5 if qu > ql else lo
    return torch.tensor(out, dtype=torch.float64).to(dtype)


def normalize(a):
    """processing.py:51-55."""
    a = a - a.min()
    return a / (a.max() + 1e-8)


def standardize(a):
    """processing.py:58-61: clamp to the inter-quartile range, then min-max normalise."""
    return normalize(torch.clamp(a, quantile(a, 0.25), quantile(a, 0.75) + 1e-10))


def spectral_flux(spec):
    """processing.py:88-89."""
    return torch.diff(spec, dim=0, append=torch.zeros((1, spec.shape[1])))


def onset_envelope(flux):
    """processing.py:93-98: half-wave rectified flux summed over bins, clamped to its 2.5 % .. 97.5 % range, unit range."""
    u = torch.sum(0.5 * (flux + torch.abs(flux)), dim=1)
    u = torch.clamp(u, quantile(u, 0.025), quantile(u, 0.975))
    u = u - u.min()
    return u / u.max()
