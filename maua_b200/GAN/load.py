"""Checkpoint loaders: mirror of maua/GAN/load.py:18-207 feeding the sm_100a networks.

Pure host code.  ``load_network(path, for_inference)`` tries the same four loaders in the same order and raises the same
aggregated error when none succeeds (:192-207).  Differences, all forced by what this build is:

* NVIDIA ``.pkl`` files (:129-164) are pickles of ``torch_utils.persistence`` objects whose reconstruction executes the
  training source embedded in the pickle and needs the un-vendored ``dnnlib`` / ``legacy`` modules
  (maua-maua-maua/nvGAN @ 7809c05, absent).  ``load_nvidia`` reads them WITHOUT that code: a restricted unpickler
  rebuilds each persistent object as a plain attribute bag from the ``state`` the format carries
  (``_reconstruct_persistent_obj(meta)``, meta = {type, version, module_src, class_name, state}) and the generator's
  state dict is collected from the ``_parameters`` / ``_buffers`` / ``_modules`` trees.
* The network hyper-parameters are read off the state dict (resolution, channels, kernel size, radial filters, mapping
  depth) instead of being assumed (the reference's ``load_nvidia_pt`` assumes 1024^2 / 8 mapping layers, :167-169).
* StyleGAN2 checkpoints in the NVIDIA training layout (``synthesis.b64.conv0...``, ``mapping.fc3``, ``noise_strength``)
  are mapped onto the in-tree inference layout this build implements (``synthesis.bs.4.conv0...``, ``mapping.fcs.3``,
  the same mapping the reference's converter uses at :23,65,71).  What the training network does differently stays with the
  loaded generator: every layer's ``noise_strength`` (it scales the constant AND any per-frame noise map the wrapper swaps
  in) and the standard ``x @ W.T`` mapping layers (the inference network's lrelu layers multiply by ``W`` itself).
"""
from __future__ import annotations

import io
import math
import pickle
import re
import traceback
import types
from functools import partial

import torch

from .networks import stylegan2, stylegan3


# ---------------------------------------------------------------------------------------------------------------------
# rosinality -> in-tree layout (maua/GAN/load.py:18-127)
# ---------------------------------------------------------------------------------------------------------------------
# The converter is a key map, stated as a table: (pattern of the rosinality key, target key(s), value transform).  In a
# target, {b} is the block index and {c} the conv index the pattern's layer number n stands for (``convs.n`` is conv n % 2 of
# block n // 2 + 1, ``to_rgbs.n`` belongs to block n + 1, ``noises.noise_n`` to conv (n - 1) % 2 of block (n - 1) // 2 + 1;
# reference :61-112), {p} is the captured ``weight`` / ``bias``.
def _drop_lead(v):
    return v.squeeze(0)


def _drop_lead2(v):
    return v.squeeze(0).squeeze(0)


def _rgb_bias(v):
    return v.squeeze(-1).squeeze(-1).squeeze(0)


def _keep(v):
    return v


_CONV_OF = lambda n: dict(b=n // 2 + 1, c=n % 2)          # noqa: E731
_RGB_OF = lambda n: dict(b=n + 1, c=0)                     # noqa: E731
_NOISE_OF = lambda n: dict(b=(n - 1) // 2 + 1, c=(n - 1) % 2) if n else dict(b=0, c=1)   # noqa: E731
_FIRST = lambda n: dict(b=0, c=1)                          # noqa: E731

_ROSINALITY_TABLE = [
    # first block (4x4): constant input, one conv, one ToRGB
    (r"input\.input", _FIRST, ["synthesis.bs.0.const"], _drop_lead),
    (r"conv1\.conv\.weight", _FIRST, ["synthesis.bs.0.conv1.weight"], _drop_lead),
    (r"conv1\.activate\.bias", _FIRST, ["synthesis.bs.0.conv1.bias"], _keep),
    (r"conv1\.conv\.modulation\.(?P<p>weight|bias)", _FIRST, ["synthesis.bs.0.conv1.affine.{p}"], _keep),
    (r"conv1\.noise\.weight", _FIRST, ["synthesis.bs.0.conv1.noise_strength"], _drop_lead),
    (r"to_rgb1\.conv\.weight", _FIRST, ["synthesis.bs.0.torgb.weight"], _drop_lead),
    (r"to_rgb1\.bias", _FIRST, ["synthesis.bs.0.torgb.bias"], _rgb_bias),
    (r"to_rgb1\.conv\.modulation\.(?P<p>weight|bias)", _FIRST, ["synthesis.bs.0.torgb.affine.{p}"], _keep),
    # mapping network, noise maps
    (r"style\.(?P<n>\d+)\.(?P<p>weight|bias)", lambda n: dict(b=n - 1, c=0), ["mapping.fcs.{b}.{p}"], _keep),
    (r"noises\.noise_(?P<n>\d+)", _NOISE_OF, ["synthesis.bs.{b}.conv{c}.noise_const"], _drop_lead2),
    # the two convs of every later block
    (r"convs\.(?P<n>\d+)\.conv\.weight", _CONV_OF, ["synthesis.bs.{b}.conv{c}.weight"], _drop_lead),
    (r"convs\.(?P<n>\d+)\.activate\.bias", _CONV_OF, ["synthesis.bs.{b}.conv{c}.bias"], _keep),
    (r"convs\.(?P<n>\d+)\.conv\.modulation\.(?P<p>weight|bias)", _CONV_OF, ["synthesis.bs.{b}.conv{c}.affine.{p}"], _keep),
    (r"convs\.(?P<n>\d+)\.noise\.weight", _CONV_OF, ["synthesis.bs.{b}.conv{c}.noise_strength"], _drop_lead),
    (r"convs\.(?P<n>\d+)\.conv\.blur\.kernel", _CONV_OF,
     ["synthesis.bs.{b}.conv0.resample_filter", "synthesis.bs.{b}.conv1.resample_filter"], "blur"),
    # ToRGB of every later block
    (r"to_rgbs\.(?P<n>\d+)\.conv\.weight", _RGB_OF, ["synthesis.bs.{b}.torgb.weight"], _drop_lead),
    (r"to_rgbs\.(?P<n>\d+)\.bias", _RGB_OF, ["synthesis.bs.{b}.torgb.bias"], _rgb_bias),
    (r"to_rgbs\.(?P<n>\d+)\.conv\.modulation\.(?P<p>weight|bias)", _RGB_OF, ["synthesis.bs.{b}.torgb.affine.{p}"], _keep),
    (r"to_rgbs\.(?P<n>\d+)\.upsample\.kernel", _RGB_OF, ["synthesis.bs.{b}.resample_filter"], "blur"),
]
_ROSINALITY_RULES = [(re.compile(pat), where, targets, fn) for pat, where, targets, fn in _ROSINALITY_TABLE]
_STRICT_PREFIXES = ("convs.", "to_rgbs.")   # an unknown key under these aborts the conversion (reference :88,109)


def rosinality_to_inference_state(state_dict, blur_scale=4.0):
    """rosinality ``g_ema`` state dict -> (state in the in-tree layout incl. ``*.noise_strength`` entries, max_res,
    num_map); the key map of load_rosinality2ada (:18-113)."""
    source = state_dict["g_ema"]
    if tuple(source["input.input"].shape) == (1,):
        raise NotImplementedError("rosinality checkpoints with a learned-affine input (no const) are not supported")
    out, max_res, num_map = {}, 4, 1
    for key, val in source.items():
        for pat, where, targets, fn in _ROSINALITY_RULES:
            m = pat.fullmatch(key)
            if m is None:
                continue
            groups = m.groupdict()
            n = int(groups.get("n") or 0)
            fields = dict(where(n), p=groups.get("p"))
            value = val / blur_scale if fn == "blur" else fn(val)
            for t in targets:
                out[t.format(**fields)] = value
            if key.startswith("style."):
                num_map = max(num_map, n)
            if key.startswith("convs."):
                max_res = max(max_res, 2 ** (3 + n // 2))
            break
        else:
            if key.startswith(_STRICT_PREFIXES):
                raise Exception(f"Key {key} not recognized!")
    # the 4x4 block borrows the first blur kernel (:46-47)
    first_blur = source["convs.0.conv.blur.kernel"] / blur_scale
    out["synthesis.bs.0.resample_filter"] = first_blur
    out["synthesis.bs.0.conv1.resample_filter"] = first_blur
    out["mapping.w_avg"] = state_dict["latent_avg"] if "latent_avg" in state_dict else torch.zeros(512)
    return out, max_res, num_map


def _take_noise_strengths(state):
    """Split the ``*.noise_strength`` entries off a state dict (they are attributes of the layers here, not buffers)."""
    return {k[: -len(".noise_strength")]: float(state.pop(k)) for k in [k for k in state if k.endswith(".noise_strength")]}


def _finish_sg2(G, strengths, standard_fc):
    """Training-layout semantics on the in-tree network: per-layer noise strength (the training network scales the noise map,
    constant or supplied per frame, by it; the inference network of the reference adds it unscaled, inference/ops.py:184) and
    the standard ``x @ W.T`` mapping layers (the inference network's lrelu branch multiplies by ``W`` itself,
    inference/stylegan2.py:57)."""
    for prefix, value in strengths.items():
        G.get_submodule(prefix).noise_strength = value
    if standard_fc:
        for fc in G.mapping.fcs:
            fc.standard_matmul = True
    return G


def load_rosinality2ada(path, blur_scale=4.0, for_inference=False):
    """maua/GAN/load.py:18-127.  ``for_inference=True`` reproduces the reference's inference network (noise added unscaled,
    mapping layers with its transposed-weight quirk); ``for_inference=False`` gives the semantics of the training network the
    reference instantiates in that case (noise strength applied, standard mapping layers) on the same in-tree network."""
    state_dict = torch.load(path, map_location="cpu", weights_only=False)
    state_nv, max_res, num_map = rosinality_to_inference_state(state_dict, blur_scale)
    strengths = _take_noise_strengths(state_nv)
    z_dim = w_dim = state_nv["mapping.fcs.0.weight"].shape[1] if "mapping.fcs.0.weight" in state_nv else 512
    G = stylegan2.Generator(z_dim, 0, w_dim, max_res, 3, mapping_kwargs=dict(num_layers=num_map), **_sg2_channels(state_nv))
    G.load_state_dict(state_nv)
    return G if for_inference else _finish_sg2(G, strengths, standard_fc=True)


# ---------------------------------------------------------------------------------------------------------------------
# NVIDIA state dicts (maua/GAN/load.py:129-189)
# ---------------------------------------------------------------------------------------------------------------------
def _sg2_channels(state):
    """channel_base / channel_max that reproduce the checkpoint's channel table (min(channel_base // res, channel_max))."""
    ch = {}
    for k, v in state.items():
        m = re.fullmatch(r"synthesis\.bs\.(\d+)\.conv1\.weight", k)
        if m:
            ch[4 * 2 ** int(m.group(1))] = v.shape[0]
    cmax = max(ch.values())
    base = max(r * c for r, c in ch.items())
    if any(min(base // r, cmax) != c for r, c in ch.items()):
        raise ValueError(f"channel table {ch} is not of the form min(channel_base // res, channel_max)")
    return dict(channel_base=base, channel_max=cmax)


def nvidia_sg2_to_inference_state(state):
    """NVIDIA training layout -> in-tree layout: b{res} -> bs.{log2(res) - 2}, mapping.fc{i} -> mapping.fcs.{i}; the
    ``noise_strength`` entries keep their (renamed) keys and are split off by the caller."""
    out = {}
    for k, v in state.items():
        m = re.fullmatch(r"synthesis\.b(\d+)\.(.*)", k)
        if m:
            k = f"synthesis.bs.{int(math.log2(int(m.group(1)))) - 2}.{m.group(2)}"
        m = re.fullmatch(r"mapping\.fc(\d+)\.(weight|bias)", k)
        if m:
            k = f"mapping.fcs.{m.group(1)}.{m.group(2)}"
        out[k] = v
    return out


def generator_from_state(state):
    """Build the generator a NVIDIA-style state dict (``synthesis.*`` / ``mapping.*``) belongs to and load it."""
    state = {k: (v.detach() if torch.is_tensor(v) else torch.as_tensor(v)) for k, v in state.items()}
    if "synthesis.input.freqs" in state:  # StyleGAN3
        names = sorted({k.split(".")[1] for k in state if re.match(r"synthesis\.L\d+_", k)}, key=lambda n: int(n.split("_")[0][1:]))
        last = names[-1]
        res, img_channels = int(last.split("_")[1]), int(last.split("_")[2])
        chans = [int(n.split("_")[2]) for n in names[:-1]]
        conv_kernel = state[f"synthesis.{names[0]}.weight"].shape[-1]
        radial = any(state[k].ndim == 2 for k in state if k.endswith(".down_filter"))
        w_dim = state[f"synthesis.{names[0]}.affine.weight"].shape[1]
        cmax = max(chans)
        # upstream: channels = rint(min(channel_base / 2 / cutoff, channel_max)); cutoff doubles towards the output
        kw = dict(channel_max=cmax, conv_kernel=conv_kernel, use_radial_filters=radial, num_layers=len(names) - 1)
        found = None
        for base in (32768, 65536, 16384, 8192, 4096, 2048, 1024):
            g = stylegan3.sg3_geometry(stylegan3.sg3_cfg(w_dim=w_dim, img_resolution=res, img_channels=img_channels, channel_base=base, **kw))
            if [l["name"] for l in g["layers"]] == names:
                found = base
                break
        if found is None:
            raise ValueError(f"cannot infer channel_base for layers {names}")
        n_map = len({k.split(".")[1] for k in state if re.match(r"mapping\.fc\d+\.weight", k)})
        z_dim = state["mapping.fc0.weight"].shape[1]
        c_dim = state["mapping.embed.weight"].shape[1] if "mapping.embed.weight" in state else 0
        if c_dim:
            z_dim -= w_dim
        G = stylegan3.Generator(z_dim, c_dim, w_dim, res, img_channels, mapping_kwargs=dict(num_layers=n_map),
                                channel_base=found, **kw)
        G.load_state_dict(state)
        return G
    training_layout = any(re.match(r"synthesis\.b\d+\.", k) for k in state)
    if training_layout:
        state = nvidia_sg2_to_inference_state(state)
    strengths = _take_noise_strengths(state)
    if "synthesis.bs.0.const" not in state:
        raise ValueError("not a StyleGAN2 / StyleGAN3 generator state dict")
    n_blocks = len({k.split(".")[2] for k in state if k.startswith("synthesis.bs.")})
    res = 4 * 2 ** (n_blocks - 1)
    img_channels = state["synthesis.bs.0.torgb.weight"].shape[0]
    w_dim = state["synthesis.bs.0.conv1.affine.weight"].shape[1]
    n_map = len({k.split(".")[2] for k in state if re.match(r"mapping\.fcs\.\d+\.weight", k)})
    z_dim = state["mapping.fcs.0.weight"].shape[1]
    c_dim = state["mapping.embed.weight"].shape[1] if "mapping.embed.weight" in state else 0
    if c_dim:
        z_dim -= state["mapping.embed.weight"].shape[0]
    G = stylegan2.Generator(z_dim, c_dim, w_dim, res, img_channels, mapping_kwargs=dict(num_layers=n_map), **_sg2_channels(state))
    G.load_state_dict(state)
    return _finish_sg2(G, strengths, standard_fc=training_layout)


def load_nvidia_pt(path, for_inference=False, **unused):
    """maua/GAN/load.py:166-189: ``torch.load(path)["G_ema"]`` is a generator state dict."""
    state = torch.load(path, map_location="cpu", weights_only=False)["G_ema"]
    return generator_from_state(state)


class _Bag:
    """Attribute bag standing in for a torch_utils.persistence object (no training code is executed)."""


def _reconstruct_persistent_obj(meta):
    meta = types.SimpleNamespace(**meta) if isinstance(meta, dict) else meta
    obj = _Bag()
    state = getattr(meta, "state", None)
    if isinstance(state, dict):
        obj.__dict__.update(state)
    obj._orig_class_name = getattr(meta, "class_name", None)
    return obj


class _PersistenceUnpickler(pickle.Unpickler):
    """legacy._LegacyUnpickler (nvGAN legacy.py) without dnnlib: persistent objects become _Bag trees.  Only the exact
    globals a tensor / array / OrderedDict needs to rebuild itself resolve; anything else (``builtins.eval``,
    ``torch.load``, ``numpy.load``, ``os.system`` ...) is refused, so a crafted pickle cannot run code through this loader."""

    _ALLOWED = {
        ("collections", "OrderedDict"),
        ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_parameter"),
        ("torch._utils", "_rebuild_parameter_with_state"), ("torch._tensor", "_rebuild_from_type_v2"),
        ("torch.storage", "_load_from_bytes"), ("torch", "Size"), ("torch", "device"), ("torch", "Tensor"),
        ("torch.nn.parameter", "Parameter"),
        ("numpy.core.multiarray", "_reconstruct"), ("numpy.core.multiarray", "scalar"),
        ("numpy._core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "scalar"),
        ("numpy", "ndarray"), ("numpy", "dtype"), ("_codecs", "encode"),
        ("builtins", "dict"), ("builtins", "list"), ("builtins", "tuple"), ("builtins", "set"), ("builtins", "frozenset"),
        ("builtins", "int"), ("builtins", "float"), ("builtins", "bool"), ("builtins", "str"), ("builtins", "bytes"),
        ("builtins", "complex"), ("builtins", "slice"), ("builtins", "range"), ("builtins", "bytearray"),
    }
    _ALLOWED_TORCH_NAMES = re.compile(r"(\w+Storage|float\d+|bfloat16|half|double|u?int\d+|long|short|bool|complex\d+)")

    def find_class(self, module, name):
        if module == "torch_utils.persistence" and name == "_reconstruct_persistent_obj":
            return _reconstruct_persistent_obj
        if module == "dnnlib.util" and name == "EasyDict":
            return dict
        if (module, name) in self._ALLOWED or (module == "torch" and self._ALLOWED_TORCH_NAMES.fullmatch(name)):
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"global '{module}.{name}' is not allowed in a generator pickle")


def _bag_state_dict(obj, prefix=""):
    out = {}
    for k, v in (getattr(obj, "_parameters", None) or {}).items():
        if v is not None:
            out[prefix + k] = v.detach() if torch.is_tensor(v) else torch.as_tensor(v)
    for k, v in (getattr(obj, "_buffers", None) or {}).items():
        if v is not None and k not in (getattr(obj, "_non_persistent_buffers_set", None) or ()):
            out[prefix + k] = v.detach() if torch.is_tensor(v) else torch.as_tensor(v)
    for k, m in (getattr(obj, "_modules", None) or {}).items():
        if m is not None:
            out.update(_bag_state_dict(m, prefix + k + "."))
    return out


def load_nvidia(path, for_inference=None):
    """maua/GAN/load.py:129-164: the ``G_ema`` entry of a NVIDIA network pickle."""
    with open(path, "rb") as f:
        data = _PersistenceUnpickler(io.BytesIO(f.read())).load()
    G_persistence = data["G_ema"]
    state = G_persistence.state_dict() if hasattr(G_persistence, "state_dict") else _bag_state_dict(G_persistence)
    return generator_from_state(state)


def load_network(path, for_inference=False):
    """maua/GAN/load.py:192-207."""
    errors = {}
    for name, loader in [
        ("NVIDIA StyleGAN3 loader", load_nvidia),
        ("NVIDIA non-persistence loader", load_nvidia_pt),
        ("Rosinality StyleGAN2 to ADA-PT converter", load_rosinality2ada),
        ("Rosinality StyleGAN2 to Inference converter", partial(load_rosinality2ada, for_inference=True)),
    ]:
        try:
            return loader(path, for_inference=for_inference)
        except Exception:
            errors[name] = traceback.format_exc()
    error_str = "\n".join([f"\n{k}:\n{e}\n" for k, e in errors.items()])
    raise Exception(f"Error loading checkpoint! None of the converters succeeded:\n{error_str}")
