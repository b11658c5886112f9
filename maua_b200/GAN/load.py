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
  the same mapping the reference's converter uses at :23,65,71); the inference network adds its noise unscaled
  (inference/ops.py:184), so ``noise_strength`` is folded into ``noise_const``.
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
# rosinality -> ADA / inference layout (maua/GAN/load.py:18-127)
# ---------------------------------------------------------------------------------------------------------------------
def rosinality_to_inference_state(state_dict, blur_scale=4.0):
    """Key-for-key restatement of load_rosinality2ada's mapping for ``for_inference=True`` (:18-113): returns
    (state_nv, max_res, num_map)."""
    state_ros = state_dict["g_ema"]
    state_nv = {}
    nv_key = "bs.0"
    if tuple(state_ros["input.input"].shape) != (1,):
        state_nv[f"synthesis.{nv_key}.const"] = state_ros["input.input"].squeeze(0)
    else:
        raise NotImplementedError("rosinality checkpoints with a learned-affine input (no const) are not supported")
    state_nv[f"synthesis.{nv_key}.conv1.noise_const"] = state_ros["noises.noise_0"].squeeze(0).squeeze(0)
    state_nv[f"synthesis.{nv_key}.conv1.weight"] = state_ros["conv1.conv.weight"].squeeze(0)
    state_nv[f"synthesis.{nv_key}.conv1.bias"] = state_ros["conv1.activate.bias"]
    state_nv[f"synthesis.{nv_key}.conv1.affine.weight"] = state_ros["conv1.conv.modulation.weight"]
    state_nv[f"synthesis.{nv_key}.conv1.affine.bias"] = state_ros["conv1.conv.modulation.bias"]
    state_nv[f"synthesis.{nv_key}.torgb.weight"] = state_ros["to_rgb1.conv.weight"].squeeze(0)
    state_nv[f"synthesis.{nv_key}.torgb.bias"] = state_ros["to_rgb1.bias"].squeeze(-1).squeeze(-1).squeeze(0)
    state_nv[f"synthesis.{nv_key}.torgb.affine.weight"] = state_ros["to_rgb1.conv.modulation.weight"]
    state_nv[f"synthesis.{nv_key}.torgb.affine.bias"] = state_ros["to_rgb1.conv.modulation.bias"]
    state_nv[f"synthesis.{nv_key}.resample_filter"] = state_ros["convs.0.conv.blur.kernel"] / blur_scale
    state_nv[f"synthesis.{nv_key}.conv1.resample_filter"] = state_ros["convs.0.conv.blur.kernel"] / blur_scale

    max_res, num_map = 4, 1
    for key, val in state_ros.items():
        if key.startswith("style"):
            _, num, weight_or_bias = key.split(".")
            state_nv[f"mapping.fcs.{int(num) - 1}.{weight_or_bias}"] = val
            num_map = max(num_map, int(num))
        if key.startswith("noises"):
            n = int(key.split("_")[1])
            if n == 0:
                continue
            state_nv[f"synthesis.bs.{(n - 1) // 2 + 1}.conv{(n - 1) % 2}.noise_const"] = val.squeeze(0).squeeze(0)
        if key.startswith("convs"):
            n = int(key.split(".")[1])
            r = 2 ** (3 + n // 2)
            nv_block = f"synthesis.bs.{(n // 2) + 1}"
            ros_name = ".".join(key.split(".")[2:])
            if ros_name == "conv.weight":
                state_nv[f"{nv_block}.conv{n % 2}.weight"] = val.squeeze(0)
            elif ros_name == "activate.bias":
                state_nv[f"{nv_block}.conv{n % 2}.bias"] = val
            elif ros_name == "conv.modulation.weight":
                state_nv[f"{nv_block}.conv{n % 2}.affine.weight"] = val
            elif ros_name == "conv.modulation.bias":
                state_nv[f"{nv_block}.conv{n % 2}.affine.bias"] = val
            elif ros_name == "noise.weight":
                pass  # the inference layout has no noise_strength (:39,84)
            elif ros_name == "conv.blur.kernel":
                state_nv[f"{nv_block}.conv0.resample_filter"] = val / blur_scale
                state_nv[f"{nv_block}.conv1.resample_filter"] = val / blur_scale
            else:
                raise Exception(f"Key {key} not recognized!")
            max_res = max(max_res, r)
        if key.startswith("to_rgbs"):
            n = int(key.split(".")[1])
            nv_block = f"synthesis.bs.{n + 1}"
            ros_name = ".".join(key.split(".")[2:])
            if ros_name == "conv.weight":
                state_nv[f"{nv_block}.torgb.weight"] = val.squeeze(0)
            elif ros_name == "bias":
                state_nv[f"{nv_block}.torgb.bias"] = val.squeeze(-1).squeeze(-1).squeeze(0)
            elif ros_name == "conv.modulation.weight":
                state_nv[f"{nv_block}.torgb.affine.weight"] = val
            elif ros_name == "conv.modulation.bias":
                state_nv[f"{nv_block}.torgb.affine.bias"] = val
            elif ros_name == "upsample.kernel":
                state_nv[f"{nv_block}.resample_filter"] = val / blur_scale
            else:
                raise Exception(f"Key {key} not recognized!")
    state_nv["mapping.w_avg"] = state_dict["latent_avg"] if "latent_avg" in state_dict else torch.zeros(512)
    return state_nv, max_res, num_map


def load_rosinality2ada(path, blur_scale=4.0, for_inference=False):
    """maua/GAN/load.py:18-127.  Both values of for_inference give the inference-layout network (the only StyleGAN2
    network of this build)."""
    state_dict = torch.load(path, map_location="cpu", weights_only=False)
    state_nv, max_res, num_map = rosinality_to_inference_state(state_dict, blur_scale)
    z_dim = w_dim = state_nv["mapping.fcs.0.weight"].shape[1] if "mapping.fcs.0.weight" in state_nv else 512
    G = stylegan2.Generator(z_dim, 0, w_dim, max_res, 3, mapping_kwargs=dict(num_layers=num_map), **_sg2_channels(state_nv))
    G.load_state_dict(state_nv)
    return G


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
    """NVIDIA training layout -> in-tree inference layout: b{res} -> bs.{log2(res) - 2}, mapping.fc{i} -> mapping.fcs.{i},
    noise_const * noise_strength folded (the inference net adds noise unscaled, inference/ops.py:184)."""
    out = {}
    for k, v in state.items():
        m = re.fullmatch(r"synthesis\.b(\d+)\.(.*)", k)
        if m:
            k = f"synthesis.bs.{int(math.log2(int(m.group(1)))) - 2}.{m.group(2)}"
        m = re.fullmatch(r"mapping\.fc(\d+)\.(weight|bias)", k)
        if m:
            k = f"mapping.fcs.{m.group(1)}.{m.group(2)}"
        out[k] = v
    for k in [k for k in out if k.endswith(".noise_strength")]:
        strength = out.pop(k)
        nc = k[: -len("noise_strength")] + "noise_const"
        out[nc] = out[nc] * strength
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
    if any(re.match(r"synthesis\.b\d+\.", k) for k in state):
        state = nvidia_sg2_to_inference_state(state)
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
    return G


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
    """legacy._LegacyUnpickler (nvGAN legacy.py) without dnnlib: persistent objects become _Bag trees; only torch /
    numpy / collections / builtins globals needed to rebuild tensors are allowed."""

    _ALLOWED_PREFIXES = ("torch", "numpy", "collections", "builtins", "_codecs")

    def find_class(self, module, name):
        if module == "torch_utils.persistence" and name == "_reconstruct_persistent_obj":
            return _reconstruct_persistent_obj
        if module == "dnnlib.util" and name == "EasyDict":
            return dict
        if module.split(".")[0] in self._ALLOWED_PREFIXES:
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
