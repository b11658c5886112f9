"""Checkpoint loaders (maua/GAN/load.py:18-207) on synthetic checkpoints written in the three formats the
reference reads: rosinality ``{"g_ema": ...}``, NVIDIA ``{"G_ema": state_dict}`` .pt, and NVIDIA persistence pickles.
Pure host code: no GPU, no NVIDIA training code."""
import pickle
import sys
import types

import pytest
import torch

from maua_b200.GAN import load as L
from maua_b200.GAN.networks import stylegan2, stylegan3

SG2_KW = dict(channel_base=256, channel_max=16)


def small_sg2(seed=0):
    torch.manual_seed(seed)
    return stylegan2.Generator(512, 0, 512, 16, 3, mapping_kwargs=dict(num_layers=3), **SG2_KW)


def to_rosinality(G, blur_scale=4.0):
    """Inverse of the reference's key table (maua/GAN/load.py:23-113) for a generator in the inference layout."""
    s = G.state_dict()
    ros = {"input.input": s["synthesis.bs.0.const"].unsqueeze(0)}
    ros["noises.noise_0"] = s["synthesis.bs.0.conv1.noise_const"][None, None]
    ros["conv1.conv.weight"] = s["synthesis.bs.0.conv1.weight"].unsqueeze(0)
    ros["conv1.activate.bias"] = s["synthesis.bs.0.conv1.bias"]
    ros["conv1.conv.modulation.weight"] = s["synthesis.bs.0.conv1.affine.weight"]
    ros["conv1.conv.modulation.bias"] = s["synthesis.bs.0.conv1.affine.bias"]
    ros["conv1.noise.weight"] = torch.ones(1)
    ros["to_rgb1.conv.weight"] = s["synthesis.bs.0.torgb.weight"].unsqueeze(0)
    ros["to_rgb1.bias"] = s["synthesis.bs.0.torgb.bias"][None, :, None, None]
    ros["to_rgb1.conv.modulation.weight"] = s["synthesis.bs.0.torgb.affine.weight"]
    ros["to_rgb1.conv.modulation.bias"] = s["synthesis.bs.0.torgb.affine.bias"]
    n_blocks = len(G.synthesis.bs)
    for b in range(1, n_blocks):
        for c in range(2):
            n = 2 * (b - 1) + c
            p = f"synthesis.bs.{b}.conv{c}"
            ros[f"convs.{n}.conv.weight"] = s[f"{p}.weight"].unsqueeze(0)
            ros[f"convs.{n}.activate.bias"] = s[f"{p}.bias"]
            ros[f"convs.{n}.conv.modulation.weight"] = s[f"{p}.affine.weight"]
            ros[f"convs.{n}.conv.modulation.bias"] = s[f"{p}.affine.bias"]
            ros[f"convs.{n}.noise.weight"] = torch.ones(1)
            ros[f"noises.noise_{n + 1}"] = s[f"{p}.noise_const"][None, None]
            if c == 0:
                ros[f"convs.{n}.conv.blur.kernel"] = s[f"{p}.resample_filter"] * blur_scale
        p = f"synthesis.bs.{b}.torgb"
        ros[f"to_rgbs.{b - 1}.conv.weight"] = s[f"{p}.weight"].unsqueeze(0)
        ros[f"to_rgbs.{b - 1}.bias"] = s[f"{p}.bias"][None, :, None, None]
        ros[f"to_rgbs.{b - 1}.conv.modulation.weight"] = s[f"{p}.affine.weight"]
        ros[f"to_rgbs.{b - 1}.conv.modulation.bias"] = s[f"{p}.affine.bias"]
        ros[f"to_rgbs.{b - 1}.upsample.kernel"] = s[f"synthesis.bs.{b}.resample_filter"] * blur_scale
    for i in range(len(G.mapping.fcs)):
        ros[f"style.{i + 1}.weight"] = s[f"mapping.fcs.{i}.weight"]
        ros[f"style.{i + 1}.bias"] = s[f"mapping.fcs.{i}.bias"]
    return {"g_ema": ros, "latent_avg": s["mapping.w_avg"]}


def assert_same_state(a, b):
    sa, sb = a.state_dict(), b.state_dict()
    assert set(sa) == set(sb)
    for k in sa:
        assert torch.allclose(sa[k], sb[k]), k


def test_rosinality_checkpoint_round_trip(tmp_path):
    G = small_sg2()
    G.mapping.w_avg.copy_(torch.randn(512))
    path = tmp_path / "ros.pt"
    torch.save(to_rosinality(G), path)
    for loader in (L.load_rosinality2ada, L.load_network):
        G2 = loader(str(path), for_inference=True)
        assert isinstance(G2, stylegan2.Generator) and G2.img_resolution == 16 and G2.mapping.num_layers == 3
        assert_same_state(G, G2)


def to_nvidia_sg2(G):
    s = G.state_dict()
    out = {}
    for k, v in s.items():
        if k.startswith("synthesis.bs."):
            parts = k.split(".")
            k = ".".join(["synthesis", f"b{4 * 2 ** int(parts[2])}"] + parts[3:])
        if k.startswith("mapping.fcs."):
            k = k.replace("mapping.fcs.", "mapping.fc")
        out[k] = v.clone()
    for k in [k for k in out if k.endswith("noise_const")]:
        out[k[: -len("noise_const")] + "noise_strength"] = torch.tensor(0.5)
    return out


def test_nvidia_pt_stylegan2_training_layout(tmp_path):
    G = small_sg2(1)
    path = tmp_path / "nv2.pt"
    torch.save({"G_ema": to_nvidia_sg2(G)}, path)
    G2 = L.load_network(str(path))
    assert isinstance(G2, stylegan2.Generator)
    assert_same_state(G, G2)
    # what the training network does differently stays with the loaded generator: the trained noise strength of every layer
    # (it scales the constant and any per-frame map the wrapper swaps in) ...
    layers = [m for m in G2.synthesis.modules() if isinstance(m, stylegan2.SynthesisLayer)]
    assert layers and all(m.noise_strength == 0.5 for m in layers)
    name = "bs.1.conv0.noise_const"
    nc = G2.synthesis.get_buffer(name)
    assert torch.equal(G2.synthesis._prepare_param(name, nc.clone()), nc * 0.5)
    assert G2.synthesis._param_key(name, nc)[-1] == 0.5
    # ... and the standard x @ W.T mapping layers: the checkpoint's mapping must give what plain F.linear + lrelu gives
    z = torch.randn(4, 512)
    x = z * (z.square().mean(1, keepdim=True) + 1e-8).rsqrt()
    sd = to_nvidia_sg2(G)
    for i in range(3):
        w, b = sd[f"mapping.fc{i}.weight"], sd[f"mapping.fc{i}.bias"]
        x = torch.nn.functional.leaky_relu(torch.nn.functional.linear(x, w * (0.01 / 512 ** 0.5), b * 0.01), 0.2) * 2 ** 0.5
    got = G2.mapping(z)
    assert torch.allclose(got[:, 0], x, atol=1e-5), float((got[:, 0] - x).abs().max())
    assert not torch.allclose(G.mapping(z)[:, 0], x, atol=1e-3)   # the inference-layout quirk multiplies by W itself


def test_rosinality_training_semantics(tmp_path):
    """for_inference=False: the reference builds the training network, which applies ``noise.weight`` and standard FCs."""
    G = small_sg2(5)
    ck = to_rosinality(G)
    for k in ck["g_ema"]:
        if k.endswith("noise.weight"):
            ck["g_ema"][k] = torch.full((1,), 0.25)
    path = tmp_path / "ros_train.pt"
    torch.save(ck, path)
    G_train = L.load_rosinality2ada(str(path), for_inference=False)
    G_inf = L.load_rosinality2ada(str(path), for_inference=True)
    assert_same_state(G_train, G_inf)
    assert all(fc.standard_matmul for fc in G_train.mapping.fcs) and not any(fc.standard_matmul for fc in G_inf.mapping.fcs)
    assert G_train.synthesis.bs[0].conv1.noise_strength == 0.25 and G_train.synthesis.bs[2].conv0.noise_strength == 0.25
    assert G_inf.synthesis.bs[0].conv1.noise_strength == 1.0


def test_rosinality_unknown_key_aborts(tmp_path):
    ck = to_rosinality(small_sg2(6))
    ck["g_ema"]["convs.0.conv.surprise"] = torch.zeros(1)
    with pytest.raises(Exception, match="not recognized"):
        L.rosinality_to_inference_state(ck)


@pytest.mark.parametrize("config", ["T", "R"])
def test_nvidia_pt_stylegan3(tmp_path, config):
    torch.manual_seed(2)
    kw = dict(channel_base=2048, channel_max=32)
    if config == "R":
        kw.update(conv_kernel=1, use_radial_filters=True, channel_base=4096, channel_max=64)
    G = stylegan3.Generator(512, 0, 512, 256, 3, mapping_kwargs=dict(num_layers=2), **kw)
    path = tmp_path / "nv3.pt"
    torch.save({"G_ema": G.state_dict()}, path)
    G2 = L.load_network(str(path))
    assert isinstance(G2, stylegan3.Generator) and G2.img_resolution == 256
    assert G2.synthesis.layer_names == G.synthesis.layer_names
    assert_same_state(G, G2)
    # what the wrappers read from a loaded network (wrappers/stylegan3.py:38-40)
    assert (G2.synthesis.w_dim, G2.synthesis.num_ws, G2.mapping.z_dim) == (512, 16, 512)


def test_nvidia_persistence_pickle_without_nvidia_code(tmp_path):
    """A pickle in the torch_utils.persistence format: every module reduces to
    _reconstruct_persistent_obj(dict(type='class', version, module_src, class_name, state=module.__dict__))."""
    G = small_sg2(3)
    fake = types.ModuleType("torch_utils.persistence")

    def _reconstruct_persistent_obj(meta):  # only ever resolved by name at load time
        raise AssertionError("the loader must not execute the pickled module source")

    _reconstruct_persistent_obj.__module__ = "torch_utils.persistence"
    _reconstruct_persistent_obj.__qualname__ = "_reconstruct_persistent_obj"
    fake._reconstruct_persistent_obj = _reconstruct_persistent_obj
    pkg = types.ModuleType("torch_utils")
    pkg.persistence = fake
    sys.modules["torch_utils"], sys.modules["torch_utils.persistence"] = pkg, fake

    class Persistent:
        def __init__(self, module):
            state = dict(module.__dict__)
            state["_modules"] = {k: Persistent(m) for k, m in module._modules.items() if m is not None}
            state = {k: v for k, v in state.items() if k in ("_parameters", "_buffers", "_modules", "z_dim", "c_dim", "w_dim",
                                                            "img_resolution", "img_channels", "num_ws")}
            self.meta = dict(type="class", version=6, module_src="raise RuntimeError('never executed')",
                             class_name=type(module).__name__, state=state)

        def __reduce__(self):
            return (fake._reconstruct_persistent_obj, (self.meta,))

    try:
        nv = small_sg2(3)
        nv.load_state_dict(G.state_dict())
        path = tmp_path / "network.pkl"
        with open(path, "wb") as f:
            pickle.dump({"G": None, "D": None, "G_ema": Persistent(nv), "training_set_kwargs": None}, f)
    finally:
        del sys.modules["torch_utils"], sys.modules["torch_utils.persistence"]
    G2 = L.load_nvidia(str(path))
    assert_same_state(G, G2)
    assert L.load_network(str(path)).img_resolution == 16


def test_unpickler_refuses_foreign_globals(tmp_path):
    path = tmp_path / "evil.pkl"
    with open(path, "wb") as f:
        pickle.dump({"G_ema": types.SimpleNamespace(a=1)}, f)
    with pytest.raises(pickle.UnpicklingError):
        L.load_nvidia(str(path))
    with pytest.raises(Exception, match="None of the converters succeeded"):
        L.load_network(str(path))


class _Evil:
    def __reduce__(self):
        return (eval, ("{'G_ema': 'pwned'}",))


def test_unpickler_refuses_code_execution(tmp_path):
    """A pickle whose reduce calls eval / torch.load / numpy.load must not get those globals (they sit under the module
    prefixes a tensor pickle needs, so prefixes are not enough: the loader whitelists exact names)."""
    path = tmp_path / "eval.pkl"
    with open(path, "wb") as f:
        pickle.dump(_Evil(), f)
    with pytest.raises(pickle.UnpicklingError, match="builtins.eval"):
        L.load_nvidia(str(path))
    for module, name in [("torch", "load"), ("numpy", "load"), ("torch.hub", "load"), ("os", "system"), ("builtins", "exec")]:
        with pytest.raises(pickle.UnpicklingError):
            L._PersistenceUnpickler(__import__("io").BytesIO(b"")).find_class(module, name)


def test_wrapper_loads_a_model_file(tmp_path):
    from maua_b200.GAN.wrappers.stylegan3 import StyleGAN3

    torch.manual_seed(4)
    G = stylegan3.Generator(512, 0, 512, 256, 3, mapping_kwargs=dict(num_layers=2), channel_base=2048, channel_max=32)
    path = tmp_path / "sg3.pt"
    torch.save({"G_ema": G.state_dict()}, path)
    gen = StyleGAN3(model_file=str(path))
    assert gen.synthesizer.G_synth.img_resolution == 256 and gen.synthesizer.output_size == (256, 256)
    assert torch.equal(gen.mapper.G_map.fc0.weight, G.mapping.fc0.weight)
    assert gen.synthesizer.avg_shift.shape == (4,)
