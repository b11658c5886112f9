"""Host-side facade: same names / arguments / error behaviour as maua.GAN.wrappers (no GPU needed)."""
import numpy as np
import pytest
import torch

from maua_b200.GAN.networks import stylegan3 as N
from maua_b200.GAN.wrappers import MauaGenerator, MauaMapper, MauaSynthesizer, get_generator_class
from maua_b200.GAN.wrappers.stylegan3 import StyleGAN3, StyleGAN3Synthesizer, make_transform_mat
from oracle import sg3 as O


@pytest.fixture(scope="module")
def gen():
    torch.manual_seed(0)
    return get_generator_class("stylegan3")(model_file=None)


def test_class_surface(gen):
    assert isinstance(gen, StyleGAN3) and isinstance(gen, MauaGenerator)
    assert isinstance(gen.mapper, MauaMapper) and isinstance(gen.synthesizer, MauaSynthesizer)
    assert (gen.z_dim, gen.c_dim, gen.w_dim, gen.res) == (512, 0, 512, 1024)
    assert gen.synthesizer.num_ws == 16 and gen.synthesizer.output_size == (1024, 1024)
    assert set(gen.synthesizer.modulation_targets) == {"latent_w", "latent_w_plus", "translation", "rotation"}
    assert gen.synthesizer.avg_shift.tolist() == [1.0, 0.0, 0.0, 0.0]
    with pytest.raises(Exception):
        get_generator_class("biggan")


def test_seed_to_latent_convention(gen):
    z = gen.get_z_latents("1-4,7")
    assert z.shape == (4, 512)
    assert np.array_equal(z[0].numpy(), np.random.RandomState(1).randn(1, 512)[0])
    assert np.array_equal(z[3].numpy(), np.random.RandomState(7).randn(1, 512)[0])


def test_state_dict_keys_match_upstream_layout(gen):
    keys = set(gen.synthesizer.G_synth.state_dict().keys())
    okeys = set(O.make_synthesis("T", 1024, seed=0).state_dict().keys())
    assert keys == okeys
    assert "L3_52_512.affine.weight" in keys and "input.freqs" in keys and "L13_1024_32.up_filter" in keys


def test_random_init_order_matches_oracle():
    kw = dict(channel_base=2048, channel_max=32)
    onet = O.make_synthesis("T", img_resolution=128, seed=5, **kw)
    torch.manual_seed(5)
    net = N.SynthesisNetwork(w_dim=512, img_resolution=128, img_channels=3, **kw)
    for k, v in onet.state_dict().items():
        assert torch.equal(v, net.state_dict()[k]), k


def test_make_transform_mat_takes_pinv_path():
    with pytest.warns(UserWarning):
        m = make_transform_mat(torch.tensor([[0.1, -0.2]]), torch.tensor([30.0]))
    assert m.shape == (3, 3) and torch.isfinite(m).all()


def test_no_cpu_fallback(gen):
    with pytest.raises(RuntimeError):
        gen.synthesizer.G_synth(torch.randn(1, 16, 512))
    with pytest.raises(RuntimeError):
        next(gen.render({"latents": torch.randn(2, 16, 512)}, device="cpu"))
    # a non-native output size is a host-side decision (hook module + size); rendering it still needs the GPU
    S = StyleGAN3Synthesizer(None, False, (512, 768), "stretch", 0)
    assert S.G_synth.output_hw() == (768, 512) and len(S._hook_handles) == 1
    with pytest.raises(RuntimeError):
        S(latents=torch.randn(1, 16, 512))
    S.refresh_model_hooks()
    assert S.G_synth.output_hw() == (1024, 1024) and S._hook_handles == []


def test_mapping_network_matches_oracle():
    torch.manual_seed(3)
    m = N.MappingNetwork(z_dim=512, c_dim=0, w_dim=512, num_ws=16)
    torch.manual_seed(3)
    om = O.MappingNetwork(z_dim=512, c_dim=0, w_dim=512, num_ws=16)
    z = torch.randn(3, 512)
    assert torch.allclose(m(z), om(z), atol=1e-5)
    assert torch.allclose(m(z, truncation_psi=0.7), om(z, truncation_psi=0.7), atol=1e-5)
