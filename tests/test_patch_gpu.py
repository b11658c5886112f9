"""Patch.forward on the device against the same composition of oracle functions (oracle/selfsup.py, oracle/noise.py),
driven by a CPU generator so both sides draw the same permutations and noise fields."""
import pytest
import torch

from oracle import selfsup as OS

pytestmark = pytest.mark.gpu


def test_forward_matches_oracle_composition(cuda):
    from maua_b200.audiovisual.audioreactive import patch as P

    T = 96
    torch.manual_seed(0)
    dims = {"chromagram": 12, "tonnetz": 6, "mfcc": 20, "spectral_contrast": 7}
    feats = {k: torch.rand(T, dims.get(k, 1)) for k in P.ALLFEATS}
    segs = {(name, k): torch.randint(0, k, (T,)) for name in P.ALLFEATS for k in (2, 4, 6, 8)}
    palette = torch.randn(24, 18, 32)
    p = P.Patch({k: v.to(cuda) for k, v in feats.items()}, {k: v.to(cuda) for k, v in segs.items()}, tempo=110.0, fps=24, seed=11,
                device="cpu")
    latents, noise = p.forward(palette.to(cuda), downscale_factor=4)
    assert latents.shape == (T, 18, 32) and len(noise) == 17

    # oracle side: the same draws in the same order (patch.py:128-153)
    rng = torch.Generator().manual_seed(11)
    base = torch.randperm(len(palette), generator=rng)[: p.n_base_latents]
    want = OS.spline_loop_latents(palette[base], T)
    for sub in p.latent_patches:
        want = OS.latent_patch(rng, want, palette, segs, feats, 110.0, 24, **sub)
    err = float((latents.cpu() - want).abs().max())
    print(f"Patch.forward latents vs oracle composition: max abs err {err:.2e} over {len(p.latent_patches)} sub-patches")
    assert err <= 2e-4
    # the noise sequencers are lazy: evaluate one batch of every layer
    for n, mod in zip([4, 8, 8, 16, 16, 32, 32, 64, 64, 128, 128, 256, 256, 512, 512, 1024, 1024], noise):
        out = mod(8, 4)
        assert out.shape == (4, round(n / 4), round(n / 4)) and torch.isfinite(out).all()
    again, _ = p.forward(palette.to(cuda), downscale_factor=4)
    assert torch.equal(again, latents)
