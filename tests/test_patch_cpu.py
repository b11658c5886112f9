"""Random patch generator (maua_b200/audiovisual/audioreactive/patch.py) against the tables the reference's own Patch
class drew for the same seeds on a CPU generator (tests/golden/patch.json, written by make_patch_golden.py)."""
import json
import os

import torch

from maua_b200.audiovisual.audioreactive import patch as P

GOLD = os.path.join(os.path.dirname(__file__), "golden", "patch.json")


def make(seed, T=64):
    features = {k: torch.zeros(T, 1) for k in P.ALLFEATS}
    segmentations = {(name, k): torch.zeros(T, dtype=torch.long) for name in P.ALLFEATS for k in (2, 4, 6, 8, 12, 16)}
    return P.Patch(features, segmentations, tempo=120.0, fps=24, seed=seed, device="cpu")


def test_subpatch_tables_match_reference_draws():
    for rec in json.load(open(GOLD)):
        p = make(rec["seed"])
        assert (p.n_base_latents, p.sigma_base_noise, p.loops_base_noise) == (rec["n_base_latents"], rec["sigma_base_noise"], rec["loops_base_noise"])
        assert 2 <= len(p.latent_patches) < 20 and 2 <= len(p.noise_patches) < 20
        p.update_intensity(0.7)
        assert p.latent_patches == rec["latent_patches"] and p.noise_patches == rec["noise_patches"]
        assert repr(p) == rec["repr"]


def test_save_load_round_trip(tmp_path):
    p = make(42)
    path = tmp_path / "patch.json"
    p.save(str(path))
    state = json.load(open(path))
    assert set(state) == {"seed", "latent_patches", "noise_patches", "n_base_latents", "sigma_base_noise", "loops_base_noise"}
    q = P.Patch.load(str(path), p.features, p.segmentations, 120.0, 24, "cpu")
    assert q.seed == 42 and q.latent_patches == p.latent_patches and q.noise_patches == p.noise_patches
    assert (q.n_base_latents, q.sigma_base_noise, q.loops_base_noise) == (p.n_base_latents, p.sigma_base_noise, p.loops_base_noise)


def test_pickle_drops_and_rebuilds_the_generator():
    import pickle

    p = make(7)
    q = pickle.loads(pickle.dumps(p))
    assert q.latent_patches == p.latent_patches and isinstance(q.rng, torch.Generator) and q.rng.initial_seed() == 7


def test_forward_needs_cuda():
    import pytest

    with pytest.raises(RuntimeError):
        make(42).forward(torch.randn(20, 18, 16))
