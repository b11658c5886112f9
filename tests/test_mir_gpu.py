"""retrieve_music_information end to end on the device (features, tempo / beats, Laplacian segmentations) feeding the
random Patch generator: the pre-pass of selfsupervised/sample.py:52-77 without any host DSP library."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_music_information_drives_a_random_patch(cuda):
    from maua_b200.audiovisual.audioreactive.mir import retrieve_music_information
    from maua_b200.audiovisual.audioreactive.patch import ALLFEATS, Patch

    fps, seconds = 24, 20
    sr = 1024 * fps
    t = torch.arange(seconds * sr) / sr
    # 120 BPM clicks (in real time) over two alternating tonal sections
    clicks = (torch.remainder(t, 0.5) < 0.01).float() * torch.randn(len(t)) * 0.8
    tone = torch.where(torch.remainder(t, 10.0) < 5.0, torch.sin(2 * torch.pi * 220 * t), torch.sin(2 * torch.pi * 392 * t))
    audio = 0.3 * tone + clicks
    ks = (2, 4, 6)
    features, segmentations, tempo = retrieve_music_information(audio, sr, ks=ks, device=cuda)
    T = seconds * fps
    assert list(features) == ALLFEATS[:4] + ["spectral_flatness", "rms", "drop_strength", "onsets"]
    for k, v in features.items():
        assert v.shape[0] == T and v.is_cuda and torch.isfinite(v).all(), k
        assert float(v.min()) >= 0 and float(v.max()) <= 1 + 1e-6
    assert set(segmentations) == {(n, k) for n in list(features) + ["rosa"] for k in ks}
    for (n, k), seg in segmentations.items():
        assert seg.shape == (T,) and int(seg.max()) < k
    # 2 clicks per second at 24 frames/s = a 12-frame period; in the reference's units (21.53 frames/s assumed) that reads 107.7 BPM
    assert abs(tempo - 60.0 * (22050 / 1024) / 12.0) < 1e-6, tempo
    # the two tonal sections (A B A B, 5 s each) separate in the chroma segmentation with k = 2
    seg = segmentations[("chromagram", 2)].cpu().numpy()
    mids = [int(np.bincount(seg[int((5 * i + 1.5) * fps): int((5 * i + 3.5) * fps)]).argmax()) for i in range(4)]
    assert mids[0] == mids[2] and mids[1] == mids[3] and mids[0] != mids[1], mids

    # ... and in the whole-track ("rosa") segmentation built from constant-Q / MFCC recurrence
    rosa = segmentations[("rosa", 2)].cpu().numpy()
    rmids = [int(np.bincount(rosa[int((5 * i + 1.5) * fps): int((5 * i + 3.5) * fps)]).argmax()) for i in range(4)]
    assert rmids[0] == rmids[2] and rmids[1] == rmids[3] and rmids[0] != rmids[1], rmids

    patch = Patch(features, segmentations, tempo, fps=fps, seed=3, device="cpu")
    palette = torch.randn(30, 18, 512, device=cuda)
    latents, noise = patch.forward(palette, downscale_factor=8)
    assert latents.shape == (T, 18, 512) and torch.isfinite(latents).all() and len(noise) == 17
    assert noise[5].forward(0, 4).shape == (4, 4, 4)
