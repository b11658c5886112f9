"""Laplacian segmentation (maua_b200/audiovisual/audioreactive/segment.py) against the fixture the reference's own
rosa/segment.py agreed with (tests/golden/segment.pt, make_segment_golden.py), plus first-principles checks."""
import os

import torch

from maua_b200.audiovisual.audioreactive import segment as S

GOLD = os.path.join(os.path.dirname(__file__), "golden", "segment.pt")


def decided(soft, margin):
    top2 = soft.topk(2, dim=1).values
    return (top2[:, 0] - top2[:, 1]) > margin


def test_segmentation_matches_fixture():
    g = torch.load(GOLD, weights_only=False)
    segs = S.laplacian_segmentation(g["envelope"], g["beats"], ks=g["ks"])
    for k, got, soft, hard in zip(g["ks"], segs, g["soft"], g["hard"]):
        assert got.shape == (len(g["envelope"]), k)
        assert torch.allclose(got, soft, atol=1e-5)
        m = decided(soft, 1e-4)
        assert torch.equal(got.argmax(1)[m], hard[m])
        assert torch.allclose(got.sum(1), torch.ones(len(got)), atol=1e-5)   # soft one-hot memberships


def test_block_structure_is_recovered():
    """Five alternating blocks A B A C B: k = 3 must give A, B, C three different labels and both A (and both B) blocks
    the same one."""
    torch.manual_seed(1)
    protos = torch.randn(3, 8) * 2
    order = [0, 1, 0, 2, 1]
    env = torch.cat([protos[i].expand(60, 8) for i in order]) + 0.05 * torch.randn(300, 8)
    seg = S.laplacian_segmentation(env, list(range(4, 300, 5)), ks=[3])[0].argmax(1)
    mids = [int(seg[30 + 60 * b]) for b in range(5)]
    assert mids[0] == mids[2] and mids[1] == mids[4] and len({mids[0], mids[1], mids[3]}) == 3


def test_normalized_laplacian_and_helpers():
    a = torch.tensor([[0.0, 2.0, 0.0], [2.0, 5.0, 0.0], [0.0, 0.0, 0.0]])   # a self-loop and an isolated node
    lap = S.normalized_laplacian(a)
    assert torch.allclose(lap, torch.tensor([[1.0, -1.0, 0.0], [-1.0, 1.0, 0.0], [0.0, 0.0, 1.0]]))
    x = torch.arange(12.0).reshape(3, 4)
    assert torch.equal(S._shear(S._shear(x, 1), -1), x)
    assert torch.equal(S._shear(x, 1)[:, 1], torch.roll(x[:, 1], 1))
    segs = S.segmentations_from_features({"f": torch.randn(120, 4)}, list(range(5, 120, 6)), ks=(2, 4))
    assert set(segs) == {("f", 2), ("f", 4)} and segs[("f", 4)].shape == (120,) and int(segs[("f", 4)].max()) < 4


def test_rosa_helpers_match_scipy_and_separate_clusters():
    """Host pieces of laplacian_segmentation_rosa: the eigenvector median filter has scipy.ndimage's boundary rule, the dB
    conversion is librosa's amplitude_to_db(ref=max) formula, the seeded hard k-means recovers well separated clusters."""
    import numpy as np
    import scipy.ndimage

    from maua_b200.audiovisual.audioreactive import segment as S

    g = torch.Generator().manual_seed(5)
    x = torch.randn(23, 6, generator=g)
    want = scipy.ndimage.median_filter(x.numpy(), size=(9, 1))
    assert np.array_equal(S._median_filter_rows(x, 9).numpy(), want)

    mag = torch.rand(30, 12, generator=g) * 3
    mag[0, 0] = 1e-9
    db = S._amplitude_to_db_max(mag)
    ref = 20 * np.log10(np.maximum(mag.numpy(), 1e-5)) - 20 * np.log10(mag.numpy().max())
    ref = np.maximum(ref, ref.max() - 80.0)
    assert np.allclose(db.numpy(), ref, atol=1e-4) and float(db.max()) == 0.0 and float(db.min()) >= -80.0

    centres = torch.tensor([[0.0, 0.0], [10.0, 0.0], [0.0, 10.0]])
    truth = torch.arange(60) % 3
    pts = centres[truth] + 0.3 * torch.randn(60, 2, generator=g)
    lab = S.hard_k_means(pts, 3)
    assert len(set(lab.tolist())) == 3
    for j in range(3):   # every true cluster maps onto exactly one label
        assert len(set(lab[truth == j].tolist())) == 1
