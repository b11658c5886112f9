"""Pins maua_b200's Laplacian segmentation against the REFERENCE's own rosa/segment.py and writes
tests/golden/segment.pt.  Runs only where /root/reference exists.  segment.py imports torch_geometric.utils.get_laplacian
(absent, un-pinned) and, further down, librosa / sklearn for a second function that is not exercised: both are stubbed,
get_laplacian by the published definition (self-loops removed, L = I - D^-1/2 A D^-1/2 as COO edges, fill value 1).
    python tests/golden/make_segment_golden.py
"""
import importlib.util
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/maua/audiovisual/audioreactive/selfsupervised/features/rosa/segment.py"
assert os.path.exists(REF)


def get_laplacian(edge_index, edge_weight, normalization="sym"):
    assert normalization == "sym"
    keep = edge_index[0] != edge_index[1]
    edge_index, edge_weight = edge_index[:, keep], edge_weight[keep]
    n = int(edge_index.max()) + 1
    deg = torch.zeros(n, dtype=edge_weight.dtype).scatter_add_(0, edge_index[0], edge_weight)
    dis = deg.pow(-0.5)
    dis[torch.isinf(dis)] = 0
    w = -dis[edge_index[0]] * edge_weight * dis[edge_index[1]]
    loops = torch.arange(n)
    return torch.cat([edge_index, torch.stack([loops, loops])], dim=1), torch.cat([w, torch.ones(n, dtype=w.dtype)])


tg = types.ModuleType("torch_geometric"); tgu = types.ModuleType("torch_geometric.utils")
tgu.get_laplacian = get_laplacian
tg.utils = tgu
sys.modules["torch_geometric"], sys.modules["torch_geometric.utils"] = tg, tgu
for name in ("librosa", "sklearn"):
    if name not in sys.modules:
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
spec = importlib.util.spec_from_file_location("ref_segment", REF)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

from maua_b200.audiovisual.audioreactive import segment as S  # noqa: E402

torch.manual_seed(0)
T = 420
# a feature track with block structure (so the segmentation is not arbitrary) plus noise
blocks = torch.randn(5, 12)
labels = torch.tensor([0] * 70 + [1] * 60 + [2] * 80 + [0] * 70 + [3] * 60 + [4] * 80)
envelope = blocks[labels] + 0.15 * torch.randn(T, 12)
beats = list(range(5, T, 6))

ra, rb = ref.recurrence_matrix(envelope[::6].clone(), width=3, sym=True), S.recurrence_matrix(envelope[::6].clone(), width=3, sym=True)
assert torch.equal(ra, rb), "recurrence_matrix"
assert torch.equal(ref.timelag_median_filter(ra), S.timelag_median_filter(rb)), "timelag_median_filter"
x = torch.randn(9, 40)
assert torch.equal(ref.median_filter1d(x, k=9, s=1, p=4), S.median_filter1d(x, k=9, s=1, p=4))
ks = [2, 4, 6, 8]
sa, sb = ref.laplacian_segmentation(envelope, beats, ks=ks), S.laplacian_segmentation(envelope, beats, ks=ks)
for k, a, b in zip(ks, sa, sb):
    assert a.shape == b.shape == (T, k)
    assert torch.allclose(a, b, atol=1e-5), (k, (a - b).abs().max())
    top2 = a.topk(2, dim=1).values
    decided = (top2[:, 0] - top2[:, 1]) > 1e-4           # frames whose segment is not a floating-point coin toss
    assert torch.equal(a.argmax(1)[decided], b.argmax(1)[decided]), k
    print(f"k={k}: soft memberships max diff {(a - b).abs().max():.1e}, {int((~decided).sum())} of {T} frames undecided")
torch.save(dict(envelope=envelope, beats=beats, ks=ks, hard=[s.argmax(1) for s in sb], soft=sb), os.path.join(ROOT, "tests", "golden", "segment.pt"))
print("maua_b200 segment.py == reference rosa/segment.py (recurrence, time-lag filter, segmentation for k =", ks, "); wrote tests/golden/segment.pt")
print("segments found for k=4:", torch.unique_consecutive(sb[1].argmax(1)).tolist())
