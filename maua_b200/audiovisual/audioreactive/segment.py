"""Laplacian structural segmentation of a feature envelope: mirror of
maua/audiovisual/audioreactive/selfsupervised/features/rosa/segment.py:7-190, the method of McFee & Ellis
(2014) as the reference restates it in torch: beat-synchronous medians -> k-nearest-neighbour recurrence affinity, cleaned
by a median filter along the time-lag diagonals -> sequence affinity between consecutive beats -> balanced combination ->
symmetric normalised graph Laplacian -> leading eigenvectors -> soft k-means memberships, one segmentation per k.

Part of the once-per-track pre-pass: the matrices are (number of beats)^2, a few hundred rows.  Dense torch linear algebra
(eigh, topk, median) on whatever device the envelope lives on; nothing here is on the render hot path.  The reference
takes its normalised Laplacian from torch_geometric.utils.get_laplacian (absent, un-pinned): restated as
L = I - D^-1/2 A D^-1/2 with self-loops removed first and isolated nodes kept at L_ii = 1, which is what that function
returns for normalization="sym".  Everything else is pinned against the reference's own functions
(tests/golden/make_segment_golden.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _pairwise_distance(x):
    """sqrt(sum (x_i - x_j)^2 + 1e-8): the reference adds the epsilon under the root (segment.py:7-20)."""
    diff = x.unsqueeze(1) - x.unsqueeze(0)
    return (diff.pow(2).sum(2) + 1e-8).sqrt()


def recurrence_matrix(data, k=None, width=1, sym=False, bandwidth=None):
    """Affinity recurrence matrix (segment.py:23-60): each point keeps links to its k nearest neighbours outside the
    +-width band, optionally only mutual ones, weighted exp(-distance / bandwidth)."""
    t = data.shape[0]
    data = data.flatten(1)
    if k is None:
        k = 2 * np.ceil(np.sqrt(t - 2 * width + 1)) if t > 2 * width + 1 else 2
    k = int(k)
    rec = _pairwise_distance(data)
    for d in range(-width + 1, width):
        torch.diagonal(rec, offset=d).fill_(0)
    rec = rec + (rec == 0).float() * 1e20                       # banned links can never be among the nearest
    nearest = torch.topk(rec, k, dim=0, largest=False)
    rec = torch.scatter(torch.zeros_like(rec), dim=0, index=nearest.indices, src=nearest.values)
    if sym:
        rec = rec.minimum(rec.T)                                # mutual neighbours only
    if bandwidth is None:
        bandwidth = torch.median(rec.max(axis=1).values)
        if not bandwidth > 0:   # degenerate track (the reference divides by zero here)
            bandwidth = rec.max().clamp_min(1e-12)
    rec = rec * (1 - (rec < 0).float())
    rec = torch.exp(rec / (-1 * bandwidth))
    return rec * (1 - (rec >= 1).float())                       # exp(0) = 1 marks "no link"


def median_filter1d(x, k=3, s=1, p=1):
    """Running median along the last axis of a [rows, n] matrix, reflect padded (segment.py:63-67)."""
    x = F.pad(x.unsqueeze(0), (p, p, 0, 0), mode="reflect").squeeze(0)
    return x.unfold(1, k, s).median(dim=-1).values


def _shear(x, factor):
    """Column i rolled by factor * i: turns time-time into time-lag coordinates and back (segment.py:70-74)."""
    n = x.shape[0]
    rows = (torch.arange(n, device=x.device).unsqueeze(1) - factor * torch.arange(x.shape[1], device=x.device).unsqueeze(0)) % n
    return torch.gather(x, 0, rows)


def timelag_median_filter(rec):
    """7-tap median along the diagonals of a recurrence matrix (segment.py:77-84)."""
    t = rec.shape[0]
    lag = _shear(F.pad(rec, (0, 0, 0, t), mode="constant"), factor=-1)
    return _shear(median_filter1d(lag, k=7, s=1, p=3), factor=1)[:t]


def normalized_laplacian(a):
    """Symmetric normalised Laplacian of a dense weighted adjacency matrix, torch_geometric get_laplacian("sym") semantics."""
    a = a - torch.diag(torch.diagonal(a))                       # remove_self_loops
    deg = a.sum(dim=1)
    inv_sqrt = deg.pow(-0.5)
    inv_sqrt[torch.isinf(inv_sqrt)] = 0
    return torch.eye(a.shape[0], device=a.device, dtype=a.dtype) - inv_sqrt[:, None] * a * inv_sqrt[None, :]


def init_plus_plus(ds, k):
    """k-means++ seeding with the reference's fixed RandomState(42 + idx) draws (segment.py:87-104); ds: numpy [n, d]."""
    centroids = [ds[0]]
    for idx in range(1, k):
        dist_sq = np.array([min(np.inner(c - x, c - x) for c in centroids) for x in ds])
        cumulative = (dist_sq / (dist_sq.sum() + 1e-8)).cumsum()
        r = np.random.RandomState(42 + idx).rand()
        hit = np.nonzero(r < cumulative)[0]
        centroids.append(ds[hit[0] if len(hit) else len(cumulative) - 1])
    return np.array(centroids)


def soft_k_means(data, k, num_iter, cluster_temp=5):
    """Soft k-means on the unit sphere (segment.py:107-130) -> (centres, memberships [n, k], similarities)."""
    data = data / torch.norm(data, p=2, dim=1, keepdim=True).clamp_min(1e-30)
    mu = torch.tensor(init_plus_plus(data.cpu().detach().numpy(), k)).to(data)
    for _ in range(num_iter):
        r = torch.softmax(cluster_temp * (data @ mu.t()), 1)
        mu = (r.t() @ data) / r.sum(dim=0)[:, None]
    dist = data @ mu.t()
    return mu, torch.softmax(cluster_temp * dist, 1), dist


def laplacian_segmentation(envelope, beats, ks=(2, 4, 6, 8, 12, 16)):
    """envelope [T, C], beats: frame indices -> list of soft segmentations [T, k], one per k (segment.py:133-190)."""
    beats = list(beats)
    if envelope.dim() == 1:
        envelope = envelope[:, None]
    # guards for degenerate input (silent frames give NaN chroma / tonnetz rows; the reference's NaNs would abort eigh)
    envelope = torch.nan_to_num(envelope.float(), nan=0.0, posinf=0.0, neginf=0.0)
    bounds = zip([0] + beats, beats + [len(envelope)])
    csync = torch.stack([torch.median(envelope[a:b], dim=0).values for a, b in bounds], dim=0)

    rf = timelag_median_filter(recurrence_matrix(csync, width=3, sym=True))
    path_distance = torch.sum(torch.diff(csync, dim=0) ** 2, dim=1)
    sigma = torch.median(path_distance)
    if not sigma > 0:           # more than half of the beat-to-beat steps are exactly zero
        sigma = path_distance.mean().clamp_min(1e-12)
    path_sim = torch.exp(-path_distance / sigma)
    r_path = torch.diag(path_sim, diagonal=1) + torch.diag(path_sim, diagonal=-1)

    deg_path, deg_rec = r_path.sum(dim=1), rf.sum(dim=1)
    mu = deg_path.dot(deg_path + deg_rec) / torch.sum((deg_path + deg_rec) ** 2).clamp_min(1e-30)
    lap = torch.nan_to_num(normalized_laplacian(mu * rf + (1 - mu) * r_path))
    try:
        _, evecs = torch.linalg.eigh(lap)
    except Exception:
        evecs = torch.linalg.eig(lap)[1].real
    evecs = median_filter1d(evecs.T, k=9, s=1, p=4).T
    cnorm = torch.cumsum(evecs ** 2, dim=1) ** 0.5

    out = []
    for k in ks:
        _, member, _ = soft_k_means(evecs[:, :k] / cnorm[:, k - 1:k].clamp_min(1e-30), k=k, num_iter=100)
        out.append(F.interpolate(member.T[None], size=envelope.shape[0], mode="nearest").squeeze().T)
    return out


def segmentations_from_features(features, beats, ks=(2, 4, 6, 8, 12, 16)):
    """{(feature name, k): hard segment index per frame}: the loop of retrieve_music_information (mir.py:34-38)."""
    out = {}
    for name, feature in features.items():
        for k, seg in zip(ks, laplacian_segmentation(feature, beats, ks=ks)):
            out[(name, k)] = seg.argmax(1)
    return out
