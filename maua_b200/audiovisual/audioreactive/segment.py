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


# ---- the ("rosa", k) segmentations of retrieve_music_information -------------------------------------------------------
BINS_PER_OCTAVE = 12 * 3
N_OCTAVES = 7


def _amplitude_to_db_max(mag, amin=1e-5, top_db=80.0):
    """librosa.amplitude_to_db(S, ref=np.max): 20 log10(max(S, amin) / max(S.max(), amin)), floored top_db below the maximum."""
    ref = mag.max().clamp_min(amin)
    db = 20.0 * torch.log10(mag.clamp_min(amin)) - 20.0 * torch.log10(ref)
    return db.clamp_min(db.max() - top_db)


def _median_filter_rows(x, k=9):
    """scipy.ndimage.median_filter(x, size=(k, 1)) with its default "reflect" boundary (the edge sample is repeated:
    d c b a | a b c d), unlike F.pad's reflect which median_filter1d above uses for the reference's own torch functions."""
    p = k // 2
    xp = torch.cat([x[:p].flip(0), x, x[-p:].flip(0)], dim=0)
    return xp.unfold(0, k, 1).median(dim=-1).values


def hard_k_means(data, k, num_iter=100):
    """Lloyd's algorithm with the seeded k-means++ start of init_plus_plus -> cluster index per row (the reference calls an
    unseeded sklearn.cluster.KMeans(n_clusters=k).fit_predict here, segment.py:255: its labels vary from run to run)."""
    mu = torch.tensor(init_plus_plus(data.detach().cpu().double().numpy(), k)).to(data)
    labels = None
    for _ in range(num_iter):
        new = torch.cdist(data, mu).argmin(1)
        if labels is not None and torch.equal(new, labels):
            break
        labels = new
        for j in range(k):
            sel = labels == j
            if sel.any():
                mu[j] = data[sel].mean(0)
    return labels


def laplacian_segmentation_rosa(audio, sr, out_size, ks=(2, 4, 6, 8, 16), beats=None):
    """Pattern-recurrence segmentation of the TRACK (not of a feature envelope): mirror of laplacian_segmentation_rosa,
    rosa/segment.py:201-258, which follows librosa's "Laplacian segmentation" example -> int64 [out_size, len(ks)].

    The reference runs this one through librosa / scipy / sklearn on the host (librosa.cqt, beat_track, util.sync,
    segment.recurrence_matrix, feature.mfcc, csgraph.laplacian, KMeans).  librosa is absent and un-pinned, so the stages are
    the package's own device functions for the same quantities -- PARITY UNPINNED: 7-octave, 36-bins-per-octave constant-Q
    magnitudes in dB below the maximum (csrc/chroma.cu), beats from the package's onset envelope and beat tracker with
    librosa's default tempo prior (the reference lets librosa derive its own onset envelope), beat-synchronous medians
    (constant-Q) and means (20 MFCCs), mutual k-NN recurrence affinity + time-lag median filter, balanced combination with
    the MFCC path affinity, symmetric normalised Laplacian, 9-row median filter of the eigenvectors, cumulative
    normalisation and k-means (seeded here, unseeded there).  `beats` overrides the internal beat tracker."""
    from . import beat as _beat
    from .chroma import cqt_magnitude
    from .features import mfcc as _mfcc, onsets as _onsets

    audio = torch.as_tensor(audio)
    if not audio.is_cuda:
        raise RuntimeError("laplacian_segmentation_rosa: the track must live on the GPU (no CPU path)")
    audio = audio.float()
    cdb = _amplitude_to_db_max(cqt_magnitude(audio, sr, hop_length=1024, n_bins=N_OCTAVES * BINS_PER_OCTAVE,
                                             bins_per_octave=BINS_PER_OCTAVE)).T.contiguous()          # [T, 252]
    n_frames = cdb.shape[0]
    if beats is None:
        env = _onsets(audio, sr).squeeze().cpu().numpy()
        bpm = _beat.tempo(env, hop_length=1024)                      # librosa.beat.beat_track defaults: start_bpm 120, std 1
        beats = [int(b) for b in _beat.beat_track(env, bpm, hop_length=1024, trim=False)]
    beats = sorted({int(b) for b in beats if 0 < int(b) < n_frames})   # util.sync pads the boundaries with 0 and the end
    edges = [0] + beats + [n_frames]
    csync = torch.stack([cdb[a:b].median(dim=0).values for a, b in zip(edges[:-1], edges[1:])], dim=0)
    mf = _mfcc(audio, sr, n_mfcc=20)[:n_frames]
    msync = torch.stack([mf[a:b].mean(dim=0) for a, b in zip(edges[:-1], edges[1:])], dim=0)
    nb = csync.shape[0]
    if nb < 4:   # nothing to segment: one label everywhere
        return torch.zeros(out_size, len(ks), dtype=torch.long, device=audio.device)

    rf = timelag_median_filter(recurrence_matrix(csync, width=min(3, max((nb - 2) // 2, 1)), sym=True))
    path_distance = torch.sum(torch.diff(msync, dim=0) ** 2, dim=1)
    sigma = torch.median(path_distance)
    if not sigma > 0:
        sigma = path_distance.mean().clamp_min(1e-12)
    path_sim = torch.exp(-path_distance / sigma)
    r_path = torch.diag(path_sim, diagonal=1) + torch.diag(path_sim, diagonal=-1)
    deg_path, deg_rec = r_path.sum(dim=1), rf.sum(dim=1)
    mu = deg_path.dot(deg_path + deg_rec) / torch.sum((deg_path + deg_rec) ** 2).clamp_min(1e-30)
    lap = torch.nan_to_num(normalized_laplacian(mu * rf + (1 - mu) * r_path))
    _, evecs = torch.linalg.eigh(lap.double())
    evecs = _median_filter_rows(evecs.float(), k=min(9, 2 * ((nb - 1) // 2) + 1))
    cnorm = torch.cumsum(evecs ** 2, dim=1) ** 0.5

    out = []
    for k in ks:
        kk = min(int(k), nb)
        x = evecs[:, :kk] / cnorm[:, kk - 1:kk].clamp_min(1e-30)
        labels = hard_k_means(x, kk).float()
        out.append(F.interpolate(labels[None, None, :], size=out_size, mode="nearest").squeeze(0).squeeze(0))
    return torch.stack(out, dim=1).long()
