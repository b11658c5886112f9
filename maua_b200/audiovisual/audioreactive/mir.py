"""retrieve_music_information: mirror of maua/audiovisual/audioreactive/selfsupervised/mir.py:24-45.

features (device kernels, selfsupervised.extract_features) -> tempo and beats from the onset envelope (beat.py, host) ->
Laplacian segmentations per feature and k (segment.py) -> post-processed features, plus the ("rosa", k) segmentations of the
whole track (segment.laplacian_segmentation_rosa: the reference's librosa / sklearn pipeline on the package's own device
functions; parity unpinned, see its docstring).
"""
from __future__ import annotations

from . import beat as _beat
from . import segment as _segment
from . import selfsupervised as _ss

AFEATS = _ss.ALLFEATS
UNITFEATS = _ss.UNITFEATS


def retrieve_music_information(audio, sr, ks=(2, 4, 6, 8, 12, 16), device="cuda"):
    """audio: float32 [n] (n a multiple of 1024: resample to sr = 1024 * fps first, sample.py:29-30) ->
    (features {name: [T, C] in [0, 1]}, segmentations {(name, k): int64 [T]}, tempo)."""
    audio = audio.to(device)
    raw = _ss.extract_features(audio, sr, postprocess=False)
    tempo, beats = _beat.tempo_and_beats(raw["onsets"].squeeze().cpu().numpy())
    segmentations = _segment.segmentations_from_features(raw, beats, ks=ks)
    n_frames = next(iter(raw.values())).shape[0]
    rosa = _segment.laplacian_segmentation_rosa(audio, sr, n_frames, ks=ks)          # mir.py:40-41
    for i, k in enumerate(ks):
        segmentations[("rosa", k)] = rosa[:, i]
    features = {k: _ss.postprocess_feature(v) for k, v in raw.items()}
    return features, segmentations, tempo
