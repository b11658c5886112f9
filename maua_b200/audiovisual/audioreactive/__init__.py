"""Audio-reactive feature / envelope / latent functions on the device (mirror of
maua.audiovisual.audioreactive and its torch-native twin selfsupervised.features.audio)."""
from .features import mel_filterbank, onset_peaks, onsets, onsets_rms, percussive, rms  # noqa: F401
