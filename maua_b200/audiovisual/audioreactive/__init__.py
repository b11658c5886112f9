"""Audio-reactive feature / envelope / latent functions on the device (mirror of
maua.audiovisual.audioreactive and its torch-native twin selfsupervised.features.audio)."""
from .chroma import chroma_cens, chroma_cqt, chromagram, cqt_magnitude, estimate_tuning  # noqa: F401
from .features import (harmonic, mel_filterbank, mfcc, onset_peaks, onsets, onsets_rms, percussive, rms,  # noqa: F401
                       spectral_contrast, spectral_flatness)
from .latent import multi_weighted, select_modulo, single_weighted, slerp_loops, spline_loops, tempo_loops  # noqa: F401,E402
from . import noise  # noqa: F401,E402
from .signal import compress, expand, gaussian_filter, normalize, percentile, percentile_clip, resample  # noqa: F401,E402
from . import selfsupervised  # noqa: F401,E402  (torch-native twins: latent_patch, noise_patch, salience_weighted, ...)
from .audio import band_pass, high_pass, load_audio, low_pass  # noqa: F401,E402  (classic maua.audiovisual.audioreactive.audio)
from .mir_classic import pitch_dominance, pulse, spectral_max, tempo, tonnetz, volume  # noqa: F401,E402  (classic mir.py; its chroma() lives in mir_classic: `chroma` is the constant-Q module here)
from .util import info, plot_signals, plot_spectra  # noqa: F401,E402
