"""The ``maua.*`` import surface: every module path and name the reference's entry points, base patches and
render loop use resolves, with the reference's signatures.  Host-only: nothing here computes."""
import inspect

import pytest


def params(fn):
    return list(inspect.signature(fn).parameters)


def test_reference_module_paths_resolve():
    from maua.audiovisual import audioreactive as ar
    from maua.audiovisual.generate import generate_audiovisal_from_patch, main  # noqa: F401
    from maua.audiovisual.patches.base import MauaPatch, get_patch_from_file  # noqa: F401
    from maua.audiovisual.patches.base.stylegan2 import StyleGAN2Patch
    from maua.audiovisual.patches.base.stylegan3 import StyleGAN3Patch
    from maua.audiovisual.render import get_output_class
    from maua.audiovisual.render.ffmpeg import FFMPEG  # noqa: F401
    from maua.audiovisual.render.memmap import MemMap  # noqa: F401
    from maua.GAN.load import load_network  # noqa: F401
    from maua.GAN.wrappers import MauaGenerator, MauaMapper, MauaSynthesizer, get_generator_class  # noqa: F401
    from maua.GAN.wrappers.stylegan import StyleGAN, StyleGANMapper  # noqa: F401
    from maua.GAN.wrappers.stylegan2 import StyleGAN2, StyleGAN2Mapper, StyleGAN2Synthesizer  # noqa: F401
    from maua.GAN.wrappers.stylegan3 import StyleGAN3, StyleGAN3Synthesizer  # noqa: F401
    from maua.GAN.wrappers.inference.stylegan2 import Generator, SynthesisNetwork  # noqa: F401
    from maua.ops.image import resample  # noqa: F401
    from maua.ops.video import VideoWriter, write_video  # noqa: F401

    import maua_b200.audiovisual.patches.base as real

    # one class hierarchy: a patch written against maua.* is a MauaPatch of this build (get_patch_from_file checks issubclass)
    assert issubclass(StyleGAN3Patch, real.MauaPatch) and issubclass(StyleGAN2Patch, real.MauaPatch) and MauaPatch is real.MauaPatch
    assert get_generator_class("stylegan3") is StyleGAN3 and get_generator_class("stylegan2") is StyleGAN2
    assert get_output_class("ffmpeg") is FFMPEG and get_output_class("memmap") is MemMap
    # the classic ar namespace = audio.py + latent.py + mir.py + signal.py + util.py (audioreactive/__init__.py:30-34)
    for name in ["load_audio", "harmonic", "percussive", "low_pass", "high_pass", "band_pass", "onsets", "volume", "chroma", "tonnetz",
                 "pulse", "tempo", "spectral_max", "pitch_dominance", "resample", "normalize", "percentile", "percentile_clip",
                 "compress", "expand", "gaussian_filter", "single_weighted", "multi_weighted", "select_modulo", "slerp_loops",
                 "spline_loops", "tempo_loops", "info", "plot_signals", "plot_spectra"]:
        assert callable(getattr(ar, name)), name


def test_classic_signatures_match_the_reference():
    from maua.audiovisual import audioreactive as ar

    assert params(ar.load_audio) == ["audio_file", "offset", "duration", "cache"]                    # audio.py:15
    assert params(ar.harmonic) == ["audio", "sr", "margin"] and params(ar.percussive) == ["audio", "sr", "margin"]   # :84,90
    assert params(ar.low_pass) == ["audio", "sr", "fmax", "db_per_octave"]                          # :96
    assert params(ar.high_pass) == ["audio", "sr", "fmin", "db_per_octave"]                         # :102
    assert params(ar.band_pass) == ["audio", "sr", "fmin", "fmax", "db_per_octave"]                 # :108
    assert params(ar.onsets) == ["audio", "sr", "type", "prepercussive"]                            # mir.py:17
    assert params(ar.volume) == ["audio", "sr"]                                                     # :65
    assert params(ar.chroma) == ["audio", "sr", "type", "nearest_neighbor", "preharmonic", "notes"]  # :81
    assert params(ar.tonnetz) == ["audio", "sr", "type", "nearest_neighbor", "preharmonic"]         # :126
    assert params(ar.pulse) == ["audio", "sr", "prior", "type", "prepercussive"]                    # :163
    assert inspect.signature(ar.low_pass).parameters["fmax"].default == 200
    assert inspect.signature(ar.band_pass).parameters["db_per_octave"].default == 12


def test_torch_native_twin_paths():
    from maua.audiovisual.audioreactive.selfsupervised.features import audio as fa
    from maua.audiovisual.audioreactive.selfsupervised.features.efficient_quantile import quantile
    from maua.audiovisual.audioreactive.selfsupervised.features.processing import (clamp_peaks_percentile, emphasize,  # noqa: F401
                                                                                   gaussian_filter, normalize, onset_envelope,
                                                                                   spectral_flux, standardize)
    from maua.audiovisual.audioreactive.selfsupervised.latent import latent_patch, spline_loop_latents  # noqa: F401
    from maua.audiovisual.audioreactive.selfsupervised.mir import retrieve_music_information, salience_weighted  # noqa: F401
    from maua.audiovisual.audioreactive.selfsupervised.noise import Blend, Loop, Multiply, noise_patch  # noqa: F401
    from maua.audiovisual.audioreactive.selfsupervised.patch import Patch  # noqa: F401
    from maua.audiovisual.audioreactive.selfsupervised.sample import generate, load_audio  # noqa: F401

    assert params(fa.onsets) == ["audio", "sr"] and params(fa.harmonic) == ["audio", "margin"]       # features/audio.py:13,27
    assert params(fa.pulse) == ["audio", "sr"] and params(quantile) == ["tensor", "q"]
    for name in ["rms", "drop_strength", "chromagram", "tonnetz", "mfcc", "spectral_contrast", "spectral_flatness"]:
        assert callable(getattr(fa, name)), name


def test_host_tensors_need_a_gpu_not_a_fallback():
    import torch

    from maua.audiovisual import audioreactive as ar

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by tests/test_maua_alias_gpu.py")
    with pytest.raises(RuntimeError, match="CUDA"):
        ar.gaussian_filter(torch.zeros(16, 1), 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        ar.low_pass(torch.zeros(4096).numpy(), 48000)


def test_cli_arguments():
    from maua.audiovisual.generate import main

    with pytest.raises(SystemExit):
        main(["--help"])
