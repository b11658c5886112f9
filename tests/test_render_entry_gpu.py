"""End-to-end entry point (generate_audiovisal_from_patch) with a user patch file and the MemMap / FFMPEG sinks."""
import os
import wave

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _write_wav(path, seconds):
    from maua_b200.workload import sine_sweep

    y, sr = sine_sweep(seconds, tremolo_hz=4.0)
    with wave.open(path, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(sr)
        w.writeframes((np.clip(y, -1, 1) * 32767).astype("<i2").tobytes())


def test_generate_from_patch_memmap_and_raw_sink(cuda, tmp_path):
    from maua_b200.audiovisual.generate import generate_audiovisal_from_patch
    from maua_b200.audiovisual.render.ffmpeg import FFMPEG

    wav = str(tmp_path / "sweep.wav")
    _write_wav(wav, 2.0)
    torch.manual_seed(0)
    video, (audio, sr) = generate_audiovisal_from_patch(
        audio_file=wav, model_file=None, patch_file="tests/patches/sweep_patch.py", patch_name="SweepPatch", renderer="memmap",
        renderer_kwargs=dict(cache_file=str(tmp_path / "frames.npy"), batch_size=4), fps=8, out_size=(1024, 1024),
        resize_strategy="pad-zero", resize_layer=0)
    assert video.shape == (16, 3, 1024, 1024) and video.dtype == np.uint8 and sr == 48000
    assert os.path.exists(tmp_path / "frames.npy")
    assert 20 < float(video.mean()) < 235 and float(video[0].std()) > 5

    # the raw rgb24 sink (ffmpeg binary absent in this image) receives H*W*3 bytes per frame
    from tests.patches.sweep_patch import SweepPatch
    torch.manual_seed(0)
    patch = SweepPatch(None, wav, fps=8)
    patch.process_audio()
    inputs = patch.process_synthesizer_inputs(None)
    out = str(tmp_path / "out.mp4")
    FFMPEG(out, fps=8, batch_size=8)(patch.synthesizer, inputs, lambda v: v)
    raw = out + ".rgb24"
    if os.path.exists(raw):
        assert os.path.getsize(raw) == 16 * 1024 * 1024 * 3
        first = np.fromfile(raw, dtype=np.uint8, count=1024 * 1024 * 3).reshape(1024, 1024, 3)
        # same generator seed -> same frames as the memmap render (round vs truncate: off by at most 1)
        assert int(np.abs(first.astype(np.int16) - video[0].transpose(1, 2, 0).astype(np.int16)).max()) <= 1


def test_generate_with_output_size(cuda, tmp_path):
    """out_size / resize_strategy / resize_layer of generate_audiovisal_from_patch reach the synthesizer's output-size hook
    (maua/audiovisual/generate.py:27-32 -> wrappers/stylegan3.py:62-79): a 1536x1024 (W x H) render stretched at layer 0."""
    from maua_b200.audiovisual.generate import generate_audiovisal_from_patch

    wav = str(tmp_path / "sweep.wav")
    _write_wav(wav, 2.0)
    torch.manual_seed(0)
    video, _ = generate_audiovisal_from_patch(
        audio_file=wav, model_file=None, patch_file="tests/patches/sweep_patch.py", patch_name="SweepPatch", renderer="memmap",
        renderer_kwargs=dict(cache_file=str(tmp_path / "wide.npy"), batch_size=4), fps=8, out_size=(1536, 1024),
        resize_strategy="stretch", resize_layer=0)
    assert video.shape == (16, 3, 1024, 1536) and video.dtype == np.uint8
    assert 20 < float(video.mean()) < 235 and float(video[0].std()) > 5


def test_async_frame_downloader_round_trip(cuda):
    """Frames written by consecutive batches arrive intact and in order through the side-stream pinned ring."""
    import torch

    from maua_b200.audiovisual.render._loop import AsyncFrameDownloader

    dl = AsyncFrameDownloader((2, 8, 8, 3), cuda, depth=2)
    got = []
    for i in range(5):
        buf = dl.device_buffer(i)
        buf.fill_(i + 1)
        dl.download(i)
        if i >= 1:
            got.append(int(dl.host(i - 1)[0, 0, 0, 0]))
    got.append(int(dl.host(4)[-1, -1, -1, -1]))
    dl.synchronize()
    assert got == [1, 2, 3, 4, 5]
