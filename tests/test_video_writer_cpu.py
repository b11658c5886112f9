"""VideoWriter / write_video (render/video.py; reference maua/ops/video.py:107-155 + tensor2bytes ops/io.py:47-70): the
byte stream against the reference's conversion formula, frame order, value ranges.  Host tensors and an in-memory sink."""
import io

import numpy as np
import pytest
import torch

from maua_b200.audiovisual.render.video import VideoWriter, tensor2bytes_device, write_video


def reference_bytes(frame, value_range=(0, 1)):
    """tensor2bytes, ops/io.py:56-70, for one [1, C, H, W] frame."""
    mn, mx = value_range
    return frame.squeeze(0).permute(1, 2, 0).clamp(mn, mx).sub(mn).div(mx - mn).mul(255).round().byte().numpy().tobytes()


def test_stream_equals_reference_conversion_frame_by_frame():
    torch.manual_seed(0)
    video = torch.rand(7, 3, 6, 8) * 1.4 - 0.2        # values outside [0, 1] get clamped
    sink = io.BytesIO()
    write_video(video, "unused.mp4", fps=12, sink=sink)
    assert sink.getvalue() == b"".join(reference_bytes(f[None]) for f in video)


def test_value_range_and_batched_writes():
    torch.manual_seed(1)
    frames = torch.rand(5, 3, 4, 4) * 2 - 1
    sink = io.BytesIO()
    with VideoWriter("unused.mp4", (4, 4), fps=24, value_range=(-1, 1), sink=sink) as vw:
        vw.write(frames[:2])
        vw.write(frames[2])            # [C, H, W] is promoted to one frame
        vw.write(frames[3:])
    assert vw.frames_written == 5
    assert sink.getvalue() == b"".join(reference_bytes(f[None], (-1, 1)) for f in frames)
    assert tensor2bytes_device(frames, (-1, 1)).shape == (5, 4, 4, 3)


def test_numpy_input_and_odd_sizes():
    sink = io.BytesIO()
    arr = np.random.RandomState(0).rand(3, 3, 2, 2).astype(np.float32)
    write_video(arr, "unused.mp4", sink=sink)
    assert len(sink.getvalue()) == 3 * 2 * 2 * 3
    with pytest.raises(RuntimeError):
        with VideoWriter("unused.mp4", (3, 3), fps=24, sink=io.BytesIO()) as vw:
            vw.write(torch.rand(1, 3, 3, 3))      # odd host frames: no CPU resampler


def test_raw_file_sink_without_ffmpeg(tmp_path):
    import shutil

    if shutil.which("ffmpeg") is not None:
        pytest.skip("ffmpeg present: the encoder owns the output")
    out = str(tmp_path / "clip.mp4")
    write_video(torch.zeros(2, 3, 4, 6), out, fps=8)
    assert (tmp_path / "clip.mp4.rgb24").stat().st_size == 2 * 4 * 6 * 3
