"""The frame sink worker (render/_sink.py; reference: maua/ops/video.py:15-128 WriteWorker/VideoWriter): order, byte
count, back-pressure and error propagation, with numpy buffers standing in for the pinned ring (no GPU needed)."""
import io
import time

import numpy as np
import pytest

from maua_b200.audiovisual.render._sink import RingWriter


def test_frames_arrive_in_order_and_complete():
    sink = io.BytesIO()
    ring = [np.zeros((4, 2, 3, 3), dtype=np.uint8) for _ in range(3)]
    w = RingWriter(sink, ring)
    rng = np.random.RandomState(0)
    want = b""
    for i in range(11):
        n = 4 if i < 10 else 1  # ragged last batch
        frames = rng.randint(0, 256, size=(n, 2, 3, 3)).astype(np.uint8)
        k = w.acquire()
        ring[k][:n] = frames
        w.submit(k, n)
        want += frames.tobytes()
    w.close()
    assert sink.getvalue() == want and w.bytes_written == len(want)


class SlowSink:
    def __init__(self):
        self.n = 0

    def write(self, b):
        time.sleep(0.02)
        self.n += 1


def test_back_pressure_bounds_the_ring():
    sink = SlowSink()
    ring = [np.zeros((1, 8), dtype=np.uint8) for _ in range(2)]
    w = RingWriter(sink, ring)
    t0 = time.time()
    for _ in range(10):
        k = w.acquire()
        w.submit(k, 1)
    waited = time.time() - t0
    w.close()
    assert sink.n == 10
    assert waited >= 0.02 * 7  # the producer had to wait for slots: at most two batches are ever in flight


class BrokenSink:
    def write(self, b):
        raise BrokenPipeError("ffmpeg went away")


def test_sink_errors_surface_on_the_render_thread():
    w = RingWriter(BrokenSink(), [np.zeros((1, 8), dtype=np.uint8) for _ in range(2)])
    k = w.acquire()
    w.submit(k, 1)
    with pytest.raises(BrokenPipeError):
        for _ in range(5):
            k = w.acquire()
            w.submit(k, 1)
            time.sleep(0.01)
        w.close()
