"""Frame sink worker: the step right after the render path.

The reference pushes every finished frame through a Queue to one ``WriteWorker`` thread that converts it with
``tensor2bytes`` and writes it to ffmpeg's stdin (maua/ops/video.py:15-128).  Here the conversion already happened on the
device; what is left is the device->host copy and the pipe write, which run off the render thread: frames land in a ring
of pinned host buffers by asynchronous copies, and the worker waits for each copy's CUDA event and writes the bytes while
the GPU renders the next batches.  The render thread only blocks when every ring slot is still waiting to be written
(back-pressure instead of the reference's unbounded queue, ops/video.py:113-115)."""
from __future__ import annotations

import queue
import threading

_RING_CACHE = {}            # (shape, depth) -> [list of pinned uint8 buffers, in_use]
_RING_LOCK = threading.Lock()


def acquire_pinned_ring(shape, depth):
    """A ring of `depth` pinned uint8 host buffers of `shape`, reused between renders: cudaHostAlloc of 3 x 50 MB costs tens of
    milliseconds, more than two batches of a 1024^2 render.  Returns (ring, release); a ring that is still in use by another
    render is never handed out twice (a fresh one is allocated instead and not cached)."""
    import torch

    key = (tuple(shape), int(depth))
    with _RING_LOCK:
        ent = _RING_CACHE.get(key)
        if ent is not None and not ent[1]:
            ent[1] = True
            return ent[0], lambda: ent.__setitem__(1, False)
    ring = [torch.empty(shape, dtype=torch.uint8).pin_memory() for _ in range(depth)]
    with _RING_LOCK:
        if key not in _RING_CACHE:
            ent = [ring, True]
            _RING_CACHE[key] = ent
            return ring, lambda: ent.__setitem__(1, False)
    return ring, lambda: None


class RingWriter:
    """ring: list of writable host buffers (pinned uint8 tensors or numpy arrays, indexable by [:n]);
    sink: object with write(bytes-like)."""

    def __init__(self, sink, ring):
        self.sink, self.ring = sink, ring
        self.free = queue.Queue()
        for k in range(len(ring)):
            self.free.put(k)
        self.work = queue.Queue()
        self.error = None
        self.bytes_written = 0
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        while True:
            item = self.work.get()
            if item is None:
                return
            k, n, event = item
            try:
                if self.error is None:
                    if event is not None:
                        event.synchronize()  # the asynchronous device->host copy into slot k has landed
                    buf = self.ring[k][:n]
                    buf = buf.numpy() if hasattr(buf, "numpy") else buf
                    view = memoryview(buf).cast("B")
                    self.sink.write(view)
                    self.bytes_written += view.nbytes
            except BaseException as e:  # surfaced on the render thread by acquire() / close()
                self.error = e
            finally:
                self.free.put(k)

    def acquire(self):
        """Index of a ring slot that is safe to overwrite (blocks while the writer is behind)."""
        if self.error is not None:
            raise self.error
        return self.free.get()

    def submit(self, k, n, event=None):
        """Slot k holds n frames once `event` (a CUDA event recorded after the copy, or None) has completed."""
        self.work.put((k, n, event))

    def close(self):
        self.work.put(None)
        self.thread.join()
        if self.error is not None:
            raise self.error
