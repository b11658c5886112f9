"""Where the end-to-end render loses time against the device-timed steps: e2e_probe.py [frames]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maua_b200.GAN.wrappers import get_generator_class
from maua_b200.audiovisual.render.ffmpeg import FFMPEG
from maua_b200.audiovisual.render import _loop

n = int(sys.argv[1]) if len(sys.argv) > 1 else 720
B = 16
dev = torch.device("cuda")
torch.manual_seed(0)
G = get_generator_class("stylegan3")(model_file=None).to(dev)
net = G.synthesizer.G_synth
lat = torch.randn(n, 16, 512)
lat_pinned = lat.pin_memory()


class Sink:
    n = 0
    def write(self, v): self.n += len(v)
    def close(self): pass


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best


out = torch.empty(B, 1024, 1024, 3, device=dev, dtype=torch.uint8)
lat_dev = lat.to(dev)
def bare():
    for i in range(0, n, B):
        net(lat_dev[i:i + B], out_fmt="u8", out=out)
def bare_h2d():
    for i in range(0, n, B):
        net(lat_pinned[i:i + B].to(dev, non_blocking=True), out_fmt="u8", out=out)
def bare_unit():
    for i in range(0, n, B):
        f = net(lat_dev[i:i + B], out_fmt="f32_unit")
        _loop.frames_to_rgb24(f, out=out)
def ffmpeg_pinned():
    FFMPEG(None, batch_size=B, sink=Sink())(G.synthesizer, {"latents": lat_pinned}, lambda v: v)
def ffmpeg_pageable():
    FFMPEG(None, batch_size=B, sink=Sink())(G.synthesizer, {"latents": lat}, lambda v: v)
def ffmpeg_pure():
    pp = lambda v: v
    pp.pure = True
    FFMPEG(None, batch_size=B, sink=Sink())(G.synthesizer, {"latents": lat_pinned}, pp)
def ffmpeg_pure_half():
    pp = lambda v: v
    pp.pure = True
    FFMPEG(None, batch_size=B, sink=Sink())(G.synthesizer, {"latents": lat_pinned[:n // 2]}, pp)
def wrapper_only():
    for i in range(0, n, B):
        G.synthesizer(latents=lat_dev[i:i + B], out_fmt="f32_unit")

for name, fn in [("net() u8, device latents", bare), ("net() u8, pinned latents + H2D", bare_h2d), ("net() f32_unit + rgb24 kernel", bare_unit),
                 ("wrapper forward f32_unit", wrapper_only), ("FFMPEG.__call__ pinned inputs", ffmpeg_pinned), ("FFMPEG.__call__ pinned, pure postprocess", ffmpeg_pure), ("  same, half the frames (ms per FULL step count)", ffmpeg_pure_half), ("FFMPEG.__call__ pageable inputs", ffmpeg_pageable)]:
    t = timed(fn)
    print(f"{name:40s} {1000 * t / (n / B):7.3f} ms/step  {n / t:7.1f} frames/s")
