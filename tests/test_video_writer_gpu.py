"""The CUDA branch of VideoWriter / write_video and of the FFMPEG renderer's sink path (render/video.py, render/ffmpeg.py;
reference maua/ops/video.py:107-155, ops/io.py:47-70): the bytes that reach the sink against the reference's tensor2bytes
formula, through the pinned ring and the writer thread; and a generator loaded from a checkpoint file rendered on the device."""
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def reference_bytes(frame, value_range=(0, 1)):
    mn, mx = value_range
    return frame.squeeze(0).permute(1, 2, 0).clamp(mn, mx).sub(mn).div(mx - mn).mul(255).round().byte().numpy().tobytes()


def test_cuda_frames_reach_the_sink_as_tensor2bytes(cuda):
    from maua_b200.audiovisual.render.video import VideoWriter, write_video

    torch.manual_seed(0)
    video = torch.rand(9, 3, 48, 64) * 1.4 - 0.2
    # values exactly on a rounding boundary differ between round-half-even (torch) and the kernel's rintf: none here
    sink = io.BytesIO()
    write_video(video.to(cuda), "unused.mp4", fps=12, sink=sink)
    assert sink.getvalue() == b"".join(reference_bytes(f[None]) for f in video)
    sink = io.BytesIO()
    with VideoWriter("unused.mp4", (64, 48), fps=24, value_range=(-1, 1), sink=sink, ring_depth=2) as vw:
        vw.write((video[:4] * 2 - 1).to(cuda))
        vw.write((video[4] * 2 - 1).to(cuda))
        vw.write((video[5:] * 2 - 1).to(cuda))
    assert vw.frames_written == 9
    assert sink.getvalue() == b"".join(reference_bytes((f * 2 - 1)[None], (-1, 1)) for f in video)


def test_odd_sizes_are_resampled_on_the_device(cuda):
    from maua_b200.audiovisual.render.video import VideoWriter

    sink = io.BytesIO()
    with VideoWriter("unused.mp4", (31, 21), fps=24, sink=sink) as vw:
        vw.write(torch.rand(2, 3, 21, 31, device=cuda))
    assert len(sink.getvalue()) == 2 * 22 * 32 * 3


def test_ffmpeg_renderer_sink_bytes_match_the_frames(cuda):
    """FFMPEG.__call__ with a caller-supplied sink: every frame arrives once, in order, as rgb24 of clamp((x + 1) / 2)."""
    from maua_b200.audiovisual.render.ffmpeg import FFMPEG
    from maua_b200.GAN.networks import stylegan3 as N
    from maua_b200.GAN.wrappers.stylegan3 import StyleGAN3Synthesizer

    torch.manual_seed(0)
    net = N.SynthesisNetwork(w_dim=512, img_resolution=256, img_channels=3, channel_base=4096, channel_max=64)
    S = StyleGAN3Synthesizer.__new__(StyleGAN3Synthesizer)
    torch.nn.Module.__init__(S)
    S.G_synth, S._hook_handles = net, []
    lat = torch.randn(11, net.num_ws, 512)
    sink = io.BytesIO()
    seen = []
    r = FFMPEG(None, fps=24, batch_size=4, sink=sink)
    r(S, {"latents": lat}, lambda v: (seen.append(v.shape[0]), v)[1])
    assert seen == [4, 4, 3] and r.frames_written == 11
    raw = np.frombuffer(sink.getvalue(), dtype=np.uint8).reshape(11, 256, 256, 3)
    want = net(lat.to(cuda), out_fmt="u8").cpu().numpy()
    assert int(np.abs(raw.astype(np.int16) - want.astype(np.int16)).max()) <= 1      # (x+1)/2 in fp32 then *255 vs fused
    # a postprocess declared pure that hands batch 0 back untouched: the later batches leave the network as rgb24 (postprocess
    # is not called again) and the byte stream is the same
    sink3, seen3 = io.BytesIO(), []
    pure = lambda v: (seen3.append(v.shape[0]), v)[1]  # noqa: E731
    pure.pure = True
    FFMPEG(None, fps=24, batch_size=4, sink=sink3)(S, {"latents": lat}, pure)
    assert seen3 == [4]
    raw3 = np.frombuffer(sink3.getvalue(), dtype=np.uint8).reshape(11, 256, 256, 3)
    assert int(np.abs(raw3.astype(np.int16) - raw.astype(np.int16)).max()) <= 1 and np.array_equal(raw3[4:], want[4:])
    # a postprocess that changes the frame size (force_output_size) flows through the same conversion
    sink2 = io.BytesIO()
    FFMPEG(None, fps=24, batch_size=4, sink=sink2)(S, {"latents": lat[:4]}, lambda v: v[:, :, ::2, ::2].contiguous())
    assert len(sink2.getvalue()) == 4 * 128 * 128 * 3


def test_loaded_checkpoint_renders_like_the_source_network(cuda, tmp_path):
    """N4: a generator written to disk in the NVIDIA .pt layout, read back by load_network, rendered on the device: the same
    pixels as the network it came from (StyleGAN3), and for StyleGAN2 in the training layout the per-layer noise strength
    reaches the kernels."""
    from maua_b200.GAN import load as L
    from maua_b200.GAN.networks import stylegan2 as N2, stylegan3 as N3

    torch.manual_seed(3)
    G = N3.Generator(512, 0, 512, 256, 3, mapping_kwargs=dict(num_layers=2), channel_base=4096, channel_max=64)
    path = tmp_path / "sg3.pt"
    torch.save({"G_ema": G.state_dict()}, path)
    G2 = L.load_network(str(path))
    z = torch.randn(2, 512)
    ws = G.mapping(z)
    assert torch.equal(ws, G2.mapping(z))
    assert torch.equal(G.synthesis(ws.to(cuda)), G2.synthesis(ws.to(cuda)))

    torch.manual_seed(4)
    S = N2.Generator(512, 0, 512, 32, 3, mapping_kwargs=dict(num_layers=2), channel_base=1024, channel_max=32)
    sd = {}
    for k, v in S.state_dict().items():
        if k.startswith("synthesis.bs."):
            parts = k.split(".")
            k = ".".join(["synthesis", f"b{4 * 2 ** int(parts[2])}"] + parts[3:])
        sd[k.replace("mapping.fcs.", "mapping.fc")] = v.clone()
    for k in [k for k in sd if k.endswith("noise_const")]:
        sd[k[: -len("noise_const")] + "noise_strength"] = torch.tensor(0.0)       # trained strength 0: the noise must vanish
    p2 = tmp_path / "sg2.pt"
    torch.save({"G_ema": sd}, p2)
    S2 = L.load_network(str(p2))
    w = torch.randn(2, S.synthesis.num_ws, 512, device=cuda)
    with_noise = S.synthesis(w).clone()
    silent = S2.synthesis(w).clone()
    assert not torch.equal(with_noise, silent)
    for m in S.synthesis.modules():
        if isinstance(m, N2.SynthesisLayer):
            m.noise_const.zero_()
    assert torch.equal(S.synthesis(w), silent)
