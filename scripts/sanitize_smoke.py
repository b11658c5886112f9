"""Small invocations of every kernel family, for compute-sanitizer (scripts/sanitize.sh): sizes are chosen to hit the edge
paths (partial 16-channel groups, columns split into segments, image edges that are not tile multiples, odd audio lengths)
while staying fast under the sanitizer's ~30x slow-down."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from maua_b200 import ops
from maua_b200.GAN.networks import stylegan2 as N2, stylegan3 as N3
from maua_b200.audiovisual import audioreactive as ar
from maua_b200.audiovisual.audioreactive import selfsupervised as ss
from maua_b200.audiovisual.render._loop import frames_to_rgb24
from oracle import sg3 as O

dev = torch.device("cuda")
torch.manual_seed(0)
for cfg, kw in (("T", dict(channel_base=2048, channel_max=40)), ("R", dict(channel_base=4096, channel_max=72, **{k: v for k, v in N3.SG3_R_KWARGS.items() if k not in ("channel_base", "channel_max")}))):
    net = N3.SynthesisNetwork(w_dim=512, img_resolution=128, img_channels=3, **kw)
    for B in (1, 3):
        out = net(torch.randn(B, net.num_ws, 512, device=dev), out_fmt="u8")
    print("sg3", cfg, tuple(out.shape), float(out.float().mean()))
net2 = N2.SynthesisNetwork(w_dim=512, img_resolution=64, img_channels=3, channel_base=2048, channel_max=48)
print("sg2", float(net2(torch.randn(2, net2.num_ws, 512, device=dev)).mean()))
fu, fd = O.design_lowpass_filter(12, 8.0, 9.0, 64.0), O.design_lowpass_filter(12, 8.0, 9.0, 64.0)
for layout in ("0", "1"):
    os.environ["MB_FLRELU_TEST_NHWC"] = layout
    for C, H, W in ((19, 70, 45), (35, 33, 100), (3, 150, 40)):
        x = torch.randn(2, C, H, W, device=dev)
        y = ops.filtered_lrelu(x, fu.to(dev), fd.to(dev), torch.randn(C, device=dev), up=2, down=2, padding=[9, 8, 9, 8], clamp=256)
    print("filtered_lrelu layout", layout, tuple(y.shape))
x = torch.randn(1, 40, 37, 41, device=dev)
for impl in (0, 8, 10, 11, 12):   # product dispatch, epilogue groups, 34-pixel patch, stacked pixel-major, stacked cout-major tiles
    print("conv impl", impl, tuple(ops.modulated_conv2d(x, torch.randn(17, 40, 3, 3, device=dev), torch.randn(1, 40, device=dev), impl=impl).shape))
# StyleGAN2 output-size hooks (mb_sg2_set_resize): bicubic stretch on conv1 of the 16^2 block, reflect padding on conv0 of the 32^2 block
from collections import OrderedDict
from maua_b200.GAN.wrappers.stylegan2 import StyleGAN2Synthesizer
S = StyleGAN2Synthesizer.__new__(StyleGAN2Synthesizer)
torch.nn.Module.__init__(S)
S.G_synth, S._hook_handles, S._warp_hooks = net2, [], OrderedDict()
S.layer_names = [f"bs.{c // 2}.conv{1 if r == 4 else c % 2}" for c, r in enumerate(sorted(net2.block_resolutions * 2))]
for size, strategy, layer in (((96, 80), "stretch", 5), ((80, 72), "pad-reflect-out", 6)):
    S.change_output_resolution(size, strategy, layer)
    print("sg2 resize", strategy, tuple(S.forward(torch.randn(2, net2.num_ws, 512, device=dev)).shape))
S.change_output_resolution((64, 64), "stretch", 0)
# RRDBNet (csrc/rrdb.cu): channels-last epilogue of the stacked tiles, both conv variants
from maua_b200.super.image.models.realesrgan import RRDBNet
for cms in ("0", "1"):
    os.environ["MB_RRDB_CMS"] = cms
    up = RRDBNet(num_block=1)
    print("rrdb cms", cms, tuple(up(torch.rand(1, 3, 23, 37, device=dev), out_fmt="u8").shape))
f = ops.setup_filter().to(dev)
print("upfirdn2d", tuple(ops.upfirdn2d(torch.randn(2, 3, 9, 13, device=dev), f, up=2, padding=(2, 1, 2, 1), gain=4).shape),
      tuple(ops.bias_act(torch.randn(2, 3, 9, 13, device=dev), torch.randn(3, device=dev), act="lrelu", clamp=1.0).shape))
sr = 24576
y = torch.from_numpy((0.3 * np.random.default_rng(0).standard_normal(sr * 3)).astype(np.float32)).to(dev)
print("audio", ar.onsets_rms(y, sr)[0].shape, ar.chromagram(y, sr).shape, ss.pulse(y, sr).shape, float(ss.quantile(y, 0.25)))
print("sosfilt", ar.low_pass(y[:50001], sr, 200, 12).shape, "rgb24", frames_to_rgb24(torch.rand(2, 3, 30, 44, device=dev)).shape)
torch.cuda.synchronize()
print("sanitize smoke done")
